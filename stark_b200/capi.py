"""ctypes binding of the C-ABI in include/stark_b200.h (libstark_b200.so, hand-written CUDA for sm_100a).

This is plumbing for tests and bench.py: numpy host buffers in, numpy host buffers out.  There is no CPU fallback --
if the shared library is missing or no GPU is visible the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libstark_b200.so")
_lib = None


class SBError(RuntimeError):
    pass


class Fetch(C.Structure):
    _fields_ = [("array", C.c_int32), ("conn_col", C.c_int32), ("first_slot", C.c_int32), ("stride", C.c_int32)]


class Op(C.Structure):
    """sb_op: one operation of a symx sequence (include/stark_b200.h, user potentials)."""
    _fields_ = [("type", C.c_int32), ("dst", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("cond", C.c_int32), ("pad", C.c_int32), ("constant", C.c_double)]


def ops_array(ints, consts):
    """(n x 5 int32 [type, dst, a, b, cond], n float64 constants) -> ctypes array of sb_op."""
    arr = (Op * len(ints))()
    for i, (row, c) in enumerate(zip(ints, consts)):
        arr[i] = Op(int(row[0]), int(row[1]), int(row[2]), int(row[3]), int(row[4]), 0, float(c))
    return arr


class ContactMesh(C.Structure):
    _fields_ = [
        ("physical_system", C.c_int32), ("rigid_body", C.c_int32), ("n_vertices", C.c_int32), ("n_triangles", C.c_int32),
        ("n_edges", C.c_int32), ("vertex_global", C.POINTER(C.c_int32)), ("vertices_local", C.POINTER(C.c_double)),
        ("triangles", C.POINTER(C.c_int32)), ("edges", C.POINTER(C.c_int32)), ("contact_thickness", C.c_double),
    ]


class ContactBindings(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("soft_v1", "soft_x0", "soft_X", "rb_v1", "rb_w1", "rb_t0", "rb_q0", "dt")]


class NewtonSettings(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_int32), ("min_iterations", C.c_int32),
        ("residual_tolerance_abs", C.c_double), ("residual_tolerance_rel", C.c_double), ("step_tolerance", C.c_double),
        ("max_iterations_as_success", C.c_int32),
        ("step_cap", C.c_double),
        ("enable_armijo_backtracking", C.c_int32),
        ("line_search_armijo_beta", C.c_double),
        ("max_backtracking_armijo_iterations", C.c_int32), ("max_backtracking_invalid_state_iterations", C.c_int32),
        ("projection_mode", C.c_int32),
        ("projection_eps", C.c_double),
        ("project_to_pd_use_mirroring", C.c_int32), ("project_on_demand_countdown", C.c_int32),
        ("ppn_tightening_factor", C.c_double), ("ppn_release_factor", C.c_double),
        ("linear_solver", C.c_int32), ("cg_max_iterations", C.c_int32),
        ("cg_abs_tolerance", C.c_double), ("cg_rel_tolerance", C.c_double),
        ("cg_stop_on_indefiniteness", C.c_int32),
        ("bailout_residual", C.c_double),
        ("contact_enabled", C.c_int32),
        ("skip_converged_state_check", C.c_int32),
        ("intersection_test_enabled", C.c_int32),
    ]


class NewtonStats(C.Structure):
    _fields_ = [
        ("result", C.c_int32), ("newton_iterations", C.c_int32), ("cg_iterations", C.c_int32),
        ("ls_cap_iterations", C.c_int32), ("ls_max_iterations", C.c_int32), ("ls_inv_iterations", C.c_int32), ("ls_bt_iterations", C.c_int32),
        ("n_hessians", C.c_int64), ("n_projected_hessians", C.c_int64),
        ("n_evaluations", C.c_int32),
        ("last_residual", C.c_double), ("last_energy", C.c_double),
        ("residuals", C.c_double * 64),
        ("gpu_ms", C.c_double),
    ]


# every symbol include/stark_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "sb_create", "sb_destroy", "sb_last_error", "sb_get_stream", "sb_synchronize", "sb_launch_count",
    "sb_array_create", "sb_array_upload", "sb_array_download", "sb_array_rows", "sb_array_fill", "sb_array_axpy", "sb_array_copy", "sb_array_download_async", "sb_download_wait", "sb_host_register", "sb_host_unregister",
    "sb_dof_add", "sb_dof_total", "sb_dofs_get", "sb_dofs_set",
    "sb_potential_create", "sb_potential_set_connectivity", "sb_potential_info", "sb_kernel_names",
    "sb_eval", "sb_eval_prelaunch", "sb_grad_get", "sb_potential_get_element_output", "sb_potential_get_block_rows", "sb_potential_get_hessians",
    "sb_project_to_pd", "sb_assemble", "sb_bcsr_info", "sb_bcsr_get",
    "sb_solve_pcg", "sb_solve_llt", "sb_llt_stats", "sb_llt_order", "sb_du_get", "sb_dofs_save", "sb_dofs_apply_step", "sb_du_scale",
    "sb_contact_init", "sb_contact_add_mesh", "sb_contact_blacklist", "sb_contact_set_friction", "sb_contact_set_params",
    "sb_contact_update", "sb_contact_update_friction", "sb_contact_begin_time_step", "sb_contact_count_intersections", "sb_contact_get_proximity",
    "sb_contact_get_vertices", "sb_contact_set_vertices", "sb_contact_detect", "sb_contact_potential",
    "sb_newton_timer_begin", "sb_newton_default_settings", "sb_newton_solve", "sb_profile_potential",
    "sb_profile_stages", "sb_profile_report",
    "sb_dist_init", "sb_dist_connect", "sb_dist_connect_ptrs", "sb_dist_local_base", "sb_dist_stats", "sb_dist_plan", "sb_dist_set_enabled",
    "sb_potential_create_user", "sb_user_codegen", "sb_user_compile",
]


def load():
    """dlopen libstark_b200.so (raises if it has not been built -- see __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SBError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    lib.sb_last_error.restype = C.c_char_p
    lib.sb_kernel_names.restype = C.c_char_p
    lib.sb_profile_report.restype = C.c_char_p
    lib.sb_get_stream.restype = C.c_void_p
    lib.sb_dist_local_base.restype = C.c_void_p
    lib.sb_dist_local_base.argtypes = [C.c_void_p]
    lib.sb_dist_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_void_p]
    lib.sb_dist_connect.argtypes = [C.c_void_p, C.c_void_p]
    lib.sb_dist_connect_ptrs.argtypes = [C.c_void_p, C.c_void_p]
    lib.sb_launch_count.restype = C.c_int64
    lib.sb_destroy.restype = None
    lib.sb_newton_default_settings.restype = None
    lib.sb_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    _lib = lib
    return lib


def kernel_names():
    return [n for n in load().sb_kernel_names().decode().split("\n") if n]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Context:
    """One sb_context.  `stream` may be a raw cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream)."""

    def __init__(self, device=0, stream=None):
        self.lib = load()
        h = C.c_void_p()
        r = self.lib.sb_create(C.byref(h), int(device), C.c_void_p(stream) if stream else None)
        if r != 0:
            raise SBError(f"sb_create failed with status {r}: no usable CUDA device (this library has no CPU fallback)")
        self.h = h

    @classmethod
    def borrow(cls, handle):
        """Wrap an sb_context owned by someone else (e.g. a host-layer scene): close() does not destroy it."""
        self = cls.__new__(cls)
        self.lib = load()
        self.h = C.c_void_p(handle)
        self._borrowed = True
        return self

    def close(self):
        if self.h and not getattr(self, "_borrowed", False):
            self.lib.sb_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, r):
        if r != 0:
            raise SBError(f"status {r}: {self.lib.sb_last_error(self.h).decode()}")

    # ---- arrays / dofs
    def array(self, label, stride, data=None):
        out = C.c_int()
        self._ck(self.lib.sb_array_create(self.h, label.encode(), int(stride), C.byref(out)))
        if data is not None:
            self.upload(out.value, data)
        return out.value

    def upload(self, array, data):
        d = _f64(data)
        stride = d.shape[1] if d.ndim == 2 else 1
        n_rows = d.shape[0] if d.ndim >= 1 else 1
        d = d.reshape(-1)
        self._ck(self.lib.sb_array_upload(self.h, int(array), _p(d, C.c_double), int(n_rows)))
        return stride

    def download(self, array, n_rows, stride):
        out = np.empty((n_rows, stride), dtype=np.float64)
        self._ck(self.lib.sb_array_download(self.h, int(array), _p(out, C.c_double), int(n_rows)))
        return out

    def dof_add(self, array):
        out = C.c_int()
        self._ck(self.lib.sb_dof_add(self.h, int(array), C.byref(out)))
        return out.value

    def ndofs(self):
        out = C.c_int()
        self._ck(self.lib.sb_dof_total(self.h, C.byref(out)))
        return out.value

    def dofs_get(self):
        u = np.empty(self.ndofs(), dtype=np.float64)
        self._ck(self.lib.sb_dofs_get(self.h, _p(u, C.c_double)))
        return u

    def dofs_set(self, u):
        u = _f64(u)
        self._ck(self.lib.sb_dofs_set(self.h, _p(u, C.c_double)))

    # ---- potentials
    def potential(self, name, conn_stride, fetch):
        arr = (Fetch * len(fetch))(*[Fetch(*map(int, f)) for f in fetch])
        out = C.c_int()
        self._ck(self.lib.sb_potential_create(self.h, name.encode(), int(conn_stride), arr, len(fetch), C.byref(out)))
        return out.value

    def potential_user(self, name, conn_stride, fetch, n_in, block_slots, ops_p, ops_pgh):
        """A potential without a built-in kernel: its symx operation sequences go through the Sequence -> CUDA -> NVRTC back-end."""
        arr = (Fetch * len(fetch))(*[Fetch(*map(int, f)) for f in fetch])
        slots = (C.c_int32 * len(block_slots))(*[int(x) for x in block_slots])
        p, pgh = ops_array(*ops_p), ops_array(*ops_pgh)
        out = C.c_int()
        self._ck(self.lib.sb_potential_create_user(self.h, name.encode(), int(conn_stride), arr, len(fetch), int(n_in), len(block_slots), slots,
                                                   p, len(p), pgh, len(pgh), C.byref(out)))
        return out.value

    def set_connectivity(self, pot, conn):
        c = _i32(conn)
        n = c.shape[0] if c.ndim == 2 else 0
        self._ck(self.lib.sb_potential_set_connectivity(self.h, int(pot), _p(c, C.c_int32), int(n)))

    def potential_info(self, pot):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.sb_potential_info(self.h, int(pot), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    # ---- evaluation
    def eval(self, mode="PGH"):
        E, r = C.c_double(), C.c_double()
        self._ck(self.lib.sb_eval(self.h, 2 if mode == "PGH" else 0, C.byref(E), C.byref(r)))
        return (E.value, r.value) if mode == "PGH" else E.value

    def grad(self):
        g = np.empty(self.ndofs(), dtype=np.float64)
        self._ck(self.lib.sb_grad_get(self.h, _p(g, C.c_double)))
        return g

    def element_output(self, pot):
        n_in, n, ne = self.potential_info(pot)
        out = np.empty((ne, 1 + n + n * n), dtype=np.float64)
        self._ck(self.lib.sb_potential_get_element_output(self.h, int(pot), _p(out, C.c_double)))
        return out

    def block_rows(self, pot):
        n_in, n, ne = self.potential_info(pot)
        out = np.empty((ne, n // 3), dtype=np.int32)
        self._ck(self.lib.sb_potential_get_block_rows(self.h, int(pot), _p(out, C.c_int32)))
        return out

    def hessians(self, pot):
        n_in, n, ne = self.potential_info(pot)
        out = np.empty((ne, n, n), dtype=np.float64)
        self._ck(self.lib.sb_potential_get_hessians(self.h, int(pot), _p(out, C.c_double)))
        return out

    # ---- projection / assembly / solve
    def project_to_pd(self, grad_threshold, eps=1e-10, mirror=False):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int()
        self._ck(self.lib.sb_project_to_pd(self.h, C.c_double(grad_threshold), C.c_double(eps), int(mirror), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, bool(c.value)

    def assemble(self):
        self._ck(self.lib.sb_assemble(self.h))

    def bcsr(self):
        nbr, nnzb = C.c_int(), C.c_int64()
        self._ck(self.lib.sb_bcsr_info(self.h, C.byref(nbr), C.byref(nnzb)))
        rows = np.empty(nbr.value + 1, dtype=np.int64)
        cols = np.empty(nnzb.value, dtype=np.int32)
        vals = np.empty(nnzb.value * 9, dtype=np.float32)
        self._ck(self.lib.sb_bcsr_get(self.h, _p(rows, C.c_int64), _p(cols, C.c_int32), _p(vals, C.c_float)))
        return rows, cols, vals

    def solve_pcg(self, abs_tol, rel_tol=1e-4, max_iterations=10000, stop_on_indefiniteness=True):
        it, ok, dg, di = C.c_int(), C.c_int(), C.c_double(), C.c_double()
        self._ck(self.lib.sb_solve_pcg(self.h, C.c_double(abs_tol), C.c_double(rel_tol), int(max_iterations), int(stop_on_indefiniteness),
                                       C.byref(it), C.byref(ok), C.byref(dg), C.byref(di)))
        return dict(iterations=it.value, ok=bool(ok.value), du_dot_grad=dg.value, du_inf=di.value)

    def solve_llt(self):
        ok, dg, di = C.c_int(), C.c_double(), C.c_double()
        self._ck(self.lib.sb_solve_llt(self.h, C.byref(ok), C.byref(dg), C.byref(di)))
        return dict(ok=bool(ok.value), du_dot_grad=dg.value, du_inf=di.value)

    def llt_stats(self):
        o = (C.c_double * 5)()
        self._ck(self.lib.sb_llt_stats(self.h, o))
        return dict(tiles=int(o[0]), tile_rows=int(o[1]), factor_bytes=int(o[2]), orderings=int(o[3]), analyses=int(o[4]))

    def du(self):
        d = np.empty(self.ndofs(), dtype=np.float64)
        self._ck(self.lib.sb_du_get(self.h, _p(d, C.c_double)))
        return d

    def dofs_save(self):
        self._ck(self.lib.sb_dofs_save(self.h))

    def dofs_apply_step(self, step):
        self._ck(self.lib.sb_dofs_apply_step(self.h, C.c_double(step)))

    # ---- contact
    def contact_init(self, **b):
        cb = ContactBindings(**{k: int(v) for k, v in b.items()})
        self._ck(self.lib.sb_contact_init(self.h, C.byref(cb)))

    def contact_add_mesh(self, physical_system, rigid_body, vertex_global, vertices_local, triangles, edges, contact_thickness):
        tri = _i32(triangles).reshape(-1, 3)
        edg = _i32(edges).reshape(-1, 2)
        m = ContactMesh()
        m.physical_system, m.rigid_body = int(physical_system), int(rigid_body)
        keep = [tri, edg]
        if physical_system == 0:
            vg = _i32(vertex_global)
            keep.append(vg)
            m.n_vertices = len(vg)
            m.vertex_global = _p(vg, C.c_int32)
        else:
            vl = _f64(vertices_local).reshape(-1, 3)
            keep.append(vl)
            m.n_vertices = len(vl)
            m.vertices_local = _p(vl, C.c_double)
        m.n_triangles, m.n_edges = len(tri), len(edg)
        m.triangles, m.edges = _p(tri, C.c_int32), _p(edg, C.c_int32)
        m.contact_thickness = float(contact_thickness)
        out = C.c_int()
        self._ck(self.lib.sb_contact_add_mesh(self.h, C.byref(m), C.byref(out)))
        return out.value

    def contact_blacklist(self, a, b):
        self._ck(self.lib.sb_contact_blacklist(self.h, int(a), int(b)))

    def contact_set_friction(self, a, b, mu):
        self._ck(self.lib.sb_contact_set_friction(self.h, int(a), int(b), C.c_double(mu)))

    def contact_set_params(self, contact_stiffness, friction_stick_slide_threshold=0.1, point_triangle=True, edge_edge=True, friction=True):
        self._ck(self.lib.sb_contact_set_params(self.h, C.c_double(contact_stiffness), C.c_double(friction_stick_slide_threshold),
                                                int(point_triangle), int(edge_edge), int(friction)))

    def contact_update(self):
        self._ck(self.lib.sb_contact_update(self.h))

    def contact_update_friction(self):
        self._ck(self.lib.sb_contact_update_friction(self.h))

    def contact_count_intersections(self):
        out = C.c_int()
        self._ck(self.lib.sb_contact_count_intersections(self.h, C.byref(out)))
        return out.value

    def contact_set_vertices(self, group, xyz):
        x = _f64(xyz)
        self._ck(self.lib.sb_contact_set_vertices(self.h, int(group), _p(x, C.c_double)))

    def contact_get_vertices(self, group, n):
        x = np.empty((n, 3), dtype=np.float64)
        self._ck(self.lib.sb_contact_get_vertices(self.h, int(group), _p(x, C.c_double)))
        return x

    def contact_detect(self, enlargement, with_intersections=True):
        self._ck(self.lib.sb_contact_detect(self.h, C.c_double(enlargement), int(with_intersections)))

    def contact_proximity(self, kind):
        n, w = C.c_int(), C.c_int()
        self._ck(self.lib.sb_contact_get_proximity(self.h, int(kind), None, None, 0, C.byref(n), C.byref(w)))
        ids = np.empty((n.value, w.value), dtype=np.int32)
        dist = np.empty(n.value, dtype=np.float64)
        if n.value:
            self._ck(self.lib.sb_contact_get_proximity(self.h, int(kind), _p(ids, C.c_int32), _p(dist, C.c_double), n.value, C.byref(n), C.byref(w)))
        return ids, dist

    def contact_potential(self, name):
        out = C.c_int()
        self._ck(self.lib.sb_contact_potential(self.h, name.encode(), C.byref(out)))
        return out.value

    # ---- newton
    def newton_default_settings(self):
        s = NewtonSettings()
        self.lib.sb_newton_default_settings(C.byref(s))
        return s

    def newton_solve(self, settings):
        st = NewtonStats()
        self._ck(self.lib.sb_newton_solve(self.h, C.byref(settings), C.byref(st)))
        return st

    def launch_count(self):
        return int(self.lib.sb_launch_count(self.h))

    def synchronize(self):
        self._ck(self.lib.sb_synchronize(self.h))

    def stream(self):
        return self.lib.sb_get_stream(self.h)
