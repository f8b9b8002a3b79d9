"""ctypes binding of the C++ host layer (libstark_b200_host.so: stark_b200/host/*.cpp) -- the scene-level calls a user of
the reference makes (build scene, run_one_time_step, read stats).  Plumbing only."""
import ctypes as C
import os

import numpy as np

from . import capi

HOST_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libstark_b200_host.so")
_host = None


def load():
    global _host
    if _host is None:
        capi.load()
        if not os.path.exists(HOST_LIB_PATH):
            raise capi.SBError(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build()")
        h = C.CDLL(HOST_LIB_PATH)
        h.sbh_scene_create.restype = C.c_void_p
        h.sbh_scene_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p]
        h.sbh_scene_destroy.argtypes = [C.c_void_p]
        h.sbh_scene_step.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        h.sbh_scene_residuals.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        h.sbh_scene_totals.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        h.sbh_scene_positions.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        h.sbh_scene_potential.argtypes = [C.c_void_p, C.c_char_p]
        h.sbh_scene_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.c_int]
        h.sbh_scene_connectivity.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int32), C.c_int]
        h.sbh_scene_sync.argtypes = [C.c_void_p]
        h.sbh_scene_set_linear_solver.argtypes = [C.c_void_p, C.c_int]
        h.sbh_scene_context.restype = C.c_void_p
        h.sbh_scene_context.argtypes = [C.c_void_p]
        _host = h
    return _host


STEP_FIELDS = ["keep_going", "accepted", "result", "newton_iterations", "cg_iterations", "evaluations", "dt", "runtime_s", "solve_s",
               "first_residual", "ls_inv", "ls_bt", "time", "ndofs", "contact_stiffness", "solve_gpu_ms"]
TOTAL_FIELDS = ["nodes", "tets", "ndofs", "h2d_bytes", "d2h_bytes", "launches", "newton_iterations", "solve_s"]


class Scene:
    """One BASELINE.json configuration: 'tetdrop' (C2: n^3 Soft_Rubber tet grid onto a fixed rigid floor, IPC + friction),
    'tetbar' (C5: prescribed twisted bar, no contact), 'tetchain' (C4: tet block + hinged chain of boxes), 'cloth' /
    'cloth_shells' (C1 / C3: n x n Cotton_Fabric grid over a scripted rigid box; flat-bending or discrete-shell hinges)."""

    def __init__(self, name, n, ny=-1, nz=-1, dt=0.01, drop=0.003, vz=0.0, device=0, stream=None, llt=False):
        self.lib = load()
        self.h = self.lib.sbh_scene_create(name.encode(), n, ny, nz, dt, drop, vz, device, C.c_void_p(stream) if stream else None)
        if not self.h:
            raise capi.SBError(f"unknown scene {name}")
        if llt:
            self.lib.sbh_scene_set_linear_solver(self.h, 0)

    def step(self):
        out = (C.c_double * 16)()
        self.lib.sbh_scene_step(self.h, out)
        return dict(zip(STEP_FIELDS, list(out)))

    def sync(self):
        """Wait for the asynchronous read-backs of the last step: the host mirrors of positions / velocities are then current."""
        self.lib.sbh_scene_sync(self.h)

    def residuals(self):
        buf = (C.c_double * 64)()
        n = self.lib.sbh_scene_residuals(self.h, buf, 64)
        return np.array(buf[:n])

    def totals(self):
        out = (C.c_double * 8)()
        self.lib.sbh_scene_totals(self.h, out)
        return dict(zip(TOTAL_FIELDS, list(out)))

    def array(self, label):
        """Host copy of a bound array by label (flat)."""
        n = self.lib.sbh_scene_array(self.h, label.encode(), None, 0)
        if n < 0:
            raise KeyError(label)
        buf = np.empty(n)
        self.lib.sbh_scene_array(self.h, label.encode(), buf.ctypes.data_as(C.POINTER(C.c_double)), n)
        return buf

    def connectivity(self, potential_name):
        n = self.lib.sbh_scene_connectivity(self.h, potential_name.encode(), None, 0)
        if n < 0:
            raise KeyError(potential_name)
        buf = np.empty(n, dtype=np.int32)
        self.lib.sbh_scene_connectivity(self.h, potential_name.encode(), buf.ctypes.data_as(C.POINTER(C.c_int32)), n)
        return buf

    def positions(self):
        n = int(self.totals()["nodes"])
        x = np.empty((n, 3))
        self.lib.sbh_scene_positions(self.h, x.ctypes.data_as(C.POINTER(C.c_double)), n)
        return x

    def close(self):
        if self.h:
            self.lib.sbh_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
