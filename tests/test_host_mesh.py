"""CPU-side checks of the host layer's mesh helpers (the presets' generators: S/utils/mesh_generators.cpp:100-167, 264-377,
S/utils/mesh_utils.cpp:217-252, 278-327), through exports of the host library that need no GPU.  Their outputs are also
compared bit for bit with the reference's arrays in the GPU tests (tests/test_scene_e2e.py); here: counts and invariants."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host():
    from stark_b200 import capi
    C.CDLL(capi.LIB_PATH, mode=C.RTLD_GLOBAL)
    h = C.CDLL(os.path.join(ROOT, "stark_b200", "lib", "libstark_b200_host.so"))
    h.sbh_mesh_triangle_grid.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    h.sbh_mesh_internal_angles.argtypes = [C.POINTER(C.c_int32), C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int]
    h.sbh_mesh_tet_grid.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return h


def _tri_grid(h, n0, n1, dx=0.4, dy=0.4):
    V = np.zeros(((n0 + 1) * (n1 + 1), 3))
    T = np.zeros((2 * n0 * n1, 3), dtype=np.int32)
    nt = h.sbh_mesh_triangle_grid(n0, n1, dx, dy, V.ctypes.data_as(C.POINTER(C.c_double)), T.ctypes.data_as(C.POINTER(C.c_int32)))
    assert nt == 2 * n0 * n1
    return V, T


def test_triangle_grid_counts_area_and_orientation():
    h = _host()
    for n0, n1 in [(1, 1), (8, 8), (5, 3), (32, 32)]:
        V, T = _tri_grid(h, n0, n1)
        assert V[:, 2].max() == 0.0 and abs(V[:, 0].min() + 0.2) < 1e-15 and abs(V[:, 0].max() - 0.2) < 1e-15
        a, b, c = V[T[:, 0]], V[T[:, 1]], V[T[:, 2]]
        nz = np.cross(b - a, c - a)[:, 2]
        assert (nz > 0).all()                                  # homogeneous normals (+z)
        assert abs(0.5 * nz.sum() - 0.16) < 1e-14             # the triangles tile the 0.4 x 0.4 sheet
        assert sorted(set(T.ravel())) == list(range(len(V)))   # every vertex used


def test_internal_angles_are_the_interior_edges_with_their_opposite_vertices():
    h = _host()
    for n0, n1 in [(1, 1), (2, 2), (8, 8), (7, 4)]:
        V, T = _tri_grid(h, n0, n1)
        cap = 3 * len(T)
        out = np.zeros((cap, 4), dtype=np.int32)
        nh = h.sbh_mesh_internal_angles(T.ctypes.data_as(C.POINTER(C.c_int32)), len(T), len(V), out.ctypes.data_as(C.POINTER(C.c_int32)), cap)
        # a triangulated n0 x n1 grid has 3 n0 n1 + n0 + n1 edges, 2 (n0 + n1) of them on the boundary
        assert nh == 3 * n0 * n1 - n0 - n1
        out = out[:nh]
        tris = {tuple(sorted(t)) for t in T.tolist()}
        for e0, e1, o0, o1 in out.tolist():
            assert e0 < e1 and o0 < o1 and len({e0, e1, o0, o1}) == 4
            assert tuple(sorted((e0, e1, o0))) in tris and tuple(sorted((e0, e1, o1))) in tris   # the two triangles of the hinge
        keys = out[:, 0].astype(np.int64) * len(V) + out[:, 1]
        assert (np.diff(keys) > 0).all()                       # sorted by edge, unique (the reference's order)


def test_tet_grid_counts_volume_and_surface():
    h = _host()
    for n in [(1, 1, 1), (2, 3, 4), (6, 6, 6)]:
        nx, ny, nz = n
        n_hex = nx * ny * nz
        nv_expect = (nx + 1) * (ny + 1) * (nz + 1) + n_hex          # one centre node per hexahedron
        V = np.zeros((nv_expect, 3))
        T = np.zeros((12 * n_hex, 4), dtype=np.int32)
        nv, ns = C.c_int(), C.c_int()
        nt = h.sbh_mesh_tet_grid(nx, ny, nz, 1.0, 1.5, 2.0, V.ctypes.data_as(C.POINTER(C.c_double)), T.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(nv), C.byref(ns))
        assert nt == 12 * n_hex and nv.value == nv_expect
        p = V[T]
        vol = np.einsum("ij,ij->i", np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), p[:, 3] - p[:, 0]) / 6.0
        assert (np.abs(vol) > 0).all()
        assert abs(np.abs(vol).sum() - 3.0) < 1e-12                   # the tets fill the 1 x 1.5 x 2 box
        assert ns.value == 4 * (nx * ny + ny * nz + nx * nz)          # two triangles per boundary quad
