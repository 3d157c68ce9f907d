"""Round-2 prototype (fast_lio_b200/csrc/experimental/cellmap.cu): exact k-NN through the hashed cell directory against the
product's BVH walk and the CPU oracle.  OPT-IN: runs only with FASTLIO_B200_EXPERIMENTAL=1 on a GPU box after
`python scripts/build_experimental.py` -- it has never run on a GPU (written after the round-1 budget was spent)."""
import ctypes as C
import os

import numpy as np
import pytest

from fast_lio_b200 import api
from test_oracle_golden import world_queries

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fast_lio_b200", "libfastlio_b200_experimental.so")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("FASTLIO_B200_EXPERIMENTAL") != "1" or not os.path.exists(LIB),
                                 reason="experimental prototype: opt in with FASTLIO_B200_EXPERIMENTAL=1 after scripts/build_experimental.py")]

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def _lib():
    L = C.CDLL(LIB)
    L.flx_cellmap_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_float]
    L.flx_cellmap_build.argtypes = [C.c_void_p, _f32p, C.c_int]
    L.flx_cellmap_knn.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, _f32p, _i32p, C.POINTER(C.c_float), _i32p]
    L.flx_cellmap_destroy.argtypes = [C.c_void_p]
    return L


def cell_knn(L, pts, q, cell=2.0, k=5):
    h = C.c_void_p()
    assert L.flx_cellmap_create(C.byref(h), 0, cell) == 0
    assert L.flx_cellmap_build(h, np.ascontiguousarray(pts, np.float32), len(pts)) == 0
    out_p = np.zeros((len(q), k, 4), np.float32); out_d = np.zeros((len(q), k), np.float32); out_c = np.zeros(len(q), np.int32)
    rings = np.zeros(8, np.int32); ms = C.c_float(0)
    assert L.flx_cellmap_knn(h, np.ascontiguousarray(q, np.float32), len(q), k, out_p, out_d, out_c, C.byref(ms), rings) == 0
    L.flx_cellmap_destroy(h)
    return out_p, out_d, out_c, rings, ms.value


@pytest.mark.parametrize("name", ["tiny", "small", "avia_2k_50k"])
@pytest.mark.parametrize("cell", [1.0, 2.0])
def test_cell_directory_knn_equals_bvh(problems, name, cell):
    pr = problems(name)
    q = world_queries(pr)
    t = api.KdTree(0, 0.5); t.Build(pr.map_pts)
    p_ref, d_ref, c_ref = t.Nearest_Search(q, 5)
    p, d, c, rings, ms = cell_knn(_lib(), pr.map_pts, q, cell)
    assert np.array_equal(c, c_ref) and np.array_equal(d, d_ref) and np.array_equal(p, p_ref)
    assert rings[1] > 0.9 * len(q)


def test_cell_directory_sparse_maps():
    L = _lib()
    pts = np.array([[0, 0, 0, 1], [10, 0, 0, 2], [0, 10, 0, 3]], np.float32)
    p, d, c, rings, _ = cell_knn(L, pts, np.array([[1, 1, 0, 0]], np.float32))
    assert c[0] == 3 and np.array_equal(d[0, :3], np.array([2.0, 82.0, 82.0], np.float32))
    p, d, c, rings, _ = cell_knn(L, np.zeros((0, 4), np.float32), np.array([[1, 1, 0, 0]], np.float32))
    assert c[0] == 0
