"""kNN parity: the device map against the reference's own ikd-Tree (oracle/_ref) and the port."""
import numpy as np
import pytest

from fast_lio_b200 import api, synth
from oracle import bind

pytestmark = pytest.mark.gpu


def _world_queries(pr):
    """Scan points pushed through the prior pose (float32), as h_share_model does."""
    from oracle.bind import lib
    q = np.zeros((len(pr.scan), 4), dtype=np.float32)
    L = lib()
    tmp = np.zeros(3, dtype=np.float32)
    for i in range(len(pr.scan)):
        L.oracle_transform_point(pr.x_prior, np.ascontiguousarray(pr.scan[i, :3]), tmp)
        q[i, :3] = tmp
    return q


@pytest.mark.parametrize("name", ["tiny", "small", "avia_2k_50k"])
def test_knn_matches_reference_ikdtree(problems, name):
    pr = problems(name)
    q = _world_queries(pr)
    ref = bind.KdTree(pr.map_pts, "auto")
    rp, rd, rc = ref.knn(q, 5)
    t = api.KdTree(0, 0.5)
    t.Build(pr.map_pts)
    assert t.validnum() == len(pr.map_pts)
    gp, gd, gc = t.Nearest_Search(q, 5)
    assert np.array_equal(gc, rc)
    # bit-exact squared distances (float32, same operation order as ikd_Tree.cpp:1683-1688)
    assert np.array_equal(gd, rd)
    # identical neighbour coordinates and payload wherever distances are not tied
    tie = np.zeros(len(q), dtype=bool)
    tie[:] = (np.diff(rd, axis=1) == 0).any(axis=1)
    assert np.array_equal(gp[~tie], rp[~tie])


def test_knn_fewer_than_k_points():
    t = api.KdTree(0, 0.5)
    pts = np.array([[0, 0, 0, 1], [1, 0, 0, 2], [0, 1, 0, 3]], dtype=np.float32)
    t.Build(pts)
    gp, gd, gc = t.Nearest_Search(np.array([[0.1, 0, 0, 0]], dtype=np.float32), 5)
    assert gc[0] == 3
    assert np.allclose(gp[0, 0], [0, 0, 0, 1])
    assert np.isinf(gd[0, 3:]).all()


def test_knn_empty_map():
    t = api.KdTree(0, 0.5)
    gp, gd, gc = t.Nearest_Search(np.zeros((4, 4), dtype=np.float32), 5)
    assert (gc == 0).all()
    t.Build(np.zeros((0, 4), dtype=np.float32))
    gp, gd, gc = t.Nearest_Search(np.zeros((4, 4), dtype=np.float32), 5)
    assert (gc == 0).all() and t.validnum() == 0


def test_knn_brute_force_random():
    rng = np.random.default_rng(5)
    pts = rng.uniform(-20, 20, size=(5000, 4)).astype(np.float32)
    q = rng.uniform(-25, 25, size=(300, 4)).astype(np.float32)
    t = api.KdTree(0, 0.5)
    t.Build(pts)
    gp, gd, gc = t.Nearest_Search(q, 5)
    for i in range(len(q)):
        d = ((q[i, 0] - pts[:, 0]) ** 2 + (q[i, 1] - pts[:, 1]) ** 2) + (q[i, 2] - pts[:, 2]) ** 2   # float32, same order
        order = np.argsort(d, kind="stable")[:5]
        assert np.array_equal(gd[i], d[order])


def test_handles_may_be_destroyed_in_any_order(problems):
    """The C ABI keeps a map alive while a filter or a scan front end created on it exists (Python's cyclic GC may
    finalise the wrappers in any order; a C++ caller's static destructors may too)."""
    import ctypes as C
    pr = problems("tiny")
    L = api.load()
    m, f, s = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert L.fl_map_create(C.byref(m), 0, 0.5) == 0
    assert L.fl_map_build(m, pr.map_pts, len(pr.map_pts)) == 0
    assert L.fl_filter_create(C.byref(f), m, 1000) == 0
    assert L.fl_scan_create(C.byref(s), m) == 0
    assert L.fl_map_destroy(m) == 0                      # the caller lets go first
    x = pr.x_prior.copy(); P = pr.P_prior.copy()
    assert L.fl_filter_update(f, pr.scan, len(pr.scan), x, P, pr.R, None) == 0      # the filter still has its map
    assert L.fl_filter_destroy(f) == 0
    assert L.fl_scan_destroy(s) == 0
    t = api.KdTree(0, 0.5); t.Build(pr.map_pts)           # and nothing is left pending in the CUDA error state
    assert t.validnum() == len(pr.map_pts)
