"""CPU: the round-2 cell-directory model returns the exact k-NN (same distances as a brute-force float32 search and as the
reference's ikd-Tree) and settles within the 27-cell neighbourhood for almost every query of the benchmark scene."""
import numpy as np
import pytest

from fast_lio_b200 import synth
from oracle import bind
from cell_directory_model import CellDirectoryModel
from test_oracle_golden import world_queries


def brute(pts, q, k=5):
    d = q[None, :3].astype(np.float32) - pts[:, :3]
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + d[:, 2] * d[:, 2]
    order = np.argsort(d2, kind="stable")[:k]
    return order, d2[order].astype(np.float32)


@pytest.mark.parametrize("cell", [0.5, 1.0, 2.0])
def test_rings_give_the_exact_knn(problems, cell):
    pr = problems("small")
    m = CellDirectoryModel(pr.map_pts, cell)
    q = world_queries(pr)[::25]
    rings = []
    for qq in q:
        idx, d2, r, _ = m.knn(qq)
        bi, bd = brute(pr.map_pts, qq)
        assert np.array_equal(d2, bd)
        assert set(idx.tolist()) == set(bi.tolist())
        rings.append(r)
    if cell >= 1.0:
        assert np.mean(np.array(rings) <= 1) > 0.95      # the 27-cell neighbourhood settles almost every query


def test_model_matches_reference_ikdtree(problems):
    pr = problems("tiny")
    m = CellDirectoryModel(pr.map_pts, 1.0)
    t = bind.KdTree(pr.map_pts, "auto")
    q = world_queries(pr)[:200]
    _, d_ref, cnt = t.knn(q, 5)
    for i, qq in enumerate(q):
        _, d2, _, _ = m.knn(qq)
        assert cnt[i] == 5 and np.array_equal(d2, d_ref[i])


def test_sparse_and_degenerate_maps():
    pts = np.array([[0, 0, 0, 1], [10, 0, 0, 2], [0, 10, 0, 3]], np.float32)     # fewer than k points: rings stop at the map's extent
    m = CellDirectoryModel(pts, 1.0)
    idx, d2, r, _ = m.knn(np.array([1.0, 1.0, 0.0], np.float32))
    assert len(idx) == 3 and np.array_equal(d2, np.array([2.0, 82.0, 82.0], np.float32))
    far = CellDirectoryModel(np.array([[20.5, 0.2, 0.1, 1]] * 5, np.float32) + np.arange(5, dtype=np.float32)[:, None] * 0.01, 1.0)
    idx, d2, r, _ = far.knn(np.zeros(3, np.float32))
    assert len(idx) == 5 and r >= 19                    # a far query walks rings until it reaches the points
    empty = CellDirectoryModel(np.zeros((0, 4), np.float32), 1.0)
    assert len(empty.knn(np.zeros(3, np.float32))[0]) == 0
