"""Parity of the fused device update against the CPU oracle (north-star tolerance 1e-4 m / 1e-4 rad)."""
import numpy as np
import pytest

from fast_lio_b200 import api, synth
from oracle import bind

pytestmark = pytest.mark.gpu

POS_TOL = 1e-4      # metres   (BASELINE.json north_star)
ROT_TOL = 1e-4      # radians


def rot_err(qa, qb):
    """Angle of qa^-1 * qb (x,y,z,w)."""
    qa = qa / np.linalg.norm(qa); qb = qb / np.linalg.norm(qb)
    d = abs(float(np.dot(qa, qb)))
    return 2.0 * np.arccos(min(1.0, d))


def run_both(pr, solver=0, extr=0):
    ref_tree = bind.KdTree(pr.map_pts, "auto")
    o = bind.update_iterated(ref_tree, pr.scan, pr.x_prior, pr.P_prior, pr.cfg.max_iter, pr.R, pr.limit, extr)
    t = api.KdTree(0, 0.5)
    t.Build(pr.map_pts)
    f = api.Esekf(t, max_points=len(pr.scan), max_iter=pr.cfg.max_iter, limit=pr.limit, extrinsic_est_en=bool(extr), solver=solver)
    x, P, st = f.update_iterated_dyn_share_modified(pr.scan, pr.x_prior, pr.P_prior, pr.R)
    return o, (x, P, st, f)


def check_state(o, x, P):
    assert np.abs(x[0:3] - o.x[0:3]).max() <= POS_TOL
    assert rot_err(x[3:7], o.x[3:7]) <= ROT_TOL
    assert rot_err(x[7:11], o.x[7:11]) <= ROT_TOL
    assert np.abs(x[11:] - o.x[11:]).max() <= POS_TOL
    scale = np.sqrt(np.outer(np.diag(o.P), np.diag(o.P)))
    assert (np.abs(P - o.P) / scale).max() < 1e-3


@pytest.mark.parametrize("name", ["tiny", "small", "avia_2k_50k"])
@pytest.mark.parametrize("solver", [0, 1])
def test_update_matches_oracle(problems, name, solver):
    pr = problems(name)
    o, (x, P, st, f) = run_both(pr, solver)
    logs = f.pass_logs()
    assert len(logs) == len(o.passes)
    for lg, op in zip(logs, o.passes):
        assert lg["searched"] == op["searched"]
        assert lg["effct"] == op["effct"]
        assert lg["converged"] == op["converged"]
        assert abs(lg["res_sum"] - op["res_sum"]) <= 1e-6 * max(1.0, abs(op["res_sum"]))
        assert np.allclose(lg["HtH"], op["HtH"], rtol=1e-9, atol=1e-9 * np.abs(op["HtH"]).max())
        assert np.allclose(lg["Hth"], op["Hth"], rtol=1e-9, atol=1e-9 * np.abs(op["Hth"]).max())
        assert np.abs(lg["x_after"] - op["x_after"]).max() < 1e-6
    check_state(o, x, P)
    npts = len(pr.scan)
    near, cnt = f.nearest(npts)
    assert np.array_equal(cnt, o.nearest_cnt)
    assert np.array_equal(near, o.nearest)
    assert np.array_equal(f.selected(npts), o.selected)
    assert st > 0.0


def test_update_extrinsic_est(problems):
    pr = problems("small")
    o, (x, P, st, f) = run_both(pr, 0, extr=1)
    check_state(o, x, P)
    assert np.allclose(f.pass_logs()[0]["HtH"], o.passes[0]["HtH"], rtol=1e-9, atol=1e-6)
