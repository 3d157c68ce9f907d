"""CPU tests of the scan front end: the oracle's restatements are pinned by properties and by an
independent numpy model; the local-map cube (host arithmetic in the product library) is compared with the oracle."""
import numpy as np
import pytest

from fast_lio_b200 import api, synth
from oracle import bind
from semantics import voxelgrid_model


@pytest.fixture(scope="module")
def raw():
    scene = synth._scene_for(50_000)
    return synth.make_raw_scan(scene, 6000, synth.true_state("velodyne"), seed=11)


# ------------------------------------------------------------------------------------------------ de-skew
def test_undistort_recovers_end_frame(raw):
    """The sensor moves with the motion model of the backward pass (constant body rate, constant world acceleration):
    compensated points must be the surface points seen from the end pose (IMU_Processing.hpp:330-336)."""
    pts, ts = bind.undistort(raw.xyzi, raw.offset_ms, raw.imu_pose, raw.x_end)
    order = np.argsort(raw.offset_ms, kind="stable")
    assert np.array_equal(ts, raw.offset_ms[order])
    assert np.array_equal(pts[:, 3], raw.xyzi[order][:, 3])          # intensity untouched
    err = np.abs(pts[:, :3] - raw.truth_end[order]).max(axis=1)
    assert err[1:].max() < 5e-6
    # the point at offset 0 is not older than any IMU pose: the sweep leaves it as measured (:325 uses a strict >)
    assert ts[0] == 0.0 and np.array_equal(pts[0], raw.xyzi[order][0])
    assert np.abs(raw.xyzi[order][:, :3] - raw.truth_end[order]).max() > 0.1   # and the motion did distort the scan


def test_undistort_degenerate_inputs(raw):
    one_pose = raw.imu_pose[:1]
    pts, _ = bind.undistort(raw.xyzi, raw.offset_ms, one_pose, raw.x_end)   # no segment: only the sort happens
    assert np.array_equal(pts, raw.xyzi[np.argsort(raw.offset_ms, kind="stable")])
    pts, ts = bind.undistort(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), raw.imu_pose, raw.x_end)
    assert len(pts) == 0
    still = raw.imu_pose.copy()                                       # zero rate: Exp() takes its identity branch
    still[:, 4:7] = 0.0
    pts, _ = bind.undistort(raw.xyzi, raw.offset_ms, still, raw.x_end)
    assert np.isfinite(pts).all()


def test_undistort_first_point_quirk(raw):
    """IMU samples between the previous frame's end and this frame's first point carry NEGATIVE offsets (:299).  The
    sweep breaks on the first point and re-tests it against every earlier segment (:343), so it is compensated again
    by each of them.  The restatement keeps that."""
    poses = np.vstack([raw.imu_pose[:1], raw.imu_pose[:1], raw.imu_pose[1:]])
    poses[1, 0] = -0.004
    pts, ts = bind.undistort(raw.xyzi, raw.offset_ms, poses, raw.x_end)
    order = np.argsort(raw.offset_ms, kind="stable")
    x0 = raw.xyzi[order][:1].copy()
    once, _ = bind.undistort(x0, ts[:1], poses[1:3], raw.x_end)       # the only segment whose head (-4 ms) is older than t = 0
    assert np.array_equal(pts[0], once[0]) and not np.array_equal(pts[0], x0[0])
    twice_in = np.vstack([x0, x0])
    late = np.array([ts[0] + 7.0, ts[0] + 7.0], dtype=np.float32)      # a first point at 7 ms sees heads at 5, -4 and 0 ms...
    got, _ = bind.undistort(twice_in, late, poses, raw.x_end)
    assert not np.array_equal(got[0], got[1])                          # ...so point 0 is compensated repeatedly, point 1 once


# ------------------------------------------------------------------------------------------------ voxel grid
@pytest.mark.parametrize("leaf", [0.5, 0.2, 1.0])
def test_voxelgrid_matches_numpy_model(raw, leaf):
    pts, _ = bind.undistort(raw.xyzi, raw.offset_ms, raw.imu_pose, raw.x_end)
    got = bind.voxelgrid(pts, leaf)
    want = voxelgrid_model(pts, leaf)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert len(got) < len(pts)


def test_voxelgrid_properties(raw):
    pts = raw.xyzi
    out = bind.voxelgrid(pts, 0.5)
    # one output per occupied cell, inside its cell, in ascending cell index
    inv = np.float32(2.0)
    cells = np.floor(pts[:, :3] * inv).astype(np.int64)
    assert len(out) == len(np.unique(cells, axis=0))
    oc = np.floor(out[:, :3].astype(np.float64) * 2.0 + 1e-4).astype(np.int64)
    lo = cells.min(axis=0)
    div = cells.max(axis=0) - lo + 1
    key = (oc[:, 0] - lo[0]) + (oc[:, 1] - lo[1]) * div[0] + (oc[:, 2] - lo[2]) * div[0] * div[1]
    assert (np.diff(key) >= 0).all()
    # idempotent on its own output up to cell membership: every centroid is alone in its cell
    again = bind.voxelgrid(out, 0.5)
    assert len(again) == len(out)
    assert abs(float(out[:, 3].mean()) - float(pts[:, 3].mean())) < 5.0


def test_voxelgrid_edge_cases():
    assert len(bind.voxelgrid(np.zeros((0, 4), np.float32), 0.5)) == 0
    one = np.array([[1.25, -3.5, 0.75, 9.0]], np.float32)
    assert np.array_equal(bind.voxelgrid(one, 0.5), one)
    same_cell = np.array([[0.1, 0.1, 0.1, 1.0], [0.2, 0.3, 0.4, 3.0], [0.4, 0.2, 0.3, 5.0]], np.float32)
    out = bind.voxelgrid(same_cell, 0.5)
    assert out.shape == (1, 4) and np.array_equal(out, voxelgrid_model(same_cell, 0.5))
    neg = np.array([[-0.1, -0.1, -0.1, 1.0], [0.1, 0.1, 0.1, 2.0], [-0.6, -0.1, -0.1, 3.0]], np.float32)
    assert np.array_equal(bind.voxelgrid(neg, 0.5), voxelgrid_model(neg, 0.5)) and len(bind.voxelgrid(neg, 0.5)) == 3
    # PCL's "leaf size is too small for the input dataset" exit returns the input unchanged
    far = np.array([[0, 0, 0, 1.0], [300, 300, 300, 2.0], [1, 1, 1, 3.0]], np.float32)      # 3001^3 cells > INT_MAX
    assert np.array_equal(bind.voxelgrid(far, 0.1), far)
    assert np.array_equal(voxelgrid_model(far, 0.1), far)


# ------------------------------------------------------------------------------------------------ local-map cube
def _walk(seed, n, step):
    rng = np.random.default_rng(seed)
    drift = rng.normal(0, 1, 3)
    drift /= np.linalg.norm(drift)
    pos = np.array([3.0, -2.0, 0.5])
    for _ in range(n):
        pos = pos + step * (drift + 0.5 * rng.normal(0, 1, 3))
        yield pos.copy()


@pytest.mark.parametrize("cube_len,det_range", [(1000.0, 100.0), (1000.0, 450.0), (200.0, 300.0), (2000.0, 260.0)])
def test_localmap_cube_matches_oracle(cube_len, det_range):
    """fl_localmap_segment (host arithmetic, no GPU needed when no map is attached) vs the restated
    lasermap_fov_segment, box for box, over a random walk that crosses the cube several times."""
    ours = api.LocalMap(cube_len, det_range)
    ref = bind.LocalMap(cube_len, det_range)
    moved = 0
    for pos in _walk(4, 400, cube_len / 40.0):
        b_ref = ref.segment(pos)
        b_ours, n_deleted = ours.segment(pos, None)
        assert n_deleted == 0
        assert b_ours.shape == b_ref.shape and np.array_equal(b_ours, b_ref)
        assert np.array_equal(ours.box(), ref.box())
        moved += len(b_ref)
        box = ref.box()
        assert np.allclose(box[3:] - box[:3], cube_len, rtol=1e-5)
    assert moved > 0


def test_localmap_cube_semantics():
    lm = bind.LocalMap(1000.0, 100.0)
    assert len(lm.segment([0.0, 0.0, 0.0])) == 0                      # first call only centres the cube (:238-245)
    assert np.array_equal(lm.box(), np.array([-500, -500, -500, 500, 500, 500], np.float32))
    assert len(lm.segment([100.0, 0.0, 0.0])) == 0                    # far from every face: nothing moves
    boxes = lm.segment([360.0, 0.0, 0.0])                             # within 1.5 * det_range of the +x face
    step = max((1000.0 - 2 * 1.5 * 100.0) * 0.5 * 0.9, 100.0 * 0.5)   # :256
    assert boxes.shape == (1, 6)
    assert np.array_equal(boxes[0], np.array([-500, -500, -500, -500 + step, 500, 500], np.float32))   # the slab left behind
    assert np.array_equal(lm.box(), np.array([-500 + step, -500, -500, 500 + step, 500, 500], np.float32))
    with pytest.raises(api.FastLioError):
        api.LocalMap(-1.0, 100.0)
