"""Raw scan -> state: de-skew + voxel grid + iEKF update (+ map_incremental), device-resident chain vs the CPU chain
(oracle front end + the reference's ikd-Tree + the restated update).  Prints one JSON line.
usage: frontend_bench.py [n_raw=120000] [reps=200] [cpu_reps=5]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_lio_b200 import api, synth
from oracle import bind

n_raw = int(sys.argv[1]) if len(sys.argv) > 1 else 120_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
cpu_reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = synth.CONFIGS["velodyne_30k_1m"]
pr = synth.make_problem(cfg.name)
raw = synth.make_raw_scan(pr.scene, n_raw, pr.x_true, seed=31)
api.host_register(raw.xyzi); api.host_register(raw.offset_ms)

g = api.KdTree(0, 0.5); g.Build(pr.map_pts)
f = api.Esekf(g, max_points=n_raw, max_iter=cfg.max_iter, solver=1)
s = api.Scan(g)
def chain():
    s.upload(raw.xyzi, raw.offset_ms)
    s.undistort(raw.imu_pose, raw.x_end)
    n = s.voxel_downsample(0.5)
    x, P, _ = s.update(f, pr.x_prior, pr.P_prior, pr.R)
    return n, x
for _ in range(min(10, reps)):
    n_down, x_gpu = chain()
stages = {}
def timed(name, fn):
    t0 = time.perf_counter(); r = fn(); f._L.fl_filter_sync(f.h); stages[name] = stages.get(name, 0.0) + time.perf_counter() - t0; return r
t_all = time.perf_counter()
for _ in range(reps):
    n_down, x_gpu = chain()
t_all = (time.perf_counter() - t_all) / reps
for _ in range(max(1, reps // 4)):                                    # same chain with a sync after each stage, for the breakdown
    timed("upload", lambda: s.upload(raw.xyzi, raw.offset_ms))
    timed("undistort", lambda: s.undistort(raw.imu_pose, raw.x_end))
    timed("voxelgrid", lambda: s.voxel_downsample(0.5))
    timed("update", lambda: s.update(f, pr.x_prior, pr.P_prior, pr.R))
gpu = {"ms_per_scan": 1e3 * t_all, "scans_per_s": 1.0 / t_all, "n_down": int(n_down),
       "stage_ms_synced": {k: 1e3 * v / max(1, reps // 4) for k, v in stages.items()}}

if cpu_reps < 2:                                                # GPU side only (profiling runs)
    print(json.dumps({"workload": f"raw {n_raw} pts", "gpu": gpu}))
    sys.exit(0)
t = bind.KdTree(pr.map_pts, "auto")
c = {}
x_cpu = None
for k in range(cpu_reps):
    t0 = time.perf_counter(); pts, _ = bind.undistort(raw.xyzi, raw.offset_ms, raw.imu_pose, raw.x_end)
    t1 = time.perf_counter(); down = bind.voxelgrid(pts, 0.5)
    t2 = time.perf_counter(); o = bind.update_iterated(t, down, pr.x_prior, pr.P_prior, cfg.max_iter, pr.R, nthreads=min(64, os.cpu_count()))
    t3 = time.perf_counter()
    if k >= 1:
        for name, dt in (("undistort", t1 - t0), ("voxelgrid", t2 - t1), ("update", t3 - t2)):
            c[name] = c.get(name, 0.0) + dt
    x_cpu = o.x
n_c = cpu_reps - 1
cpu = {"ms_per_scan": 1e3 * sum(c.values()) / n_c, "scans_per_s": n_c / sum(c.values()), "stage_ms": {k: 1e3 * v / n_c for k, v in c.items()},
       "threads": {"undistort": 1, "voxelgrid": 1, "update": min(64, os.cpu_count())}, "n_down": int(len(down))}
print(json.dumps({"workload": f"raw {n_raw} pts -> voxel 0.5 m -> update vs 1M-pt map ({cfg.name})", "gpu": gpu, "cpu": cpu,
                  "speedup": gpu["scans_per_s"] / cpu["scans_per_s"], "state_diff_max": float(np.abs(x_gpu - x_cpu).max())}))
