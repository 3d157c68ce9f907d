"""BASELINE config 4: Livox-Avia-like stream, 24k points/scan against a ~1M-point map, sensor moving
0.1 m/scan; every scan: [Delete_Point_Boxes when the window moves] + update + map_incremental
(+ the two Add_Points).  Prints one JSON line: device pipeline vs CPU reference pipeline."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_lio_b200 import api, synth
from oracle import bind
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from semantics import map_incremental

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n_cpu = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = synth.CONFIGS["avia_stream_24k"]
map_pts, scene = synth.make_map(cfg.n_map, seed=1)
x0 = synth.true_state(cfg.lidar, 0)
xp, P0 = synth.make_prior(x0, seed=3)
scans = [synth.make_scan(scene, cfg.n_scan, synth.true_state(cfg.lidar, k), seed=50 + k) for k in range(n_scans)]
box = np.array([[-1e4, -1e4, -1e4, -scene.extent + 20.0, 1e4, 1e4]], dtype=np.float32)

g = api.KdTree(0, 0.5); g.Build(map_pts)
f = api.Esekf(g, max_points=cfg.n_scan, max_iter=cfg.max_iter, solver=1)
x, P = xp.copy(), P0.copy()
t_upd = t_map = t_del = 0.0
for k, scan in enumerate(scans):
    P = P + np.eye(23) * 1e-4
    if k == n_scans // 2:
        t0 = time.perf_counter(); g.Delete_Point_Boxes(box); t_del += time.perf_counter() - t0
    t0 = time.perf_counter(); x, P, _ = f.update_iterated_dyn_share_modified(scan, x, P, 0.001); t1 = time.perf_counter()
    f.map_incremental(0.5, True); t2 = time.perf_counter()
    if k >= 2:
        t_upd += t1 - t0; t_map += t2 - t1
n_t = n_scans - 2
gpu = {"update_ms": 1e3 * t_upd / n_t, "map_incremental_add_ms": 1e3 * t_map / n_t, "delete_boxes_ms": 1e3 * t_del,
       "scans_per_s": n_t / (t_upd + t_map), "pos_err_m": float(np.abs(x[:3] - synth.true_state(cfg.lidar, n_scans - 1)[:3]).max()),
       "map_stats": g.stats(), "validnum": g.validnum()}

r = bind.KdTree(map_pts, "reference", downsample=0.5)
x, P = xp.copy(), P0.copy()
c_upd = c_map = 0.0
for k, scan in enumerate(scans[:n_cpu]):
    P = P + np.eye(23) * 1e-4
    t0 = time.perf_counter(); o = bind.update_iterated(r, scan, x, P, cfg.max_iter, 0.001, 0.001, 0, nthreads=os.cpu_count()); t1 = time.perf_counter()
    x, P = o.x, o.P
    to_add, no_need = map_incremental(scan, x, o.nearest, o.nearest_cnt, 0.5, True)
    t2 = time.perf_counter(); r.add(to_add, True); r.add(no_need, False); t3 = time.perf_counter()
    if k >= 1:
        c_upd += t1 - t0; c_map += t3 - t2
n_c = n_cpu - 1
cpu = {"update_ms": 1e3 * c_upd / n_c, "add_points_ms": 1e3 * c_map / n_c, "scans_per_s": n_c / (c_upd + c_map), "cores": os.cpu_count()}
print(json.dumps({"workload": "avia_stream_24k", "gpu": gpu, "cpu_reference": cpu, "speedup": gpu["scans_per_s"] / cpu["scans_per_s"]}))
