"""A/B on the bench workloads: {split kernels of round 1, fused persistent k_update} x {BVH walk, cell directory}.
Device time of the kNN phase of one searching pass (L2 flushed / warm) and of a whole scan."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_lio_b200 import api, synth

out = {}
for name in sys.argv[1:] or ["velodyne_30k_1m", "ouster64_131k_5m"]:
    pr = synth.make_problem(name)
    t = api.KdTree(0, 0.5)
    t.Build(pr.map_pts)
    row = {"dir": t.dir_stats(), "map": t.stats()}
    for fused in (0, 1):
        for search in (0, 1):
            f = api.Esekf(t, max_points=len(pr.scan), max_iter=pr.cfg.max_iter, search=search, fused=fused)
            f.upload_scan(pr.scan); f.upload_state(pr.x_prior, pr.P_prior, pr.R)
            f.time_search_pass(5, True)
            t.dir_stats()
            cold = f.time_search_pass(50, True) / 50
            walked = t.dir_stats()["walked"] / 50.0
            warm = f.time_search_pass(50, False) / 50
            step = f.time_resident(200, True) / 200
            step_warm = f.time_resident(200, False) / 200
            x, P, n = f.download_state()
            row[("fused" if fused else "split") + "/" + ("cells" if search else "bvh")] = {
                "search_us_flushed": round(1e3 * cold, 2), "search_us_warm": round(1e3 * warm, 2), "walked_per_pass": walked,
                "scan_us": round(1e3 * step, 2), "scan_us_warm": round(1e3 * step_warm, 2), "passes": n, "launches": f.gpu_launches()}
    out[name] = row
    print(name, json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/search_ab.json", "w"), indent=1)
