"""ncu driver: a few launches of the kNN phase alone (fl_filter_time_search_pass), fused or split."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_lio_b200 import api, synth
name = sys.argv[1] if len(sys.argv) > 1 else "velodyne_30k_1m"
fused = int(sys.argv[2]) if len(sys.argv) > 2 else 1
search = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pr = synth.make_problem(name)
t = api.KdTree(0, 0.5); t.Build(pr.map_pts)
f = api.Esekf(t, max_points=len(pr.scan), max_iter=pr.cfg.max_iter, search=search, fused=fused)
f.upload_scan(pr.scan); f.upload_state(pr.x_prior, pr.P_prior, pr.R)
print("search pass us (flushed):", 1e3 * f.time_search_pass(6, True) / 6, "warm:", 1e3 * f.time_search_pass(6, False) / 6)
