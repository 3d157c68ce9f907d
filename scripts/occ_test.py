import sys, os
sys.path.insert(0, ".")
from fast_lio_b200 import api, synth
name = sys.argv[1]
pr = synth.make_problem(name)
t = api.KdTree(0, 0.5); t.Build(pr.map_pts)
f = api.Esekf(t, max_points=len(pr.scan), max_iter=4)
f.upload_scan(pr.scan); f.upload_state(pr.x_prior, pr.P_prior, pr.R)
n = 300
print(name, "occ", os.environ.get("FASTLIO_B200_SEARCH_OCC", "4"), "k_search us flushed %.2f warm %.2f | scan us flushed %.1f warm %.1f" % (
    1e3 * f.time_search_pass(n, True) / n, 1e3 * f.time_search_pass(n, False) / n, 1e3 * f.time_resident(n, True) / n, 1e3 * f.time_resident(n, False) / n))
