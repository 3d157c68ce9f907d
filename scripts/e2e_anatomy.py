"""Host-side anatomy of fl_filter_update (enqueue uploads / enqueue passes / wait), averaged over calls."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_lio_b200 import api, synth
pr = synth.make_problem(sys.argv[1] if len(sys.argv) > 1 else "velodyne_30k_1m")
t = api.KdTree(0, 0.5); t.Build(pr.map_pts)
f = api.Esekf(t, max_points=len(pr.scan), max_iter=pr.cfg.max_iter)
api.host_register(pr.scan)
L = api.load()
L.fl_filter_debug_prof.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
prof = (C.c_longlong * 16)()
acc = np.zeros(4)
n = 300
for i in range(n + 20):
    f.time_e2e(pr.scan, pr.x_prior, pr.P_prior, pr.R, 1)
    if i >= 20:
        L.fl_filter_debug_prof(f.h, prof)
        acc += np.array(list(prof)[12:16], dtype=np.float64)
print("us per call: enqueue uploads %.1f, enqueue passes %.1f, wait %.1f, total %.1f" % tuple(acc / n / 1e3))
sec, _, _ = f.time_e2e(pr.scan, pr.x_prior, pr.P_prior, pr.R, 1000)
print("back-to-back native: %.1f us/scan" % (sec / 1000 * 1e6))
sts = []
for i in range(220):
    x, P, st = f.update_iterated_dyn_share_modified(pr.scan, pr.x_prior, pr.P_prior, pr.R)
    if i >= 20:
        sts.append(st)
print("device time (events around uploads + passes): median %.1f us" % (1e6 * float(np.median(sts))))
f.upload_scan(pr.scan); f.upload_state(pr.x_prior, pr.P_prior, pr.R)
print("resident warm: %.1f us/scan, flushed: %.1f" % (1e3 * f.time_resident(500, False) / 500, 1e3 * f.time_resident(500, True) / 500))
