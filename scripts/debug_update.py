import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_lio_b200 import api, synth
from oracle import bind
name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
solver = int(sys.argv[2]) if len(sys.argv) > 2 else 0
pr = synth.make_problem(name)
rt = bind.KdTree(pr.map_pts, "port")
o = bind.update_iterated(rt, pr.scan, pr.x_prior, pr.P_prior, pr.cfg.max_iter, pr.R, pr.limit, 0)
t = api.KdTree(0, 0.5); t.Build(pr.map_pts)
f = api.Esekf(t, max_points=len(pr.scan), max_iter=pr.cfg.max_iter, solver=solver)
x, P, st = f.update_iterated_dyn_share_modified(pr.scan, pr.x_prior, pr.P_prior, pr.R)
logs = f.pass_logs()
print("passes gpu", len(logs), "oracle", len(o.passes))
for i, (lg, op) in enumerate(zip(logs, o.passes)):
    print(i, "searched", lg["searched"], op["searched"], "effct", lg["effct"], op["effct"], "conv", lg["converged"], op["converged"],
          "res", lg["res_sum"], op["res_sum"])
    print("   HtH maxdiff", np.abs(lg["HtH"] - op["HtH"]).max(), "Hth maxdiff", np.abs(lg["Hth"] - op["Hth"]).max())
    print("   x_after diff", np.abs(lg["x_after"] - op["x_after"]).max(), "\n", np.round(lg["x_after"] - op["x_after"], 9), "\n prior-> oracle step", np.round(op["x_after"] - pr.x_prior, 6))
print("final x diff", np.abs(x - o.x).max(), "P diff", np.abs(P - o.P).max())
