"""Round-2 prototype bench (needs a GPU; never run so far): k-NN of the benchmark scan through the hashed directory of leaf
buckets (fast_lio_b200/csrc/experimental/cellmap.cu) against the product's BVH walk -- equality of the results, kernel time,
and how many queries settle in the 27-cell neighbourhood.
usage: python scripts/build_experimental.py && python scripts/experimental_cellmap_bench.py [workload] [cell_m ...]"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from fast_lio_b200 import api, synth
from test_experimental_cellmap import _lib, cell_knn

name = sys.argv[1] if len(sys.argv) > 1 else "velodyne_30k_1m"
cells = [float(c) for c in sys.argv[2:]] or [1.0, 2.0]
pr = synth.make_problem(name)
x = pr.x_prior
R = synth.quat_to_mat(x[3:7]); Rl = synth.quat_to_mat(x[7:11])
q = np.zeros((len(pr.scan), 4), np.float32)
q[:, :3] = ((pr.scan[:, :3].astype(np.float64) @ Rl.T + x[11:14]) @ R.T + x[0:3]).astype(np.float32)
t = api.KdTree(0, 0.5); t.Build(pr.map_pts)
f = api.Esekf(t, max_points=len(pr.scan), max_iter=pr.cfg.max_iter)
f.upload_scan(pr.scan); f.upload_state(pr.x_prior, pr.P_prior, pr.R)
bvh_us = 1e3 * f.time_search_pass(100, False) / 100
p_ref, d_ref, c_ref = t.Nearest_Search(q, 5)
L = _lib()
out = {"workload": name, "k_search_bvh_us_l2_warm": bvh_us, "cells": {}}
for cell in cells:
    best = None
    for rep in range(5):
        p, d, c, rings, ms = cell_knn(L, pr.map_pts, q, cell)
        best = ms if best is None else min(best, ms)
    out["cells"][str(cell)] = {"kernel_us_best_of_5": 1e3 * best, "equal_counts": bool(np.array_equal(c, c_ref)),
                               "equal_d2": bool(np.array_equal(d, d_ref)), "equal_points": bool(np.array_equal(p, p_ref)),
                               "settled_at_ring": {str(r): int(n) for r, n in enumerate(rings) if n}}
print(json.dumps(out))
