"""BASELINE config 5 (dense-map stress: 200k points/scan vs a 20M-point map) once, on the GPU(s) at hand, with the CPU oracle
beside it: per-pass decisions and the final state must agree (north-star tolerance 1e-4 m / 1e-4 rad).
  python scripts/dense_check.py [workload]                       one GPU
  torchrun --nproc-per-node G scripts/dense_check.py [workload]  G GPUs (scan sharded, map replicated, peer-memory exchange)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_lio_b200 import api, synth

name = sys.argv[1] if len(sys.argv) > 1 else "dense_200k_20m"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t0 = time.perf_counter(); pr = synth.make_problem(name); t_gen = time.perf_counter() - t0
t0 = time.perf_counter(); t = api.KdTree(local, 0.5); t.Build(pr.map_pts); t_build = time.perf_counter() - t0
n = len(pr.scan)
f = api.Esekf(t, max_points=n, max_iter=pr.cfg.max_iter)
if world > 1:
    handles = [None] * world
    dist.all_gather_object(handles, f.p2p_handle())
    f.p2p_connect(world, rank, b"".join(handles))
    f.set_shard(*api.shard_range(n, world, rank))
x, P, st = f.update_iterated_dyn_share_modified(pr.scan, pr.x_prior, pr.P_prior, pr.R)
logs = f.pass_logs()
if world == 1:
    near, cnt = f.nearest(n)
    sel = f.selected(n)
f.upload_scan(pr.scan); f.upload_state(pr.x_prior, pr.P_prior, pr.R)
f.time_resident(5, True)
ms = f.time_resident(50, True) / 50
ms_search = f.time_search_pass(20, True) / 20
out = {"workload": name, "n_gpus": world, "n_map": len(pr.map_pts), "n_scan": n, "build_s (host->device, k-d partition, directory)": round(t_build, 3),
       "ms_per_scan": ms, "scans_per_s": 1e3 / ms, "search_phase_ms": ms_search, "passes": len(logs), "effct": [l["effct"] for l in logs],
       "map": t.stats(), "directory": t.dir_stats()}
if rank == 0 and (len(sys.argv) < 3 or sys.argv[2] != "noref"):
    from oracle import bind
    t0 = time.perf_counter(); ref = bind.KdTree(pr.map_pts, "auto"); t_ref_build = time.perf_counter() - t0
    t0 = time.perf_counter(); o = bind.update_iterated(ref, pr.scan, pr.x_prior, pr.P_prior, pr.cfg.max_iter, pr.R, pr.limit, 0, nthreads=min(64, os.cpu_count())); t_ref = time.perf_counter() - t0
    qa, qb = x[3:7] / np.linalg.norm(x[3:7]), o.x[3:7] / np.linalg.norm(o.x[3:7])
    out["cpu_reference"] = {"ikd_tree_build_s": round(t_ref_build, 2), "update_ms": round(1e3 * t_ref, 1), "threads": min(64, os.cpu_count())}
    out["parity"] = {"pos_err": float(np.abs(x[:3] - o.x[:3]).max()), "rot_err": float(2 * np.arccos(min(1.0, abs(float(qa @ qb))))),
                     "other_err": float(np.abs(x[7:] - o.x[7:]).max()),
                     "passes_equal": [(l["searched"], l["effct"], l["converged"]) for l in logs] == [(p["searched"], p["effct"], p["converged"]) for p in o.passes]}
    if world == 1:
        out["parity"]["nearest_equal"] = bool(np.array_equal(near, o.nearest) and np.array_equal(cnt, o.nearest_cnt))
        out["parity"]["selected_equal"] = bool(np.array_equal(sel, o.selected))
    out["parity"]["ok"] = bool(out["parity"]["pos_err"] <= 1e-4 and out["parity"]["rot_err"] <= 1e-4 and out["parity"]["passes_equal"])
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
