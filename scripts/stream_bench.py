"""BASELINE config 4 as written: a Livox-Avia-like stream of N scans, 24k points each, against a ~1M-point map, the sensor
advancing 0.1 m per scan.  EVERY scan runs the reference's per-scan map pipeline around the update (laserMapping.cpp:865-1019):

    lasermap_fov_segment()  -> Delete_Point_Boxes when the local-map cube slides      (fl_localmap_segment)
    update_iterated_dyn_share_modified                                                 (fl_filter_update, host buffers)
    map_incremental()       -> Add_Points(PointToAdd, true) + Add_Points(NoNeedDownsample, false)   (fl_filter_map_incremental)

  python scripts/stream_bench.py [n_scans] [n_cpu_scans]                     one GPU
  torchrun --nproc-per-node G scripts/stream_bench.py [n_scans] 0            G GPUs: scan points sharded for the update, every
                                                                             rank applies the same map mutations; replicas compared

Prints one JSON line: per-scan latency distribution (p50 / p90 / p99 / max -- re-pack and directory re-list spikes included), the
breakdown, the map's bookkeeping, and (one GPU) the CPU reference pipeline on the first scans."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_lio_b200 import api, synth

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_cpu = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

cfg = synth.CONFIGS["avia_stream_24k"]
map_pts, scene = synth.make_map(cfg.n_map, seed=1)
x0 = synth.true_state(cfg.lidar, 0)
xp, P0 = synth.make_prior(x0, seed=3)
CUBE_LEN, DET_RANGE = 240.0, 60.0           # the cube slides ~every 300 scans (laserMapping.cpp:229-277 with these parameters)

g = api.KdTree(local, 0.5); g.Build(map_pts)
f = api.Esekf(g, max_points=cfg.n_scan, max_iter=cfg.max_iter)
if world > 1:
    handles = [None] * world
    dist.all_gather_object(handles, f.p2p_handle())
    f.p2p_connect(world, rank, b"".join(handles))
    f.set_shard(*api.shard_range(cfg.n_scan, world, rank))
cube = api.LocalMap(CUBE_LEN, DET_RANGE)

x, P = xp.copy(), P0.copy()
lat, t_seg, t_upd, t_map = [], [], [], []
deleted_total = added_total = n_slides = 0
for k in range(n_scans):
    scan = synth.make_scan(scene, cfg.n_scan, synth.true_state(cfg.lidar, k), seed=50 + k)
    P = P + np.eye(23) * 1e-4                                # crude prediction: carry the state, inflate the covariance
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    pos_lid = x[0:3] + synth.quat_to_mat(x[3:7]) @ x[11:14]  # pos + rot * offset_T_L_I  (laserMapping.cpp:236)
    boxes, ndel = cube.segment(pos_lid, g)
    nb = len(boxes)
    t1 = time.perf_counter()
    x, P, _ = f.update_iterated_dyn_share_modified(scan, x, P, 0.001)
    t2 = time.perf_counter()
    n_add, n_no, added = f.map_incremental(0.5, True)
    t3 = time.perf_counter()
    lat.append(t3 - t0); t_seg.append(t1 - t0); t_upd.append(t2 - t1); t_map.append(t3 - t2)
    deleted_total += ndel; added_total += n_add + n_no; n_slides += nb > 0

lat_ms = 1e3 * np.array(lat[3:])
out = {"workload": "avia_stream_24k", "n_gpus": world, "n_scans": n_scans, "points_per_scan": cfg.n_scan,
       "latency_ms": {"p50": float(np.percentile(lat_ms, 50)), "p90": float(np.percentile(lat_ms, 90)), "p99": float(np.percentile(lat_ms, 99)),
                      "max": float(lat_ms.max()), "mean": float(lat_ms.mean())},
       "scans_per_s": float(1e3 / lat_ms.mean()),
       "breakdown_ms_mean": {"lasermap_fov_segment+Delete_Point_Boxes": float(1e3 * np.mean(t_seg[3:])), "update (host buffers)": float(1e3 * np.mean(t_upd[3:])),
                             "map_incremental + 2x Add_Points": float(1e3 * np.mean(t_map[3:]))},
       "worst_scans": [int(i) + 3 for i in np.argsort(lat_ms)[-5:][::-1]],
       "cube_slides": int(n_slides), "points_deleted": int(deleted_total), "points_offered_to_Add_Points": int(added_total),
       "pos_err_m": float(np.abs(x[:3] - synth.true_state(cfg.lidar, n_scans - 1)[:3]).max()),
       "map": {"validnum": g.validnum(), "size": g.size(), **g.stats(), "directory": g.dir_stats()}}
if world > 1:
    fl = g.flatten()
    fl = fl[np.lexsort((fl[:, 3], fl[:, 2], fl[:, 1], fl[:, 0]))]        # the replicas hold the same SET; slot order (hence flatten order) differs
    digest = torch.tensor([float(g.validnum()), float(np.cumsum(fl[:, :3].astype(np.float64), axis=0)[-1].sum()), float(np.abs(x).sum())], device="cuda", dtype=torch.float64)
    ref = digest.clone(); dist.broadcast(ref, src=0)
    same = torch.tensor([1.0 if bool((digest == ref).all().item()) else 0.0], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out["replicas_identical"] = bool(same.item() == 1.0)
elif n_cpu > 1:
    from oracle import bind
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from semantics import map_incremental
    r = bind.KdTree(map_pts, "reference", downsample=0.5)
    lm = bind.LocalMap(CUBE_LEN, DET_RANGE)
    x, P = xp.copy(), P0.copy()
    c_lat = []
    for k in range(n_cpu):
        scan = synth.make_scan(scene, cfg.n_scan, synth.true_state(cfg.lidar, k), seed=50 + k)
        P = P + np.eye(23) * 1e-4
        t0 = time.perf_counter()
        boxes = lm.segment(x[0:3] + synth.quat_to_mat(x[3:7]) @ x[11:14])
        if len(boxes): r.delete_boxes(boxes)
        o = bind.update_iterated(r, scan, x, P, cfg.max_iter, 0.001, 0.001, 0, nthreads=min(64, os.cpu_count()))
        x, P = o.x, o.P
        to_add, no_need = map_incremental(scan, x, o.nearest, o.nearest_cnt, 0.5, True)
        r.add(to_add, True); r.add(no_need, False)
        c_lat.append(time.perf_counter() - t0)
    out["cpu_reference"] = {"scans": n_cpu, "latency_ms_median": float(1e3 * np.median(c_lat[1:])), "threads": min(64, os.cpu_count()),
                            "what": "reference ikd-Tree (oracle/_ref) + restated update + numpy map_incremental, same stream"}
    out["speedup_vs_cpu_reference"] = out["cpu_reference"]["latency_ms_median"] / out["latency_ms"]["p50"]
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
