"""torchrun --nproc-per-node N scripts/multi_gpu_check.py : every rank ends with the single-GPU state."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from fast_lio_b200 import api, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pr = synth.make_problem(sys.argv[1] if len(sys.argv) > 1 else "avia_2k_50k")
t = api.KdTree(local, 0.5); t.Build(pr.map_pts)
# single-GPU result on this rank
f1 = api.Esekf(t, max_points=len(pr.scan), max_iter=pr.cfg.max_iter)
x1, P1, _ = f1.update_iterated_dyn_share_modified(pr.scan, pr.x_prior, pr.P_prior, pr.R)
# sharded
f = api.Esekf(t, max_points=len(pr.scan), max_iter=pr.cfg.max_iter)
comm = sys.argv[2] if len(sys.argv) > 2 else "p2p"
if comm == "p2p":
    handles = [None] * world
    dist.all_gather_object(handles, f.p2p_handle())
    f.p2p_connect(world, rank, b"".join(handles))
else:
    uid = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    f.comm_init(world, rank, uid[0])
lo, hi = api.shard_range(len(pr.scan), world, rank)
f.set_shard(lo, hi)
x, P, _ = f.update_iterated_dyn_share_modified(pr.scan, pr.x_prior, pr.P_prior, pr.R)
dx = float(np.abs(x - x1).max()); dP = float(np.abs(P - P1).max())
xs = torch.tensor(x, device="cuda"); ref = xs.clone(); dist.broadcast(ref, src=0)
same = bool((xs == ref).all().item())
print(f"[{comm}] rank {rank}/{world}: shard [{lo},{hi}) |x - x_1gpu|={dx:.3e} |P - P_1gpu|={dP:.3e} identical_to_rank0={same} passes={len(f.pass_logs())}", flush=True)
assert dx < 1e-9 and dP < 1e-9 and same
# the replicated map mutation: every rank classifies the whole scan (the neighbours of the other ranks' shards are recomputed
# locally) and applies the same Add_Points -- the replicas must stay identical
res = f.map_incremental(0.5, True)
fl = t.flatten()
order = np.lexsort((fl[:, 3], fl[:, 2], fl[:, 1], fl[:, 0]))
digest = torch.tensor([float(t.validnum()), float(np.float64(fl[order, :3].astype(np.float64).sum())), float(res[0]), float(res[1]), float(res[2])], device="cuda", dtype=torch.float64)
ref = digest.clone(); dist.broadcast(ref, src=0)
print(f"[{comm}] rank {rank}: map_incremental {res}, map now {t.validnum()} points, replica identical to rank 0: {bool((digest == ref).all().item())}", flush=True)
assert bool((digest == ref).all().item())
dist.barrier(); dist.destroy_process_group()
