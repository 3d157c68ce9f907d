"""CPU, world_size 2 over gloo: the N>1 host logic -- scan sharding + one all-reduce of the 92
normal-equation doubles per pass -- gives the same sums as the unsharded pass."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fast_lio_b200 import api, synth
from oracle import bind


def pack_red(p):
    """The 92-double payload of one pass: upper triangle of H^T H, H^T h, effct, sum |res|."""
    iu = np.triu_indices(12)
    return np.concatenate([p["HtH"][iu], p["Hth"], [p["effct"], p["res_sum"]]])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pr = synth.make_problem("tiny")
    lo, hi = api.shard_range(len(pr.scan), world, rank)
    tree = bind.KdTree(pr.map_pts, "port")
    # first pass of the update on this rank's shard only (max_iter=1 -> two passes; take pass 0)
    o = bind.update_iterated(tree, pr.scan[lo:hi], pr.x_prior, pr.P_prior, 1, pr.R, pr.limit, 0, nthreads=1)
    red = torch.from_numpy(pack_red(o.passes[0]))
    dist.all_reduce(red, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put(red.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_normal_equations_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    red = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    pr = synth.make_problem("tiny")
    tree = bind.KdTree(pr.map_pts, "port")
    full = bind.update_iterated(tree, pr.scan, pr.x_prior, pr.P_prior, 1, pr.R, pr.limit, 0, nthreads=1)
    ref = pack_red(full.passes[0])
    assert red[90] == ref[90]                                   # effct_feat_num is an exact count
    assert np.allclose(red, ref, rtol=1e-12, atol=1e-9)         # FP64 sums: only the order differs
