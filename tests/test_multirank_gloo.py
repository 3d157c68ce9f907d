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


def _worker_small_m(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pr = synth.make_problem("tiny")
    scan = pr.scan[:14]
    lo, hi = api.shard_range(len(scan), world, rank)
    tree = bind.KdTree(pr.map_pts, "port")
    o = bind.update_iterated(tree, scan[lo:hi], pr.x_prior, pr.P_prior, 1, pr.R, pr.limit, 0, nthreads=1)
    red = torch.from_numpy(pack_red(o.passes[0]))
    dist.all_reduce(red, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put(red.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_fewer_than_23_rows_split_over_two_ranks():
    """with fewer than 23 effective rows the reference takes K = P H^T (H P H^T / R + I)^-1 / R
    (esekfom.hpp:1715-1744), which needs the rows themselves -- and those are split over the ranks.  The kernels therefore solve
    every m >= 1 from the all-reduced 12x12 sums through the matrix-inversion lemma (update.cuh sol_pass):
    dx = P[:, :12] (I + H^T H P11 / R)^-1 H^T h / R on the first pass.  Same step as the reference's small-m branch on the
    whole scan (the oracle), whichever way the rows were sharded."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_small_m, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    red = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    pr = synth.make_problem("tiny")
    tree = bind.KdTree(pr.map_pts, "port")
    full = bind.update_iterated(tree, pr.scan[:14], pr.x_prior, pr.P_prior, 1, pr.R, pr.limit, 0, nthreads=1)
    p0 = full.passes[0]
    assert 0 < p0["effct"] < 23 and red[90] == p0["effct"]        # the whole scan took the small-m branch
    HtH = np.zeros((12, 12)); HtH[np.triu_indices(12)] = red[:78]
    HtH = HtH + np.triu(HtH, 1).T
    Hth = red[78:90]
    P = pr.P_prior
    v = np.linalg.solve(np.eye(12) + HtH @ P[:12, :12] / pr.R, Hth)
    dx = P[:, :12] @ v / pr.R
    want = pr.x_prior.copy()
    bind.lib().oracle_state_boxplus(want, np.ascontiguousarray(dx))
    assert np.abs(want - p0["x_after"]).max() < 1e-9
