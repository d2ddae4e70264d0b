"""Batch-sharded gradient evaluation on CPU: world_size-2 ``gloo`` processes, each
running the ORACLE on its column shard with the prior scaled by 1/G, allreduce(sum), compared with
the single-process gradient on the whole batch.  This pins the host-side sharding rules the GPU path
uses (workloads.shard_columns / sweep.shard_bounds, prior_scale, global ``scal`` and sliced noise for
the VAE); the device allreduce itself is NCCL and is exercised by bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, which, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import torch
    import torch.distributed as dist
    import nabla_oracle as orc
    from nabla_jl_b200 import workloads as wl
    from nabla_jl_b200.sweep import shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if which == "mlp":
            dims, B = (12, 9, 7, 5), 22
            W, b, X, Y = wl.mlp_inputs(dims, B, np.float64, seed=0)
            lo, hi = shard_bounds(B, rank, world)
            assert (wl.shard_columns(X, rank, world) == X[:, lo:hi]).all()
            obj = wl.mlp_objective(orc, X[:, lo:hi], Y[:, lo:hi], 1e-3, 3, prior_scale=1.0 / world)
            params = (*W, *b)
        else:
            dx, dz, dh, B = 9, 3, 6, 14
            params, X, noise = wl.vae_inputs(dx=dx, dz=dz, dh=dh, B=B, seed=3)
            lo, hi = shard_bounds(B, rank, world)
            # scal uses the GLOBAL batch; the noise is generated globally and sliced
            obj = wl.vae_objective(orc, X[:, lo:hi], noise[:, lo:hi], 1e-3, 60000 / B, dz, prior_scale=1.0 / world)
        y, grads = orc.nabla(obj, get_output=True)(*params)
        flat = np.concatenate([np.atleast_1d(np.asarray(orc.unbox(y), dtype=np.float64))] +
                              [np.asarray(g, dtype=np.float64).ravel() for g in grads])
        t = torch.from_numpy(flat.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if rank == 0:
            q.put(t.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("which", ["mlp", "vae"])
def test_sharded_gradient_equals_single_rank(which):
    import torch.multiprocessing as mp
    import nabla_oracle as orc
    from nabla_jl_b200 import workloads as wl
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, which, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if which == "mlp":
        W, b, X, Y = wl.mlp_inputs((12, 9, 7, 5), 22, np.float64, seed=0)
        y, grads = orc.nabla(wl.mlp_objective(orc, X, Y, 1e-3, 3), get_output=True)(*W, *b)
    else:
        params, X, noise = wl.vae_inputs(dx=9, dz=3, dh=6, B=14, seed=3)
        y, grads = orc.nabla(wl.vae_objective(orc, X, noise, 1e-3, 60000 / 14, 3), get_output=True)(*params)
    ref = np.concatenate([np.atleast_1d(np.asarray(orc.unbox(y), dtype=np.float64))] +
                         [np.asarray(g, dtype=np.float64).ravel() for g in grads])
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-12)


def _overlapped_worker(rank, world, port, which, q):
    """The decomposition OverlappedShardedGradient (sweep.py) uses, on the oracle: the batch-independent term is swept
    FIRST with the seed 1/G, its sensitivities pre-fill the reverse tape of the data sweep (the reference's pre-filled
    slot path, src/finite_differencing.jl:57-85 / core.jl:152-153), every data-term rule accumulates, then one sum over
    ranks; the objective is data + shared/G per rank."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import torch
    import torch.distributed as dist
    import nabla_oracle as orc
    from nabla_jl_b200 import workloads as wl
    from nabla_jl_b200.sweep import shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if which == "mlp":
            dims, B = (12, 9, 7, 5), 22
            W, b, X, Y = wl.mlp_inputs(dims, B, np.float64, seed=0)
            lo, hi = shard_bounds(B, rank, world)
            f_data = wl.mlp_data_objective(orc, X[:, lo:hi], Y[:, lo:hi], 3)
            f_shared = wl.mlp_prior_objective(orc, 1e-3, 3)
            params = (*W, *b)
        else:
            dx, dz, dh, B = 9, 3, 6, 14
            params, X, noise = wl.vae_inputs(dx=dx, dz=dz, dh=dh, B=B, seed=3)
            lo, hi = shard_bounds(B, rank, world)
            f_data = wl.vae_data_objective(orc, X[:, lo:hi], noise[:, lo:hi], 60000 / B, dz)
            f_shared = wl.vae_shared_objective(orc, 1e-3, 60000 / B, dz)
        # sweep A: shared term, seed 1/G
        t2 = orc.Tape()
        leaves2 = [orc.Leaf(t2, p) for p in params]
        y2 = f_shared(*leaves2)
        rt2 = orc.propagate(t2, orc.reverse_tape(y2, 1.0 / world))
        pre = [np.array(orc.unthunk(rt2[l]), dtype=np.float64, copy=True) if rt2.isassigned(l) else None for l in leaves2]
        # sweep B: data term accumulates into the pre-filled slots
        t = orc.Tape()
        leaves = [orc.Leaf(t, p) for p in params]
        y = f_data(*leaves)
        rt = orc.reverse_tape(y, orc.oneslike(orc.unbox(y)))
        for l, g in zip(leaves, pre):
            if g is not None:
                rt[l.pos] = g
        orc.propagate(t, rt)
        grads = [np.asarray(orc.unthunk(rt[l]), dtype=np.float64) for l in leaves]
        ytot = float(orc.unbox(y)) + float(orc.unbox(y2)) / world
        flat = np.concatenate([[ytot]] + [g.ravel() for g in grads])
        tt = torch.from_numpy(flat.copy())
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        if rank == 0:
            q.put(tt.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("which", ["mlp", "vae"])
def test_shared_term_first_with_prefilled_slots_equals_single_rank(which):
    import torch.multiprocessing as mp
    import nabla_oracle as orc
    from nabla_jl_b200 import workloads as wl
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_overlapped_worker, args=(r, world, port, which, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if which == "mlp":
        W, b, X, Y = wl.mlp_inputs((12, 9, 7, 5), 22, np.float64, seed=0)
        y, grads = orc.nabla(wl.mlp_objective(orc, X, Y, 1e-3, 3), get_output=True)(*W, *b)
    else:
        params, X, noise = wl.vae_inputs(dx=9, dz=3, dh=6, B=14, seed=3)
        y, grads = orc.nabla(wl.vae_objective(orc, X, noise, 1e-3, 60000 / 14, 3), get_output=True)(*params)
    ref = np.concatenate([np.atleast_1d(np.asarray(orc.unbox(y), dtype=np.float64))] +
                         [np.asarray(g, dtype=np.float64).ravel() for g in grads])
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-12)


def test_shard_bounds_cover_the_batch():
    from nabla_jl_b200.sweep import shard_bounds
    for B in (1, 7, 8, 4096, 32768):
        for G in (1, 2, 3, 4, 8):
            edges = [shard_bounds(B, r, G) for r in range(G)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(edges[i][1] == edges[i + 1][0] for i in range(G - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
