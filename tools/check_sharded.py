"""Multi-GPU parity check (run under torchrun): batch-sharded MLP gradient with the NCCL allreduce
equals the single-rank gradient of the whole batch (rtol 1e-10 Float64; summation order differs)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import nabla_jl_b200 as nb
from nabla_jl_b200 import workloads as wl

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = nb.Context(device=local)
nb.set_context(ctx)


def exchange(raw):
    obj = [raw]
    dist.broadcast_object_list(obj, src=0)
    return obj[0]


comm = nb.Communicator(ctx, rank, world, exchange)
dims, B = (784, 500, 300, 10), 512
W, b, X, Y = wl.mlp_inputs(dims, B, np.float64, seed=0)
lo, hi = nb.shard_bounds(B, rank, world)
Xd, Yd = ctx.asarray(X[:, lo:hi]), ctx.asarray(Y[:, lo:hi])
params = [ctx.asarray(p) for p in (*W, *b)]
for capture in (False, True):
    sg = nb.ShardedGradient(lambda ps: wl.mlp_objective(nb, Xd, Yd, 1e-3, 3, prior_scale=ps), params, comm,
                            capture=capture)
    y, grads = sg.step()
    ctx.sync()
    if rank == 0:
        Xf, Yf = ctx.asarray(X), ctx.asarray(Y)
        yr, gr = nb.nabla(wl.mlp_objective(nb, Xf, Yf, 1e-3, 3), get_output=True)(*params)
        err = max(np.linalg.norm(g.numpy() - r.numpy()) / np.linalg.norm(r.numpy()) for g, r in zip(grads, gr))
        yerr = abs(float(y.numpy()) - float(nb.unbox(yr).numpy())) / abs(float(nb.unbox(yr).numpy()))
        print(f"world={world} capture={capture}: max rel grad err {err:.3e}, objective rel err {yerr:.3e}", flush=True)
        assert err < 1e-10 and yerr < 1e-10
    dist.barrier()
# overlapped variant: per-parameter allreduce on the communicator stream while the sweep continues
og = nb.OverlappedShardedGradient(wl.mlp_data_objective(nb, Xd, Yd, 3), wl.mlp_prior_objective(nb, 1e-3, 3), params, comm)
for it in range(2):
    y, grads = og.step()
ctx.sync()
if rank == 0:
    Xf, Yf = ctx.asarray(X), ctx.asarray(Y)
    yr, gr = nb.nabla(wl.mlp_objective(nb, Xf, Yf, 1e-3, 3), get_output=True)(*params)
    err = max(np.linalg.norm(g.numpy() - r.numpy()) / np.linalg.norm(r.numpy()) for g, r in zip(grads, gr))
    yerr = abs(float(y.numpy()) - float(nb.unbox(yr).numpy())) / abs(float(nb.unbox(yr).numpy()))
    print(f"world={world} overlapped: max rel grad err {err:.3e}, objective rel err {yerr:.3e}", flush=True)
    assert err < 1e-10 and yerr < 1e-10
dist.barrier()
# Float32 with the deferred / fused product launches (fusion.py): sharded vs single-rank, both drivers
W, b, X, Y = wl.mlp_inputs((784, 512, 256, 10), 2048, np.float32, seed=1)
lo, hi = nb.shard_bounds(2048, rank, world)
Xd, Yd = ctx.asarray(X[:, lo:hi]), ctx.asarray(Y[:, lo:hi])
params = [ctx.asarray(p) for p in (*W, *b)]
sg = nb.ShardedGradient(lambda ps: wl.mlp_objective(nb, Xd, Yd, 1e-3, 3, prior_scale=ps), params, comm)
og = nb.OverlappedShardedGradient(wl.mlp_data_objective(nb, Xd, Yd, 3), wl.mlp_prior_objective(nb, 1e-3, 3), params, comm)
for name, drv in (("sharded", sg), ("overlapped", og)):
    for it in range(2):
        y, grads = drv.step()
    ctx.sync()
    if rank == 0:
        Xf, Yf = ctx.asarray(X), ctx.asarray(Y)
        yr, gr = nb.nabla(wl.mlp_objective(nb, Xf, Yf, 1e-3, 3), get_output=True)(*params)
        err = max(np.linalg.norm(g.numpy() - r.numpy()) / np.linalg.norm(r.numpy()) for g, r in zip(grads, gr))
        print(f"world={world} f32 fused {name}: max rel grad err {err:.3e}", flush=True)
        assert err < 1e-4
    dist.barrier()
# VAE (examples/vae.jl:67-90): global `scal`, noise generated globally and sliced by columns, the constant D and the
# prior in the shared term; eager and as a CUDA-graph replay of the whole step (NCCL collectives inside the graph)
vp, VX, noise = wl.vae_inputs(B=1024, seed=3)
Bv = 1024
lo, hi = nb.shard_bounds(Bv, rank, world)
VXd, Nd = ctx.asarray(np.asfortranarray(VX[:, lo:hi])), ctx.asarray(np.asfortranarray(noise[:, lo:hi]))
vparams = [ctx.asarray(p) for p in vp]
scal = 60000.0 / Bv
for capture in (False, True):
    og = nb.OverlappedShardedGradient(wl.vae_data_objective(nb, VXd, Nd, scal, 10), wl.vae_shared_objective(nb, 1e-3, scal, 10),
                                      vparams, comm, capture=capture)
    for it in range(3):
        y, grads = og.step()
    ctx.sync()
    if rank == 0:
        yr, gr = nb.nabla(wl.vae_objective(nb, ctx.asarray(VX), ctx.asarray(noise), 1e-3, scal, 10), get_output=True)(*vparams)
        err = max(np.linalg.norm(g.numpy() - r.numpy()) / np.linalg.norm(r.numpy()) for g, r in zip(grads, gr))
        yerr = abs(float(y.numpy()) - float(nb.unbox(yr).numpy())) / abs(float(nb.unbox(yr).numpy()))
        print(f"world={world} VAE f64 overlapped capture={capture}: max rel grad err {err:.3e}, objective rel err {yerr:.3e}", flush=True)
        assert err < 1e-10 and yerr < 1e-10
    dist.barrier()
    og.close()
# the wide-MLP step as one CUDA graph (what bench.py --capture replays), Float32, against the eager step
W, b, X, Y = wl.mlp_inputs((784, 1024, 1024, 10), 2048, np.float32, seed=1)
lo, hi = nb.shard_bounds(2048, rank, world)
Xd, Yd = ctx.asarray(np.asfortranarray(X[:, lo:hi])), ctx.asarray(np.asfortranarray(Y[:, lo:hi]))
params = [ctx.asarray(p) for p in (*W, *b)]
res = {}
for capture in (False, True):
    og = nb.OverlappedShardedGradient(wl.mlp_data_objective(nb, Xd, Yd, 3), wl.mlp_prior_objective(nb, 1e-3, 3), params, comm,
                                      capture=capture)
    for it in range(3):
        y, grads = og.step()
    ctx.sync()
    res[capture] = [g.numpy() for g in grads]
    og.close()
if rank == 0:
    err = max(np.linalg.norm(a - r) / np.linalg.norm(r) for a, r in zip(res[True], res[False]))
    print(f"world={world} f32 MLP captured step vs eager step: max rel diff {err:.3e}", flush=True)
    assert err < 1e-6
dist.barrier()
comm.close()
dist.destroy_process_group()
if rank == 0:
    print("sharded ok")
