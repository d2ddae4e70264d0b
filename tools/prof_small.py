"""Profiling driver for the small Float64 configs (C1 MLP 784-500-300-10, C4 VAE, B = 4096): a few eager
gradient evaluations so `ncu` can list every launch.  Usage: python tools/prof_small.py [mlp|vae] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import workloads as wl  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "mlp"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = nb.get_context()
if which == "mlp":
    W, b, X, Y = wl.mlp_inputs((784, 500, 300, 10), 4096, np.float64, seed=0)
    Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
    params = [ctx.asarray(p) for p in (*W, *b)]
    obj = wl.mlp_objective(nb, Xd, Yd, 1e-3, 3)
else:
    vp, VX, noise = wl.vae_inputs(B=4096, seed=3)
    VXd, Nd = ctx.asarray(VX), ctx.asarray(noise)
    params = [ctx.asarray(p) for p in vp]
    obj = wl.vae_objective(nb, VXd, Nd, 1e-3, 60000 / 4096, 10)
for _ in range(reps):
    g = nb.nabla(obj)(*params)
ctx.sync()
print("ok", ctx.stats()["kernel_launches"])
