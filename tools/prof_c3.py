"""Profiling driver for the HBM-bound kernels of config C3 (s = sum(tanh.(A .+ b)), 8192 x 12288):
runs each kernel of the forward and reverse sweep a few times so `ncu` can capture them.
Usage (on the GPU box):  python tools/prof_c3.py [f32|f64] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import kernels as K  # noqa: E402

dt = np.float64 if (len(sys.argv) > 1 and sys.argv[1] == "f64") else np.float32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = nb.get_context()
m, n = 8192, 12288
rng = np.random.default_rng(2)
tile = ctx.asarray(rng.standard_normal((m, 64)).astype(dt))
A = ctx.empty((m, n), dt)
K.bcast_reduce(None, [tile.reshape((m, 64, 1))], A.reshape((m, 64, n // 64)))
b = ctx.asarray(rng.standard_normal(m).astype(dt))
brow = ctx.asarray(rng.standard_normal((1, n)).astype(dt))
t1, y, tb1, Abar = (ctx.empty((m, n), dt) for _ in range(4))
bbar, browbar = ctx.empty((m,), dt), ctx.empty((1, n), dt)
one, sout = ctx.scalar(1.0, dt), ctx.scalar(0.0, dt)
add, tanh = nb.compile_scalar(nb.add, 2), nb.compile_scalar(nb.tanh, 1)
fused = nb.compile_scalar(lambda p, q: nb.tanh(p + q), 2)
for _ in range(reps):
    K.bcast_fwd(add, [A, b], t1)
    K.bcast_fwd(tanh, [t1], y)
    K.bcast_reduce(None, [y], sout)
    K.bcast_pullback(tanh, one, y, [t1], [tb1], [False])
    K.bcast_pullback(add, tb1, None, [A, b], [None, bbar], [False, False])
    K.bcast_pullback(add, tb1, None, [A, brow], [None, browbar], [False, False])
    K.bcast_pullback(fused, one, None, [A, b], [Abar, bbar], [False, False])
ctx.sync()
print("ok", ctx.stats()["kernel_launches"])
