"""Profiling driver for the composite-lambda broadcast pullback (bytecode interpreter path)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb
from nabla_jl_b200 import kernels as K
ctx = nb.get_context()
dt = np.float32
m, n = 8192, 12288
rng = np.random.default_rng(2)
tile = ctx.asarray(rng.standard_normal((m, 64)).astype(dt))
A = ctx.empty((m, n), dt)
K.bcast_reduce(None, [tile.reshape((m, 64, 1))], A.reshape((m, 64, n // 64)))
b = ctx.asarray(rng.standard_normal(m).astype(dt))
Abar, bbar, one = ctx.empty((m, n), dt), ctx.empty((m,), dt), ctx.scalar(1.0, dt)
fused = nb.compile_scalar(lambda p, q: nb.tanh(p + q), 2)
for _ in range(3):
    K.bcast_pullback(fused, one, None, [A, b], [Abar, bbar], [False, False])
ctx.sync()
print("ok")
