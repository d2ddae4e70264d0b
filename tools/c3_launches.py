"""One C3 gradient evaluation at 8192 x 12288 Float32 (for an ncu launch list: which kernels run, how long)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import kernels as K  # noqa: E402
from nabla_jl_b200 import workloads as wl  # noqa: E402

ctx = nb.get_context()
dt = np.float32
m, n = 8192, 12288
rng = np.random.default_rng(2)
tile = rng.standard_normal((m, 64)).astype(dt)
A = ctx.empty((m, n), dt)
K.bcast_reduce(None, [ctx.asarray(tile).reshape((m, 64, 1))], A.reshape((m, 64, n // 64)))
b = ctx.asarray(rng.standard_normal(m).astype(dt))
f = nb.nabla(wl.bcast_sum(nb), get_output=True)
for _ in range(3):
    y, g = f(A, b)
ctx.sync()
n0 = ctx.stats()["kernel_launches"]
y, g = f(A, b)
ctx.sync()
print("launches per gradient:", ctx.stats()["kernel_launches"] - n0)
