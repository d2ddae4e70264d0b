"""One 4096^3 Float64 product through the Ozaki split (for ncu)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import kernels as K  # noqa: E402

ctx = nb.get_context()
rng = np.random.default_rng(0)
m = n = k = 4096
A, B = ctx.asarray(rng.standard_normal((m, k))), ctx.asarray(rng.standard_normal((k, n)))
C = ctx.empty((m, n), np.float64)
for _ in range(2):
    K.gemm("N", "N", 1.0, A, B, 0.0, C, m, n, k)
ctx.sync()
print("ok")
