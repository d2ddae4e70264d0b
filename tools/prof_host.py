"""Host-side (Python) cost of one eager gradient evaluation of the C1 MLP (Float64, B = 4096): cProfile top entries."""
import cProfile
import os
import pstats
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import workloads as wl  # noqa: E402

ctx = nb.get_context()
dims, B = (784, 500, 300, 10), 4096
W, b, X, Y = wl.mlp_inputs(dims, B, np.float64, seed=0)
Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
params = [ctx.asarray(p) for p in (*W, *b)]
f = nb.nabla(wl.mlp_objective(nb, Xd, Yd, 1e-3, 3))
for _ in range(5):
    f(*params)
ctx.sync()
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    f(*params)
ctx.sync()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
st.sort_stats("tottime").print_stats(22)
