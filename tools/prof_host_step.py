"""Host-side (Python) cost of one sharded step of the wide MLP (OverlappedShardedGradient, single rank): wall time of the
enqueue and cProfile top entries.  The batch is small so the device is never the bottleneck.
Usage: python tools/prof_host_step.py [batch]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import workloads as wl  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = nb.get_context()
dims = (784, 8192, 8192, 8192, 8192, 10)
W, b, X, Y = wl.mlp_inputs(dims, B, np.float32, seed=4)
params = [ctx.asarray(p) for p in (*W, *b)]
Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
comm = nb.Communicator(ctx, 0, 1, None)
sg = nb.OverlappedShardedGradient(wl.mlp_data_objective(nb, Xd, Yd, 5), wl.mlp_prior_objective(nb, 1e-3, 5), params, comm)
for _ in range(5):
    sg.step()
ctx.sync()
n = 30
t0 = time.perf_counter()
for _ in range(n):
    sg.step()
t1 = time.perf_counter()
ctx.sync()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / n:.2f} ms per step, with final sync {1e3 * (t2 - t0) / n:.2f} ms", flush=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    sg.step()
ctx.sync()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(35)
st.sort_stats("tottime").print_stats(30)
