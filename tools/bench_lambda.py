"""Throughput of composite scalar functions in broadcast (the interpreter path): forward and fused pullback,
CUDA-event timed, as a fraction of the measured HBM copy peak.  Usage: python tools/bench_lambda.py [f32|f64]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import kernels as K  # noqa: E402

ctx = nb.get_context()
dt = np.float64 if (len(sys.argv) > 1 and sys.argv[1] == "f64") else np.float32
s = np.dtype(dt).itemsize
peak = 6552.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                             "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
m, n = 8192, 12288
N = m * n
rng = np.random.default_rng(2)
tile = ctx.asarray(rng.standard_normal((m, 64)).astype(dt))
A = ctx.empty((m, n), dt)
K.bcast_reduce(None, [tile.reshape((m, 64, 1))], A.reshape((m, 64, n // 64)))
B2 = ctx.empty((m, n), dt)
K.bcast_fwd(nb.compile_scalar(lambda p: 0.5 * p + 1.25, 1), [A], B2)
b = ctx.asarray(rng.standard_normal(m).astype(dt))
Y, Abar, Bbar, bbar = ctx.empty((m, n), dt), ctx.empty((m, n), dt), ctx.empty((m, n), dt), ctx.empty((m,), dt)
one = ctx.scalar(1.0, dt)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(reps):
        fn()
    ctx.record(e1)
    ctx.sync()
    return ctx.elapsed_ms(e0, e1) / reps


cases = [
    ("(a,b)->tanh(a+b), b::(m,)  pullback a,b", lambda p, q: nb.tanh(p + q), [A, b], [Abar, bbar], 2 * N * s),
    ("(a,b)->tanh(a+b), same shape pullback a,b", lambda p, q: nb.tanh(p + q), [A, B2], [Abar, Bbar], 4 * N * s),
    ("(a,b)->tanh(a)*b+exp(-a)/(1+b*b) pullback a,b (8 ops)",
     lambda p, q: nb.tanh(p) * q + nb.exp(-p) / (1.0 + q * q), [A, B2], [Abar, Bbar], 4 * N * s),
    ("x->x*x*x+2x (4 ops) pullback", lambda p: p * p * p + 2.0 * p, [A], [Abar], 2 * N * s),
]
for name, f, ops, outs, bytes_pull in cases:
    prog = nb.compile_scalar(f, len(ops))
    bytes_fwd = (sum(o.size for o in ops) + N) * s
    t_f = timeit(lambda: K.bcast_fwd(prog, ops, Y))
    t_p = timeit(lambda: K.bcast_pullback(prog, one, None, ops, outs, [False] * len(outs)))
    t_s = timeit(lambda: K.bcast_pullback(prog, one, Y, ops, outs, [False] * len(outs)))
    bytes_s = bytes_pull + N * s
    print(f"{np.dtype(dt).name} {name}: fwd {t_f:.3f} ms = {bytes_fwd / t_f / 1e6:.0f} GB/s ({bytes_fwd / t_f / 1e6 / peak:.0%}), "
          f"pullback {t_p:.3f} ms = {bytes_pull / t_p / 1e6:.0f} GB/s ({bytes_pull / t_p / 1e6 / peak:.0%}), "
          f"with saved y {t_s:.3f} ms = {bytes_s / t_s / 1e6:.0f} GB/s ({bytes_s / t_s / 1e6 / peak:.0%})", flush=True)
