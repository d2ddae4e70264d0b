"""C1 (MLP 784-500-300-10) and C4 (VAE) at B = 4096, Float64: CUDA-graph replay time of one gradient evaluation
(CUDA events over `reps` replays) and launches per replay.  Usage: python tools/bench_small.py [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import workloads as wl  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = nb.get_context()


def timed(fn, n):
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(n):
        fn()
    ctx.record(e1)
    ctx.sync()
    return ctx.elapsed_ms(e0, e1) / n


W, b, X, Y = wl.mlp_inputs((784, 500, 300, 10), 4096, np.float64, seed=0)
params = [ctx.asarray(p) for p in (*W, *b)]
obj = wl.mlp_objective(nb, ctx.asarray(X), ctx.asarray(Y), 1e-3, 3)
vp, VX, noise = wl.vae_inputs(B=4096, seed=3)
vparams = [ctx.asarray(p) for p in vp]
vobj = wl.vae_objective(nb, ctx.asarray(VX), ctx.asarray(noise), 1e-3, 60000 / 4096, 10)
for name, f, p, flops in (("c1_mlp", obj, params, 2.0 * 4096 * (3 * 545000 - 392000)),
                          ("c4_vae", vobj, vparams, 2.0 * 4096 * (2 * 799000 + 407000))):
    cap = nb.CapturedGradient(f, p, ctx)
    for _ in range(5):
        cap()
    best = min(timed(cap, reps) for _ in range(3))
    print(f"{name}: {best * 1e3:.1f} us per evaluation, {cap.launches_per_replay} launches, "
          f"{flops / best / 1e9:.2f} TFLOP/s", flush=True)
