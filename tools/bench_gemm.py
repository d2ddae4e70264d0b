"""GEMM throughput probe: the wide-MLP shapes (forward NN, Abar NT, Bbar TN) in each math mode.
Usage: python tools/bench_gemm.py [B] [modes]   (env NB_GEMM_BN=128|256 overrides the tile width)"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import kernels as K  # noqa: E402

ctx = nb.get_context()
Bc = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["tf32", "3xtf32", "bf16x3"]
D = int(os.environ.get("D", "8192"))
DT = np.float64 if os.environ.get("DTYPE", "f32") == "f64" else np.float32
if DT == np.float64:
    modes = ["f64"]
rng = np.random.default_rng(0)
W = ctx.asarray(rng.standard_normal((D, D)).astype(DT))
X = ctx.asarray(rng.standard_normal((D, Bc)).astype(DT))
Y = ctx.empty((D, Bc), DT)
Wb = ctx.empty((D, D), DT)
Xb = ctx.empty((D, Bc), DT)
shapes = {
    "fwd  NN W*X        ": lambda: K.gemm("N", "N", 1.0, W, X, 0.0, Y, D, Bc, D),
    "Abar NT Ybar*X'    ": lambda: K.gemm("N", "T", 1.0, Y, X, 0.0, Wb, D, D, Bc),
    "Bbar TN W'*Ybar    ": lambda: K.gemm("T", "N", 1.0, W, Y, 0.0, Xb, D, Bc, D),
}
flops = 2.0 * D * D * Bc
if os.environ.get("SHAPES"):
    sel = os.environ["SHAPES"].split(",")
    shapes = {k: v for k, v in shapes.items() if any(k.lower().startswith(x) for x in sel)}
TAIL = int(os.environ.get("TAIL", "0"))  # report only the last TAIL reps (power-capped steady state)
for mode in modes:
    ctx.math_mode = {"3xtf32": nb.MATH_3XTF32, "tf32": nb.MATH_TF32, "fp32": nb.MATH_FP32, "bf16x3": nb.MATH_BF16X3, "fp16x3": nb.MATH_FP16X3, "f64": nb.MATH_DEFAULT}[mode]
    for name, fn in shapes.items():
        for _ in range(2):
            fn()
        reps = int(os.environ.get("REPS", "5"))
        evs = [ctx.event() for _ in range(reps + 1)]
        ctx.record(evs[0])
        for i in range(reps):
            fn()
            ctx.record(evs[i + 1])
        ctx.sync()
        ts = [ctx.elapsed_ms(evs[i], evs[i + 1]) for i in range(reps)]
        if TAIL:
            ts = ts[-TAIL:]
        ts = sorted(ts)
        reps = len(ts)
        ms = sum(ts) / reps
        print(f"{mode:7s} {name} B={Bc}: avg {ms:8.3f} ms  min {ts[0]:8.3f}  med {ts[reps // 2]:8.3f}  max {ts[-1]:8.3f}  "
              f"{flops / ms / 1e9:8.1f} TFLOP/s (algorithmic)", flush=True)
