"""K-chunk sweep of the 16-bit split GEMM: the tensor core adds every tcgen05.mma into the FP32 TMEM accumulator with
truncation, so one product's error grows linearly with the K range accumulated before the epilogue folds it into its
register sum (correctly rounded).  Prints error (8192 x 512 x 8192, vs NumPy Float64) and time (8192 x 32768 x 8192)
per chunk length and mode.  Run on a GPU box."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb
from nabla_jl_b200 import kernels as K

ctx = nb.get_context()
rng = np.random.default_rng(0)
m, n, k = 8192, 512, 8192
A = rng.standard_normal((m, k)).astype(np.float32)
B = rng.standard_normal((k, n)).astype(np.float32)
ref = A.astype(np.float64) @ B.astype(np.float64)
Ad, Bd = ctx.asarray(A), ctx.asarray(B)
out = ctx.empty((m, n), np.float32)
N2 = int(os.environ.get("SWEEP_N", "32768"))
B2 = ctx.asarray(rng.standard_normal((k, N2)).astype(np.float32))
out2 = ctx.empty((m, N2), np.float32)
for mode, nm in ((nb.MATH_FP16X3, "fp16x3"), (nb.MATH_BF16X3, "bf16x3"), (nb.MATH_3XTF32, "3xtf32")):
    for chunk in (2048, 1024, 512, 256, 128, 64):
        os.environ["NB_GEMM_KCHUNK"] = str(chunk)
        K.gemm("N", "N", 1.0, Ad, Bd, 0.0, out, m, n, k, math=mode)
        err = np.linalg.norm(out.numpy() - ref) / np.linalg.norm(ref)
        for _ in range(2):
            K.gemm("N", "N", 1.0, Ad, B2, 0.0, out2, m, N2, k, math=mode)
        e0, e1 = ctx.event(), ctx.event()
        ctx.record(e0)
        reps = 5
        for _ in range(reps):
            K.gemm("N", "N", 1.0, Ad, B2, 0.0, out2, m, N2, k, math=mode)
        ctx.record(e1)
        ctx.sync()
        ms = ctx.elapsed_ms(e0, e1) / reps
        print(f"{nm} kchunk={chunk}: rel err {err:.2e}  {ms:.3f} ms  {2.0 * m * N2 * k / ms / 1e9:.1f} TFLOP/s", flush=True)
