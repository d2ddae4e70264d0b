"""Profiling driver for the skinny-shape products of the wide MLP (csrc/gemm_skinny.cu) and the streaming forward
epilogue: each shape a few times, so `ncu -k regex:k_skinny|k_gemm_smallk|k_bias_act|k_pack` can capture them, and
CUDA-event times per shape against the HBM bound.  Usage: python tools/prof_skinny.py [batch]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import kernels as K  # noqa: E402
from nabla_jl_b200.fusion import nb_gemm_epilogue  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32768
dt = np.float64 if "--f64" in sys.argv else np.float32
s = np.dtype(dt).itemsize
ctx = nb.get_context()
rng = np.random.default_rng(0)
H = ctx.asarray(np.tanh(rng.standard_normal((8192, 64))).astype(dt))
H4 = ctx.empty((8192, B), dt)
K.bcast_reduce(None, [H.reshape((8192, 64, 1))], H4.reshape((8192, 64, B // 64)))
W5 = ctx.asarray((0.1 * rng.standard_normal((10, 8192))).astype(dt))
b5 = ctx.asarray(rng.standard_normal(10).astype(dt))
Z5 = ctx.asarray(rng.standard_normal((10, B)).astype(dt))
logits = ctx.empty((10, B), dt)
W5bar = ctx.empty((10, 8192), dt)
H4bar = ctx.empty((8192, B), dt)
ops = nb._lib.OPCODES


def fused(tA, tB, m, n, k, A, lda, Bm, ldb, out, kind, op, bias=None, y=None, math=nb.MATH_BF16X3):
    epi = nb_gemm_epilogue()
    epi.kind, epi.op = kind, ops[op]
    epi.bias = bias.ptr if bias is not None else None
    epi.y, epi.ldy = (y.ptr, m) if y is not None else (None, 0)
    epi.emit_split = 1
    nb._lib.check(ctx.L.nb_gemm_fused(ctx.h, tA.encode(), tB.encode(), m, n, k, 1.0, A.ptr, lda, Bm.ptr, ldb, out.ptr, m,
                                      out.nb_dtype, math, C.byref(epi)), ctx.h)


cases = {
    "NN_10xBx8192 logits=logistic(W5*H4+b5)": (lambda: fused("N", "N", 10, B, 8192, W5, 10, H4, 8192, logits, 1, "logistic", bias=b5,
                                                             math=nb.MATH_DEFAULT), 8192 * B * s),
    "NT_10x8192xB W5bar=Z5bar*H4'": (lambda: K.gemm("N", "T", 1.0, Z5, H4, 0.0, W5bar, 10, 8192, B, math=nb.MATH_BF16X3), 8192 * B * s),
    "TN_8192xBx10 H4bar=(W5'*Z5bar).*tanh'(H4) + split": (lambda: fused("T", "N", 8192, B, 10, W5, 10, Z5, 10, H4bar, 2, "tanh", y=H4),
                                                          8192 * B * (2 * s + (4 if dt == np.float32 else 0))),
}
peak = 6552.0
for name, (fn, nbytes) in cases.items():
    for _ in range(3):
        fn()
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    reps = 10
    for _ in range(reps):
        fn()
    ctx.record(e1)
    ctx.sync()
    ms = ctx.elapsed_ms(e0, e1) / reps
    print(f"{name}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s  {nbytes / ms / 1e6 / peak:.2f} of HBM peak", flush=True)
