"""Device check of the 16-bit split modes: per-product error of FP16X3 / BF16X3 / mixed-format products against
NumPy Float64, and the whole-gradient error of the width-8192 MLP (B = 512) against the Float64 oracle for the
policies selectable with NB_F32_FWD_MATH / NB_F32_ADJ_MATH / NB_GEMM_MIXED_FMT.  Run on a GPU box."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nabla_jl_b200 as nb
from nabla_jl_b200 import kernels as K, workloads as wl

ctx = nb.get_context()
rng = np.random.default_rng(0)


def rel(a, r):
    return float(np.linalg.norm((np.asarray(a, np.float64) - r).ravel()) / np.linalg.norm(r.ravel()))


def product(mode, m, n, k, tA="N", tB="N", scaleA=1.0, scaleB=1.0, prime_a=None, prime_b=None):
    A = (scaleA * rng.standard_normal((k, m) if tA == "T" else (m, k))).astype(np.float32)
    B = (scaleB * rng.standard_normal((n, k) if tB == "T" else (k, n))).astype(np.float32)
    Ad, Bd = ctx.asarray(A), ctx.asarray(B)
    out = ctx.empty((m, n), np.float32)
    for prime, X, t in ((prime_a, Ad, tA), (prime_b, Bd, tB)):
        if prime is not None:   # put a split of this operand in the cache in another format first
            r, c = X.shape
            eye = ctx.asarray(np.eye(c, 128, dtype=np.float32))
            tmp = ctx.empty((r, 128), np.float32)
            K.gemm("N", "N", 1.0, X, eye, 0.0, tmp, r, 128, c, math=prime)
    K.gemm(tA, tB, 1.0, Ad, Bd, 0.0, out, m, n, k, math=mode)
    ref = (A.T if tA == "T" else A).astype(np.float64) @ (B.T if tB == "T" else B).astype(np.float64)
    return rel(out.numpy(), ref)


if "--products" in sys.argv or len(sys.argv) == 1:
    for (m, n, k) in ((512, 512, 512), (2304, 2304, 1024), (8192, 512, 8192)):
        for tA, tB in (("N", "N"), ("N", "T"), ("T", "N")):
            e = {nm: product(md, m, n, k, tA, tB) for nm, md in (("fp16x3", nb.MATH_FP16X3), ("bf16x3", nb.MATH_BF16X3),
                                                                  ("3xtf32", nb.MATH_3XTF32))}
            e["bf16x3 A:fp16 cached"] = product(nb.MATH_BF16X3, m, n, k, tA, tB, prime_a=nb.MATH_FP16X3)
            e["bf16x3 B:fp16 cached"] = product(nb.MATH_BF16X3, m, n, k, tA, tB, prime_b=nb.MATH_FP16X3)
            print(f"{m}x{n}x{k} {tA}{tB}", {k_: f"{v:.2e}" for k_, v in e.items()}, flush=True)
    print("tiny / huge magnitudes (fp16x3):", f"{product(nb.MATH_FP16X3, 512, 512, 512, scaleA=1e-20, scaleB=1e12):.2e}",
          f"{product(nb.MATH_FP16X3, 512, 512, 512, scaleA=1e15, scaleB=1e-9):.2e}")

if "--mlp" in sys.argv or len(sys.argv) == 1:
    import nabla_oracle as orc
    dims = (784, 8192, 8192, 8192, 8192, 10)
    B = 512
    W, b, X, Y = wl.mlp_inputs(dims, B, np.float32, seed=4)
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    t0 = time.time()
    yr, gr = orc.nabla(wl.mlp_objective(orc, f64(X), f64(Y), 1e-3, 5), get_output=True)(*[f64(p) for p in (*W, *b)])
    print("oracle", time.time() - t0, "s", flush=True)
    Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
    pol = {"default (fwd fp16x3, adj bf16x3 + cached fp16)": (None, None),
           "fwd 3xtf32, adj bf16x3": (nb.MATH_3XTF32, None),
           "all fp16x3": (nb.MATH_FP16X3, nb.MATH_FP16X3),
           "all 3xtf32": (nb.MATH_3XTF32, nb.MATH_3XTF32),
           "all bf16x3": (nb.MATH_BF16X3, nb.MATH_BF16X3),
           "all fp32 simt": (nb.MATH_FP32, nb.MATH_FP32)}
    for name, (f, a) in pol.items():
        ctx._fwd_math, ctx._adj_math = f, a
        y, g = nb.nabla(wl.mlp_objective(nb, Xd, Yd, 1e-3, 5), get_output=True)(*W, *b)
        errs = [rel(x.numpy(), r) for x, r in zip(g, gr)]
        print(name, "y", f"{abs(float(nb.to_host(y)) - float(orc.unbox(yr))) / abs(float(orc.unbox(yr))):.1e}",
              "max", f"{max(errs):.2e}", [f"{e:.1e}" for e in errs], flush=True)
