"""Float64 GEMM through the Ozaki split (kind::i8) against NumPy and against DMMA: accuracy and time."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import kernels as K  # noqa: E402

ctx = nb.get_context()
rng = np.random.default_rng(0)
SHAPES = ((4096, 4096, 4096), (4100, 4200, 4000), (8192, 2048, 8192))
if os.environ.get("OZ_SHAPES"):   # e.g. OZ_SHAPES=4100x4200x4000,4096x4096x4096
    SHAPES = tuple(tuple(int(v) for v in s.split("x")) for s in os.environ["OZ_SHAPES"].split(","))
ITERS = int(os.environ.get("OZ_ITERS", "3"))
for (m, n, k) in SHAPES:
    for tA in os.environ.get("OZ_TA", "NT"):
        for tB in os.environ.get("OZ_TB", "NT"):
            A = rng.standard_normal((k, m) if tA == "T" else (m, k))
            B = rng.standard_normal((n, k) if tB == "T" else (k, n))
            C0 = rng.standard_normal((m, n))
            Ad, Bd = ctx.asarray(A), ctx.asarray(B)
            ref = 0.7 * ((A.T if tA == "T" else A) @ (B.T if tB == "T" else B)) + 0.5 * C0
            res = {}
            for mode, name in ((nb.MATH_DEFAULT, "ozaki"), (nb.MATH_FP32, "dmma")):
                ctx.math_mode = mode
                Cd = ctx.asarray(C0)
                K.gemm(tA, tB, 0.7, Ad, Bd, 0.5, Cd, m, n, k)
                got = Cd.numpy()
                err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
                emax = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
                e0, e1 = ctx.event(), ctx.event()
                ctx.record(e0)
                for _ in range(ITERS):
                    K.gemm(tA, tB, 0.7, Ad, Bd, 0.0, Cd, m, n, k)
                ctx.record(e1)
                ctx.sync()
                ms = ctx.elapsed_ms(e0, e1) / ITERS
                res[name] = (err, emax, ms)
            ctx.math_mode = nb.MATH_DEFAULT
            print(f"{tA}{tB} {m}x{n}x{k}: " + "  ".join(
                f"{nm}: rel {e:.2e} max {x:.2e} {ms:.2f} ms = {2.0 * m * n * k / ms / 1e9:.1f} TFLOP/s" for nm, (e, x, ms) in res.items()),
                flush=True)
