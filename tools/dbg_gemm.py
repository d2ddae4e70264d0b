"""Bring-up diagnostics for the tcgen05 GEMM: structured operands that expose layout/descriptor
mistakes (which of M / N / K is mapped wrongly), per transpose case and math mode."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import kernels as K  # noqa: E402

ctx = nb.get_context()
m, n, k = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (256, 256, 64)))
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["tf32", "3xtf32"]
rng = np.random.default_rng(0)


def run(tA, tB, opA, opB):
    A = np.ascontiguousarray(opA.T) if tA == "T" else opA
    B = np.ascontiguousarray(opB.T) if tB == "T" else opB
    Ad, Bd = ctx.asarray(A.astype(np.float32)), ctx.asarray(B.astype(np.float32))
    Cd = ctx.full((m, n), -7.0, np.float32)
    K.gemm(tA, tB, 1.0, Ad, Bd, 0.0, Cd, m, n, k)
    return Cd.numpy().astype(np.float64)


cases = {
    "ones": (np.ones((m, k)), np.ones((k, n))),
    "rowidx": (np.tile(np.arange(m)[:, None] % 64, (1, k)).astype(float), np.ones((k, n))),
    "colidx": (np.ones((m, k)), np.tile(np.arange(n)[None, :] % 64, (k, 1)).astype(float)),
    "kidx": (np.tile(np.arange(k)[None, :] % 8 == 0, (m, 1)).astype(float), np.tile((np.arange(k) % 32)[:, None], (1, n)).astype(float)),
    "random": (rng.standard_normal((m, k)), rng.standard_normal((k, n))),
}
for mode in modes:
    ctx.math_mode = {"3xtf32": nb.MATH_3XTF32, "tf32": nb.MATH_TF32, "fp32": nb.MATH_FP32, "bf16x3": nb.MATH_BF16X3}[mode]
    for tA in "NT":
        for tB in "NT":
            for name, (opA, opB) in cases.items():
                ref = opA @ opB
                got = run(tA, tB, opA, opB)
                err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
                line = f"{mode:7s} {tA}{tB} {name:7s} err={err:9.3e} |got|={np.linalg.norm(got):10.4g} |ref|={np.linalg.norm(ref):10.4g}"
                if err > 1e-2:
                    line += f" nz={np.count_nonzero(got)} untouched={(got == -7.0).sum()} nan={np.isnan(got).sum()}"
                    line += " got[0:3,0:4]=" + np.array2string(got[0:3, 0:4], precision=3).replace("\n", "")
                    line += " ref[0:3,0:4]=" + np.array2string(ref[0:3, 0:4], precision=3).replace("\n", "")
                print(line, flush=True)
