"""Accuracy of the one-launch Ozaki variant (mid-size Float64 products) against NumPy; run with NB_OZAKI_MIN_MNK=0."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import kernels as K  # noqa: E402

ctx = nb.get_context()
rng = np.random.default_rng(1)
for (m, n, k) in ((500, 4096, 784), (300, 4096, 500), (1100, 1300, 2000), (2048, 4096, 2048), (130, 129, 257)):
    for tA in "NT":
        for tB in "NT":
            A = rng.standard_normal((k, m) if tA == "T" else (m, k))
            B = rng.standard_normal((n, k) if tB == "T" else (k, n))
            C0 = rng.standard_normal((m, n))
            ref = 0.7 * ((A.T if tA == "T" else A) @ (B.T if tB == "T" else B)) - 0.5 * C0
            Cd = ctx.asarray(C0)
            n0 = ctx.stats()["kernel_launches"]
            K.gemm(tA, tB, 0.7, ctx.asarray(A), ctx.asarray(B), -0.5, Cd, m, n, k)
            nl = ctx.stats()["kernel_launches"] - n0
            got = Cd.numpy()
            err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            print(f"{tA}{tB} {m}x{n}x{k}: launches {nl} rel err {err:.2e} max {np.max(np.abs(got - ref)) / np.max(np.abs(ref)):.2e}", flush=True)
            assert err < 1e-11
print("ok")
