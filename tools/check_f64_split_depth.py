"""How much of the rtol 1e-10 budget a CHAIN of Float64 split products uses: whole gradient of a five-layer Float64 MLP
whose dozen 4096^3 products all run on the multi-launch Ozaki split, against the same gradient with every product on
DMMA (exact FP64 FMAs; within 8e-14 of the CPU oracle — tests/test_gpu_parity.py pins both against the oracle).
Measured: seven slices 9e-13, six slices (NB_OZAKI_SLICES=6) 8e-11."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import workloads as wl  # noqa: E402

dims, B = (784, 4096, 4096, 4096, 4096, 10), 4096
W, b, X, Y = wl.mlp_inputs(dims, B, np.float64, seed=4)
nl = len(dims) - 1
ctx = nb.get_context()
Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
grads = {}
for mode, name in ((nb.MATH_FP32, "dmma"), (nb.MATH_DEFAULT, "split")):
    ctx.math_mode = mode
    grads[name] = [g.numpy() for g in nb.nabla(wl.mlp_objective(nb, Xd, Yd, 1e-3, nl))(*W, *b)]
errs = [np.linalg.norm(g - r) / np.linalg.norm(r) for g, r in zip(grads["split"], grads["dmma"])]
print("split vs dmma, max rel grad diff", f"{max(errs):.2e}", ["%.1e" % e for e in errs], flush=True)
