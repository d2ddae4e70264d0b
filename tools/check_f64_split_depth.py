"""Whole-gradient error of a deeper Float64 MLP whose products all run on the multi-launch Ozaki split (4096^3),
against the Float64 CPU oracle: how much of the 1e-10 budget a chain of split products uses."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nabla_jl_b200 as nb  # noqa: E402
import nabla_oracle as orc  # noqa: E402
from nabla_jl_b200 import workloads as wl  # noqa: E402

dims, B = (784, 4096, 4096, 4096, 4096, 10), 4096
W, b, X, Y = wl.mlp_inputs(dims, B, np.float64, seed=4)
nl = len(dims) - 1
t0 = time.time()
ref = orc.nabla(wl.mlp_objective(orc, X, Y, 1e-3, nl))(*W, *b)
print(f"oracle {time.time() - t0:.1f} s", flush=True)
ctx = nb.get_context()
Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
for mode, name in ((nb.MATH_DEFAULT, "split"), (nb.MATH_FP32, "dmma")):
    ctx.math_mode = mode
    got = nb.nabla(wl.mlp_objective(nb, Xd, Yd, 1e-3, nl))(*W, *b)
    errs = [np.linalg.norm(g.numpy() - r) / np.linalg.norm(r) for g, r in zip(got, ref)]
    print(name, "max rel grad err", f"{max(errs):.2e}", ["%.1e" % e for e in errs], flush=True)
