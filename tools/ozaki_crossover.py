"""Where the Ozaki split starts to beat DMMA for Float64 products (run with NB_OZAKI_MIN_MNK=0 to force the split)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import kernels as K  # noqa: E402

ctx = nb.get_context()
rng = np.random.default_rng(0)
for (m, n, k) in ((512, 4096, 784), (512, 4096, 512), (1024, 1024, 1024), (1024, 4096, 1024), (2048, 2048, 2048),
                  (2048, 4096, 2048), (4096, 4096, 1024), (4096, 4096, 4096)):
    A, B = ctx.asarray(rng.standard_normal((m, k))), ctx.asarray(rng.standard_normal((k, n)))
    Cd = ctx.empty((m, n), np.float64)
    row = []
    for mode, name in ((nb.MATH_DEFAULT, "ozaki"), (nb.MATH_FP32, "dmma")):
        ctx.math_mode = mode
        for _ in range(2):
            K.gemm("N", "N", 1.0, A, B, 0.0, Cd, m, n, k)
        e0, e1 = ctx.event(), ctx.event()
        ctx.record(e0)
        for _ in range(10):
            K.gemm("N", "N", 1.0, A, B, 0.0, Cd, m, n, k)
        ctx.record(e1)
        ctx.sync()
        row.append(f"{name} {ctx.elapsed_ms(e0, e1) / 10 * 1e3:8.1f} us")
    ctx.math_mode = nb.MATH_DEFAULT
    print(f"{m}x{n}x{k} (mnk {m * n * k:.1e}): " + "   ".join(row), flush=True)
