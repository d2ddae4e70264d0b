"""BASELINE config 3: s = sum(tanh.(A .+ b)) with its full reverse sweep at
1e6 .. 1e9 elements, Float32 and Float64, b::(m,) and b::(1 x n).  CUDA-event time of the whole gradient evaluation
through the tape (forward `.+`, `tanh.`, `sum`; reverse tanh-pullback reading the saved y, bias sensitivity with the
fused unbroadcast; Ā aliases t̄1), and the HBM fraction against the bytes those launches have to move:
8·N·s (+ vector terms).  At 1e9 elements the result is also checked through size-independent properties.
Usage: python tools/bench_c3_sweep.py [max_elems]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nabla_jl_b200 as nb  # noqa: E402
from nabla_jl_b200 import kernels as K  # noqa: E402
from nabla_jl_b200 import workloads as wl  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ctx = nb.get_context()
peak = 6552.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
max_elems = float(sys.argv[1]) if len(sys.argv) > 1 else 2e9
out = {}
for dt in (np.float32, np.float64):
    s = np.dtype(dt).itemsize
    for (m, n) in ((1024, 1024), (4096, 2560), (8192, 12288), (32768, 32768)):
        N = m * n
        if N > max_elems:
            continue
        rng = np.random.default_rng(2)
        tile = rng.standard_normal((m, 64)).astype(dt)
        A = ctx.empty((m, n), dt)
        K.bcast_reduce(None, [ctx.asarray(tile).reshape((m, 64, 1))], A.reshape((m, 64, n // 64)))
        for row in (False, True):
            bh = rng.standard_normal((1, n) if row else (m,)).astype(dt)
            b = ctx.asarray(bh)
            f = nb.nabla(wl.bcast_sum(nb), get_output=True)
            for _ in range(3):
                y, (gA, gb) = f(A, b)
            reps = 10 if N < 5e8 else 4
            e0, e1 = ctx.event(), ctx.event()
            ctx.record(e0)
            for _ in range(reps):
                y, (gA, gb) = f(A, b)
            ctx.record(e1)
            ctx.sync()
            ms = ctx.elapsed_ms(e0, e1) / reps
            nbytes = 8 * N * s + 3 * bh.size * s
            key = f"{np.dtype(dt).name}_{m}x{n}_{'brow' if row else 'bcol'}"
            out[key] = {"elements": N, "ms_per_gradient": ms, "algorithmic_bytes": nbytes,
                        "GB/s": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak}
            print(key, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in out[key].items()}, flush=True)
            if N >= 1e9 and not row:
                # parity at full size through size-independent properties (tests/test_gpu_parity.py does 1e8)
                ref_tile = 1.0 - np.tanh(tile.astype(np.float64) + bh[:, None].astype(np.float64)) ** 2
                gh = gA.numpy()[:, -64:].astype(np.float64)
                e_a = np.linalg.norm(gh - ref_tile) / np.linalg.norm(ref_tile)
                e_b = np.linalg.norm(gb.numpy() - ref_tile.sum(1) * (n // 64)) / np.linalg.norm(ref_tile.sum(1) * (n // 64))
                s_ref = np.tanh(tile.astype(np.float64) + bh[:, None]).sum() * (n // 64)
                e_s = abs(float(y.val.numpy()) - s_ref) / abs(s_ref)
                tol = 1e-4 if dt == np.float32 else 1e-10
                print(f"  parity at {N:.2e} elements: Abar {e_a:.2e}  bbar {e_b:.2e}  s {e_s:.2e}  (tol {tol:.0e})", flush=True)
                assert max(e_a, e_b, e_s) < tol
                out[key]["parity"] = {"Abar": e_a, "bbar": e_b, "s": e_s, "tol": tol}
            del y, gA, gb, b
        del A
print(json.dumps(out))
