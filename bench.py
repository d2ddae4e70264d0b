#!/usr/bin/env python
"""bench.py — gradient evaluations / second of the batch-sharded examples (BASELINE.json metric), with the
per-kernel roofline figures, a gradient parity check at the benchmarked size and rank count, and the CPU oracle
timed beside it.

One "step" = one gradient evaluation of the objective on the global batch: forward pass + full reverse sweep
producing the objective and EVERY parameter sensitivity, and (N>1) the NCCL allreduce(sum) of those sensitivities.

  python bench.py [--gpus N] [--steps K] [--warmup W]     wide MLP 784-8192^4-10, Float32, global batch 32768
                                                           (BASELINE configs[4], the config the metric is quoted on)
  python bench.py --workload vae  [--dtype f64]            examples/vae.jl (configs[3]), batch 4096, sharded the same way
  python bench.py --workload mlp  [--dtype f64]            examples/mlp.jl 784-500-300-10 (configs[0]), batch 4096
  python bench.py --scaling weak --per-gpu-batch 8192      per-GPU batch fixed (default: strong, global batch fixed)
  python bench.py --impl reference ...                     the CPU oracle (NumPy port of the reference) on the host cores
  (N>1: launched under torchrun, one rank per GPU)

``value``   : inputs resident in HBM, CUDA events on the sweep stream, max over ranks.
``e2e``     : the same step through the public API with HOST (pinned) batch buffers: H2D of this rank's data shard,
              sweep, allreduce, D2H of the objective — all inside the timed region; ``e2e_with_gradients`` also copies
              every parameter sensitivity (what ∇(f)(args...) returns, core.jl:214-229) back to pinned host memory.
``parity``  : after the timed regions the allreduced gradient of the benchmarked step is compared on rank 0 with an
              independent single-rank evaluation of the WHOLE global batch in column chunks of a different size, in Float64
              on the device (a Float32 run is measured against the Float64 gradient of the same inputs); the run FAILS
              above 1e-4 (Float32) / 1e-10 (Float64), and every rank's copy of the reduced gradient must be bit-identical.
``roofline``: the dominant kernel (dense GEMM sensitivities) timed per launch with CUDA events in a second, eager timed
              region of the same steps; ``extra`` carries the HBM-bound kernels (C3 broadcast pullback) and the small
              configs (C1 MLP, C2 quadratic forms, C4 VAE; CUDA-graph replay).
"""
import argparse
import json
import os
import subprocess
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAM = 1e-3


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"],
                "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


# ---------------------------------------------------------------------------------------------
# workloads: the BASELINE configurations as (parameters, column-sharded data, data term, shared term)
# ---------------------------------------------------------------------------------------------
class MLPWorkload:
    """examples/mlp.jl mlp_log_joint; log-likelihood = data term, log-prior = shared term."""

    def __init__(self, dims, label):
        self.dims, self.label = tuple(dims), label
        self.nl = len(dims) - 1

    def inputs(self, B, dtype):
        from nabla_jl_b200 import workloads as wl
        W, b, X, Y = wl.mlp_inputs(self.dims, B, dtype, seed=4 if len(self.dims) > 4 else 0)
        return [*W, *b], [X, Y]

    def data_objective(self, ns, data, B_global):
        from nabla_jl_b200 import workloads as wl
        return wl.mlp_data_objective(ns, data[0], data[1], self.nl)

    def shared_objective(self, ns, B_global):
        from nabla_jl_b200 import workloads as wl
        return wl.mlp_prior_objective(ns, LAM, self.nl)

    def full_objective(self, ns, data, B_global):
        from nabla_jl_b200 import workloads as wl
        return wl.mlp_objective(ns, data[0], data[1], LAM, self.nl)

    def flops(self, B):
        """forward GEMMs + Ā and B̄ GEMMs, minus the first layer's never-forced Wᵀ·Ȳ (X is not a Node: core.jl:151)"""
        d = self.dims
        params = sum(d[i] * d[i + 1] for i in range(len(d) - 1))
        return 2.0 * B * (3 * params - d[0] * d[1])

    def nparams(self):
        d = self.dims
        return sum(d[i] * d[i + 1] + d[i + 1] for i in range(len(d) - 1))

    def describe(self):
        return (f"{self.label} {'-'.join(map(str, self.dims))} (examples/mlp.jl mlp_log_joint), batch-sharded gradient "
                "+ NCCL allreduce")


class VAEWorkload:
    """examples/vae.jl:67-90 (L = 1): scal * (loglik - KL) = data term, log-prior + constant = shared term."""

    dx, dz, dh = 784, 10, 500
    label = "VAE"

    def inputs(self, B, dtype):
        from nabla_jl_b200 import workloads as wl
        params, X, noise = wl.vae_inputs(self.dx, self.dz, self.dh, B, dtype, seed=3)
        return list(params), [X, noise]

    def data_objective(self, ns, data, B_global):
        from nabla_jl_b200 import workloads as wl
        return wl.vae_data_objective(ns, data[0], data[1], 60000.0 / B_global, self.dz)

    def shared_objective(self, ns, B_global):
        from nabla_jl_b200 import workloads as wl
        return wl.vae_shared_objective(ns, LAM, 60000.0 / B_global, self.dz)

    def full_objective(self, ns, data, B_global):
        from nabla_jl_b200 import workloads as wl
        return wl.vae_objective(ns, data[0], data[1], LAM, 60000.0 / B_global, self.dz)

    def flops(self, B):
        return 2.0 * B * (2 * 799000 + 407000)

    def nparams(self):
        return self.dh * self.dx + 2 * self.dz * self.dh + self.dh * self.dz + self.dx * self.dh

    def describe(self):
        return "VAE 784-500-10-500-784 (examples/vae.jl:67-90, L = 1), batch-sharded gradient + NCCL allreduce"


def make_workload(args):
    if args.workload == "wide_mlp":
        return MLPWorkload((784, 8192, 8192, 8192, 8192, 10), "wide MLP"), "BASELINE configs[4]"
    if args.workload == "mlp":
        return MLPWorkload((784, 500, 300, 10), "MLP"), "BASELINE configs[0]"
    return VAEWorkload(), "BASELINE configs[3]"


def defaults_for(args):
    if args.batch is None:
        args.batch = 32768 if args.workload == "wide_mlp" else 4096
    if args.dtype is None:
        args.dtype = "f32" if args.workload == "wide_mlp" else "f64"
    if args.scaling == "weak":
        args.batch = args.per_gpu_batch * max(1, args.gpus)
    if args.capture is None:
        args.capture = args.workload != "wide_mlp"
    return args


def workload_config(args, n, wk, cfg):
    return {"workload": f"{wk.describe()} ({cfg})", "global_batch": args.batch, "per_gpu_batch": args.batch // n,
            "math_mode": (args.math if args.dtype == "f32" else "f64: Ozaki int8 split on tcgen05 for m*n*k >= 3e9, DMMA below"),
            "parallelism": f"dp{n} (batch columns), allreduce(sum) of {wk.nparams() / 1e6:.1f}M parameter sensitivities",
            "capture": bool(args.capture),
            "l2": ("working set (tens of GB of activations/sensitivities) >> 126 MB L2; no flush needed"
                   if args.workload == "wide_mlp" else
                   "small model: part of the working set is L2-resident between steps, as in any optimiser loop")}


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU reference is entitled to every host core."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)   # BLAS and OpenMP pools
    except Exception:
        pass


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p["num_threads"] for p in threadpool_info() if p.get("user_api") == "blas"]
        if n:
            return int(max(n))
    except Exception:
        pass
    return os.cpu_count() or 1


def _best(fn, reps):
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


class OracleTimer:
    """The NumPy oracle (op-for-op port of the reference sweep: un-fused Branches, one pullback pass per argument,
    tmp + sum! unbroadcast, add!! accumulate; OpenBLAS threads = host cores, elementwise single-threaded like Julia's
    broadcast!) timed on a BOUNDED sample of the workload.

    Small configurations (VAE, MLP at batch 4096: 0.1-0.4 s per evaluation) are timed whole.  The wide MLP at batch
    32768 (tens of seconds and ~40 GB per evaluation) is timed as  t(B) = t_shared + t_data(B): the batch-independent
    shared term (prior over 207.8 M parameters) is timed on its own, the data term — whose cost is proportional to the
    number of columns — at 4 x ``cols`` columns per step (4096 by default: the per-column cost has converged there,
    at 1024 it is ~15 % higher) and, once, at ``cols``, 2x and 4x ``cols`` (best of two each) to measure the linearity
    the extrapolation relies on (reported as ``fit``)."""

    def __init__(self, wk, args, dtype):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import nabla_oracle as orc
        self.orc, self.wk, self.args = orc, wk, args
        self.B = args.batch
        self.whole = args.workload != "wide_mlp" or self.B <= 4 * args.cpu_cols
        self.cols = self.B if self.whole else args.cpu_cols
        ncols = self.B if self.whole else min(self.B, 4 * self.cols)
        self.params, self.data = wk.inputs(ncols, dtype)
        self.step_cols = ncols      # columns of the per-step sample
        self.t_shared = None
        self.fit = None

    def _data_seconds(self, cols, reps=1):
        orc, wk = self.orc, self.wk
        data = [d[:, :cols] for d in self.data]
        f = orc.nabla(wk.data_objective(orc, data, self.B))
        return _best(lambda: f(*self.params), reps)

    def prepare(self):
        orc, wk = self.orc, self.wk
        if self.whole:
            return
        fs = orc.nabla(wk.shared_objective(orc, self.B))
        self.t_shared = _best(lambda: fs(*self.params), 3)
        pts = [(c, self._data_seconds(c, 2)) for c in (self.cols, 2 * self.cols, 4 * self.cols) if c <= self.data[0].shape[1]]
        xs, ys = np.array([p[0] for p in pts], float), np.array([p[1] for p in pts], float)
        slope = float(np.sum(xs * ys) / np.sum(xs * xs))       # through the origin: the data term is linear in B
        resid = float(np.max(np.abs(ys - slope * xs) / ys))
        self.fit = {"cols": [int(x) for x in xs], "seconds": [float(y) for y in ys], "seconds_per_column": slope,
                    "max_rel_residual": resid, "t_shared_s": self.t_shared}

    def step_seconds(self):
        """One bounded sample -> the modelled seconds of one evaluation at the full batch."""
        orc, wk = self.orc, self.wk
        if self.whole:
            f = orc.nabla(wk.full_objective(orc, self.data, self.B))
            return _best(lambda: f(*self.params), 1)
        return self.t_shared + self._data_seconds(self.step_cols) * (self.B / self.step_cols)

    def sample_text(self):
        if self.whole:
            return f"NumPy oracle, whole evaluation at B={self.B} per step (not extrapolated)"
        return (f"NumPy oracle: shared (prior) term timed whole + data term timed at {self.step_cols} columns per step, "
                f"scaled by {self.B}/{self.step_cols} (the data term is linear in the batch: measured at "
                f"{self.fit['cols']} columns, max residual {self.fit['max_rel_residual']:.1%}); extrapolated")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wk, cfg = make_workload(args)
    dtype = np.float32 if args.dtype == "f32" else np.float64
    use_all_host_threads()
    t0 = time.perf_counter()
    ot = OracleTimer(wk, args, dtype)
    ot.prepare()
    for _ in range(args.warmup):
        ot.step_seconds()
    per = [ot.step_seconds() for _ in range(args.steps)]
    sec = float(np.median(per))
    val = 1.0 / sec
    n = max(1, args.gpus)
    cb = {"value": val, "unit": "evals/s", "cores": host_threads(), "kind": "port", "sample": ot.sample_text(),
          "fit": ot.fit}
    print(json.dumps({
        "impl": "reference", "metric": "gradient evals/sec", "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic", "config": workload_config(args, n, wk, cfg), "cpu_baseline": cb,
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0}))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def timed_region(ctx, fn, steps):
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    timed_region.host_ms = (time.perf_counter() - t0) * 1e3   # host time to ENQUEUE the steps (no sync inside)
    ctx.record(e1)
    ctx.sync()
    ms = ctx.elapsed_ms(e0, e1)
    ctx.L.nb_event_destroy(e0)
    ctx.L.nb_event_destroy(e1)
    return ms


def _time_oracle(fn, reps=3):
    """Best-of-reps wall time of one oracle gradient evaluation on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nabla_oracle as orc
    best = _best(lambda: fn(orc), reps)
    return {"ms_per_eval": best * 1e3, "evals_per_s": 1.0 / best, "cores": host_threads()}


def extra_measurements(nb, ctx, peaks, quick):
    """HBM-bound kernels (C3) and the small configs (C1, C2, C4), N=1 only.  Short."""
    from nabla_jl_b200 import kernels as K, workloads as wl
    out = {}
    # C3: s = sum(tanh.(A .+ b)); kernels timed one by one, inputs (>= 400 MB) larger than L2
    for dt, nm in ((np.float32, "f32"), (np.float64, "f64")):
        m, n = 8192, 12288
        s = np.dtype(dt).itemsize
        N = m * n
        rng = np.random.default_rng(2)
        tile = ctx.asarray(rng.standard_normal((m, 64)).astype(dt))
        A = ctx.empty((m, n), dt)
        K.bcast_reduce(None, [tile.reshape((m, 64, 1))], A.reshape((m, 64, n // 64)))
        b = ctx.asarray(rng.standard_normal(m).astype(dt))
        t1 = ctx.empty((m, n), dt)
        y = ctx.empty((m, n), dt)
        tb1 = ctx.empty((m, n), dt)
        bbar = ctx.empty((m,), dt)
        Abar = ctx.empty((m, n), dt)
        one = ctx.scalar(1.0, dt)
        sout = ctx.scalar(0.0, dt)
        add = nb.compile_scalar(nb.add, 2)
        tanh = nb.compile_scalar(nb.tanh, 1)
        fused = nb.compile_scalar(lambda p, q: nb.tanh(p + q), 2)
        brow = ctx.asarray(rng.standard_normal((1, n)).astype(dt))
        browbar = ctx.empty((1, n), dt)
        cases = {
            # transposed-reduction variant (b :: 1 x n): unbroadcast over ROWS, one value per column
            "pullback_add_unbroadcast_brow": (lambda: K.bcast_pullback(add, tb1, None, [A, brow], [None, browbar],
                                                                       [False, False]), N * s + n * s),
            "fwd_add": (lambda: K.bcast_fwd(add, [A, b], t1), 2 * N * s + m * s),
            "fwd_tanh": (lambda: K.bcast_fwd(tanh, [t1], y), 2 * N * s),
            "fwd_sum": (lambda: K.bcast_reduce(None, [y], sout), N * s),
            # reverse of the tanh Branch: t̄1 = ȳ(scalar) * (1 - y²) reading the saved y
            "pullback_tanh_saved_y": (lambda: K.bcast_pullback(tanh, one, y, [t1], [tb1], [False]), 2 * N * s),
            # reverse of the `+` Branch for b: b̄ = sum over columns of t̄1 (Ā aliases t̄1: no kernel)
            "pullback_add_unbroadcast_b": (lambda: K.bcast_pullback(add, tb1, None, [A, b], [None, bbar], [False, False]),
                                           N * s + m * s),
            # the whole C3 reverse as ONE fused broadcast pullback of (a, b) -> tanh(a + b):
            # reads A, writes Ā and the unbroadcast b̄ (2·N·s + m·s, scalar ȳ)
            "pullback_fused_tanh_add": (lambda: K.bcast_pullback(fused, one, None, [A, b], [Abar, bbar], [False, False]),
                                        2 * N * s + 2 * m * s),
            # the same with the Branch's saved forward value, as the tape passes it (core.jl:76): tanh is not
            # re-evaluated; reads A and y, writes Ā and b̄ (3·N·s + 2·m·s)
            "pullback_fused_tanh_add_saved_y": (lambda: K.bcast_pullback(fused, one, y, [A, b], [Abar, bbar],
                                                                         [False, False]), 3 * N * s + 2 * m * s),
            "fwd_fused_tanh_add": (lambda: K.bcast_fwd(fused, [A, b], y), 2 * N * s + m * s),
        }
        res = {}
        for name, (fn, nbytes) in cases.items():
            for _ in range(3):
                fn()
            reps = 5 if quick else 20
            ms = timed_region(ctx, fn, reps) / reps
            gbs = nbytes / (ms * 1e-3) / 1e9
            res[name] = {"ms": ms, "algorithmic_bytes": nbytes, "GB/s": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]}
        out[f"c3_bcast_{nm}_8192x12288"] = res
        del A, t1, y, tb1, Abar
    # C1 / C4: small Float64 models, CUDA-graph replay of the whole sweep
    W, b, X, Y = wl.mlp_inputs((784, 500, 300, 10), 4096, np.float64, seed=0)
    Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
    params = [ctx.asarray(p) for p in (*W, *b)]
    obj = wl.mlp_objective(nb, Xd, Yd, LAM, 3)
    flops_c1 = 2.0 * 4096 * (3 * 545000 - 392000)
    for label, mk in (("eager", None), ("graph", True)):
        if mk is None:
            fn = lambda: nb.nabla(obj)(*params)
        else:
            cap = nb.CapturedGradient(obj, params, ctx)
            fn = cap
        for _ in range(3):
            fn()
        reps = 10 if quick else 50
        ms = timed_region(ctx, fn, reps) / reps
        out[f"c1_mlp_f64_B4096_{label}"] = {"ms_per_eval": ms, "evals_per_s": 1e3 / ms,
                                            "TFLOP/s": flops_c1 / (ms * 1e-3) / 1e12}
        if mk:
            out[f"c1_mlp_f64_B4096_{label}"]["launches_per_eval"] = cap.launches_per_replay
    out["c1_mlp_f64_B4096_cpu_oracle"] = _time_oracle(lambda orc: orc.nabla(wl.mlp_objective(orc, X, Y, LAM, 3))(*W, *b))
    # C2: quadratic forms at N = 4096, Float64
    Aq, xq, yq = wl.quad_inputs(4096)
    Aqd = ctx.asarray(Aq)
    for name, fn, args_ in (("c2_quad_symmetric_x'Ax", wl.quad_symmetric(nb, Aqd), (xq,)),
                            ("c2_quad_asymmetric_x'Ay", wl.quad_asymmetric(nb, Aqd), (xq, yq))):
        pa = [ctx.asarray(a) for a in args_]
        f = nb.CapturedGradient(fn, pa, ctx)      # CUDA-graph replay of the whole sweep
        for _ in range(3):
            f()
        reps = 10 if quick else 30
        ms = timed_region(ctx, f, reps) / reps
        nbytes = 2 * 4096 * 4096 * 8   # A streamed once forward and once in the reverse sweep
        out[name + "_f64_N4096"] = {"ms_per_eval": ms, "evals_per_s": 1e3 / ms, "GB/s": nbytes / (ms * 1e-3) / 1e9,
                                    "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                    "note": "two GEMV passes over A (134 MB, about the size of the L2); CUDA-graph replay"}
    # matrix form quad(A, B) = B'AB with Ybar = I (test/core.jl:206-218): Float64 products of 4096^3
    Bq = ctx.asarray(np.random.default_rng(5).standard_normal((4096, 4096)))
    eye = ctx.asarray(np.eye(4096))

    def quad_matrix():
        t = nb.Tape()
        A_, B_ = nb.Leaf(t, Aqd), nb.Leaf(t, Bq)
        Q = nb.matmul(nb.adjoint(B_), nb.matmul(A_, B_))
        rt = nb.nabla(Q, eye)
        t.tape.clear()
        return rt
    for _ in range(2):
        quad_matrix()
    reps = 3 if quick else 5
    ms = timed_region(ctx, quad_matrix, reps) / reps
    out["c2_quad_matrix_B'AB_f64_N4096"] = {"ms_per_eval": ms, "TFLOP/s": 6 * 2.0 * 4096 ** 3 / (ms * 1e-3) / 1e12,
                                            "note": "2 forward + 4 reverse Float64 products (4096^3 each: Ozaki int8 slice products on the tensor cores, csrc/ozaki.cu) and 3 transposes per evaluation"}
    del Aqd, Bq, eye
    vp, VX, noise = wl.vae_inputs(B=4096, seed=3)
    VXd, Nd = ctx.asarray(VX), ctx.asarray(noise)
    vparams = [ctx.asarray(p) for p in vp]
    vobj = wl.vae_objective(nb, VXd, Nd, LAM, 60000 / 4096, 10)
    cap = nb.CapturedGradient(vobj, vparams, ctx)
    for _ in range(3):
        cap()
    reps = 10 if quick else 50
    ms = timed_region(ctx, cap, reps) / reps
    out["c4_vae_f64_B4096_cpu_oracle"] = _time_oracle(
        lambda orc: orc.nabla(wl.vae_objective(orc, VX, noise, LAM, 60000 / 4096, 10))(*vp))
    out["c4_vae_f64_B4096_graph"] = {"ms_per_eval": ms, "evals_per_s": 1e3 / ms,
                                     "TFLOP/s": 2.0 * 4096 * (2 * 799000 + 407000) / (ms * 1e-3) / 1e12,
                                     "launches_per_eval": cap.launches_per_replay}
    return out


def reference_gradient(nb, ctx, wk, params, data_host, B_global, dtype, chunk):
    """Independent single-rank evaluation of the whole global batch in FLOAT64 on the device (the Float64 path is
    itself pinned to the CPU oracle at 1e-10 by tests/): the data term in column chunks (other GEMM shapes and another
    summation order than the sharded step), accumulated on the device, plus the shared term once.  A Float32 run is
    compared with the Float64 evaluation of the same (Float32-representable) inputs — the true error of the
    benchmarked gradient, not the difference of two roundings; a Float64 run with DMMA-only products in chunks."""
    from nabla_jl_b200.core import accumulate, unthunk
    old = ctx.math_mode
    ctx.math_mode = nb.MATH_DEFAULT if dtype == np.float32 else nb.MATH_FP32
    f64 = np.float64
    try:
        p64 = [ctx.asarray(p.numpy().astype(f64)) for p in params] if dtype == np.float32 else params
        tot, ytot = None, 0.0
        for lo in range(0, B_global, chunk):
            hi = min(B_global, lo + chunk)
            dd = [ctx.asarray(np.asfortranarray(d[:, lo:hi]).astype(f64)) for d in data_host]
            y, g = nb.nabla(wk.data_objective(nb, dd, B_global), get_output=True)(*p64)
            g = [unthunk(x) for x in g]
            ytot += float(nb.to_host(y))
            if tot is None:
                tot = [x.copy() if x.shared else x for x in g]
            else:
                tot = [accumulate(a, x) for a, x in zip(tot, g)]
            del dd, g
        y, g = nb.nabla(wk.shared_objective(nb, B_global), get_output=True)(*p64)
        ytot += float(nb.to_host(y))
        tot = [accumulate(a, unthunk(x)) for a, x in zip(tot, g)]
        ctx.sync()
        return ytot, tot
    finally:
        ctx.math_mode = old


def rel_err_device(nb, ctx, got, ref):
    """|got - ref| / |ref| in Float64 on the host, one parameter at a time (the arrays stay on the device otherwise)."""
    errs = []
    for a, r in zip(got, ref):
        ah, rh = a.numpy().astype(np.float64).ravel(), r.numpy().astype(np.float64).ravel()
        errs.append(float(np.linalg.norm(ah - rh) / max(np.linalg.norm(rh), 1e-300)))
    return errs


def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.gpus
    if world != n:
        if world == 1 and n > 1:
            raise SystemExit(f"--gpus {n} needs torchrun with {n} ranks (WORLD_SIZE={world})")
        n = world
        args.gpus = n
        if args.scaling == "weak":
            args.batch = args.per_gpu_batch * n
    import nabla_jl_b200 as nb
    wk, cfg = make_workload(args)
    peaks = load_peaks()
    dist = None
    if n > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ctx = nb.Context(device=local)
    nb.set_context(ctx)
    modes = {"default": nb.MATH_DEFAULT, "fp16x3": nb.MATH_FP16X3, "3xtf32": nb.MATH_3XTF32, "tf32": nb.MATH_TF32,
             "fp32": nb.MATH_FP32, "bf16x3": nb.MATH_BF16X3}
    ctx.math_mode = modes[args.math] if args.dtype == "f32" else nb.MATH_DEFAULT
    dtype = np.float32 if args.dtype == "f32" else np.float64
    B = args.batch

    def exchange(raw):
        obj = [raw]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    comm = nb.Communicator(ctx, rank, n, exchange)
    params_h, data_h = wk.inputs(B, dtype)
    lo, hi = nb.shard_bounds(B, rank, n)
    shard_h = [np.asfortranarray(d[:, lo:hi]) for d in data_h]
    if rank != 0 or args.no_parity:
        del data_h
    params = [ctx.asarray(p) for p in params_h]
    del params_h
    shard_d = [ctx.asarray(s) for s in shard_h]
    pinned = [ctx.pinned_empty(s.shape, dtype) for s in shard_h]
    for p, s in zip(pinned, shard_h):
        p.a[...] = s
    del shard_h
    yp = ctx.pinned_empty((), dtype)
    sg = nb.OverlappedShardedGradient(wk.data_objective(nb, shard_d, B), wk.shared_objective(nb, B), params, comm,
                                      capture=args.capture)

    def step():
        return sg.step()

    def step_e2e():
        for d, p in zip(shard_d, pinned):
            d.copy_from_pinned(p)
        y, _ = sg.step()
        y.copy_to_pinned(yp)
        ctx.sync()          # the host reads the objective every step

    gp = []

    def step_e2e_grads():
        for d, p in zip(shard_d, pinned):
            d.copy_from_pinned(p)
        y, g = sg.step()
        y.copy_to_pinned(yp)
        if not gp:
            gp.extend(ctx.pinned_empty(x.shape, dtype) for x in g)
        for x, p in zip(g, gp):
            x.copy_to_pinned(p)
        ctx.sync()

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    barrier()
    # ---- headline: K steps, inputs resident -----------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    stats0 = ctx.stats()
    barrier()
    ms = timed_region(ctx, step, args.steps)
    host_ms = timed_region.host_ms
    barrier()
    launches = ctx.stats()["kernel_launches"] - stats0["kernel_launches"]
    mallocs = ctx.stats()["cuda_mallocs"] - stats0["cuda_mallocs"]
    clocks = sampler.stop() if sampler else None
    ms = max_over_ranks(ms)
    # ---- the same steps again, eager, every GEMM launch bracketed by CUDA events (roofline) --------------------
    sg_prof = sg if not args.capture else nb.OverlappedShardedGradient(
        wk.data_objective(nb, shard_d, B), wk.shared_objective(nb, B), params, comm, capture=False)
    for _ in range(2):
        sg_prof.step()
    barrier()
    ctx.profile_begin()
    ms_prof = timed_region(ctx, lambda: sg_prof.step(), args.steps)
    barrier()
    recs = ctx.profile_end()
    ms_prof = max_over_ranks(ms_prof)
    # ---- end-to-end: host buffers in, objective out (and, second figure, every gradient out) -------------------
    for _ in range(2):
        step_e2e()
    barrier()
    l0 = ctx.stats()["kernel_launches"]
    ms_e2e = max_over_ranks(timed_region(ctx, step_e2e, args.steps))
    launches_e2e = ctx.stats()["kernel_launches"] - l0
    barrier()
    e2e_g = None
    if not args.no_e2e_grads:
        for _ in range(2):
            step_e2e_grads()
        barrier()
        ms_e2e_g = max_over_ranks(timed_region(ctx, step_e2e_grads, args.steps))
        barrier()
        e2e_g = {"value": 1e3 / (ms_e2e_g / args.steps), "unit": "evals/s", "ms_per_step": ms_e2e_g / args.steps,
                 "h2d_bytes_per_step": int(sum(p.nbytes for p in pinned)),
                 "d2h_bytes_per_step": int(yp.nbytes + sum(p.nbytes for p in gp)),
                 "note": "also copies every parameter sensitivity to pinned host memory on every rank, serially after "
                         "the step (not overlapped with the sweep)"}
        gp.clear()
    # ---- parity of the benchmarked step -----------------------------------------------------------------------
    parity = None
    if not args.no_parity:
        y, grads = sg.step()
        ctx.sync()
        # every rank must hold the same reduced gradient: compare CRCs of the raw bytes
        same = True
        if dist is not None:
            crc = [zlib.crc32(g.numpy().tobytes()) for g in grads]
            allc = [None] * n
            dist.all_gather_object(allc, crc)
            same = all(c == allc[0] for c in allc)
        if rank == 0:
            chunk = min(8192, B)
            if chunk == B // n:      # never the shard size itself: other GEMM shapes, another summation order
                chunk = max(1, chunk // 2)
            yr, gref = reference_gradient(nb, ctx, wk, params, data_h, B, dtype, chunk)
            errs = rel_err_device(nb, ctx, grads, gref)
            yerr = abs(float(y.numpy()) - yr) / max(abs(yr), 1e-300)
            tol = 1e-4 if dtype == np.float32 else 1e-10
            parity = {"parity_rel_err": max(errs), "objective_rel_err": yerr, "tolerance": tol,
                      "ranks_bit_identical": bool(same),
                      "reference": f"single-rank evaluation of the global batch in {chunk}-column chunks, "
                                   f"{'in Float64 on the device (Ozaki / DMMA products)' if dtype == np.float32 else 'DMMA-only Float64 products'}",
                      "per_parameter": errs}
            del gref
        del grads
    ms_step = ms / args.steps
    value = 1e3 / ms_step
    # ---- roofline of the dominant kernel family: the dense GEMM sensitivities ----------------------------------
    by_tag = {}
    for tag, work, unit, t in recs:
        if unit == "FLOP":
            d = by_tag.setdefault(tag, [0.0, 0, work, []])
            d[0] += t
            d[1] += 1
            d[3].append(t)
    roof = None
    phase_ms = {}
    for tag, work, unit, t in recs:
        if unit == "ms":
            phase_ms[tag] = phase_ms.get(tag, 0.0) + t
    gemm_ms_total = sum(v[0] for v in by_tag.values())
    if by_tag:
        flop = sum(v[2] * v[1] for v in by_tag.values())
        achieved = flop / (gemm_ms_total * 1e-3) / 1e12
        if args.dtype == "f32":
            ratio = {"default": 1.0, "fp16x3": 1.0, "bf16x3": 1.0, "3xtf32": 0.5, "tf32": 0.5, "fp32": 0.5}[args.math]
            peak = peaks["bf16_sustained"] * ratio
            passes = 1.0 if args.math in ("tf32", "fp32") else 3.0
            peak_note = (f"{peaks['source']} bf16 sustained {peaks['bf16_sustained']} TFLOP/s x {ratio} (kind::f16 for the "
                         "16-bit split modes FP16x3 / BF16x3, kind::tf32 at half that rate otherwise); achieved counts "
                         "ALGORITHMIC flops 2mnk: the split modes needed for rtol 1e-4 issue three tensor passes per "
                         "product, so they top out at 1/3 of this peak (tensor_work_frac = 3 x frac)")
        else:
            slices = min(7, max(2, int(os.environ.get("NB_OZAKI_SLICES", "7"))))
            npass = slices * (slices + 1) // 2
            peak, passes = 2.0 * peaks["bf16_sustained"] / npass, 1.0
            peak_note = (f"{peaks['source']} bf16 sustained {peaks['bf16_sustained']} TFLOP/s x 2 (kind::i8, assumed: "
                         f"MEASURED_PEAKS has no int8 figure) / {npass} int8 passes per Float64 product (Ozaki split, "
                         f"{slices} slices); small products on DMMA")
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r2_gemm_traffic.json")
        if args.workload == "wide_mlp" and args.dtype == "f32" and n == 1 and B == 32768 and os.path.exists(tpath):
            t = json.load(open(tpath))["per_gemm_call"]["NN_8192x32768x8192"]
            traffic = t["dram_read_bytes"] + t["dram_write_bytes"]
        nl = sum(v[1] for v in by_tag.values())
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_note": "DRAM bytes of one forward 8192x32768x8192 launch of the same kernel (FP16x3, 128-long K "
                                "chunks) from one ncu --set full capture, profiles/r2_gemm_traffic.json; algorithmic "
                                "bytes 2.42e9" if traffic else None,
                "kernel": "nb_gemm / nb_gemm_fused (dense GEMM sensitivities: forward A*B, Abar=Ybar*B', Bbar=A'*Ybar)",
                "peak_note": peak_note, "tensor_work_frac": passes * achieved / peak,
                "launches": nl, "share_of_step": gemm_ms_total / ms_prof,
                "timed_in": "second (eager) timed region of the same steps", "ms_per_step_in_that_region": ms_prof / args.steps,
                "avg_launch_ms": gemm_ms_total / max(1, nl),
                "per_shape": {k: {"launches": v[1], "avg_ms": v[0] / v[1], "min_ms": min(v[3]), "max_ms": max(v[3]),
                                  "TFLOP/s": v[2] / (v[0] / v[1] * 1e-3) / 1e12} for k, v in by_tag.items()}}
    line = {
        "metric": "gradient evals/sec", "value": value, "unit": "evals/s", "n_gpus": n,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args, n, wk, cfg),
        "TFLOP/s_algorithmic": wk.flops(B) / (ms_step * 1e-3) / 1e12,
        "clocks": clocks, "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_ms / args.steps,
        "non_gemm_ms_per_step": (ms_prof - gemm_ms_total) / args.steps,
        "step_breakdown_ms": {"gemm": gemm_ms_total / args.steps,
                              "shared_sweep": phase_ms.get("shared_sweep", 0.0) / args.steps,
                              "join_wait": phase_ms.get("join_wait", 0.0) / args.steps,
                              "other": (ms_prof - gemm_ms_total - phase_ms.get("shared_sweep", 0.0)
                                        - phase_ms.get("join_wait", 0.0)) / args.steps,
                              "note": "CUDA-event brackets on the sweep stream in the eager profiling region; join_wait = "
                                      "the part of the NCCL exchange not hidden behind the sweep"},
        "pool": {"cuda_mallocs_in_timed_regions": int(mallocs), "high_water_GB": ctx.stats()["pool_high_water"] / 1e9},
        "e2e": {"value": 1e3 / (ms_e2e / args.steps), "unit": "evals/s",
                "h2d_bytes_per_step": int(sum(p.nbytes for p in pinned)), "d2h_bytes_per_step": int(yp.nbytes),
                "ms_per_step": ms_e2e / args.steps, "gpu_launches": int(launches_e2e)},
        "e2e_with_gradients": e2e_g,
        "parity": parity,
        "notes": ("operand splits (hi/lo 16-bit copies of W, X and the activations) are cached per source array and "
                  "invalidated on every write: the parameters are not updated between steps here, so their splits "
                  "(~0.3 ms of HBM traffic per step for 207.8M parameters) are computed once, not per step")
        if args.dtype == "f32" and not args.capture else None,
        "roofline": roof,
    }
    failed = None
    if parity is not None and (parity["parity_rel_err"] > parity["tolerance"] or not parity["ranks_bit_identical"]
                               or parity["objective_rel_err"] > parity["tolerance"]):
        failed = f"gradient parity check failed: {parity}"
    if rank == 0 and n == 1 and not args.no_cpu:
        use_all_host_threads()
        ot = OracleTimer(wk, args, dtype)
        ot.prepare()
        sec = min(ot.step_seconds() for _ in range(2))
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "evals/s", "cores": host_threads(), "kind": "port",
                                "sample": ot.sample_text() + "; OpenBLAS threads = cores, elementwise single-threaded",
                                "fit": ot.fit}
        del ot
    if rank == 0 and n == 1 and not args.no_extra:
        del sg, sg_prof, params, shard_d
        try:
            line["extra"] = extra_measurements(nb, ctx, peaks, args.quick)
        except Exception as e:  # the headline number stands on its own
            line["extra"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    barrier()
    for drv in (locals().get("sg"), locals().get("sg_prof")):
        if drv is not None:
            drv.close()      # a captured step holds NCCL kernels: destroy the graph before the communicator
    comm.close()
    if dist is not None:
        dist.destroy_process_group()
    if failed and rank == 0:
        raise SystemExit(failed)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="wide_mlp", choices=["wide_mlp", "vae", "mlp"])
    ap.add_argument("--batch", type=int, default=None, help="GLOBAL batch (columns); default 32768 (wide_mlp) / 4096")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--per-gpu-batch", type=int, default=8192, help="--scaling weak: columns per GPU")
    ap.add_argument("--dtype", default=None, choices=["f32", "f64"])
    ap.add_argument("--math", default="default", choices=["default", "fp16x3", "bf16x3", "3xtf32", "tf32", "fp32"])
    ap.add_argument("--capture", action="store_true", default=None,
                    help="replay the whole per-rank step as one CUDA graph (default for the small, launch-bound "
                         "workloads vae / mlp; the wide MLP runs eager: its step is 99 %% GEMM time)")
    ap.add_argument("--no-capture", dest="capture", action="store_false")
    ap.add_argument("--cpu-cols", type=int, default=1024, help="columns per CPU-oracle sample (wide MLP)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-e2e-grads", action="store_true")
    ap.add_argument("--quick", action="store_true")
    args = defaults_for(ap.parse_args())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
