#!/usr/bin/env python
"""bench.py — gradient evaluations / second of the batch-sharded wide MLP (BASELINE.json configs[4],
the configuration the metric is quoted on at 1/2/4/8 B200), plus the per-kernel roofline figures.

One "step" = one gradient evaluation of examples/mlp.jl's ``mlp_log_joint`` on the global batch:
forward pass + full reverse sweep producing the objective and every parameter sensitivity, and (N>1)
the NCCL allreduce(sum) of those sensitivities.  Global batch is FIXED and split over the ranks
(strong scaling), so ``value`` is whole-job evaluations/second.

  python bench.py [--gpus N] [--steps K] [--warmup W]          (N>1: launched under torchrun)
  python bench.py --impl reference ...                         CPU oracle on the host cores

``value``  : inputs resident in HBM, CUDA events on the sweep stream, max over ranks.
``e2e``    : the same step through the public API with HOST (pinned) batch buffers: H2D of this rank's
             X/Y shard, sweep, allreduce, D2H of the objective — all inside the timed region.
``roofline``: the dominant kernel (dense GEMM sensitivities) timed per launch with CUDA events
             inside the timed region; ``extra`` carries the HBM-bound kernels (C3 broadcast pullback)
             and the small configs (C1 MLP, C4 VAE, CUDA-graph replay).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDE_DIMS = (784, 8192, 8192, 8192, 8192, 10)
LAM = 1e-3


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"],
                "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


def mlp_flops(dims, B):
    """FLOPs of one gradient evaluation: forward GEMMs + Ā and B̄ GEMMs, minus the first layer's
    never-forced Wᵀ·Ȳ (X is not a Node: core.jl:151)."""
    params = sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))
    return 2.0 * B * (3 * params - dims[0] * dims[1])


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
_ORACLE_INPUTS = {}


def oracle_eval_seconds(dims, B_full, dtype, sample_cols, reps):
    """Time the NumPy oracle (op-for-op restatement of the reference sweep) on two column samples and
    extrapolate affinely to B_full: t(B) = a + c*B (the prior terms do not scale with B)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nabla_oracle as orc
    from nabla_jl_b200 import workloads as wl
    key = (tuple(dims), max(sample_cols), np.dtype(dtype).name)
    if key not in _ORACLE_INPUTS:   # generating 208M normals takes seconds: do it once, untimed
        _ORACLE_INPUTS[key] = wl.mlp_inputs(dims, max(sample_cols), dtype, seed=4)
    W, b, X, Y = _ORACLE_INPUTS[key]
    nl = len(dims) - 1
    ts = []
    for Bs in sample_cols:
        obj = wl.mlp_objective(orc, X[:, :Bs], Y[:, :Bs], LAM, nl)
        best = float("inf")
        for _ in range(reps):
            t0 = time.perf_counter()
            orc.nabla(obj)(*W, *b)
            best = min(best, time.perf_counter() - t0)
        ts.append(best)
    (b1, b2), (t1, t2) = sample_cols, ts
    c = max((t2 - t1) / (b2 - b1), 0.0)
    a = max(t1 - c * b1, 0.0)
    return a + c * B_full, {"cols": list(sample_cols), "seconds": ts}


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dtype = np.float32 if args.dtype == "f32" else np.float64
    B = args.batch
    cols = (args.cpu_cols, 3 * args.cpu_cols)
    use_all_host_threads()
    for _ in range(args.warmup):
        oracle_eval_seconds(WIDE_DIMS, B, dtype, cols, 1)
    t0 = time.perf_counter()
    per = []
    for _ in range(args.steps):
        sec, detail = oracle_eval_seconds(WIDE_DIMS, B, dtype, cols, 1)
        per.append(sec)
    sec = float(np.median(per))
    val = 1.0 / sec
    cb = {"value": val, "unit": "evals/s", "cores": host_threads(), "kind": "port",
          "sample": f"oracle timed at B={cols[0]} and B={cols[1]} columns per step, affine fit "
                    f"t(B)=a+c*B extrapolated to B={B} (extrapolated)"}
    print(json.dumps({
        "impl": "reference", "metric": "gradient evals/sec", "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic", "config": workload_config(args, args.gpus), "cpu_baseline": cb,
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0}))


def workload_config(args, n):
    return {"workload": "wide MLP 784-8192-8192-8192-8192-10 (examples/mlp.jl mlp_log_joint), "
                        "batch-sharded gradient + NCCL allreduce (BASELINE configs[4])",
            "dims": list(WIDE_DIMS), "global_batch": args.batch, "per_gpu_batch": args.batch // n,
            "math_mode": args.math, "parallelism": f"dp{n} (batch columns), allreduce(sum) of 207.8M "
                                                   "parameter sensitivities",
            "l2": "working set (~26 GB of activations/sensitivities at B=32768) >> 126 MB L2; no flush needed"}


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
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn(orc)
        best = min(best, time.perf_counter() - t0)
    return {"ms_per_eval": best * 1e3, "evals_per_s": 1.0 / best, "cores": host_threads()}


def extra_measurements(nb, ctx, peaks, quick):
    """HBM-bound kernels (C3) and the small configs (C1, C4), N=1 only.  Short."""
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
                                            "TFLOP/s": mlp_flops((784, 500, 300, 10), 4096) / (ms * 1e-3) / 1e12}
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
    # matrix form quad(A, B) = B'AB with Ybar = I (test/core.jl:206-218): 8 DMMA GEMMs of 4096^3
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


def run_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.gpus
    if world != n:
        if world == 1 and n > 1:
            raise SystemExit(f"--gpus {n} needs torchrun with {n} ranks (WORLD_SIZE={world})")
        n = world
    import nabla_jl_b200 as nb
    from nabla_jl_b200 import workloads as wl
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
    ctx.math_mode = {"3xtf32": nb.MATH_3XTF32, "tf32": nb.MATH_TF32, "fp32": nb.MATH_FP32, "bf16x3": nb.MATH_BF16X3}[args.math] \
        if args.dtype == "f32" else nb.MATH_DEFAULT
    dtype = np.float32 if args.dtype == "f32" else np.float64
    dims, B = WIDE_DIMS, args.batch
    nl = len(dims) - 1

    def exchange(raw):
        obj = [raw]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    comm = nb.Communicator(ctx, rank, n, exchange)
    W, b, X, Y = wl.mlp_inputs(dims, B, dtype, seed=4)
    lo, hi = nb.shard_bounds(B, rank, n)
    Xs, Ys = np.asfortranarray(X[:, lo:hi]), np.asfortranarray(Y[:, lo:hi])
    del X, Y
    params = [ctx.asarray(p) for p in (*W, *b)]
    del W, b
    Xd, Yd = ctx.asarray(Xs), ctx.asarray(Ys)
    Xp, Yp = ctx.pinned_empty(Xs.shape, dtype), ctx.pinned_empty(Ys.shape, dtype)
    Xp.a[...] = Xs
    Yp.a[...] = Ys
    yp = ctx.pinned_empty((), dtype)
    # data term (log-likelihood of this rank's columns) + shared term (log-prior): gradients of the data
    # term are allreduced per parameter as soon as they are final, overlapped with the rest of the sweep
    sg = nb.OverlappedShardedGradient(wl.mlp_data_objective(nb, Xd, Yd, nl), wl.mlp_prior_objective(nb, LAM, nl),
                                      params, comm)

    def step():
        return sg.step()

    def step_e2e():
        Xd.copy_from_pinned(Xp)
        Yd.copy_from_pinned(Yp)
        y, _ = sg.step()
        y.copy_to_pinned(yp)
        ctx.sync()          # the host reads the objective every step

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
    sampler = ClockSampler(local) if rank == 0 else None
    stats0 = ctx.stats()
    launches0 = stats0["kernel_launches"]
    ctx.profile_begin()
    barrier()
    ms = timed_region(ctx, step, args.steps)
    host_ms = timed_region.host_ms
    barrier()
    recs = ctx.profile_end()
    launches = ctx.stats()["kernel_launches"] - launches0
    clocks = sampler.stop() if sampler else None
    ms = max_over_ranks(ms)
    # end-to-end: host buffers in, objective out
    for _ in range(2):
        step_e2e()
    barrier()
    ms_e2e = max_over_ranks(timed_region(ctx, step_e2e, args.steps))
    barrier()

    ms_step = ms / args.steps
    value = 1e3 / ms_step
    # roofline of the dominant kernel: the hidden-layer GEMMs (largest total time among gemm tags)
    by_tag = {}
    for tag, work, unit, t in recs:
        if unit == "FLOP":
            d = by_tag.setdefault(tag, [0.0, 0, work, []])
            d[0] += t
            d[1] += 1
            d[3].append(t)
    roof = None
    gemm_ms_total = sum(v[0] for v in by_tag.values())
    if by_tag:
        # aggregate all big GEMM launches (they are the same kernel): achieved = sum FLOP / sum time
        flop = sum(v[2] * v[1] for v in by_tag.values())
        achieved = flop / (gemm_ms_total * 1e-3) / 1e12
        ratio = {"bf16x3": 1.0, "3xtf32": 0.5, "tf32": 0.5, "fp32": 0.5}[args.math] if args.dtype == "f32" else 0.5
        peak = peaks["bf16_sustained"] * ratio
        f64_peak_note = None
        if args.dtype == "f64":
            # Large Float64 products run as S(S+1)/2 = 28 exact int8 slice products on the tensor cores
            # (csrc/ozaki.cu, S = 7 slices): the bound is the kind::i8 pipe at twice the bf16 rate, i.e.
            # 2 x bf16 / 28 in algorithmic Float64 FLOPs (MEASURED_PEAKS.json has no int8 or FP64 figure; DMMA, used
            # for the small products, is ~36 TFLOP/s)
            slices = min(7, max(2, int(os.environ.get("NB_OZAKI_SLICES", "7"))))
            passes = slices * (slices + 1) // 2
            peak, ratio = 2.0 * peaks["bf16_sustained"] / passes, None
            f64_peak_note = (f"{peaks['source']} bf16 sustained {peaks['bf16_sustained']} TFLOP/s x 2 (kind::i8) / {passes} "
                             f"int8 passes per Float64 product (Ozaki split, {slices} slices); small products on DMMA")
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")
        if args.math == "bf16x3" and args.dtype == "f32" and n == 1 and B == 32768 and os.path.exists(tpath):
            # one `ncu` capture of the same kernel at the dominant shape (never taken during a timed run)
            t = json.load(open(tpath))["per_gemm_call"]["NN_8192x32768x8192"]
            traffic = t["dram_read_bytes"] + t["dram_write_bytes"]
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_note": "DRAM bytes per GEMM launch (8192x32768x8192) from profiles/r1_gemm_traffic.json; "
                                "algorithmic bytes 2.42e9" if traffic else None,
                "kernel": "nb_gemm (dense GEMM sensitivities: forward A*B, Abar=Ybar*B', Bbar=A'*Ybar)",
                "peak_note": f64_peak_note or f"{peaks['source']} bf16 sustained {peaks['bf16_sustained']} TFLOP/s x {ratio} "
                             "(kind::f16 for bf16x3, kind::tf32 at half that rate otherwise); achieved counts "
                             "ALGORITHMIC flops 2mnk: the split modes needed for rtol 1e-4 (bf16x3, 3xtf32) "
                             "issue three tensor passes per product, so they top out at 1/3 of this peak "
                             "(tensor_work_frac = 3 x frac)",
                "tensor_work_frac": (3.0 if (args.dtype == "f32" and args.math in ("bf16x3", "3xtf32")) else 1.0) * achieved / peak,
                "launches": sum(v[1] for v in by_tag.values()),
                "share_of_step": gemm_ms_total / ms,
                "avg_launch_ms": gemm_ms_total / max(1, sum(v[1] for v in by_tag.values())),
                "per_shape": {k: {"launches": v[1], "avg_ms": v[0] / v[1], "min_ms": min(v[3]), "max_ms": max(v[3]),
                                  "TFLOP/s": v[2] / (v[0] / v[1] * 1e-3) / 1e12} for k, v in by_tag.items()}}
    line = {
        "metric": "gradient evals/sec", "value": value, "unit": "evals/s", "n_gpus": n,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args, n),
        "TFLOP/s_algorithmic": mlp_flops(dims, B) / (ms_step * 1e-3) / 1e12,
        "clocks": clocks, "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_ms / args.steps,
        "pool": {"cuda_mallocs_in_timed_regions": ctx.stats()["cuda_mallocs"] - stats0["cuda_mallocs"],
                 "high_water_GB": ctx.stats()["pool_high_water"] / 1e9},
        "e2e": {"value": 1e3 / (ms_e2e / args.steps), "unit": "evals/s",
                "h2d_bytes_per_step": int(Xp.nbytes + Yp.nbytes), "d2h_bytes_per_step": int(yp.nbytes),
                "ms_per_step": ms_e2e / args.steps},
        "roofline": roof,
    }
    if rank == 0 and n == 1 and not args.no_cpu:
        use_all_host_threads()
        sec, detail = oracle_eval_seconds(dims, B, dtype, (args.cpu_cols, 3 * args.cpu_cols), 2)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "evals/s", "cores": host_threads(), "kind": "port",
                                "sample": f"NumPy oracle (op-for-op port of the reference sweep) at B={args.cpu_cols} "
                                          f"and B={3 * args.cpu_cols} columns, affine fit extrapolated to B={B} "
                                          "(extrapolated); OpenBLAS threads = cores, elementwise single-threaded",
                                "detail": detail}
    if rank == 0 and n == 1 and not args.no_extra:
        del sg, params, Xd, Yd
        try:
            line["extra"] = extra_measurements(nb, ctx, peaks, args.quick)
        except Exception as e:  # the headline number stands on its own
            line["extra"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    barrier()
    comm.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32768, help="GLOBAL batch (columns)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--math", default="bf16x3", choices=["bf16x3", "3xtf32", "tf32", "fp32"])
    ap.add_argument("--cpu-cols", type=int, default=128)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
