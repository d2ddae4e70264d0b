"""Parity of the CUDA path (through the C ABI of libnabla_cuda.so) against the CPU oracle.

The cases mirror the reference's own tests for the hot path:
  test/sensitivities/functional/functional.jl   broadcast / map, every scalar function, Ref/scalar mixes
  test/sensitivities/functional/reducedim.jl    sum / mean / mapreduce with dims
  test/sensitivities/functional/reduce.jl       mapfoldl / mapfoldr
  test/sensitivities/linalg/blas.jl             gemm (4 transposes, +-alpha), gemv
  test/sensitivities/linalg/generic.jl          *, /, \\, dot
  test/core.jl                                  API known answers, in-place accumulate
Both the fresh-slot and the pre-filled-slot (accumulate) variants of src/finite_differencing.jl:44-85
are exercised for every family.

Tolerances (BASELINE.json north_star): rtol 1e-10 for Float64, 1e-4 for Float32.  Float32 device
results are compared with the oracle evaluated in Float64 on the same (Float32-representable)
inputs, so the 1e-4 bounds the device's own error rather than the difference of two roundings."""
import math
import os
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = {np.float64: 1e-10, np.float32: 1e-4}


def _np(x):
    import nabla_jl_b200 as nb
    x = nb.unthunk(nb.unbox(x))
    return x.numpy() if hasattr(x, "numpy") else np.asarray(x)


def close(got, ref, dtype=np.float64, scale=1.0):
    """Norm-wise relative error plus an element-wise check scaled by the largest reference entry."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"shape {got.shape} vs {ref.shape}"
    if ref.size == 0:
        return
    rtol = RTOL[dtype] * scale
    den = max(float(np.linalg.norm(ref.ravel())), 1e-300)
    err = float(np.linalg.norm((got - ref).ravel())) / den
    assert err <= rtol, f"relative error {err:.3e} > {rtol:.1e}"
    np.testing.assert_allclose(got, ref, rtol=rtol * 10, atol=rtol * 10 * float(np.max(np.abs(ref))))


def run_both(nb, orc, build, xs, ybar=None, prefill_seed=None, dtype=np.float64):
    """Evaluate ``build(ns, *leaves)`` on the device and on the oracle.

    Returns (y_dev, grads_dev, y_ref, grads_ref).  With ``prefill_seed`` the leaf slots of the reverse
    tape are pre-filled with random values first, so every rule goes through its accumulate path
    (compute_Dv_update, src/finite_differencing.jl:57-85)."""
    out = []
    for ns in (nb, orc):
        t = ns.Tape()
        if ns is orc:  # ground truth in Float64 on the same input values
            xs_ns = [np.asarray(x, dtype=np.float64) if np.ndim(x) else float(x) for x in xs]
        else:
            xs_ns = xs
        leaves = [ns.Leaf(t, x) for x in xs_ns]
        y = build(ns, *leaves)
        yv = ns.unbox(y)
        if ybar is None:
            yb = None
        else:
            yb = np.asarray(np.asarray(ybar, dtype=dtype), dtype=dtype if ns is nb else np.float64)
        if prefill_seed is None:
            rt = ns.nabla(y) if yb is None else ns.nabla(y, yb if yb.ndim else float(yb))
        else:
            if ns is nb:
                seed = nb.oneslike(y) if yb is None else nb.as_device(yb, yv.ctx, yv.dtype)
            else:
                seed = (1.0 if yb is None else (yb if yb.ndim else float(yb)))
            rt = ns.reverse_tape(y, seed)
            rng = np.random.default_rng(prefill_seed)
            for l, x in zip(leaves, xs):
                pre = rng.standard_normal(np.shape(x)).astype(dtype)
                if ns is orc:
                    pre = pre.astype(np.float64)
                if ns is nb:
                    rt[l.pos] = nb.get_context().asarray(pre) if np.ndim(x) else nb.as_device(dtype(pre))
                else:
                    rt[l.pos] = pre.copy() if np.ndim(x) else float(pre)
            ns.propagate(y.tape, rt)
        grads = [np.asarray(_np(rt[l])) if rt.isassigned(l) else None for l in leaves]
        out += [np.asarray(_np(yv)), grads]
    return out


def check_case(nb, orc, build, xs, ybar=None, dtype=np.float64, scale=1.0, accumulate=True):
    xs = [np.asarray(x, dtype=dtype) if np.ndim(x) else dtype(x) for x in xs]
    seeds = [None, 99] if accumulate else [None]
    for seed in seeds:
        yd, gd, yr, gr = run_both(nb, orc, build, xs, ybar, seed, dtype)
        close(yd, yr, dtype, scale)
        for a, b in zip(gd, gr):
            assert (a is None) == (b is None)
            if a is not None:
                close(a, b, dtype, scale)


# ---------------------------------------------------------------------------------------------
# context / allocator (K7)
# ---------------------------------------------------------------------------------------------
def test_context_pool_refcount_and_reuse(nb, ctx):
    import ctypes as C
    L = ctx.L
    p = C.c_void_p()
    nb._lib.check(L.nb_alloc(ctx.h, 1 << 20, C.byref(p)), ctx.h)
    n = C.c_int32()
    L.nb_refcount(ctx.h, p, C.byref(n))
    assert n.value == 1
    L.nb_retain(ctx.h, p)
    L.nb_refcount(ctx.h, p, C.byref(n))
    assert n.value == 2
    L.nb_release(ctx.h, p)
    L.nb_release(ctx.h, p)
    L.nb_refcount(ctx.h, p, C.byref(n))
    assert n.value == 0
    q = C.c_void_p()
    before = ctx.stats()["cuda_mallocs"]
    nb._lib.check(L.nb_alloc(ctx.h, 1 << 20, C.byref(q)), ctx.h)
    assert q.value == p.value, "a released block of the same size class is handed out again"
    assert ctx.stats()["cuda_mallocs"] == before
    L.nb_release(ctx.h, q)
    assert L.nb_release(ctx.h, q) != 0  # double free is an error, not a crash


def test_h2d_d2h_roundtrip_and_fill(nb, ctx):
    rng = np.random.default_rng(0)
    for dt in (np.float32, np.float64):
        for shape in ((), (7,), (5, 3), (4, 3, 2), (2, 3, 4, 5), (0,), (3, 0)):
            a = rng.standard_normal(shape).astype(dt)
            d = ctx.asarray(a)
            assert d.shape == a.shape and d.dtype == a.dtype
            np.testing.assert_array_equal(d.numpy(), a)
        np.testing.assert_array_equal(ctx.full((3, 5), 2.5, dt).numpy(), np.full((3, 5), 2.5, dt))
    with pytest.raises(TypeError):
        nb.DevArray(ctx, (3,), np.int32)


# ---------------------------------------------------------------------------------------------
# family (1): broadcast forward + fused pullback/unbroadcast
# ---------------------------------------------------------------------------------------------
def _domain(orc, name):
    """domain1 (src/finite_differencing.jl:175-179): widest interval of the reference's test points
    on which f is finite."""
    pts = [-math.pi + .1, -.5 * math.pi + .1, -.9, -.1, .1, .9, .5 * math.pi - .1, math.pi - .1]
    f = getattr(orc, "abs_" if name == "abs" else name)
    sets, cur = [], []
    with np.errstate(all="ignore"):
        for p in pts:
            if np.isfinite(f(np.float64(p))):
                cur.append(p)
            else:
                if cur:
                    sets.append(cur)
                cur = []
    if cur:
        sets.append(cur)
    best = max(sets, key=lambda s: max(s) - min(s))
    return min(best), max(best)


def _fn(ns, name):
    return getattr(ns, {"abs": "abs_", "max": "max_", "min": "min_", "pow": "pow_"}.get(name, name))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_unary_catalogue(nb, orc, dtype):
    # test/sensitivities/functional/functional.jl:15-26 — every unary scalar function, forward value
    # and broadcast gradient, fresh and accumulate.
    for name in orc.UNARY_NAMES:
        lo, hi = _domain(orc, name)
        pad = 0.02 * (hi - lo)
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        x = rng.uniform(lo + pad, hi - pad, 203)
        if name == "abs":
            x = x[np.abs(x) > 1e-3]
        ybar = rng.standard_normal(x.shape)
        try:
            check_case(nb, orc, lambda ns, a: ns.broadcast(_fn(ns, name), a), [x], ybar, dtype)
        except AssertionError as e:
            raise AssertionError(f"{name} ({np.dtype(dtype).name}): {e}") from e


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_binary_catalogue_and_scalar_mixes(nb, orc, dtype):
    # test/sensitivities/functional/functional.jl:126-184 — array∘array, array∘scalar, scalar∘array
    rng = np.random.default_rng(7)
    n = 131
    x, y = rng.uniform(0.3, 2.0, n), rng.uniform(0.3, 2.0, n)
    ybar = rng.standard_normal(n)
    for name in orc.BINARY_NAMES:
        try:
            check_case(nb, orc, lambda ns, a, b: ns.broadcast(_fn(ns, name), a, b), [x, y], ybar, dtype)
            check_case(nb, orc, lambda ns, a, b: ns.broadcast(_fn(ns, name), a, b), [0.7, y], ybar, dtype)
            check_case(nb, orc, lambda ns, a, b: ns.broadcast(_fn(ns, name), a, b), [x, 1.3], ybar, dtype)
            # constant (non-Node) operands get no sensitivity (core.jl:151)
            check_case(nb, orc, lambda ns, a: ns.broadcast(_fn(ns, name), a, 1.7), [x], ybar, dtype)
            check_case(nb, orc, lambda ns, b: ns.broadcast(_fn(ns, name), 0.6, b), [y], ybar, dtype)
        except AssertionError as e:
            raise AssertionError(f"{name} ({np.dtype(dtype).name}): {e}") from e


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_scalar_catalogue_against_mpmath_golden(nb, dtype):
    # tests/golden/scalar_catalogue_mpmath.json: 50-digit mpmath values of every catalogue function and its
    # derivative at the reference's test points (src/finite_differencing.jl:175-179) — no oracle, no SciPy involved.
    import json
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scalar_catalogue_mpmath.json")))

    def one(fn, consts, t, tag):
        x = np.asarray(t["x"], dtype=dtype)
        if x.size == 0:
            return
        # (Float32: the table is evaluated at the Float64 points; compare at the Float32 bar)
        leaf = nb.Leaf(nb.Tape(), x)
        y = nb.broadcast(fn, *consts, leaf)
        try:
            close(_np(nb.unbox(y)), t["f"], dtype)
            close(_np(nb.nabla(y, nb.oneslike(y))[leaf]), t["df"], dtype)
        except AssertionError as e:
            raise AssertionError(f"{tag} ({np.dtype(dtype).name}): {e}") from e

    for name, t in G["unary"].items():
        one(_fn(nb, name), (), t, name)
    for name, rows in G["second_arg_only"].items():
        for row in rows:
            one(_fn(nb, name), (row["n"],), row, f"{name}({row['n']}, x)")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_second_arg_only_catalogue(nb, orc, dtype):
    # test/sensitivities/scalar.jl:38-49 — besselj/y/i/k(n, x), polygamma(n, x): integer first argument 0..5 (a
    # constant: not differentiable), sensitivity in x; generic kernels (ragged size) and the lean ones (64 x 36).
    for name in orc.SECOND_ARG_ONLY_NAMES:
        for n in range(6):
            rng = np.random.default_rng(zlib.crc32(name.encode()) + n)
            for shape in ((203,), (64, 36)):
                x = rng.uniform(0.15, 3.0, shape)
                if name in ("besselj", "besseli") and len(shape) == 1:
                    x = x * rng.choice([-1.0, 1.0], shape)
                ybar = rng.standard_normal(shape)
                try:
                    check_case(nb, orc, lambda ns, a: ns.broadcast(_fn(ns, name), n, a), [x], ybar, dtype)
                except AssertionError as e:
                    raise AssertionError(f"{name}({n}, x) {shape} ({np.dtype(dtype).name}): {e}") from e
    # a Node in the order position gets DiffRules' NaN; x keeps its sensitivity
    ctx = nb.get_context()
    nu, x = np.full(50, 2.0), np.linspace(0.5, 2.5, 50)
    got = nb.nabla(lambda a, b: nb.sum(nb.broadcast(nb.besselk, a, b)))(nu.astype(dtype), x.astype(dtype))
    assert np.all(np.isnan(_np(got[0])))
    ref = orc.nabla(lambda a, b: orc.sum(orc.broadcast(orc.besselk, a, b)))(nu, x)
    close(_np(got[1]), ref[1], dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_catalogue_on_lean_kernels(nb, orc, dtype):
    # The same scalar catalogue at vector-friendly sizes (multiples of the 128-bit vector, 16-byte aligned
    # buffers), where every single primitive runs on the compile-time-specialised lean kernels
    # (bcast_lean.cuh, bcast_lean_cold.inc): flat for the unary functions, flat / tall / warp-per-column
    # with the fused unbroadcast for the binary ones.  Also sum(f, A) (reducedim.jl:23-39) on the same kernels.
    for name in orc.UNARY_NAMES:
        lo, hi = _domain(orc, name)
        pad = 0.02 * (hi - lo)
        rng = np.random.default_rng(zlib.crc32(name.encode()) + 1)
        x = rng.uniform(lo + pad, hi - pad, (64, 36))
        if name == "abs":
            x = np.where(np.abs(x) > 1e-3, x, 0.5)
        ybar = rng.standard_normal(x.shape)
        try:
            check_case(nb, orc, lambda ns, a: ns.broadcast(_fn(ns, name), a), [x], ybar, dtype)
            # the sum of an odd function over a symmetric domain nearly cancels: widen the tolerance by its
            # condition number (sum |f| / |sum f|) so the check bounds the kernel's error, not the cancellation
            fx = np.asarray(_fn(orc, name)(x.astype(dtype).astype(np.float64)))
            cond = float(np.abs(fx).sum() / max(abs(fx.sum()), 1e-300))
            check_case(nb, orc, lambda ns, a: ns.sum(_fn(ns, name), a), [x], None, dtype, accumulate=False,
                       scale=max(1.0, cond / 30.0))
        except AssertionError as e:
            raise AssertionError(f"{name} ({np.dtype(dtype).name}): {e}") from e
    rng = np.random.default_rng(8)
    A = rng.uniform(0.3, 2.0, (256, 64))
    ybar = rng.standard_normal(A.shape)
    for name in orc.BINARY_NAMES:
        for other in (rng.uniform(0.3, 2.0, (256, 64)), rng.uniform(0.3, 2.0, (256,)),
                      rng.uniform(0.3, 2.0, (1, 64)), 1.3):
            try:
                check_case(nb, orc, lambda ns, a, b: ns.broadcast(_fn(ns, name), a, b), [A, other], ybar, dtype)
                check_case(nb, orc, lambda ns, a, b: ns.broadcast(_fn(ns, name), b, a), [A, other], ybar, dtype)
            except AssertionError as e:
                raise AssertionError(f"{name} {np.shape(other)} ({np.dtype(dtype).name}): {e}") from e


SHAPES = [
    ((5, 7), (5, 7)),        # same shape
    ((5, 7), (5,)),          # bias: reduce over columns (functional.jl:61-63)
    ((5, 7), (5, 1)),
    ((5, 7), (1, 7)),        # row vector: reduce over rows
    ((5, 7), ()),            # scalar operand: reduce over everything (functional.jl:84-87)
    ((5,), (1, 7)),          # both operands broadcast (outer)
    ((300, 70), (300,)),     # k_tall / KEEP_ROWS
    ((300, 70), (1, 70)),    # k_wcol / KEEP_COLS
    ((10, 4096), (10,)),     # k_short with many columns (MLP output layer)
    ((10, 4096), (1, 4096)),
    ((33, 129), (33,)),      # sizes that are not multiples of the vector width
    ((1031,), (1031,)),      # k_flat with a ragged tail
    ((4, 3, 5), (4, 1, 5)),  # 3-D with a reduced middle dimension (k_nd)
    ((4, 3, 5), (1, 3)),
    ((2, 3, 4, 5), (2, 1, 4, 1)),
    ((6, 5, 4), (6, 5)),     # trailing dims missing
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shapes", SHAPES, ids=[f"{a}x{b}" for a, b in SHAPES])
def test_broadcast_unbroadcast_shapes(nb, orc, dtype, shapes):
    # Appendix A.6: unbroadcast = sum over every dim where the operand has extent 1 or is missing
    rng = np.random.default_rng(11)
    sa, sb = shapes
    a = rng.standard_normal(sa) if sa else float(rng.standard_normal())
    b = rng.uniform(0.5, 1.5, sb) if sb else float(rng.uniform(0.5, 1.5))
    full = orc.jl_broadcast_shape(sa, sb)
    ybar = rng.standard_normal(full)
    for name in ("add", "mul", "div"):
        check_case(nb, orc, lambda ns, u, v: ns.broadcast(_fn(ns, name), u, v), [a, b], ybar, dtype)
    # composite scalar function (device bytecode; the reference uses ForwardDiff duals)
    f = lambda ns: (lambda p, q: ns.tanh(p) * q + ns.exp(-p) / (1.0 + q * q))
    check_case(nb, orc, lambda ns, u, v: ns.broadcast(f(ns), u, v), [a, b], ybar, dtype)


def test_ternary_broadcast(nb, orc):
    # test/sensitivities/functional/functional.jl:68-78
    rng = np.random.default_rng(0)
    f = lambda x, y, z: x * y + y * z + x * z
    x, y, z = (rng.standard_normal(5) for _ in range(3))
    t = nb.Tape()
    x_, y_, z_ = nb.Leaf(t, x), nb.Leaf(t, y), nb.Leaf(t, z)
    s = nb.broadcast(f, x_, y_, z_)
    ds = nb.nabla(s, np.ones(5))
    close(_np(s), f(x, y, z))
    close(_np(ds[x_]), y + z)
    close(_np(ds[y_]), x + z)
    close(_np(ds[z_]), y + x)
    check_case(nb, orc, lambda ns, a, b, c: ns.broadcast(f, a, b, c),
               [rng.standard_normal((6, 9)), rng.standard_normal(6), rng.standard_normal((1, 9))],
               rng.standard_normal((6, 9)))


def test_same_node_twice_and_ref(nb, orc):
    # σz .* σz (examples/vae.jl:85): store-then-accumulate inside one Branch (Appendix A.3);
    # Ref operands broadcast as scalars (functional.jl:181)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((6, 4))
    check_case(nb, orc, lambda ns, a: ns.broadcast(ns.mul, a, a), [x], rng.standard_normal((6, 4)))
    check_case(nb, orc, lambda ns, a: ns.broadcast(ns.add, a, ns.Ref(2.5)), [x], rng.standard_normal((6, 4)))
    check_case(nb, orc, lambda ns, a, b: ns.sum(ns.broadcast(ns.mul, ns.broadcast(ns.add, a, b), a)),
               [x, rng.standard_normal(6)])


def test_map_and_literal_pow_known_answers(nb, orc):
    # test/sensitivities/functional/functional.jl:208-227 (#111, #117, literal_pow)
    x = np.array([1.0, 2.0, 3.0])
    f = lambda ns: (lambda v: ns.sum(ns.broadcast(ns.mul, np.array([1.0, 2, 3]),
                                                  ns.broadcast(ns.add, v, np.array([3.0, 2, 1])))))
    (g,) = nb.nabla(f(nb))(x)
    assert g.dtype == np.float64 and g.shape == (3,)
    np.testing.assert_array_equal(g.numpy(), [1.0, 2.0, 3.0])
    y, _ = nb.nabla(f(nb), get_output=True)(x)
    assert float(_np(y)) == float(np.sum(np.array([1.0, 2, 3]) * (x + np.array([3.0, 2, 1]))))
    sq = lambda v: nb.sum(nb.broadcast(lambda a: a ** 2, v))
    np.testing.assert_array_equal(nb.nabla(sq)(x)[0].numpy(), [2.0, 4.0, 6.0])
    assert float(_np(nb.nabla(sq, get_output=True)(x)[0])) == 14.0
    rng = np.random.default_rng(5)
    a, b = rng.standard_normal((4, 6)), rng.standard_normal((4, 6))
    check_case(nb, orc, lambda ns, u, v: ns.map_(lambda p, q: p * q + ns.sin(p), u, v), [a, b],
               rng.standard_normal((4, 6)))
    with pytest.raises(nb.DimensionMismatch):
        nb.map_(nb.add, nb.Leaf(nb.Tape(), a), np.ones(4))
    # ten operands, the most the reference generates `map` for (functional.jl:10-18): every one a Node, so the pullback
    # wants ten sensitivities (more than one launch of the general kernels produces)
    ops10 = [rng.standard_normal((4, 6)) for _ in range(10)]
    f10 = lambda ns, *v: ns.map_(lambda a0, a1, a2, a3, a4, a5, a6, a7, a8, a9:
                                 a0 * a1 + a2 * a3 - a4 * a5 + ns.sin(a6) * a7 + a8 / (2.0 + a9 * a9), *v)
    check_case(nb, orc, f10, ops10, rng.standard_normal((4, 6)))
    f10b = lambda ns, *v: ns.broadcast(lambda a0, a1, a2, a3, a4, a5, a6, a7, a8, a9:
                                       (a0 + a1 + a2) * (a3 + a4) + a5 * a6 + a7 + a8 * a9, *v)
    mixed = [rng.standard_normal(sh) for sh in ((4, 6), (4,), (1, 6), (4, 6), (4, 1), (4, 6), (1, 6), (4,), (4, 6), (4, 6))]
    check_case(nb, orc, f10b, mixed, rng.standard_normal((4, 6)))


def test_empty_and_error_cases(nb, ctx):
    t = nb.Tape()
    e = nb.Leaf(t, np.zeros((0, 3)))
    y = nb.broadcast(nb.tanh, e)
    assert _np(y).shape == (0, 3)
    s = nb.sum(nb.Leaf(t, np.zeros((0,))))
    # Julia: sum of an empty Float64 array is 0.0
    assert float(_np(s)) == 0.0
    with pytest.raises(nb.DimensionMismatch):
        nb.broadcast(nb.add, nb.Leaf(t, np.ones((5, 3))), np.ones((4, 3)))
    with pytest.raises(nb.UnsupportedOp):
        nb.broadcast(lambda v: 1.0 if v > 0 else 0.0, nb.Leaf(t, np.ones(3)))
    with pytest.raises(TypeError):     # MethodError: array output without ȳ (core.jl:200-201)
        nb.nabla(lambda a, b: nb.broadcast(nb.add, a, b))(np.ones(5), np.ones(5))


# ---------------------------------------------------------------------------------------------
# family (2): reductions
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_reductions(nb, orc, dtype):
    # test/sensitivities/functional/reducedim.jl:27-67, reduce.jl, linalg/generic.jl:81-87
    rng = np.random.default_rng(5)
    r = lambda *s: rng.standard_normal(s)
    sc = 1.0
    for shape in ((10,), (10, 10), (37, 65), (300, 40), (10, 700), (5, 4, 3)):
        x = r(*shape)
        nd = len(shape)
        check_case(nb, orc, lambda ns, a: ns.sum(a), [x], None, dtype, sc)
        check_case(nb, orc, lambda ns, a: ns.mean(a), [x], None, dtype, sc)
        check_case(nb, orc, lambda ns, a: ns.sum(ns.abs2, a), [x], None, dtype, sc)
        check_case(nb, orc, lambda ns, a: ns.mean(ns.abs2, a), [x], None, dtype, sc)
        check_case(nb, orc, lambda ns, a: ns.mapfoldl(ns.exp, "+", a), [x], None, dtype, sc)
        check_case(nb, orc, lambda ns, a: ns.mapfoldr(ns.sin, "+", a), [x], None, dtype, sc)
        for d in range(1, nd + 1):
            rs = tuple(1 if i + 1 == d else s for i, s in enumerate(shape))
            yb = r(*rs)
            check_case(nb, orc, lambda ns, a: ns.sum(a, dims=d), [x], yb, dtype, sc)
            check_case(nb, orc, lambda ns, a: ns.mean(a, dims=d), [x], yb, dtype, sc)
            check_case(nb, orc, lambda ns, a: ns.sum(ns.abs2, a, dims=d), [x], yb, dtype, sc)
            check_case(nb, orc, lambda ns, a: ns.mapreduce(ns.exp, "+", a, dims=d), [x], yb, dtype, sc)
        if nd == 3:
            yb = r(shape[0], 1, 1)
            check_case(nb, orc, lambda ns, a: ns.sum(a, dims=(2, 3)), [x], yb, dtype, sc)
            check_case(nb, orc, lambda ns, a: ns.mean(a, dims=(2, 3)), [x], yb, dtype, sc)
    check_case(nb, orc, lambda ns, a, b: ns.dot(a, b), [r(257), r(257)], None, dtype, sc)
    # logsumexp-style composition through scalar Nodes (reducedim.jl:66)
    check_case(nb, orc, lambda ns, a: ns.log(ns.sum(ns.exp, a)), [r(10)], None, dtype, sc)


def test_reducedim_known_answers(nb):
    # test/sensitivities/functional/reducedim.jl:4-25, 33-61
    x = nb.Leaf(nb.Tape(), np.array([1.0, 2, 3, 4, 5]))
    s = 5.0 * nb.mapreduce(nb.abs2, "+", x, dims=1)
    np.testing.assert_array_equal(_np(nb.nabla(s, nb.oneslike(s))[x]), 5.0 * np.array([2.0, 4, 6, 8, 10]))
    x5 = nb.Leaf(nb.Tape(), np.ones((5, 4, 3)))
    for d, shp, val in ((1, (1, 4, 3), 5.0), (2, (5, 1, 3), 4.0), (3, (5, 4, 1), 3.0)):
        np.testing.assert_array_equal(_np(nb.sum(x5, dims=d)), np.full(shp, val))
        np.testing.assert_array_equal(_np(nb.mean(x5, dims=d)), np.ones(shp))
    assert float(_np(nb.sum(x5))) == 60.0 and float(_np(nb.mean(x5))) == 1.0
    assert nb.sum(x5).f.name == "sum" and nb.mean(x5, dims=1).f.name == "mean"
    # Issue #123
    x6 = np.arange(1.0, 11.0)
    tens = np.full(10, 10.0)
    np.testing.assert_array_equal(_np(nb.nabla(lambda x: nb.sum(nb.sum(x, dims=2)))(x6)[0]), np.ones(10))
    g = nb.nabla(lambda x, y: nb.sum(nb.broadcast(nb.add, nb.sum(x, dims=2), nb.adjoint(nb.sum(y, dims=2)))))(x6, x6)
    np.testing.assert_array_equal(_np(g[0]), tens)
    np.testing.assert_array_equal(_np(g[1]), tens)
    g = nb.nabla(lambda x, y: nb.sum(nb.broadcast(nb.add, x, nb.adjoint(y))))(x6, x6)
    np.testing.assert_array_equal(_np(g[0]), tens)
    np.testing.assert_array_equal(_np(g[1]), tens)


# ---------------------------------------------------------------------------------------------
# family (3): dense linalg
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("tA", "NT")
@pytest.mark.parametrize("tB", "NT")
def test_gemm_all_transposes(nb, orc, dtype, tA, tB):
    # test/sensitivities/linalg/blas.jl:36-47 (N = 100) plus ragged and tile-sized shapes
    rng = np.random.default_rng(123456)
    sc = 1.0
    for (m, k, n) in ((100, 100, 100), (37, 53, 29), (128, 256, 64), (10, 300, 260), (513, 40, 7), (60, 3001, 150)):
        A = rng.standard_normal((k, m) if tA == "T" else (m, k))
        B = rng.standard_normal((n, k) if tB == "T" else (k, n))
        yb = rng.standard_normal((m, n))
        al = float(rng.standard_normal())
        check_case(nb, orc, lambda ns, a, b: ns.gemm(tA, tB, a, b), [A, B], yb, dtype, sc)
        check_case(nb, orc, lambda ns, s, a, b: ns.gemm(tA, tB, s, a, b), [al, A, B], yb, dtype, sc)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_matmul_gemv_dot_and_scaling(nb, orc, dtype):
    # test/sensitivities/linalg/blas.jl:49-61, linalg/generic.jl:47-61
    rng = np.random.default_rng(3)
    r = lambda *s: rng.standard_normal(s)
    sc = 1.0
    check_case(nb, orc, lambda ns, a, b: ns.matmul(a, b), [r(5, 5), r(5, 5)], r(5, 5), dtype, sc)
    check_case(nb, orc, lambda ns, a, b: ns.matmul(a, b), [r(70, 45), r(45, 130)], r(70, 130), dtype, sc)
    check_case(nb, orc, lambda ns, a, x: ns.matmul(a, x), [r(70, 45), r(45)], r(70), dtype, sc)
    check_case(nb, orc, lambda ns, x, y: ns.matmul(ns.adjoint(x), y), [r(33), r(33)], None, dtype, sc)
    for tA in "NT":
        A, x = r(40, 40), r(40)
        check_case(nb, orc, lambda ns, a, v: ns.gemv(tA, a, v), [A, x], r(40), dtype, sc)
        check_case(nb, orc, lambda ns, s, a, v: ns.gemv(tA, s, a, v), [0.7, A, x], r(40), dtype, sc)
    A = r(12, 30)
    check_case(nb, orc, lambda ns, a, v: ns.gemv("N", a, v), [A, r(30)], r(12), dtype, sc)
    check_case(nb, orc, lambda ns, a, v: ns.gemv("T", a, v), [A, r(12)], r(30), dtype, sc)
    check_case(nb, orc, lambda ns, a, s: a / s, [r(5, 5), 1.7], r(5, 5), dtype, sc)
    check_case(nb, orc, lambda ns, s, a: ns.ldivide(s, a), [1.7, r(5, 5)], r(5, 5), dtype, sc)
    check_case(nb, orc, lambda ns, a, b: a + b, [r(6, 5), r(6, 5)], r(6, 5), dtype, sc)
    check_case(nb, orc, lambda ns, a: ns.adjoint(a), [r(6, 5)], r(5, 6), dtype, sc)


def test_quadratic_forms_and_docs_identity(nb, orc):
    # docs/src/index.md:25-39: ∇(y'(A x)) = (A'y, A x); examples/symmetric_quadratic.jl:15
    import nabla_jl_b200.workloads as wl
    A, x, y = wl.quad_inputs(300)
    Ad = nb.get_context().asarray(A)
    (g,) = nb.nabla(wl.quad_symmetric(nb, Ad))(x)
    close(_np(g), (A + A.T) @ x)
    gx, gy = nb.nabla(wl.quad_asymmetric(nb, Ad))(x, y)
    close(_np(gx), A @ y)
    close(_np(gy), A.T @ x)
    (gr,) = orc.nabla(wl.quad_symmetric(orc, A))(x)
    close(_np(g), gr)


def test_gemm_error_behaviour(nb, ctx):
    with pytest.raises(nb.DimensionMismatch):
        nb.matmul(nb.Leaf(nb.Tape(), np.ones((3, 4))), np.ones((5, 2)))
    with pytest.raises(nb.NablaCudaError):
        import ctypes as C
        a = ctx.zeros((4, 4), np.float64)
        nb._lib.check(ctx.L.nb_gemm(ctx.h, b"X", b"N", 4, 4, 4, 1.0, a.ptr, 4, a.ptr, 4, 0.0, a.ptr, 4, 1, 0), ctx.h)


# ---------------------------------------------------------------------------------------------
# core API (test/core.jl) and the docs golden vector
# ---------------------------------------------------------------------------------------------
def test_docs_autodiff_golden(nb):
    # docs/src/pages/autodiff.md:203-221, 252-258, 285-286 — scalar Nodes are 0-dim device arrays
    x, y = 0.6791074260357777, 0.8284134829000359
    t = nb.Tape()
    x_, y_ = nb.Leaf(t, x), nb.Leaf(t, y)
    z = x_ * y_ + nb.sin(x_)
    vals = [float(_np(n)) for n in t.tape]
    ref = [x, y, 0.5625817480655771, 0.6280987324705773, 1.1906804805361544]
    np.testing.assert_allclose(vals, ref, rtol=1e-15)
    rt = nb.nabla(z)
    got = [float(_np(rt[i])) for i in range(1, 6)]
    np.testing.assert_allclose(got, [1.6065471361170487, 0.6791074260357777, 1.0, 1.0, 1.0], rtol=1e-15)


def test_core_known_answers(nb):
    # test/core.jl:129-173
    g = nb.nabla(lambda a, b, c, d: a * c)(1, np.array([2.0]), 3, 4.0)
    assert float(_np(g[0])) == 3 and (_np(g[1]) == [0.0]).all() and float(_np(g[2])) == 1 and float(_np(g[3])) == 0.0
    g = nb.nabla(lambda a, b: 12)(1, 2)
    assert [float(_np(v)) for v in g] == [0.0, 0.0]
    f = lambda x, y: 2 * x + y
    x, y = 0.3, -1.2
    t = nb.Tape()
    x_, y_ = nb.Leaf(t, x), nb.Leaf(t, y)
    dz = nb.nabla(f(x_, y_))
    assert (float(_np(dz[x_])), float(_np(dz[y_]))) == (2.0, 1.0)
    z, (gx, gy) = nb.nabla(f, get_output=True)(x, y)
    np.testing.assert_allclose(float(_np(z)), f(x, y), rtol=1e-15)
    assert (float(_np(gx)), float(_np(gy))) == (2.0, 1.0)
    out = nb.nabla(lambda a: -a, get_output=True)(2)
    assert isinstance(out[0], nb.Branch) and float(_np(out[1][0])) == -1.0


def test_accumulate_mutates_only_target_slot(nb, ctx):
    # test/core.jl:206-224 — `quad` registered through the extension surface (@explicit_intercepts +
    # ∇(f, Arg{N}, ...)), its rules written with the device GEMM
    from nabla_jl_b200 import kernels as K

    def mm(tA, tB, A, B):
        m = A.shape[1] if tA == "T" else A.shape[0]
        k = A.shape[0] if tA == "T" else A.shape[1]
        n = B.shape[0] if tB == "T" else B.shape[1]
        out = ctx.empty((m, n), A.dtype)
        return K.gemm(tA, tB, 1.0, A, B, 0.0, out, m, n, k)

    @nb.explicit_intercepts("quad")
    def quad(A, B):
        return mm("T", "N", B, mm("N", "N", A, B))

    @quad.grad(1)
    def _(p, Y, Yb, A, B):
        return mm("N", "T", mm("N", "N", B, Yb), B)

    rng = np.random.default_rng(123456)
    n = 5
    A_, B_ = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    A = nb.Leaf(nb.Tape(), A_)
    B = ctx.asarray(B_)
    Q = quad(A, B)
    QQ = quad(Q, B)
    close(_np(QQ), B_.T @ (B_.T @ A_ @ B_) @ B_)
    rt = nb.nabla(QQ, np.eye(n))
    close(_np(rt[A]), B_ @ (B_ @ np.eye(n) @ B_.T) @ B_.T)
    old = [np.array(_np(v), copy=True) for v in rt.tape]
    ptr = rt.raw(1).ptr
    nb.core.propagate_branch(Q, rt)   # triggers a mutating addition into slot 1
    assert rt.raw(1).ptr == ptr, "the slot buffer is mutated in place"
    assert not np.allclose(old[0], _np(rt.tape[0]))
    close(_np(rt.tape[0]), 2 * old[0])
    for a, b in zip(old[1:], rt.tape[1:]):
        np.testing.assert_array_equal(a, _np(b))


def test_seed_is_never_mutated(nb, ctx):
    # a rule that returns ȳ itself (array +) must not let a later accumulate write into the caller's seed
    rng = np.random.default_rng(1)
    x = rng.standard_normal((4, 3))
    seed = ctx.asarray(rng.standard_normal((4, 3)))
    keep = seed.numpy().copy()
    t = nb.Tape()
    a = nb.Leaf(t, x)
    y = (a + a) + a
    rt = nb.nabla(y, seed)
    np.testing.assert_array_equal(seed.numpy(), keep)
    close(_np(rt[a]), 3 * keep)


# ---------------------------------------------------------------------------------------------
# whole workloads (BASELINE.json configs) vs the oracle
# ---------------------------------------------------------------------------------------------
def _mlp_both(nb, orc, dims, B, dtype, seed=0):
    import nabla_jl_b200.workloads as wl
    W, b, X, Y = wl.mlp_inputs(dims, B, dtype, seed=seed)
    nl = len(dims) - 1
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    ref_y, ref = orc.nabla(wl.mlp_objective(orc, f64(X), f64(Y), 1e-3, nl), get_output=True)(
        *[f64(p) for p in (*W, *b)])
    ctx = nb.get_context()
    Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
    got_y, got = nb.nabla(wl.mlp_objective(nb, Xd, Yd, 1e-3, nl), get_output=True)(*W, *b)
    return float(_np(got_y)), [_np(g) for g in got], float(orc.unbox(ref_y)), ref


@pytest.mark.parametrize("dims,B", [((20, 16, 12, 10), 64), ((784, 500, 300, 10), 256)])
def test_mlp_gradient_f64(nb, orc, dims, B):
    # C1: examples/mlp.jl:56-82 (parity unpinned in the reference; oracle is FD-checked on CPU)
    y, g, yr, gr = _mlp_both(nb, orc, dims, B, np.float64)
    np.testing.assert_allclose(y, yr, rtol=1e-10)
    for a, b in zip(g, gr):
        close(a, b)


def test_mlp_gradient_f32(nb, orc):
    y, g, yr, gr = _mlp_both(nb, orc, (784, 500, 300, 10), 256, np.float32)
    np.testing.assert_allclose(y, yr, rtol=1e-4)
    for a, b in zip(g, gr):
        close(a, b, np.float32)


def test_wide_mlp_gradient_f32_default_math(nb, orc):
    """A scaled-down C5 (wide MLP, Float32, default BF16x3 tensor-core GEMMs, K up to 2048) against the
    Float64 oracle: the whole-gradient error must stay inside 1e-4."""
    y, g, yr, gr = _mlp_both(nb, orc, (784, 2048, 2048, 10), 1024, np.float32, seed=4)
    np.testing.assert_allclose(y, yr, rtol=1e-4)
    for a, b in zip(g, gr):
        close(a, b, np.float32)


def test_vae_gradient_f64(nb, orc):
    # C4: examples/vae.jl:67-90
    import nabla_jl_b200.workloads as wl
    for (dx, dz, dh, B) in ((8, 3, 6, 5), (784, 10, 500, 128)):
        params, X, noise = wl.vae_inputs(dx=dx, dz=dz, dh=dh, B=B, seed=3)
        kw = dict(lam=1e-3, scal=60000 / B, dz=dz)
        ref = orc.nabla(lambda *p: wl.vae_log_joint(orc, X, noise, *p, **kw))(*params)
        ctx = nb.get_context()
        Xd, Nd = ctx.asarray(X), ctx.asarray(noise)
        got = nb.nabla(lambda *p: wl.vae_log_joint(nb, Xd, Nd, *p, **kw))(*params)
        for a, b in zip(got, ref):
            close(_np(a), b)


def test_vae_gradient_f32_fused(nb, orc):
    """C4 in Float32 (BF16x3 tensor products with the deferred / fused launches: `tanh.(Wh*X)`, `logistic.(Vf*h)`
    are activation-only epilogues, `exp.(Wsig*h)` feeds two consumers) against the Float64 oracle."""
    import nabla_jl_b200.workloads as wl
    dx, dz, dh, B = 784, 10, 500, 1024
    params, X, noise = wl.vae_inputs(dx=dx, dz=dz, dh=dh, B=B, dtype=np.float32, seed=3)
    kw = dict(lam=1e-3, scal=60000 / B, dz=dz)
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    ref = orc.nabla(lambda *p: wl.vae_log_joint(orc, f64(X), f64(noise), *p, **kw))(*[f64(p) for p in params])
    ctx = nb.get_context()
    Xd, Nd = ctx.asarray(X), ctx.asarray(noise)
    got = nb.nabla(lambda *p: wl.vae_log_joint(nb, Xd, Nd, *p, **kw))(*params)
    for a, b in zip(got, ref):
        close(_np(a), b, np.float32)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("row_vector", [False, True])
def test_c3_broadcast_sum(nb, orc, dtype, row_vector):
    # C3: s = sum(tanh.(A .+ b)) — Appendix A.10
    import nabla_jl_b200.workloads as wl
    for (m, n) in ((9, 6), (1024, 1024)):
        A, b = wl.c3_inputs(m, n, dtype, row_vector=row_vector)
        (y, (gA, gb)) = nb.nabla(wl.bcast_sum(nb), get_output=True)(A, b)
        (yr, (rA, rb)) = orc.nabla(wl.bcast_sum(orc), get_output=True)(A.astype(np.float64), b.astype(np.float64))
        close(_np(y), orc.unbox(yr), dtype)
        close(_np(gA), rA, dtype)
        close(_np(gb), rb, dtype)


def test_c3_full_size_properties(nb, ctx):
    """C3 at 1e8 elements (the size the metric is quoted on): size-independent properties instead of
    a CPU oracle run — Ā = 1 - y² element-wise on a sample, b̄ = row sums of Ā (checked through
    linearity: b̄·1 == sum(Ā)), and s == sum(y)."""
    import nabla_jl_b200.workloads as wl
    m, n = 8192, 12288
    rng = np.random.default_rng(2)
    b = rng.standard_normal(m).astype(np.float32)
    A = ctx.empty((m, n), np.float32)
    # fill A on the device from a small host tile (H2D of 400 MB of pageable memory is slow)
    tile = rng.standard_normal((m, 64)).astype(np.float32)
    td = ctx.asarray(tile)
    from nabla_jl_b200 import kernels as K
    A3 = A.reshape((m, 64, n // 64))
    K.bcast_reduce(None, [td.reshape((m, 64, 1))], A3)   # broadcast fill: A[:, 64k:64k+64] = tile
    y, (gA, gb) = nb.nabla(wl.bcast_sum(nb), get_output=True)(A, b)
    ref_tile = 1.0 - np.tanh(tile.astype(np.float64) + b[:, None].astype(np.float64)) ** 2
    gA_h = gA.numpy()
    close(gA_h[:, :64], ref_tile, np.float32)
    close(gA_h[:, -64:], ref_tile, np.float32)
    close(gb.numpy(), ref_tile.sum(1) * (n // 64), np.float32)
    s_ref = np.tanh(tile.astype(np.float64) + b[:, None]).sum() * (n // 64)
    np.testing.assert_allclose(float(_np(y)), s_ref, rtol=1e-4)


def test_graph_capture_replays_sweep(nb, ctx):
    """CUDA-graph capture of a whole forward + reverse sweep: the replay
    produces the same gradients as the eager sweep and follows new input data in place."""
    import nabla_jl_b200.workloads as wl
    dims, B = (20, 16, 12, 10), 64
    W, b, X, Y = wl.mlp_inputs(dims, B, np.float64, seed=0)
    Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
    obj = wl.mlp_objective(nb, Xd, Yd, 1e-3, 3)
    eager = [g.numpy() for g in nb.nabla(obj)(*W, *b)]
    cap = nb.CapturedGradient(obj, [ctx.asarray(p) for p in (*W, *b)])
    y0, g0 = cap()
    ctx.sync()
    for a, e in zip(g0, eager):
        np.testing.assert_array_equal(a.numpy(), e)
    # new data in the same buffers -> replay sees it
    X2 = np.random.default_rng(9).random(X.shape)
    Xd.copy_from_host(X2)
    _, g1 = cap()
    ctx.sync()
    ref = [g.numpy() for g in nb.nabla(obj)(*W, *b)]
    for a, e in zip(g1, ref):
        np.testing.assert_array_equal(a.numpy(), e)
    assert cap.launches_per_replay > 0


def test_capture_rejects_host_uploads_with_a_clear_error(nb, ctx):
    """A captured function that uploads pageable host data (the index vector of `getindex`) or reads a device value on
    the host cannot be recorded: the library says so instead of surfacing CUDA's stream-capture error, ends the
    capture cleanly, and the context keeps working."""
    from nabla_jl_b200._lib import NablaCudaError
    x = ctx.asarray(np.arange(1.0, 7.0))
    f = lambda v: nb.sum(nb.getindex(v, [2, 3, 3]))
    (g_eager,) = nb.nabla(f)(x)
    expected = g_eager.numpy()
    with pytest.raises(NablaCudaError, match="captur"):
        nb.CapturedGradient(f, [x], ctx)
    g = lambda v: nb.sum(nb.broadcast(nb.mul, v, float(nb.to_host(nb.sum(v)))))   # host read inside the function
    with pytest.raises(NablaCudaError, match="captur"):
        nb.CapturedGradient(g, [x], ctx)
    (g_again,) = nb.nabla(f)(x)
    np.testing.assert_array_equal(g_again.numpy(), expected)
    np.testing.assert_array_equal(expected, [0.0, 1.0, 2.0, 0.0, 0.0, 0.0])


@pytest.mark.parametrize("mode,tol", [("bf16x3", 3e-5), ("3xtf32", 2e-5), ("tf32", 3e-3), ("fp32", 2e-5)])
def test_gemm_tensor_core_modes(nb, ctx, mode, tol):
    """Float32 GEMM on tcgen05 (kind::tf32) for all four transposes, ragged M/N/K tails, alpha/beta.
    3xTF32 (default) must sit well inside rtol 1e-4; single-pass TF32 is reported at its own 1e-3
    class; the CUDA-core FP32 path is the cross-check.  Reference: NumPy Float64."""
    from nabla_jl_b200 import kernels as K
    rng = np.random.default_rng(42)
    old = ctx.math_mode
    ctx.math_mode = {"3xtf32": nb.MATH_3XTF32, "tf32": nb.MATH_TF32, "fp32": nb.MATH_FP32, "bf16x3": nb.MATH_BF16X3}[mode]
    try:
        for (m, n, k) in ((256, 256, 256), (384, 512, 1000), (132, 260, 72), (128, 2048, 4096), (1024, 136, 520), (256, 128, 16384 + 40), (10, 2048, 1030), (250, 77, 333)):
            for tA in "NT":
                for tB in "NT":
                    A = rng.standard_normal((k, m) if tA == "T" else (m, k)).astype(np.float32)
                    B = rng.standard_normal((n, k) if tB == "T" else (k, n)).astype(np.float32)
                    C0 = rng.standard_normal((m, n)).astype(np.float32)
                    opA = A.T if tA == "T" else A
                    opB = B.T if tB == "T" else B
                    Ad, Bd = ctx.asarray(A), ctx.asarray(B)
                    for alpha, beta in ((1.0, 0.0), (-0.7, 1.0), (1.3, 0.5)):
                        Cd = ctx.asarray(C0)
                        K.gemm(tA, tB, alpha, Ad, Bd, beta, Cd, m, n, k)
                        ref = alpha * (opA.astype(np.float64) @ opB.astype(np.float64)) + beta * C0.astype(np.float64)
                        got = Cd.numpy().astype(np.float64)
                        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
                        assert err < tol, f"{mode} {tA}{tB} {m}x{n}x{k} alpha={alpha} beta={beta}: rel err {err:.3e}"
    finally:
        ctx.math_mode = old


def test_overlapped_sharded_gradient_single_rank_equals_the_plain_sweep(nb, ctx):
    """shared objective (swept first, its sensitivities pre-fill the reverse tape) + data objective (accumulates into
    them: the reference's pre-filled-slot path) on one rank against the plain ∇(mlp_log_joint) sweep — the same terms
    summed in the other order, so equal to rounding; also as a CUDA-graph replay of the whole step."""
    import nabla_jl_b200.workloads as wl
    dims, B = (20, 16, 12, 10), 64
    W, b, X, Y = wl.mlp_inputs(dims, B, np.float64, seed=0)
    Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
    params = [ctx.asarray(p) for p in (*W, *b)]
    yr, gr = nb.nabla(wl.mlp_objective(nb, Xd, Yd, 1e-3, 3), get_output=True)(*params)
    comm = nb.Communicator(ctx, 0, 1, None)
    og = nb.OverlappedShardedGradient(wl.mlp_data_objective(nb, Xd, Yd, 3), wl.mlp_prior_objective(nb, 1e-3, 3),
                                      params, comm)
    y, g = og.step()
    np.testing.assert_allclose(float(_np(y)), float(_np(yr)), rtol=1e-14)
    for a, r in zip(g, gr):
        np.testing.assert_allclose(_np(a), _np(r), rtol=1e-13, atol=1e-13 * float(np.max(np.abs(_np(r)))))
    oc = nb.OverlappedShardedGradient(wl.mlp_data_objective(nb, Xd, Yd, 3), wl.mlp_prior_objective(nb, 1e-3, 3),
                                      params, comm, capture=True)
    for _ in range(2):
        y2, g2 = oc.step()
    ctx.sync()
    np.testing.assert_allclose(float(_np(y2)), float(_np(yr)), rtol=1e-14)
    for a, r in zip(g2, gr):
        np.testing.assert_allclose(_np(a), _np(r), rtol=1e-13, atol=1e-13 * float(np.max(np.abs(_np(r)))))
    oc.close()
    # the hook fires once per parameter, in descending order of each Leaf's first consumer
    t = nb.Tape()
    leaves = [nb.Leaf(t, p) for p in params]
    yn = wl.mlp_data_objective(nb, Xd, Yd, 3)(*leaves)
    order = []
    nb.propagate(t, nb.reverse_tape(yn, nb.oneslike(yn)), on_final=lambda pos, rvs: order.append(pos))
    assert sorted(order) == [l.pos for l in leaves] and order[0] in (leaves[2].pos, leaves[5].pos)


def test_split_cache_is_invalidated_by_every_writer(nb, ctx):
    """The tensor-core GEMM caches the hi/lo copies of its operands; any write to an operand (H2D copy,
    a kernel output, a GEMM output, buffer recycling) must drop the cached copy."""
    from nabla_jl_b200 import kernels as K
    rng = np.random.default_rng(8)
    m = n = k = 256
    A = rng.standard_normal((m, k)).astype(np.float32)
    B = rng.standard_normal((k, n)).astype(np.float32)
    Ad, Bd, Cd = ctx.asarray(A), ctx.asarray(B), ctx.empty((m, n), np.float32)

    def check(Ah, Bh):
        K.gemm("N", "N", 1.0, Ad, Bd, 0.0, Cd, m, n, k)
        ref = Ah.astype(np.float64) @ Bh.astype(np.float64)
        err = np.linalg.norm(Cd.numpy() - ref) / np.linalg.norm(ref)
        assert err < 1e-4, err

    check(A, B)
    check(A, B)                                   # served from the cache
    A2 = rng.standard_normal((m, k)).astype(np.float32)
    Ad.copy_from_host(A2)                         # H2D into the cached source
    check(A2, B)
    K.bcast_reduce(None, [Ad], Bd, scale=2.0)     # a kernel writes the other operand (B = 2 A2; m == k == n)
    check(A2, 2 * A2)
    K.gemm("N", "N", 1.0, Bd, Bd, 0.0, Ad, m, n, k)   # a GEMM output overwrites a cached source
    A3 = (2 * A2).astype(np.float64) @ (2 * A2).astype(np.float64)
    check(A3.astype(np.float32), 2 * A2)
    del Ad                                        # recycle the block, then reuse it with new contents
    Ad = ctx.asarray(A)
    check(A, 2 * A2)


# ---------------------------------------------------------------------------------------------
# array structure and checkpointing
# ---------------------------------------------------------------------------------------------
def test_hcat_vcat_known_answers(nb):
    # test/sensitivities/array.jl:13-19
    rng = np.random.default_rng(0)
    a, b, c = rng.random((3, 2)), rng.random(3), rng.random((3, 3))
    g = nb.nabla(lambda a, b, c: nb.sum(nb.hcat(2 * a, 3 * b, 4 * c)))(a, b, c)
    for got, ref in zip(g, (2 * np.ones((3, 2)), 3 * np.ones(3), 4 * np.ones((3, 3)))):
        np.testing.assert_array_equal(_np(got), ref)
    a, b, c = rng.random((2, 4)), rng.random((1, 4)), rng.random((3, 4))
    g = nb.nabla(lambda a, b, c: nb.sum(nb.vcat(2 * a, 3 * b, 4 * c)))(a, b, c)
    for got, ref in zip(g, (2 * np.ones((2, 4)), 3 * np.ones((1, 4)), 4 * np.ones((3, 4)))):
        np.testing.assert_array_equal(_np(got), ref)
    y = nb.hcat(nb.Leaf(nb.Tape(), a), c[:2])
    np.testing.assert_array_equal(_np(y), np.hstack([a, c[:2]]))


def test_getindex_known_answers(nb):
    # test/sensitivities/indexing.jl: Int, Vector (range), overlapping indices (#139)
    leaf = nb.Leaf(nb.Tape(), 5.0 * np.ones(5))
    y = nb.getindex(leaf, 1)
    assert float(_np(y)) == 5.0
    np.testing.assert_array_equal(_np(nb.nabla(y, 1.0)[leaf]), [1, 0, 0, 0, 0])
    x = nb.Leaf(nb.Tape(), 10.0 * np.ones(3))
    y = nb.getindex(x, nb.R(2, 3))
    np.testing.assert_array_equal(_np(y), [10, 10])
    np.testing.assert_array_equal(_np(nb.nabla(y, np.ones(2))[x]), [0, 1, 1])
    x = nb.Leaf(nb.Tape(), 10.0 * np.ones(3))
    y = nb.getindex(x, [2, 3, 3])
    np.testing.assert_array_equal(_np(y), [10, 10, 10])
    np.testing.assert_array_equal(_np(nb.nabla(y, np.ones(3))[x]), [0, 1, 2])
    with pytest.raises(IndexError):
        nb.getindex(x, 4)


def test_array_structure_parity(nb, orc):
    rng = np.random.default_rng(21)
    r = lambda *s: rng.standard_normal(s)
    check_case(nb, orc, lambda ns, a, b, c: ns.hcat(a, b, c), [r(6, 2), r(6), r(6, 4)], r(6, 7))
    check_case(nb, orc, lambda ns, a, b: ns.vcat(a, b), [r(2, 5), r(4, 5)], r(6, 5))
    check_case(nb, orc, lambda ns, a, b: ns.vcat(a, b), [r(3), r(5)], r(8))
    check_case(nb, orc, lambda ns, a: ns.getindex(a, ns.R(2, 4), ns.colon), [r(5, 6)], r(3, 6))
    check_case(nb, orc, lambda ns, a: ns.getindex(a, ns.colon, 3), [r(5, 6)], r(5))
    check_case(nb, orc, lambda ns, a: ns.getindex(a, 2, ns.R(2, 5)), [r(5, 6)], r(4))
    check_case(nb, orc, lambda ns, a: ns.getindex(a, 4, 5), [r(5, 6)], 0.7)
    check_case(nb, orc, lambda ns, a: ns.getindex(a, [7, 1, 7, 3]), [r(9)], r(4))
    check_case(nb, orc, lambda ns, a: ns.reshape(a, 5, 4), [r(2, 10)], r(5, 4))   # array.jl:5-11
    check_case(nb, orc, lambda ns, a: ns.reshape(a, (5, 4)), [r(2, 10)], r(5, 4))
    check_case(nb, orc, lambda ns, s: ns.fill(s, 4, 4), [0.3], r(4, 4))            # array.jl:21-22
    # a slice feeding further device work: minibatch columns of a data matrix
    check_case(nb, orc, lambda ns, W, X: ns.sum(ns.broadcast(ns.tanh, ns.matmul(W, ns.getindex(X, ns.colon, ns.R(3, 9))))),
               [r(4, 6), r(6, 12)])


def _ck_layer(W, h, ns=None):
    return ns.broadcast(ns.tanh, ns.matmul(W, h))


def test_checkpoint(nb, orc):
    # test/checkpointing.jl
    x = 5.0
    assert abs(float(_np(nb.checkpoint(nb.sin, (nb.as_device(x),)))) - math.sin(5.0)) < 1e-15   # device scalar in
    (g,) = nb.nabla(lambda x: nb.checkpoint(nb.sin, (x,)))(x)
    assert abs(float(_np(g)) - math.cos(5.0)) < 1e-15
    foo_ck = lambda x: nb.cos(nb.checkpoint(nb.sin, (x,)))
    foo = lambda x: nb.cos(nb.sin(x))
    assert float(_np(nb.nabla(foo_ck)(x)[0])) == float(_np(nb.nabla(foo)(x)[0]))
    y = 5.0
    with pytest.raises(RuntimeError):
        nb.checkpoint(lambda v: v + y, (x,))
    # a checkpointed layer inside a model: same gradients, the layer's intermediates are not kept
    rng = np.random.default_rng(2)
    W1, W2, X = rng.standard_normal((8, 6)), rng.standard_normal((5, 8)), rng.standard_normal((6, 7))

    Xd = nb.get_context().asarray(X)
    # kwargs variant (checkpointing.jl:37-54): `ns` travels as a keyword, so `_ck_layer` is not a closure
    g_ck = nb.nabla(lambda W1, W2: nb.sum(nb.matmul(W2, nb.checkpoint(_ck_layer, (W1, Xd), {"ns": nb}))))(W1, W2)
    g_ref = orc.nabla(lambda W1, W2: orc.sum(orc.matmul(W2, orc.checkpoint(_ck_layer, (W1, X), {"ns": orc}))))(W1, W2)
    g_plain = nb.nabla(lambda W1, W2: nb.sum(nb.matmul(W2, _ck_layer(W1, Xd, ns=nb))))(W1, W2)
    for a, b_, c in zip(g_ck, g_ref, g_plain):
        close(_np(a), b_)
        np.testing.assert_array_equal(_np(a), _np(c))


# ---------------------------------------------------------------------------------------------
# symmetric / triangular BLAS family (test/sensitivities/linalg/blas.jl:63-135)
# ---------------------------------------------------------------------------------------------
def test_symm_symv(nb, orc):
    import itertools
    rng = np.random.default_rng(123456)
    r = lambda *s: rng.standard_normal(s)
    for N in (10, 100):
        for side, ul in itertools.product("LR", "LU"):
            A, B = r(N, N), r(N, N)
            check_case(nb, orc, lambda ns, a, A, B: ns.symm(side, ul, a, A, B), [float(r()), A, B], r(N, N))
            check_case(nb, orc, lambda ns, A, B: ns.symm(side, ul, A, B), [A, B], r(N, N))
        for ul in "LU":
            check_case(nb, orc, lambda ns, a, A, x: ns.symv(ul, a, A, x), [float(r()), r(N, N), r(N)], r(N))
            check_case(nb, orc, lambda ns, A, x: ns.symv(ul, A, x), [r(N, N), r(N)], r(N))


def test_trmm_trmv(nb, orc):
    import itertools
    rng = np.random.default_rng(123456)
    r = lambda *s: rng.standard_normal(s)
    N = 10
    for side, ul, tA, dA in itertools.product("LR", "LU", "NT", "UN"):
        check_case(nb, orc, lambda ns, a, A, B: ns.trmm(side, ul, tA, dA, a, A, B), [float(r()), r(N, N), r(N, N)], r(N, N))
    for ul, tA, dA in itertools.product("LU", "NT", "UN"):
        check_case(nb, orc, lambda ns, A, b: ns.trmv(ul, tA, dA, A, b), [r(N, N), r(N)], r(N))
    check_case(nb, orc, lambda ns, a, A, B: ns.trmm("L", "U", "T", "N", a, A, B), [0.7, r(70, 70), r(70, 33)], r(70, 33))


def test_trsm_trsv(nb, orc):
    import itertools
    rng = np.random.default_rng(123456)
    r = lambda *s: rng.standard_normal(s)
    N = 10
    for side, ul, tA, dA in itertools.product("LR", "LU", "NT", "UN"):
        A = r(N, N) + 3 * np.eye(N)
        check_case(nb, orc, lambda ns, a, A, X: ns.trsm(side, ul, tA, dA, a, A, X), [float(r()), A, r(N, N)], r(N, N), scale=100.0)
    for ul, tA, dA in itertools.product("LU", "NT", "UN"):
        A = r(N, N) + np.eye(N)
        A = A.T @ A
        check_case(nb, orc, lambda ns, A, x: ns.trsv(ul, tA, dA, A, x), [A, r(N)], r(N), scale=1000.0)
    A = r(150, 150) + 12 * np.eye(150)
    check_case(nb, orc, lambda ns, a, A, X: ns.trsm("L", "L", "N", "N", a, A, X), [1.3, A, r(150, 40)], r(150, 40), scale=100.0)
    check_case(nb, orc, lambda ns, a, A, X: ns.trsm("R", "U", "T", "N", a, A, X), [1.3, A, r(40, 150)], r(40, 150), scale=100.0)


# ---------------------------------------------------------------------------------------------
# tape-level fusion around products: nb_gemm_fused + fusion.py
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode,tol", [("bf16x3", 3e-5), ("3xtf32", 2e-5)])
def test_gemm_fused_epilogue_c_abi(nb, ctx, mode, tol):
    """nb_gemm_fused through the C ABI against NumPy Float64: forward f.(A*B .+ bias) and pullback
    (A*B) .* f'(y), all four transposes, ragged tails, K beyond one chunk and beyond one launch range,
    row sums fresh and accumulated."""
    import ctypes as C
    from nabla_jl_b200 import fusion
    rng = np.random.default_rng(5)
    old = ctx.math_mode
    ctx.math_mode = {"3xtf32": nb.MATH_3XTF32, "bf16x3": nb.MATH_BF16X3}[mode]
    acts = {"tanh": (np.tanh, lambda y: 1 - y * y), "logistic": (lambda x: 1 / (1 + np.exp(-x)), lambda y: y * (1 - y)),
            "identity": (lambda x: x, lambda y: np.ones_like(y)), "exp": (np.exp, lambda y: y)}
    try:
        for (m, n, k) in ((256, 512, 256), (300, 200, 150), (128, 1030, 4200), (520, 260, 8192 + 72), (10, 2048, 1030)):
            for tA in "NT":
                for tB in "NT":
                    A = (rng.standard_normal((k, m) if tA == "T" else (m, k)) / np.sqrt(k)).astype(np.float32)
                    B = rng.standard_normal((n, k) if tB == "T" else (k, n)).astype(np.float32)
                    bias = rng.standard_normal(m).astype(np.float32)
                    P = (A.T if tA == "T" else A).astype(np.float64) @ (B.T if tB == "T" else B).astype(np.float64)
                    Ad, Bd, bd = ctx.asarray(A), ctx.asarray(B), ctx.asarray(bias)
                    for name, (f, df) in acts.items():
                        for kind in (1, 2):
                            y = np.tanh(rng.standard_normal((m, n))).astype(np.float32) * 0.9
                            yd = ctx.asarray(y)
                            rs0 = rng.standard_normal(m).astype(np.float32)
                            for acc in (0, 1):
                                out = ctx.empty((m, n), np.float32)
                                rs = ctx.asarray(rs0)
                                epi = fusion.nb_gemm_epilogue()
                                epi.kind, epi.op = kind, nb._lib.OPCODES[name]
                                epi.bias = bd.ptr if kind == 1 else None
                                epi.y, epi.ldy = (yd.ptr, m) if kind == 2 else (None, 0)
                                epi.rowsum, epi.rowsum_acc, epi.emit_split = rs.ptr, acc, 1
                                st = ctx.L.nb_gemm_fused(ctx.h, tA.encode(), tB.encode(), m, n, k, 1.0, Ad.ptr,
                                                         k if tA == "T" else m, Bd.ptr, n if tB == "T" else k,
                                                         out.ptr, m, out.nb_dtype, ctx.math_mode, C.byref(epi))
                                assert st == 0, f"status {st}"
                                ref = f(P + bias[:, None].astype(np.float64)) if kind == 1 else P * df(y.astype(np.float64))
                                got = out.numpy().astype(np.float64)
                                tag = f"{mode} {tA}{tB} {m}x{n}x{k} {name} kind={kind} acc={acc}"
                                err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
                                assert err < tol, f"{tag}: rel err {err:.3e}"
                                rref = got.sum(axis=1) + (rs0 if acc else 0)
                                rerr = np.linalg.norm(rs.numpy() - rref) / max(np.linalg.norm(rref), 1e-30)
                                assert rerr < 1e-5, f"{tag}: row sums rel err {rerr:.3e}"
                                # the emitted operand split must be what a later product of `out` uses
                                D = rng.standard_normal((n, 64)).astype(np.float32)
                                Dd, E = ctx.asarray(D), ctx.empty((m, 64), np.float32)
                                from nabla_jl_b200 import kernels as K
                                K.gemm("N", "N", 1.0, out, Dd, 0.0, E, m, 64, n)
                                eref = got @ D.astype(np.float64)
                                eerr = np.linalg.norm(E.numpy() - eref) / np.linalg.norm(eref)
                                assert eerr < tol, f"{tag}: product from the emitted split rel err {eerr:.3e}"
    finally:
        ctx.math_mode = old


def test_fused_dense_layers(nb, orc, ctx):
    """f.(W*X .+ b) chains (examples/mlp.jl:79-82) with the product launches deferred and the `.+` / `f.`
    Branches joined to their epilogues: same tape, same gradients as the unfused sweep and as the oracle,
    fewer launches, and the intermediates W*X and W*X .+ b stay unallocated."""
    import nabla_jl_b200.workloads as wl
    dims, B = (96, 256, 192, 128, 10), 512
    W, b, X, Y = wl.mlp_inputs(dims, B, np.float32, seed=12)
    nl = len(dims) - 1
    ref = orc.nabla(wl.mlp_objective(orc, X.astype(np.float64), Y.astype(np.float64), 1e-3, nl))(
        *[w.astype(np.float64) for w in W], *[v.astype(np.float64) for v in b])
    Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
    res = {}
    for fuse in (True, False):
        ctx.fuse = fuse
        try:
            n0 = ctx.stats()["kernel_launches"]
            g = nb.nabla(wl.mlp_objective(nb, Xd, Yd, 1e-3, nl))(*W, *b)
            res[fuse] = ([x.numpy() for x in g], ctx.stats()["kernel_launches"] - n0)
        finally:
            ctx.fuse = True
    for a, r in zip(res[True][0], ref):
        close(a, r, np.float32)
    for a, u in zip(res[True][0], res[False][0]):
        close(a, u, np.float32, scale=0.2)
    assert res[True][1] < res[False][1], f"fusion must save launches: {res[True][1]} vs {res[False][1]}"

    # the tape itself is unchanged: three Branches per layer, and reading a deferred intermediate works
    t = nb.Tape()
    Wn, bn = nb.Leaf(t, W[0]), nb.Leaf(t, b[0])
    z = nb.matmul(Wn, Xd)
    t1 = nb.broadcast(nb.add, z, bn)
    h = nb.broadcast(nb.tanh, t1)
    assert [type(x).__name__ for x in t.tape] == ["Leaf", "Leaf", "Branch", "Branch", "Branch"]
    assert z.val.pending is not None and t1.val.pending is not None and z.val._ptr is None
    s = nb.sum(h)
    rt = nb.nabla(s)
    Z = W[0].astype(np.float64) @ X.astype(np.float64)
    Hh = np.tanh(Z + b[0][:, None])
    close(rt[Wn].numpy(), (1 - Hh * Hh) @ X.astype(np.float64).T, np.float32)
    close(rt[bn].numpy(), (1 - Hh * Hh).sum(axis=1), np.float32)
    assert z.val._ptr is None and t1.val._ptr is None, "unread intermediates are never launched or allocated"
    close(z.val.numpy(), Z, np.float32)          # reading one forces just that product
    close(t1.val.numpy(), Z + b[0][:, None], np.float32)
    close(rt[z].numpy(), 1 - Hh * Hh, np.float32)  # the sensitivity of an intermediate Branch is readable too


def test_fusion_edge_cases(nb, orc, ctx):
    """Deferred products under the tape shapes that must NOT fuse, or fuse only partly: every one must give the
    oracle's values and gradients (fresh and accumulate variants)."""
    rng = np.random.default_rng(21)
    m, k, n = 128, 96, 160   # m*n*k >= 2^18: products are deferred
    W = (rng.standard_normal((m, k)) / np.sqrt(k)).astype(np.float32)
    X = rng.standard_normal((k, n)).astype(np.float32)
    b = rng.standard_normal(m).astype(np.float32)
    b2 = rng.standard_normal((m, 1)).astype(np.float32)
    brow = rng.standard_normal((1, n)).astype(np.float32)
    ybar = rng.standard_normal((m, n)).astype(np.float32)
    V = (rng.standard_normal((64, m)) / np.sqrt(m)).astype(np.float32)
    f32 = np.float32
    cases = {
        # constant (non-Node) bias: joins the forward epilogue, no row-sum in reverse
        "const_bias": (lambda ns, w, x: ns.broadcast(ns.tanh, ns.broadcast(ns.add, ns.matmul(w, x), b)), [W, X], ybar),
        # bias on the left, column-vector bias (m x 1), activation other than tanh
        "bias_left": (lambda ns, w, x, c: ns.broadcast(ns.exp, ns.broadcast(ns.add, c, ns.matmul(w, x))), [W, X, b], ybar),
        "bias_col": (lambda ns, w, x, c: ns.broadcast(orc_logistic(ns), ns.broadcast(ns.add, ns.matmul(w, x), c)), [W, X, b2], ybar),
        # a row bias (1 x n) broadcasts along the other dimension: must not be taken for the epilogue bias
        "bias_row": (lambda ns, w, x, c: ns.broadcast(ns.tanh, ns.broadcast(ns.add, ns.matmul(w, x), c)), [W, X, brow], ybar),
        # activation straight on the product, and a function whose derivative needs its argument (no act fusion)
        "act_only": (lambda ns, w, x: ns.broadcast(ns.tanh, ns.matmul(w, x)), [W, X], ybar),
        "sin_act": (lambda ns, w, x, c: ns.broadcast(ns.sin, ns.broadcast(ns.add, ns.matmul(w, x), c)), [W, X, b], ybar),
        # the product itself is read as well (forces it next to the fused launch)
        "z_used_twice": (lambda ns, w, x, c: ns.broadcast(ns.add, ns.broadcast(ns.tanh, ns.broadcast(ns.add, ns.matmul(w, x), c)),
                                                           ns.matmul(w, x)), [W, X, b], ybar),
        # the activation is consumed twice: its sensitivity slot is accumulated, the f' join must not happen
        "y_used_twice": (lambda ns, w, x, c: (lambda h: ns.broadcast(ns.mul, h, h))(
            ns.broadcast(ns.tanh, ns.broadcast(ns.add, ns.matmul(w, x), c))), [W, X, b], ybar),
        # two layers: the second product's X̄ joins the first layer's tanh pullback and bias row sum
        "two_layers": (lambda ns, w, x, c, v: ns.matmul(v, ns.broadcast(ns.tanh, ns.broadcast(ns.add, ns.matmul(w, x), c))),
                       [W, X, b, V], rng.standard_normal((64, n)).astype(np.float32)),
        # composite lambda after the product: not an epilogue form
        "lambda_after": (lambda ns, w, x: ns.broadcast(lambda t: ns.tanh(t) * t, ns.matmul(w, x)), [W, X], ybar),
    }
    for name, (build, xs, yb) in cases.items():
        try:
            check_case(nb, orc, build, xs, yb, f32)
        except AssertionError as e:
            raise AssertionError(f"{name}: {e}") from e
    # single-pass TF32 goes through the same fused entry point (its own 1e-3 accuracy class)
    old = ctx.math_mode
    ctx.math_mode = nb.MATH_TF32
    try:
        check_case(nb, orc, cases["two_layers"][0], cases["two_layers"][1], cases["two_layers"][2], f32, scale=30.0)
    finally:
        ctx.math_mode = old


def orc_logistic(ns):
    return lambda x: 1.0 / (1.0 + ns.exp(-x))


def test_gemm_cta_pair_kernel(nb, ctx):
    """The CTA-pair kernel (cta_group::2, 256 x 256 tiles, dynamic tile ring, register-fold epilogue) is chosen
    when a product has at least num_sms/2 pair tiles: all four transposes, ragged M / N / K tails, alpha / beta,
    K beyond one chunk and beyond one launch range, and the fused epilogue forms, against NumPy Float64."""
    import ctypes as C
    from nabla_jl_b200 import fusion
    from nabla_jl_b200 import kernels as K
    rng = np.random.default_rng(77)
    old = ctx.math_mode
    ctx.math_mode = nb.MATH_BF16X3
    try:
        for (m, n, k) in ((2304, 4352, 520), (2100, 4401, 1000), (2560, 2304, 8192 + 136)):
            assert ((m + 255) // 256) * ((n + 255) // 256) >= 74
            for tA in "NT":
                for tB in "NT":
                    A = (rng.standard_normal((k, m) if tA == "T" else (m, k)) / np.sqrt(k)).astype(np.float32)
                    B = rng.standard_normal((n, k) if tB == "T" else (k, n)).astype(np.float32)
                    P = (A.T if tA == "T" else A).astype(np.float64) @ (B.T if tB == "T" else B).astype(np.float64)
                    Ad, Bd = ctx.asarray(A), ctx.asarray(B)
                    C0 = rng.standard_normal((m, n)).astype(np.float32)
                    for alpha, beta in ((1.0, 0.0), (-0.7, 1.0)):
                        Cd = ctx.asarray(C0)
                        K.gemm(tA, tB, alpha, Ad, Bd, beta, Cd, m, n, k)
                        ref = alpha * P + beta * C0.astype(np.float64)
                        err = np.linalg.norm(Cd.numpy() - ref) / np.linalg.norm(ref)
                        assert err < 3e-5, f"{tA}{tB} {m}x{n}x{k} alpha={alpha} beta={beta}: rel err {err:.3e}"
                    if tA == "T" and tB == "T":
                        continue  # the fused forms below: three transposes are enough
                    bias = rng.standard_normal(m).astype(np.float32)
                    y = (np.tanh(rng.standard_normal((m, n))) * 0.9).astype(np.float32)
                    bd, yd = ctx.asarray(bias), ctx.asarray(y)
                    for kind in (1, 2):
                        out, rs = ctx.empty((m, n), np.float32), ctx.zeros((m,), np.float32)
                        epi = fusion.nb_gemm_epilogue()
                        epi.kind, epi.op = kind, nb._lib.OPCODES["tanh"]
                        epi.bias = bd.ptr if kind == 1 else None
                        epi.y, epi.ldy = (yd.ptr, m) if kind == 2 else (None, 0)
                        epi.rowsum, epi.rowsum_acc, epi.emit_split = rs.ptr, 0, 1
                        st = ctx.L.nb_gemm_fused(ctx.h, tA.encode(), tB.encode(), m, n, k, 1.0, Ad.ptr,
                                                 k if tA == "T" else m, Bd.ptr, n if tB == "T" else k, out.ptr, m,
                                                 out.nb_dtype, ctx.math_mode, C.byref(epi))
                        assert st == 0
                        ref = np.tanh(P + bias[:, None]) if kind == 1 else P * (1 - y.astype(np.float64) ** 2)
                        got = out.numpy().astype(np.float64)
                        err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
                        assert err < 3e-5, f"fused kind={kind} {tA}{tB} {m}x{n}x{k}: rel err {err:.3e}"
                        rref = got.sum(axis=1)
                        rerr = np.linalg.norm(rs.numpy() - rref) / np.linalg.norm(rref)
                        assert rerr < 1e-5, f"fused kind={kind} {tA}{tB} {m}x{n}x{k}: row sums rel err {rerr:.3e}"
                        D = rng.standard_normal((n, 64)).astype(np.float32)
                        E = ctx.empty((m, 64), np.float32)
                        K.gemm("N", "N", 1.0, out, ctx.asarray(D), 0.0, E, m, 64, n)   # uses the emitted split
                        eref = got @ D.astype(np.float64)
                        eerr = np.linalg.norm(E.numpy() - eref) / np.linalg.norm(eref)
                        assert eerr < 3e-5, f"fused kind={kind} {tA}{tB}: product from the emitted split {eerr:.3e}"
    finally:
        ctx.math_mode = old


def test_fused_dense_layers_on_the_pair_kernel(nb, orc, ctx):
    """The wide-MLP gradient at a size whose hidden-layer products run on the CTA-pair kernel with the fused
    epilogues (>= 74 pair tiles), against the Float64 oracle."""
    import nabla_jl_b200.workloads as wl
    dims, B = (200, 2304, 2304, 10), 4352
    y, g, yr, gr = _mlp_both(nb, orc, dims, B, np.float32, seed=5)
    np.testing.assert_allclose(y, yr, rtol=1e-4)
    for a, b in zip(g, gr):
        close(a, b, np.float32)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_kron_plus_I_copy(nb, orc, dtype):
    # src/sensitivities/linalg/generic.jl:7-51; test/sensitivities/linalg/generic.jl:64-79
    rng = np.random.default_rng(3)
    for (sa, sb) in (((3, 4), (2, 5)), ((5, 5), (5, 5)), ((1, 7), (6, 1)), ((16, 8), (8, 12))):
        A, B = rng.standard_normal(sa), rng.standard_normal(sb)
        ybar = rng.standard_normal((sa[0] * sb[0], sa[1] * sb[1]))
        check_case(nb, orc, lambda ns, a, b: ns.kron(a, b), [A, B], ybar, dtype)
        check_case(nb, orc, lambda ns, a: ns.kron(a, B.astype(dtype)), [A], ybar, dtype)
        check_case(nb, orc, lambda ns, b: ns.kron(A.astype(dtype), b), [B], ybar, dtype)
    A = rng.standard_normal((6, 6))
    ybar = rng.standard_normal((6, 6))
    check_case(nb, orc, lambda ns, a: ns.plus_I(a, 0.65), [A], ybar, dtype)
    check_case(nb, orc, lambda ns, a: ns.copy(a), [A], ybar, dtype)
    with pytest.raises(ValueError):
        nb.plus_I(nb.get_context().asarray(rng.standard_normal((3, 4)).astype(dtype)), 1.0)


def test_fused_dense_layers_inside_graph_capture(nb, orc, ctx):
    """Deferred / fused Float32 product launches inside a CUDA-graph capture: everything the captured sweep returns
    is forced before the capture ends, the replay reproduces the eager gradients and sees new data in the same
    buffers; Float64 products are never deferred (the fused entry point answers NB_ERR_UNSUPPORTED)."""
    import ctypes as C
    import nabla_jl_b200.workloads as wl
    from nabla_jl_b200 import fusion
    dims, B = (96, 256, 192, 10), 512
    W, b, X, Y = wl.mlp_inputs(dims, B, np.float32, seed=12)
    nl = len(dims) - 1
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
    params = [ctx.asarray(p) for p in (*W, *b)]
    eager = [g.numpy() for g in nb.nabla(wl.mlp_objective(nb, Xd, Yd, 1e-3, nl))(*params)]
    n0 = ctx.stats()["kernel_launches"]
    nb.nabla(wl.mlp_objective(nb, Xd, Yd, 1e-3, nl))(*params)
    eager_launches = ctx.stats()["kernel_launches"] - n0
    cap = nb.CapturedGradient(wl.mlp_objective(nb, Xd, Yd, 1e-3, nl), params)
    _, g1 = cap()
    ctx.sync()
    for a, e in zip(g1, eager):
        close(a.numpy(), e, np.float32, scale=0.2)
    assert cap.launches_per_replay > 0 and eager_launches > 0
    X2 = np.random.default_rng(9).random(X.shape).astype(np.float32)
    Xd.copy_from_host(X2)
    _, g2 = cap()
    ctx.sync()
    ref2 = orc.nabla(wl.mlp_objective(orc, f64(X2), f64(Y), 1e-3, nl))(*[f64(p) for p in (*W, *b)])
    for a, r in zip(g2, ref2):
        close(a.numpy(), r, np.float32)
    # Float64
    A64, B64 = ctx.asarray(np.ones((256, 256))), ctx.asarray(np.ones((256, 256)))
    out = ctx.empty((256, 256), np.float64)
    epi = fusion.nb_gemm_epilogue()
    epi.kind, epi.op = 1, nb._lib.OPCODES["tanh"]
    st = ctx.L.nb_gemm_fused(ctx.h, b"N", b"N", 256, 256, 256, 1.0, A64.ptr, 256, B64.ptr, 256, out.ptr, 256,
                             out.nb_dtype, ctx.math_mode, C.byref(epi))
    assert st == 4
    z = nb.matmul(nb.Leaf(nb.Tape(), np.ones((256, 256))), B64)
    assert z.val.pending is None


def test_gemm_f64_ozaki_split(nb, ctx):
    """Large Float64 products run as exact int8 slice products on the tensor cores (csrc/ozaki.cu): all four
    transposes, ragged extents, alpha / beta, against NumPy Float64 at a hundredth of the Float64 bar (1e-12 norm-wise,
    1e-11 element-wise relative to the largest entry: seven slices leave ~4e-14, so that chains of products stay far
    inside rtol 1e-10) and against the DMMA path (NB_MATH_FP32 = no tensor-core split)."""
    from nabla_jl_b200 import kernels as K
    rng = np.random.default_rng(123)
    m, n, k = 4100, 4200, 4000   # m*n*k >= 6e10: the split path
    old = ctx.math_mode
    try:
        for tA in "NT":
            for tB in "NT":
                A = rng.standard_normal((k, m) if tA == "T" else (m, k)) * np.exp(rng.uniform(-3, 3, (1, m) if tA == "T" else (m, 1)))
                B = rng.standard_normal((n, k) if tB == "T" else (k, n))
                C0 = rng.standard_normal((m, n))
                ref = 0.7 * ((A.T if tA == "T" else A) @ (B.T if tB == "T" else B)) - 0.5 * C0
                Ad, Bd = ctx.asarray(A), ctx.asarray(B)
                out = {}
                for mode in (nb.MATH_DEFAULT, nb.MATH_FP32):
                    ctx.math_mode = mode
                    n0 = ctx.stats()["kernel_launches"]
                    Cd = ctx.asarray(C0)
                    K.gemm(tA, tB, 0.7, Ad, Bd, -0.5, Cd, m, n, k)
                    out[mode] = (Cd.numpy(), ctx.stats()["kernel_launches"] - n0)
                close(out[nb.MATH_DEFAULT][0], ref, scale=0.01)
                close(out[nb.MATH_FP32][0], ref, scale=0.01)
                assert out[nb.MATH_DEFAULT][1] >= 11 and out[nb.MATH_FP32][1] <= 3   # split path: scales, slices, 7 levels
        # mid-size products: every level in ONE launch (Float64 running sum in the epilogue), seven levels
        for (mm, nn, kk) in ((1100, 1300, 2100), (700, 4096, 1100)):
            for tA in "NT":
                for tB in "NT":
                    A = rng.standard_normal((kk, mm) if tA == "T" else (mm, kk))
                    B = rng.standard_normal((nn, kk) if tB == "T" else (kk, nn))
                    C0 = rng.standard_normal((mm, nn))
                    ref = -1.3 * ((A.T if tA == "T" else A) @ (B.T if tB == "T" else B)) + C0
                    ctx.math_mode = nb.MATH_DEFAULT
                    Cd = ctx.asarray(C0)
                    n0 = ctx.stats()["kernel_launches"]
                    K.gemm(tA, tB, -1.3, ctx.asarray(A), ctx.asarray(B), 1.0, Cd, mm, nn, kk)
                    assert 8 <= ctx.stats()["kernel_launches"] - n0 <= 11   # scales, slices, guard, one split launch, idle DMMA
                    close(Cd.numpy(), ref, scale=0.01)   # 1e-12
        # non-finite input: the affected row of C is NaN (never finite garbage), the rest is untouched
        A = rng.standard_normal((m, k))
        B = rng.standard_normal((k, n))
        A[7, 11] = np.inf
        A[300, 5] = np.nan
        ctx.math_mode = nb.MATH_DEFAULT
        Cd = ctx.zeros((m, n), np.float64)
        K.gemm("N", "N", 1.0, ctx.asarray(A), ctx.asarray(B), 0.0, Cd, m, n, k)
        got = Cd.numpy()
        assert np.all(np.isnan(got[7])) and np.all(np.isnan(got[300]))
        assert np.all(np.isfinite(np.delete(got, [7, 300], axis=0)))
    finally:
        ctx.math_mode = old


def test_gemm_f64_split_range_guard(nb, ctx):
    """dgemm's error is relative to sum_k |a_ik||b_kj|, the int8 split's to max|a_i.| * max|b_.j| (csrc/ozaki.cu): a row
    dominated by one outlier (1e9 here) that meets a zero of B would keep only ~19 bits of its other entries.  The guard measures
    how far the largest entry of a line stands above the rest in the pass that finds the scales and — on the device,
    without a host synchronisation — hands such a product to DMMA.  Both split variants (one launch / one launch per
    level); with the guard disabled by its limit the same product shows the loss the guard prevents."""
    from nabla_jl_b200 import kernels as K
    rng = np.random.default_rng(321)
    for (m, n, k) in ((2048, 2048, 1024), (4100, 4200, 4000)):
        A = rng.standard_normal((m, k))
        B = rng.standard_normal((k, n))
        A[5, 0] = 1e9          # the outlier of row 5 ...
        B[0, :] = 0.0          # ... meets zeros: C[5, :] is made of the ordinary entries of the row
        ref = A @ B
        Cd = ctx.zeros((m, n), np.float64)
        K.gemm("N", "N", 1.0, ctx.asarray(A), ctx.asarray(B), 0.0, Cd, m, n, k)
        got = Cd.numpy()
        err5 = np.abs(got[5] - ref[5]).max() / np.abs(ref[5]).max()
        assert err5 < 1e-13, err5
        close(got, ref, scale=0.01)
        # ordinary data is NOT handed to DMMA: same product without the outlier stays on the split (launch count)
        A[5, 0] = 0.3
        n0 = ctx.stats()["kernel_launches"]
        K.gemm("N", "N", 1.0, ctx.asarray(A), ctx.asarray(B), 0.0, Cd, m, n, k)
        assert ctx.stats()["kernel_launches"] - n0 >= 8
        close(Cd.numpy(), A @ B, scale=0.01)


def test_mlp_gradient_f64_on_the_split_path(nb, orc):
    """A Float64 MLP whose hidden-layer products (2048^3 = 8.6e9) run on the Ozaki split: the whole gradient
    against the oracle at the Float64 bar."""
    y, g, yr, gr = _mlp_both(nb, orc, (2048, 2048, 2048, 10), 2048, np.float64, seed=7)
    np.testing.assert_allclose(y, yr, rtol=1e-10)
    for a, b in zip(g, gr):
        close(a, b, scale=0.02)


def test_mlp_gradient_f64_chain_of_split_products(nb, orc):
    """Five layers of 4096^3 Float64 products, all on the multi-launch split: the error of a CHAIN of split products
    (a dozen in sequence, forward and reverse) must stay far inside rtol 1e-10.  Measured 9e-13 with seven slices
    (DMMA: 8e-14); six slices gave 8e-11, which is why seven is the default (tools/check_f64_split_depth.py)."""
    y, g, yr, gr = _mlp_both(nb, orc, (784, 4096, 4096, 4096, 4096, 10), 4096, np.float64, seed=4)
    np.testing.assert_allclose(y, yr, rtol=1e-11)
    for a, b in zip(g, gr):
        close(a, b, scale=0.05)


# ---------------------------------------------------------------------------------------------
# value-dependent scalar functions: fmad differentiates ANY scalar f along the branch taken (core.jl:285-292)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_value_dependent_lambdas(nb, orc, dtype):
    """`x > 0 ? x : 0.1x` and friends, written with ifelse (both branches traced, the derivative of the taken one):
    values and sensitivities against the oracle's duals, whose comparisons act on values like ForwardDiff's."""
    rng = np.random.default_rng(21)
    sc = 1.0 if dtype == np.float64 else 1.0
    x = rng.standard_normal((37, 29))
    x[3, 4] = 0.0          # exactly on the breakpoint
    y = rng.standard_normal((37, 29))
    leaky = lambda ns: (lambda a: ns.ifelse(a > 0.0, a, 0.1 * a))
    check_case(nb, orc, lambda ns, a: ns.broadcast(leaky(ns), a), [x], rng.standard_normal((37, 29)), dtype, sc)
    # the branch NOT taken is NaN / Inf there (log of a negative number, 1/0): it must contribute exactly zero
    safe_log = lambda ns: (lambda a: ns.ifelse(a > 0.0, ns.log(a), 0.0 * a))
    check_case(nb, orc, lambda ns, a: ns.broadcast(safe_log(ns), a), [x], rng.standard_normal((37, 29)), dtype, sc)
    safe_inv = lambda ns: (lambda a: ns.ifelse(a > 0.0, 1.0 / ns.sqrt(a), a * a))
    check_case(nb, orc, lambda ns, a: ns.broadcast(safe_inv(ns), a), [x], rng.standard_normal((37, 29)), dtype, sc)
    # two operands, unbroadcast of the second, comparison between operands
    pick = lambda ns: (lambda a, b: ns.ifelse(a >= b, a * b, a - b))
    check_case(nb, orc, lambda ns, a, b: ns.broadcast(pick(ns), a, b), [x, y[:, :1]], rng.standard_normal((37, 29)), dtype, sc)
    # comparisons as values, step / sign / floor inside a composite
    comp = lambda ns: (lambda a: ns.step(a) * ns.exp(a) + ns.sign(a) * a + ns.floor(a) * 0.5 + (a < 0.25) * a * a)
    check_case(nb, orc, lambda ns, a: ns.broadcast(comp(ns), a), [x], rng.standard_normal((37, 29)), dtype, sc)
    # inside a reduction
    check_case(nb, orc, lambda ns, a: ns.sum(leaky(ns), a), [x], None, dtype, sc)


# ---------------------------------------------------------------------------------------------
# tape-level fusion of adjacent broadcast Branches (ewfusion.py)
# ---------------------------------------------------------------------------------------------
def _with_ew_fusion(nb, on, fn):
    from nabla_jl_b200 import ewfusion
    old = ewfusion.ENABLED
    ewfusion.ENABLED = on
    try:
        return fn()
    finally:
        ewfusion.ENABLED = old


def test_adjacent_broadcast_fusion_matches_unfused_and_oracle(nb, orc, ctx):
    """The same tapes with and without the fusion: equal to rounding (the composed program evaluates the same
    primitives; only the order in which a slot receives its contributions can change), both equal to the oracle —
    chains, a value with two consumers (must be materialised), the same Node twice, broadcast operands, reductions over
    chains, array seeds, pre-filled slots."""
    rng = np.random.default_rng(31)
    A, Bm = rng.standard_normal((23, 17)), rng.standard_normal((23, 17))
    bcol, brow = rng.standard_normal(23), rng.standard_normal((1, 17))
    Y = (rng.random((23, 17)) > 0.5) * 1.0

    def chain(ns, a, b, c):          # three adjacent broadcasts + a reduction over the chain
        t = ns.broadcast(ns.tanh, ns.broadcast(ns.add, a, c))
        u = ns.broadcast(ns.mul, t, ns.broadcast(ns.exp, b))
        return ns.sum(ns.broadcast(ns.log, ns.broadcast(ns.add, ns.broadcast(ns.abs2, u), 1e-3)))

    def two_consumers(ns, a, b, c):  # `t` feeds two Branches: it gets its own sensitivity slot
        t = ns.broadcast(ns.tanh, ns.broadcast(ns.add, a, c))
        return ns.sum(ns.broadcast(ns.mul, t, b)) + ns.sum(ns.abs2, t)

    def same_twice(ns, a, b, c):     # σ .* σ (examples/vae.jl:85) inside a chain
        s = ns.broadcast(ns.exp, ns.broadcast(ns.mul, a, 0.3))
        return ns.sum(ns.broadcast(ns.log, ns.broadcast(ns.mul, s, s))) + ns.sum(ns.broadcast(ns.add, b, c))

    def loglik(ns, a, b, c):         # the log-likelihood chain of examples/mlp.jl:68-73
        f = ns.broadcast(lambda x: 1.0 / (1.0 + ns.exp(-x)), ns.broadcast(ns.add, a, c))
        p = ns.broadcast(ns.div, f, ns.sum(f, dims=1))
        t1 = ns.broadcast(ns.mul, Y, ns.broadcast(ns.log, ns.broadcast(ns.add, p, 1e-15)))
        t2 = ns.broadcast(ns.mul, ns.broadcast(ns.sub, 1.0, Y), ns.broadcast(ns.log, ns.broadcast(ns.sub, 1.0 + 1e-15, p)))
        return ns.sum(ns.broadcast(ns.add, t1, t2)) + 0.0 * ns.sum(b)

    def array_out(ns, a, b, c):      # array-valued result with an explicit seed, unbroadcast over rows
        return ns.broadcast(ns.mul, ns.broadcast(ns.tanh, ns.broadcast(ns.add, a, brow)), ns.broadcast(ns.add, b, c))

    for build, ybar in ((chain, None), (two_consumers, None), (same_twice, None), (loglik, None),
                        (array_out, rng.standard_normal((23, 17)))):
        xs = [A, Bm, bcol]
        for on in (True, False):
            _with_ew_fusion(nb, on, lambda: check_case(nb, orc, build, xs, ybar))
        fused = _with_ew_fusion(nb, True, lambda: run_both(nb, orc, build, xs, ybar))
        plain = _with_ew_fusion(nb, False, lambda: run_both(nb, orc, build, xs, ybar))
        np.testing.assert_allclose(fused[0], plain[0], rtol=1e-13)
        for a, b in zip(fused[1], plain[1]):
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-13 * np.max(np.abs(b)))


def test_adjacent_broadcast_fusion_cuts_launches_of_c1_and_c4(nb, orc, ctx):
    """C1 (examples/mlp.jl) and C4 (examples/vae.jl) at a small batch: fewer kernel launches per evaluation with the
    fusion, gradients equal to the unfused sweep and to the oracle; also inside a CUDA-graph capture."""
    import nabla_jl_b200.workloads as wl
    W, b, X, Y = wl.mlp_inputs((784, 500, 300, 10), 256, np.float64, seed=0)
    Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
    vp, VX, noise = wl.vae_inputs(B=256, seed=3)
    VXd, Nd = ctx.asarray(VX), ctx.asarray(noise)
    for name, obj, oobj, params in (
            ("mlp", wl.mlp_objective(nb, Xd, Yd, 1e-3, 3), wl.mlp_objective(orc, X, Y, 1e-3, 3), [*W, *b]),
            ("vae", wl.vae_objective(nb, VXd, Nd, 1e-3, 60000 / 256, 10), wl.vae_objective(orc, VX, noise, 1e-3, 60000 / 256, 10),
             list(vp))):
        dparams = [ctx.asarray(p) for p in params]
        ref = oobj and orc.nabla(oobj)(*params)
        counts, grads = {}, {}
        for on in (False, True):
            def run():
                nb.nabla(obj)(*dparams)                      # warm: programs compiled, pool sized
                n0 = ctx.stats()["kernel_launches"]
                g = nb.nabla(obj)(*dparams)
                g = [x.numpy() for x in g]
                return ctx.stats()["kernel_launches"] - n0, g
            counts[on], grads[on] = _with_ew_fusion(nb, on, run)
        assert counts[True] < 0.8 * counts[False], f"{name}: {counts}"
        for a, r, o in zip(grads[True], grads[False], ref):
            np.testing.assert_allclose(a, r, rtol=1e-11, atol=1e-12 * np.max(np.abs(r)))
            close(a, o)
        cap = _with_ew_fusion(nb, True, lambda: nb.CapturedGradient(obj, dparams, ctx))
        _, g = cap()
        ctx.sync()
        for a, o in zip(g, ref):
            close(a.numpy(), o)
        cap.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_two_op_primitives(nb, orc, dtype):
    """un(bin(a, b)) lambdas — `tanh.(A .+ b)`, `log.(f .+ eps)`, `exp.(a .* b)` … — run as one fused primitive of the
    single-primitive kernels (csrc/ops.cuh nb_fop): every supported pair, array / column / scalar-constant second
    operands (flat, tall and per-column mappings), large enough for the vectorised kernels, fresh and pre-filled
    slots, against the oracle."""
    rng = np.random.default_rng(41)
    m, n = 256, 96
    un = {"tanh": (-2, 2), "exp": (-2, 2), "log": (0.5, 3), "logistic": (-3, 3), "sqrt": (0.5, 3), "abs2": (-2, 2)}
    for uname, (lo, hi) in un.items():
        for bname in ("add", "sub", "mul"):
            # operands chosen so that bin(a, b) lies in the domain of un
            if bname == "add":
                a, b = rng.uniform(lo / 2, hi / 2, (m, n)), rng.uniform(lo / 2, hi / 2, (m, n))
            elif bname == "sub":
                a, b = rng.uniform(hi * 0.75, hi, (m, n)), rng.uniform(0, hi * 0.75 - max(lo, 0), (m, n))
                if lo < 0:
                    a, b = rng.uniform(lo / 2, hi / 2, (m, n)), rng.uniform(lo / 2, hi / 2, (m, n))
            else:
                s = np.sqrt(hi)
                a, b = rng.uniform(np.sqrt(max(lo, 0.25)), s, (m, n)), rng.uniform(np.sqrt(max(lo, 0.25)), s, (m, n))
            def outer(ns, t, u=uname):   # (the oracle has no `logistic` primitive: examples/mlp.jl:42 writes it out)
                return 1.0 / (1.0 + ns.exp(-t)) if u == "logistic" else getattr(ns, u)(t)
            f = lambda ns, bn=bname: (lambda x, y: outer(ns, getattr(ns, bn)(x, y)))
            ybar = rng.standard_normal((m, n))
            sc = 4.0 if dtype == np.float32 else 1.0
            check_case(nb, orc, lambda ns, x, y: ns.broadcast(f(ns), x, y), [a, b], ybar, dtype, sc)
            check_case(nb, orc, lambda ns, x, y: ns.broadcast(f(ns), x, y), [a, b[:, 0]], ybar, dtype, sc)      # bias column
            check_case(nb, orc, lambda ns, x, y: ns.broadcast(f(ns), x, y), [a, b[:1, :]], ybar, dtype, sc)    # row vector
            c = float(b[0, 0])
            check_case(nb, orc, lambda ns, x: ns.broadcast(lambda t: outer(ns, getattr(ns, bname)(t, c)), x),
                       [a], ybar, dtype, sc)
            check_case(nb, orc, lambda ns, x, y: ns.sum(f(ns), x) if False else ns.sum(ns.broadcast(f(ns), x, y)), [a, b], None,
                       dtype, sc)


def test_two_contexts_on_two_gpus_in_one_process(nb):
    """One host thread driving a context on each of two GPUs (the Julia host model is one ctx per task): every entry
    point selects its context's device, and per-device kernel attributes (the > 48 KB shared memory of the GEMM
    kernels) are set on both.  Skipped on a one-GPU box."""
    import ctypes as C
    n = C.c_int32()
    nb._lib.lib().nb_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(5)
    ctxs = [nb.Context(device=0), nb.Context(device=1)]
    try:
        outs = []
        for dtype in (np.float32, np.float64):
            A = rng.standard_normal((1024, 768)).astype(dtype)
            B = rng.standard_normal((768, 1536)).astype(dtype)
            res = []
            for c in ctxs:       # interleave the two devices call by call
                res.append((c.asarray(A), c.asarray(B), c.empty((1024, 1536), dtype)))
            from nabla_jl_b200 import kernels as K
            for (a, b, o) in res:
                K.gemm("N", "N", 1.0, a, b, 0.0, o, 1024, 1536, 768)
            for (a, b, o) in res:
                K.bcast_fwd(nb.compile_scalar(nb.tanh, 1), [o], o)
            ref = np.tanh(A.astype(np.float64) @ B.astype(np.float64))
            for (_, _, o) in res:
                err = np.linalg.norm(o.numpy() - ref) / np.linalg.norm(ref)
                assert err < (1e-4 if dtype == np.float32 else 1e-12), err
                outs.append(o)
        del outs, res
    finally:
        for c in ctxs:
            c.close()
