"""Pins the CPU oracle against every golden vector / known-answer case the reference holds for the
hot path, and against the reference's own test method (5-point central finite differences, both the
fresh-slot and the pre-filled-slot variants: src/finite_differencing.jl:106-120).  CPU only."""
import math
import zlib

import numpy as np
import pytest

import nabla_oracle as o

bc = o.broadcast


def test_docs_autodiff_golden():
    # docs/src/pages/autodiff.md:203-221, 252-258, 285-286
    x, y = 0.6791074260357777, 0.8284134829000359
    t = o.Tape()
    x_, y_ = o.Leaf(t, x), o.Leaf(t, y)
    z = x_ * y_ + o.sin(x_)
    vals = [o.unbox(n) for n in t.tape]
    assert vals == [x, y, 0.5625817480655771, 0.6280987324705773, 1.1906804805361544]
    assert [n.f.name for n in t.tape[2:]] == ["mul", "sin", "add"]
    rt = o.nabla(z)
    assert [rt[i] for i in range(1, 6)] == [1.6065471361170487, 0.6791074260357777, 1.0, 1.0, 1.0]
    assert o.nabla(lambda a, b: a * b + o.sin(a))(x, y) == (1.6065471361170487, 0.6791074260357777)


def test_broadcastsum_integer_cases():
    # test/sensitivities/functional/functional.jl:6-12
    Z, X, Y = np.array([[1, 2], [3, 4]]), np.array([[1, 2]]), np.array([[5, 6], [7, 8]])
    f = lambda x: 2 * x
    assert (o.broadcastsum_inplace(f, False, X.copy(), Z) == [[2 + 6, 4 + 8]]).all()
    assert (o.broadcastsum_inplace(f, True, X.copy(), Z) == X + [[2 + 6, 4 + 8]]).all()
    assert (o.broadcastsum_inplace(f, False, Y.copy(), Z) == 2 * Z).all()
    assert (o.broadcastsum_inplace(f, True, Y.copy(), Z) == Y + 2 * Z).all()


def test_issue_111_117_literal_pow():
    # test/sensitivities/functional/functional.jl:208-227
    x = np.array([1.0, 2.0, 3.0])
    f = lambda x: o.sum(bc(o.mul, np.array([1.0, 2, 3]), bc(o.add, x, np.array([3.0, 2, 1]))))
    (g,) = o.nabla(f)(x)
    assert isinstance(g, np.ndarray) and g.dtype == np.float64
    assert (g == [1.0, 2.0, 3.0]).all()
    y, _ = o.nabla(f, get_output=True)(x)
    assert o.unbox(y) == float(np.sum(np.array([1.0, 2, 3]) * (x + np.array([3.0, 2, 1]))))
    f = lambda x: o.sum(bc(o.add, x, np.array([1, 2, 4]) * np.array([4, 2, 1])))
    assert (o.nabla(f)(x)[0] == np.ones(3)).all()
    f = lambda x: o.sum(bc(lambda a: a ** 2, x))
    assert (o.nabla(f)(x)[0] == [2.0, 4.0, 6.0]).all()
    assert o.unbox(o.nabla(f, get_output=True)(x)[0]) == 14.0


def test_core_known_answers():
    # test/core.jl:129-173
    assert o.nabla(lambda a, b, c, d: a * c)(1, 2, 3, 4) == (3, 0, 1, 0)
    g = o.nabla(lambda a, b, c, d: a * c)(1, np.array([2.0]), 3, 4.0)
    assert g[0] == 3 and (g[1] == [0.0]).all() and g[2] == 1 and g[3] == 0.0
    assert o.nabla(lambda a, b: 12)(1, 2) == (0, 0)
    f = lambda x, y: 2 * x + y
    with pytest.raises(TypeError):  # MethodError: array output without ȳ
        o.nabla(f)(np.random.randn(5), np.random.randn(5))
    x, y = 0.3, -1.2
    t = o.Tape()
    dz = o.nabla(f(o.Leaf(t, x), o.Leaf(t, y)))
    assert o.nabla(f)(x, y) == (dz[1], dz[2]) == (2.0, 1.0)
    z, (gx, gy) = o.nabla(f, get_output=True)(x, y)
    assert o.unbox(z) == f(x, y) and (gx, gy) == (2.0, 1.0)
    out = o.nabla(lambda a: -a, get_output=True)(2)
    assert isinstance(out[0], o.Branch) and out[1] == (-1,)


def test_reducedim_known_answers():
    # test/sensitivities/functional/reducedim.jl:4-25, 33-47, 56-61
    x = o.Leaf(o.Tape(), np.array([1.0, 2, 3, 4, 5]))
    s = 5.0 * o.mapreduce(o.abs2, "+", x, dims=1)
    np.testing.assert_allclose(o.nabla(s, o.oneslike(o.unbox(s)))[x], 5.0 * np.array([2.0, 4, 6, 8, 10]))
    x2_ = np.array([[1.0, 3.0], [2.0, 4.0]])
    x2 = o.Leaf(o.Tape(), x2_)
    s = o.mapreduce(o.abs2, "+", x2, dims=1)
    np.testing.assert_allclose(o.nabla(s, o.oneslike(o.unbox(s)))[x2], 2.0 * x2_)
    rng = np.random.default_rng(123456)
    x3_ = rng.standard_normal((5, 4))
    x3 = o.Leaf(o.Tape(), x3_)
    s = o.mapreduce(o.exp, "+", x3, dims=1)
    assert (o.nabla(s, o.oneslike(o.unbox(s)))[x3] == np.exp(x3_)).all()
    x4_ = rng.standard_normal((5, 4))
    x4 = o.Leaf(o.Tape(), x4_)
    s = o.mapreduce(lambda v: v * v, "+", x4, dims=2)
    assert (o.nabla(s, o.oneslike(o.unbox(s)))[x4] == 2 * x4_).all()
    x5_ = np.ones((5, 4, 3))
    x5 = o.Leaf(o.Tape(), x5_)
    for d, shp, val in ((1, (1, 4, 3), 5.0), (2, (5, 1, 3), 4.0), (3, (5, 4, 1), 3.0)):
        assert (o.unbox(o.sum(x5, dims=d)) == np.full(shp, val)).all()
        assert (o.unbox(o.mean(x5, dims=d)) == np.ones(shp)).all()
    assert o.unbox(o.sum(x5)) == 60.0 and o.unbox(o.mean(x5)) == 1.0
    assert o.sum(x5).f.name == "sum" and o.mean(x5, dims=1).f.name == "mean"
    xs = o.Leaf(o.Tape(), rng.standard_normal((5, 4, 3)))
    assert (o.unbox(o.sum(xs, dims=[2, 3])) == o.unbox(o.mapreduce(o.identity, "+", xs, dims=[2, 3]))).all()
    # Issue #123
    x6 = np.arange(1.0, 11.0)
    tens = np.full(10, 10.0)
    assert (o.nabla(lambda x: o.sum(o.sum(x, dims=2)))(x6)[0] == np.ones(10)).all()
    g = o.nabla(lambda x, y: o.sum(bc(o.add, o.sum(x, dims=2), o.adjoint(o.sum(y, dims=2)))))(x6, x6)
    assert (g[0] == tens).all() and (g[1] == tens).all()
    g = o.nabla(lambda x, y: o.sum(bc(o.add, x, o.adjoint(y))))(x6, x6)
    assert (g[0] == tens).all() and (g[1] == tens).all()


def test_quadratic_identity():
    # docs/src/index.md:25-39: ∇(y'(A x)) = (A'y, A x)  =>  ∇(x'Ax) = (A+A')x
    rng = np.random.default_rng(1)
    A = rng.standard_normal((7, 7))
    x, y = rng.standard_normal(7), rng.standard_normal(7)
    (g,) = o.nabla(lambda x: o.adjoint(x) * (A * x))(x)
    np.testing.assert_allclose(g, (A + A.T) @ x, rtol=1e-13)
    gx, gy = o.nabla(lambda x, y: o.adjoint(x) * (A * y))(x, y)
    np.testing.assert_allclose(gx, A @ y, rtol=1e-13)
    np.testing.assert_allclose(gy, A.T @ x, rtol=1e-13)


def test_accumulate_mutates_only_target_slot():
    # test/core.jl:206-224 with the custom `quad` primitive
    quad = o.Primitive(
        "quad", forward=lambda A, B: B.T @ A @ B,
        grads={1: lambda n, p, Y, Yb, A, B: B @ Yb @ B.T,
               2: lambda n, p, Y, Yb, A, B: A @ B @ Yb.T + A.T @ B @ Yb})
    rng = np.random.default_rng(123456)
    n = 5
    A = o.Leaf(o.Tape(), rng.standard_normal((n, n)))
    B = rng.standard_normal((n, n))
    Q = quad(A, B)
    QQ = quad(Q, B)
    rt = o.nabla(QQ, np.eye(n))
    old = [np.array(v, copy=True) for v in rt.tape]
    o.propagate_branch(Q, rt)  # triggers a mutating addition into slot 1
    assert not np.allclose(old[0], rt.tape[0])
    for a, b in zip(old[1:], rt.tape[1:]):
        np.testing.assert_allclose(a, b)


def _domain(name):
    """domain1 of src/finite_differencing.jl:175-179 over the reference's test points."""
    pts = [-math.pi + .1, -.5 * math.pi + .1, -.9, -.1, .1, .9, .5 * math.pi - .1, math.pi - .1]
    f = getattr(o, "abs_" if name == "abs" else name)
    sets, cur = [], []
    with np.errstate(all="ignore"):
        for p in pts:
            v = f(np.float64(p))
            if np.isfinite(v):
                cur.append(p)
            else:
                if cur:
                    sets.append(cur)
                cur = []
    if cur:
        sets.append(cur)
    best = max(sets, key=lambda s: max(s) - min(s))
    return min(best), max(best)


@pytest.mark.parametrize("name", o.UNARY_NAMES)
def test_unary_catalogue_matches_fd(name):
    # test/sensitivities/functional/functional.jl:15-26 (broadcast gradient == scalar derivative)
    lo, hi = _domain(name)
    rng = np.random.default_rng(zlib.crc32(name.encode()))      # (hash() is salted per process: not reproducible)
    x = rng.uniform(lo, hi, 50)
    f = getattr(o, "abs_" if name == "abs" else name)
    x_ = o.Leaf(o.Tape(), x)
    s = bc(f, x_)
    g = o.nabla(s, np.ones_like(x))[x_]
    h = 1e-6
    with np.errstate(all="ignore"):
        fd = (f(x + h) - f(x - h)) / (2 * h)
        fd2 = (f(x + 2 * h) - f(x - 2 * h)) / (4 * h)
    # the interval may contain poles (tan, gamma, digamma, ...): keep the points where the difference quotient itself
    # has converged (two step sizes agree), i.e. where it is a valid reference
    ok = np.isfinite(fd) & np.isfinite(fd2) & (np.abs(fd - fd2) <= 1e-6 * np.maximum(np.abs(fd), 1e-3))
    if name == "abs":
        ok &= np.abs(x) > 1e-3
    assert ok.sum() >= 30
    np.testing.assert_allclose(g[ok], fd[ok], rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("name", o.BINARY_NAMES)
def test_binary_catalogue_check_errs(name):
    rng = np.random.default_rng(7)
    f = getattr(o, {"max": "max_", "min": "min_", "pow": "pow_"}.get(name, name))
    x, y = rng.uniform(0.3, 2.0, 20), rng.uniform(0.3, 2.0, 20)
    vx, vy = rng.standard_normal(20), rng.standard_normal(20)
    assert o.check_errs(lambda a, b: bc(f, a, b), rng.standard_normal(20), (x, y), (vx, vy))
    # array ∘ scalar and scalar ∘ array, Ref (test/.../functional.jl:151-184)
    xs, g = 0.7, rng.standard_normal(20)
    t = o.Tape()
    a_, b_ = o.Leaf(t, xs), o.Leaf(t, y)
    s = bc(f, a_, b_)
    r1 = o.nabla(s, g)
    t2 = o.Tape()
    a2, b2 = o.Leaf(t2, np.full(20, xs)), o.Leaf(t2, y)
    r2 = o.nabla(bc(f, a2, b2), g)
    np.testing.assert_allclose(r1[a_], np.sum(r2[a2]), rtol=1e-12)
    np.testing.assert_allclose(r1[b_], r2[b2], rtol=1e-12)


def test_scalar_catalogue_against_mpmath_golden():
    """tests/golden/scalar_catalogue_mpmath.json (50-digit mpmath, generated by tests/golden/make_scalar_catalogue.py):
    every function of test/runtests.jl:28-52 and its derivative at the reference's own test points.  The oracle's
    values AND its rules (through the tape: broadcast forward, sensitivity for a seed of ones) must agree to 1e-13."""
    import json
    import os
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scalar_catalogue_mpmath.json")))
    assert set(o.UNARY_NAMES) <= set(G["unary"]) and set(o.SECOND_ARG_ONLY_NAMES) == set(G["second_arg_only"])
    for name in o.UNARY_NAMES:
        t = G["unary"][name]
        x = np.array(t["x"])
        x_ = o.Leaf(o.Tape(), x)
        s = bc(getattr(o, "abs_" if name == "abs" else name), x_)
        np.testing.assert_allclose(o.unbox(s), t["f"], rtol=1e-13, atol=0, err_msg=name)
        np.testing.assert_allclose(o.nabla(s, np.ones_like(x))[x_], t["df"], rtol=1e-13, atol=0, err_msg=name)
    for name in o.SECOND_ARG_ONLY_NAMES:
        for row in G["second_arg_only"][name]:
            x = np.array(row["x"])
            x_ = o.Leaf(o.Tape(), x)
            s = bc(getattr(o, name), row["n"], x_)
            np.testing.assert_allclose(o.unbox(s), row["f"], rtol=1e-13, atol=0, err_msg=f"{name}({row['n']}, x)")
            np.testing.assert_allclose(o.nabla(s, np.ones_like(x))[x_], row["df"], rtol=1e-13, atol=0,
                                       err_msg=f"d/dx {name}({row['n']}, x)")


@pytest.mark.parametrize("name", o.SECOND_ARG_ONLY_NAMES)
def test_second_arg_only_catalogue_matches_fd(name):
    # test/sensitivities/scalar.jl:38-49 — besseli/j/k/y and polygamma: the first argument is an integer 0..5 and is
    # not differentiable; the rule in the second argument against central differences.
    f = getattr(o, name)
    rng = np.random.default_rng(11)
    for n in range(6):
        x = rng.uniform(0.2, 3.0, 30)
        x_ = o.Leaf(o.Tape(), x)
        g = o.nabla(bc(f, n, x_), np.ones_like(x))[x_]
        h = 1e-6
        fd = (f.f(n, x + h) - f.f(n, x - h)) / (2 * h)
        np.testing.assert_allclose(g, fd, rtol=2e-5, atol=1e-7)
    # a Node in the order position gets DiffRules' NaN
    t = o.Tape()
    n_, x_ = o.Leaf(t, np.full(4, 2.0)), o.Leaf(t, np.linspace(0.5, 2.0, 4))
    r = o.nabla(bc(f, n_, x_), np.ones(4))
    assert np.all(np.isnan(r[n_])) and np.all(np.isfinite(r[x_]))


def test_ternary_broadcast_and_fmad():
    # test/sensitivities/functional/functional.jl:68-78
    rng = np.random.default_rng(0)
    f = lambda x, y, z: x * y + y * z + x * z
    x, y, z = (rng.standard_normal(5) for _ in range(3))
    t = o.Tape()
    x_, y_, z_ = o.Leaf(t, x), o.Leaf(t, y), o.Leaf(t, z)
    s = bc(f, x_, y_, z_)
    ds = o.nabla(s, np.ones(5))
    assert (o.unbox(s) == f(x, y, z)).all()
    np.testing.assert_array_equal(ds[x_], y + z)
    np.testing.assert_array_equal(ds[y_], x + z)
    np.testing.assert_array_equal(ds[z_], y + x)


@pytest.mark.parametrize("tA", "NT")
@pytest.mark.parametrize("tB", "NT")
def test_gemm_all_transposes_check_errs(tA, tB):
    # test/sensitivities/linalg/blas.jl:36-47
    rng = np.random.default_rng(123456)
    N = 12
    A, B, VA, VB = (rng.standard_normal((N, N)) for _ in range(4))
    al, val = rng.standard_normal(), rng.standard_normal()
    lam = lambda a, A, B: o.gemm(tA, tB, a, A, B)
    gam = lambda A, B: o.gemm(tA, tB, A, B)
    assert o.check_errs(lam, rng.standard_normal((N, N)), (al, A, B), (val, VA, VB))
    assert o.check_errs(gam, rng.standard_normal((N, N)), (A, B), (VA, VB))


@pytest.mark.parametrize("tA", "NT")
def test_gemv_check_errs(tA):
    # test/sensitivities/linalg/blas.jl:49-61
    rng = np.random.default_rng(3)
    N = 10
    A, VA = rng.standard_normal((N, N)), rng.standard_normal((N, N))
    x, vx = rng.standard_normal(N), rng.standard_normal(N)
    al, val = rng.standard_normal(), rng.standard_normal()
    assert o.check_errs(lambda a, A, x: o.gemv(tA, a, A, x), rng.standard_normal(N), (al, A, x), (val, VA, vx))
    assert o.check_errs(lambda A, x: o.gemv(tA, A, x), rng.standard_normal(N), (A, x), (VA, vx))


def test_reductions_check_errs():
    # test/sensitivities/functional/reducedim.jl:49-67, linalg/generic.jl:47-61,81-87
    rng = np.random.default_rng(5)
    r = lambda *s: rng.standard_normal(s)
    assert o.check_errs(o.mean, r(), r(10), r(10))
    assert o.check_errs(lambda x: o.mean(o.abs2, x), r(), r(10), r(10))
    assert o.check_errs(lambda x: o.mean(x, dims=2), r(10, 1), r(10, 10), r(10, 10))
    assert o.check_errs(lambda x: o.mean(x, dims=(2, 3)), r(10, 1, 1), r(10, 10, 10), r(10, 10, 10))
    assert o.check_errs(o.sum, r(), r(10), r(10))
    assert o.check_errs(lambda x: o.sum(o.abs2, x), r(), r(10), r(10))
    assert o.check_errs(lambda x: o.log(o.sum(o.exp, x)), r(), r(10), r(10))
    assert o.check_errs(o.dot, r(), (r(5), r(5)), (r(5), r(5)))
    assert o.check_errs(o.matmul, r(5, 5), (r(5, 5), r(5, 5)), (r(5, 5), r(5, 5)))
    assert o.check_errs(lambda A, s: A / s, r(5, 5), (r(5, 5), 1.7), (r(5, 5), 0.3))
    assert o.check_errs(lambda s, A: o.ldivide(s, A), r(5, 5), (1.7, r(5, 5)), (0.3, r(5, 5)))


def test_whole_models_check_errs():
    """MLP (examples/mlp.jl:56-82) and VAE (examples/vae.jl:67-90): no reference fixture exists
    (parity unpinned) — checked with the reference's FD method on small sizes."""
    import nabla_jl_b200.workloads as wl
    rng = np.random.default_rng(11)
    W, b, X, Y = wl.mlp_inputs((6, 5, 4, 3), 7, np.float64, seed=0)
    obj = wl.mlp_objective(o, X, Y, 1e-3, 3)
    xs = tuple(W) + tuple(b)
    vs = tuple(rng.standard_normal(p.shape) for p in xs)
    assert o.check_errs(obj, 1.0, xs, vs, eps_abs=1e-8, eps_rel=1e-6)
    params, X, noise = wl.vae_inputs(dx=8, dz=3, dh=6, B=5, seed=3)
    obj = lambda *p: wl.vae_log_joint(o, X, noise, *p, lam=1e-3, scal=60000 / 5, dz=3)
    vs = tuple(rng.standard_normal(p.shape) for p in params)
    assert o.check_errs(obj, 1.0, params, vs, eps_abs=1e-6, eps_rel=1e-6)
    # C3 analytic: Ā = 1 - tanh²(A .+ b), b̄ = Σ_j Ā[:, j]
    A, bb = wl.c3_inputs(9, 6)
    gA, gb = o.nabla(wl.bcast_sum(o))(A, bb)
    ref = 1 - np.tanh(A + bb[:, None]) ** 2
    np.testing.assert_allclose(gA, ref, rtol=1e-14)
    np.testing.assert_allclose(gb, ref.sum(1), rtol=1e-13)


def test_array_structure_and_checkpoint_known_answers():
    # test/sensitivities/array.jl:5-26, test/sensitivities/indexing.jl, test/checkpointing.jl
    rng = np.random.default_rng(0)
    a, b, c = rng.random((3, 2)), rng.random(3), rng.random((3, 3))
    g = o.nabla(lambda a, b, c: o.sum(o.hcat(2 * a, 3 * b, 4 * c)))(a, b, c)
    assert all((x == v).all() and x.shape == s for x, v, s in zip(g, (2, 3, 4), ((3, 2), (3,), (3, 3))))
    a, b, c = rng.random((2, 4)), rng.random((1, 4)), rng.random((3, 4))
    g = o.nabla(lambda a, b, c: o.sum(o.vcat(2 * a, 3 * b, 4 * c)))(a, b, c)
    assert all((x == v).all() and x.shape == s for x, v, s in zip(g, (2, 3, 4), ((2, 4), (1, 4), (3, 4))))
    leaf = o.Leaf(o.Tape(), 5.0 * np.ones(5))
    y = o.getindex(leaf, 1)
    assert o.unbox(y) == 5 and (o.nabla(y, 1.0)[leaf] == [1, 0, 0, 0, 0]).all()
    x = o.Leaf(o.Tape(), 10.0 * np.ones(3))
    y = o.getindex(x, o.R(2, 3))
    assert (o.unbox(y) == [10, 10]).all() and (o.nabla(y, np.ones(2))[x] == [0, 1, 1]).all()
    x = o.Leaf(o.Tape(), 10.0 * np.ones(3))
    y = o.getindex(x, [2, 3, 3])          # overlapping indices (#139)
    assert (o.unbox(y) == [10, 10, 10]).all() and (o.nabla(y, np.ones(3))[x] == [0, 1, 2]).all()
    xr = rng.standard_normal((2, 10))
    assert o.check_errs(lambda v: o.reshape(v, 5, 4), rng.standard_normal((5, 4)), xr, rng.standard_normal((2, 10)))
    assert o.check_errs(lambda v: o.reshape(v, (5, 4)), rng.standard_normal((5, 4)), xr, rng.standard_normal((2, 10)))
    assert o.check_errs(lambda v: o.fill(v, 4, 4), rng.standard_normal((4, 4)), rng.standard_normal(), rng.standard_normal())
    assert o.checkpoint(o.sin, (5.0,)) == math.sin(5.0)
    assert o.nabla(lambda v: o.checkpoint(o.sin, (v,)))(5.0) == (math.cos(5.0),)
    assert o.nabla(lambda v: o.cos(o.checkpoint(o.sin, (v,))))(5.0) == o.nabla(lambda v: o.cos(o.sin(v)))(5.0)
    yy = 5.0
    with pytest.raises(RuntimeError):
        o.checkpoint(lambda v: v + yy, (5.0,))


def test_structured_blas_check_errs():
    # test/sensitivities/linalg/blas.jl:63-135 — every flag permutation, the reference's FD method
    import itertools
    rng = np.random.default_rng(123456)
    r = lambda *s: rng.standard_normal(s)
    N = 10
    for side, ul in itertools.product("LR", "LU"):
        A, B, VA, VB = r(N, N), r(N, N), r(N, N), r(N, N)
        assert o.check_errs(lambda a, A, B: o.symm(side, ul, a, A, B), r(N, N), (r(), A, B), (r(), VA, VB))
        assert o.check_errs(lambda A, B: o.symm(side, ul, A, B), r(N, N), (A, B), (VA, VB))
    for ul in "LU":
        A, VA, x, vx = r(N, N), r(N, N), r(N), r(N)
        assert o.check_errs(lambda a, A, x: o.symv(ul, a, A, x), r(N), (r(), A, x), (r(), VA, vx))
        assert o.check_errs(lambda A, x: o.symv(ul, A, x), r(N), (A, x), (VA, vx))
    for side, ul, tA, dA in itertools.product("LR", "LU", "NT", "UN"):
        A, B, VA, VB = r(N, N), r(N, N), r(N, N), r(N, N)
        assert o.check_errs(lambda a, A, B: o.trmm(side, ul, tA, dA, a, A, B), r(N, N), (r(), A, B), (r(), VA, VB))
        X, VX = r(N, N), r(N, N)
        A3 = r(N, N) + 3 * np.eye(N)
        assert o.check_errs(lambda a, A, X: o.trsm(side, ul, tA, dA, a, A, X), r(N, N), (r(), A3, X), (r(), VA, VX),
                            eps_abs=1e-7, eps_rel=1e-5)
    for ul, tA, dA in itertools.product("LU", "NT", "UN"):
        A, VA, b, vb = r(N, N), r(N, N), r(N), r(N)
        assert o.check_errs(lambda A, b: o.trmv(ul, tA, dA, A, b), r(N), (A, b), (VA, vb))
        A1 = r(N, N) + np.eye(N)
        A1 = A1.T @ A1
        assert o.check_errs(lambda A, x: o.trsv(ul, tA, dA, A, x), r(N), (A1, b), (VA, vb), eps_abs=1e-7, eps_rel=1e-5)


def test_generic_linalg_kron_plus_I_copy():
    """src/sensitivities/linalg/generic.jl:7-51 — kron (the reference's index loops restated literally and compared
    with the oracle's contraction), A + λI, copy; each with the reference's own check (test/sensitivities/linalg/
    generic.jl:64-79: check_errs = allocated + in-place variants against 5-point finite differences)."""
    import nabla_oracle as o
    rng = np.random.default_rng(3)

    def loops(n, Yb, A, B):   # generic.jl:19-40, 1-based column-major walk of m = length(Y) .. 1
        (I, J), (K, L) = A.shape, B.shape
        m, Yf = Yb.size, Yb.reshape(-1, order="F")
        Ab, Bb = np.zeros_like(A), np.zeros_like(B)
        for j in reversed(range(J)):
            for l in reversed(range(L)):
                for i in reversed(range(I)):
                    for k in reversed(range(K)):
                        if n == 1:
                            Ab[i, j] += Yf[m - 1] * B[k, l]
                        else:
                            Bb[k, l] += Yf[m - 1] * A[i, j]
                        m -= 1
        return Ab if n == 1 else Bb

    for _ in range(5):
        A, B = rng.standard_normal((3, 4)), rng.standard_normal((2, 5))
        VA, VB = rng.standard_normal((3, 4)), rng.standard_normal((2, 5))
        Yb = rng.standard_normal((6, 20))
        np.testing.assert_allclose(o._grad_kron(1, None, None, Yb, A, B), loops(1, Yb, A, B), rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(o._grad_kron(2, None, None, Yb, A, B), loops(2, Yb, A, B), rtol=1e-13, atol=1e-13)
        assert o.check_errs(o.kron, Yb, (A, B), (VA, VB))
    for _ in range(5):
        A, VA = rng.standard_normal((5, 5)), rng.standard_normal((5, 5))
        assert o.check_errs(lambda X: o.plus_I(X, 0.65), VA, A, 1e-1 * rng.standard_normal((5, 5)))
        assert o.check_errs(o.copy, VA, A, 1e-1 * rng.standard_normal((5, 5)))
    # Ā of A + λI is Ȳ itself (generic.jl:43)
    t = o.Tape()
    a = o.Leaf(t, rng.standard_normal((4, 4)))
    yb = rng.standard_normal((4, 4))
    assert o.nabla(o.plus_I(a, 2.0), yb)[a] is not None
    np.testing.assert_array_equal(o.nabla(o.plus_I(a, 2.0), yb)[a], yb)
