"""Parity at the BENCHMARKED configurations (BASELINE.json configs[0..4]) — the shapes bench.py times, not scaled
down: the wide MLP at width 8192 (K = 8192 chains through four hidden layers; Float32 on the default BF16x3 tensor
path and Float64 on the Ozaki int8 split) against the Float64 oracle on a 512-column batch (the oracle's cost is
linear in the batch; the products keep their K), C1 (examples/mlp.jl) and C4 (examples/vae.jl) at B = 4096, and the
C2 quadratic forms at N = 4096 against the analytic identity of docs/src/index.md:25-39.

Tolerances are north_star's: rtol 1e-10 Float64, 1e-4 Float32 (Float32 device results against the oracle evaluated
in Float64 on the same Float32-representable inputs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WIDE = (784, 8192, 8192, 8192, 8192, 10)


def _np(x):
    import nabla_jl_b200 as nb
    x = nb.unthunk(nb.unbox(x))
    return x.numpy() if hasattr(x, "numpy") else np.asarray(x)


def _rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape
    return float(np.linalg.norm((got - ref).ravel())) / max(float(np.linalg.norm(ref.ravel())), 1e-300)


def _wide_inputs(B, dtype):
    import nabla_jl_b200.workloads as wl
    return wl.mlp_inputs(WIDE, B, dtype, seed=4)


def _oracle_wide(orc, W, b, X, Y):
    import nabla_jl_b200.workloads as wl
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    y, g = orc.nabla(wl.mlp_objective(orc, f64(X), f64(Y), 1e-3, 5), get_output=True)(*[f64(p) for p in (*W, *b)])
    return float(orc.unbox(y)), g


@pytest.fixture(scope="module")
def wide_f32_case(orc):
    """Inputs are drawn in Float32 and the oracle runs in Float64 on exactly those values; shared by the Float32
    test and (re-used as Float64 inputs) by the Float64 test, so the 1.7 GB oracle evaluation runs once."""
    W, b, X, Y = _wide_inputs(512, np.float32)
    yr, gr = _oracle_wide(orc, W, b, X, Y)
    return W, b, X, Y, yr, gr


def test_wide_mlp_8192_f32_bf16x3_vs_oracle(nb, ctx, wide_f32_case):
    """BASELINE configs[4] at its own width, Float32, NB_MATH_DEFAULT (BF16x3 on tcgen05, fused epilogues, the CTA-pair
    kernel for the 8192 x 512 x 8192 products' weight gradients 8192 x 8192 x 512)."""
    import nabla_jl_b200.workloads as wl
    W, b, X, Y, yr, gr = wide_f32_case
    assert ctx.math_mode == nb.MATH_DEFAULT
    Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
    y, g = nb.nabla(wl.mlp_objective(nb, Xd, Yd, 1e-3, 5), get_output=True)(*W, *b)
    np.testing.assert_allclose(float(_np(y)), yr, rtol=1e-4)
    errs = [_rel(_np(a), r) for a, r in zip(g, gr)]
    assert max(errs) <= 1e-4, f"per-parameter relative errors {errs}"


def test_wide_mlp_8192_f32_sharded_drivers_agree_with_oracle(nb, ctx, wide_f32_case):
    """The drivers bench.py times (data term + shared prior term, sweep.OverlappedShardedGradient on one rank) on the
    same case: identical tolerance against the oracle."""
    import nabla_jl_b200.workloads as wl
    W, b, X, Y, yr, gr = wide_f32_case
    Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
    params = [ctx.asarray(p) for p in (*W, *b)]
    comm = nb.Communicator(ctx, 0, 1, None)
    sg = nb.OverlappedShardedGradient(wl.mlp_data_objective(nb, Xd, Yd, 5), wl.mlp_prior_objective(nb, 1e-3, 5),
                                      params, comm)
    y, g = sg.step()
    np.testing.assert_allclose(float(_np(y)), yr, rtol=1e-4)
    errs = [_rel(_np(a), r) for a, r in zip(g, gr)]
    assert max(errs) <= 1e-4, f"per-parameter relative errors {errs}"


def test_wide_mlp_8192_f64_ozaki_vs_oracle(nb, ctx, orc, wide_f32_case):
    """The same configuration in Float64 (every reference example is Float64): the 8192-deep products run as exact
    int8 slice products on the tensor cores (csrc/ozaki.cu); rtol 1e-10 on every parameter sensitivity."""
    import nabla_jl_b200.workloads as wl
    W, b, X, Y, yr, gr = wide_f32_case
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    Xd, Yd = ctx.asarray(f64(X)), ctx.asarray(f64(Y))
    y, g = nb.nabla(wl.mlp_objective(nb, Xd, Yd, 1e-3, 5), get_output=True)(*[f64(p) for p in (*W, *b)])
    np.testing.assert_allclose(float(_np(y)), yr, rtol=1e-10)
    errs = [_rel(_np(a), r) for a, r in zip(g, gr)]
    assert max(errs) <= 1e-10, f"per-parameter relative errors {errs}"


def test_c1_mlp_B4096_f64_vs_oracle(nb, ctx, orc):
    """C1 (examples/mlp.jl:56-82, 784-500-300-10) at the benchmarked batch 4096, eager and as a CUDA-graph replay."""
    import nabla_jl_b200.workloads as wl
    dims, B = (784, 500, 300, 10), 4096
    W, b, X, Y = wl.mlp_inputs(dims, B, np.float64, seed=0)
    yr, gr = orc.nabla(wl.mlp_objective(orc, X, Y, 1e-3, 3), get_output=True)(*W, *b)
    Xd, Yd = ctx.asarray(X), ctx.asarray(Y)
    obj = wl.mlp_objective(nb, Xd, Yd, 1e-3, 3)
    y, g = nb.nabla(obj, get_output=True)(*W, *b)
    np.testing.assert_allclose(float(_np(y)), float(orc.unbox(yr)), rtol=1e-10)
    for a, r in zip(g, gr):
        assert _rel(_np(a), r) <= 1e-10
    cap = nb.CapturedGradient(obj, [ctx.asarray(p) for p in (*W, *b)], ctx)
    y2, g2 = cap()
    ctx.sync()
    np.testing.assert_allclose(float(_np(y2)), float(orc.unbox(yr)), rtol=1e-10)
    for a, r in zip(g2, gr):
        assert _rel(_np(a), r) <= 1e-10
    cap.close()


def test_c4_vae_B4096_f64_vs_oracle(nb, ctx, orc):
    """C4 (examples/vae.jl:67-90; dx 784, dz 10, dh 500) at batch 4096, eager and as a CUDA-graph replay."""
    import nabla_jl_b200.workloads as wl
    params, X, noise = wl.vae_inputs(B=4096, seed=3)
    kw = dict(lam=1e-3, scal=60000 / 4096, dz=10)
    yr, gr = orc.nabla(wl.vae_objective(orc, X, noise, **kw), get_output=True)(*params)
    Xd, Nd = ctx.asarray(X), ctx.asarray(noise)
    obj = wl.vae_objective(nb, Xd, Nd, **kw)
    y, g = nb.nabla(obj, get_output=True)(*params)
    np.testing.assert_allclose(float(_np(y)), float(orc.unbox(yr)), rtol=1e-10)
    for a, r in zip(g, gr):
        assert _rel(_np(a), r) <= 1e-10
    cap = nb.CapturedGradient(obj, [ctx.asarray(p) for p in params], ctx)
    y2, g2 = cap()
    ctx.sync()
    for a, r in zip(g2, gr):
        assert _rel(_np(a), r) <= 1e-10
    cap.close()


def test_c2_quadratic_forms_N4096_analytic(nb, ctx):
    """C2 at N = 4096 against the identity the reference documents (docs/src/index.md:25-39): ∇(x'Ax) = (A + A')x,
    ∇(x'Ay) = (Ay, A'x); the matrix form quad(A, B) = B'AB with Ȳ = I (test/core.jl:206-218):
    Ā = B B', B̄ = (A + A') B."""
    import nabla_jl_b200.workloads as wl
    N = 4096
    A, x, y = wl.quad_inputs(N)
    Ad = ctx.asarray(A)
    (g,) = nb.nabla(wl.quad_symmetric(nb, Ad))(x)
    assert _rel(_np(g), (A + A.T) @ x) <= 1e-10
    gx, gy = nb.nabla(wl.quad_asymmetric(nb, Ad))(x, y)
    assert _rel(_np(gx), A @ y) <= 1e-10
    assert _rel(_np(gy), A.T @ x) <= 1e-10
    Bm = np.random.default_rng(5).standard_normal((N, N))
    t = nb.Tape()
    A_, B_ = nb.Leaf(t, Ad), nb.Leaf(t, Bm)
    Q = nb.matmul(nb.adjoint(B_), nb.matmul(A_, B_))
    rt = nb.nabla(Q, np.eye(N))
    gA, gB = _np(rt[A_]), _np(rt[B_])
    assert _rel(gA, Bm @ Bm.T) <= 1e-10
    assert _rel(gB, (A + A.T) @ Bm) <= 1e-10
