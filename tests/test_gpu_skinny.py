"""The skinny-shape product kernels (csrc/gemm_skinny.cu) through the C ABI against NumPy Float64: every
transpose, small M / small N / small K, alpha and beta, ragged sizes and padded leading dimensions, and the
fused epilogues of nb_gemm_fused (bias + activation, `.* f'(y)`, row sums).  These are the output-layer products of
examples/mlp.jl (10 x B x 8192 and transposes) and the matrix*vector products of the quadratic examples."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.linalg.norm((np.asarray(got, np.float64) - ref).ravel())) / max(float(np.linalg.norm(ref.ravel())), 1e-300)


def _pad(a, ld):
    """Column-major copy of `a` with leading dimension ld >= rows (rows beyond are NaN: never read)."""
    out = np.full((ld, a.shape[1]), np.nan, dtype=a.dtype, order="F")
    out[:a.shape[0]] = a
    return out


SHAPES = [  # (m, n, k)
    (10, 4096, 1000), (10, 515, 777), (1, 4097, 513), (16, 2048, 2048), (3, 33, 5000),          # small M
    (4096, 10, 1000), (777, 3, 515), (4097, 1, 513), (2048, 16, 2048),                          # small N
    (4096, 512, 10), (1030, 257, 3), (515, 300, 16), (2048, 64, 1), (129, 131, 7),              # small K
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("tA,tB", [("N", "N"), ("N", "T"), ("T", "N"), ("T", "T")])
def test_skinny_products_all_transposes(nb, ctx, dtype, tA, tB):
    rng = np.random.default_rng(11)
    tol = 2e-6 if dtype == np.float32 else 1e-13
    for (m, n, k) in SHAPES:
        for pad, alpha, beta in ((0, 1.0, 0.0), (3, -0.7, 1.0), (4, 2.5, 0.5)):
            A = rng.standard_normal((k, m) if tA == "T" else (m, k)).astype(dtype)
            B = rng.standard_normal((n, k) if tB == "T" else (k, n)).astype(dtype)
            C0 = rng.standard_normal((m, n)).astype(dtype)
            lda, ldb, ldc = A.shape[0] + pad, B.shape[0] + pad, m + pad
            Ad, Bd, Cd = ctx.asarray(_pad(A, lda)), ctx.asarray(_pad(B, ldb)), ctx.asarray(_pad(C0, ldc))
            nb._lib.check(ctx.L.nb_gemm(ctx.h, tA.encode(), tB.encode(), m, n, k, alpha, Ad.ptr, lda, Bd.ptr, ldb, beta,
                                        Cd.ptr, ldc, Cd.nb_dtype, nb.MATH_DEFAULT), ctx.h)
            got = Cd.numpy()
            opA = (A.T if tA == "T" else A).astype(np.float64)
            opB = (B.T if tB == "T" else B).astype(np.float64)
            ref = alpha * (opA @ opB) + beta * C0.astype(np.float64)
            assert _rel(got[:m], ref) <= tol, (m, n, k, pad, alpha, beta)
            if pad:
                assert np.all(np.isnan(got[m:])), "rows beyond m must not be touched"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_skinny_fused_epilogues(nb, ctx, dtype):
    from nabla_jl_b200.fusion import nb_gemm_epilogue
    rng = np.random.default_rng(12)
    tol = 5e-6 if dtype == np.float32 else 1e-12
    ops = nb._lib.OPCODES
    for (m, n, k, tA, tB) in ((10, 4096, 1000, "N", "N"), (2048, 515, 10, "T", "N"), (4100, 1, 600, "N", "N"),
                              (1027, 300, 10, "T", "N")):
        A = (0.1 * rng.standard_normal((k, m) if tA == "T" else (m, k))).astype(dtype)
        B = rng.standard_normal((n, k) if tB == "T" else (k, n)).astype(dtype)
        bias = rng.standard_normal(m).astype(dtype)
        ysaved = np.tanh(rng.standard_normal((m, n))).astype(dtype)
        Ad, Bd, bd, yd = ctx.asarray(A), ctx.asarray(B), ctx.asarray(bias), ctx.asarray(ysaved)
        P = (A.T if tA == "T" else A).astype(np.float64) @ (B.T if tB == "T" else B).astype(np.float64)
        for kind, op, want_rowsum in ((1, "tanh", False), (1, "logistic", True), (2, "tanh", True), (2, "logistic", False),
                                      (1, "identity", True)):
            small_out = min(m, n) <= 16 and k > 16
            if want_rowsum and small_out:
                continue   # skinny outputs leave row sums to the general path (covered by test_gpu_parity)
            out = ctx.empty((m, n), dtype)
            rs = ctx.asarray(np.ones(m, dtype=dtype))
            epi = nb_gemm_epilogue()
            epi.kind, epi.op = kind, ops[op]
            epi.bias = bd.ptr if kind == 1 else None
            epi.y, epi.ldy = (yd.ptr, m) if kind == 2 else (None, 0)
            epi.rowsum = rs.ptr if want_rowsum else None
            epi.rowsum_acc = 1 if want_rowsum else 0
            epi.emit_split = 1
            nb._lib.check(ctx.L.nb_gemm_fused(ctx.h, tA.encode(), tB.encode(), m, n, k, 1.0, Ad.ptr, A.shape[0], Bd.ptr,
                                              B.shape[0], out.ptr, m, out.nb_dtype, nb.MATH_DEFAULT, C.byref(epi)), ctx.h)
            if kind == 1:
                z = P + bias.astype(np.float64)[:, None]
                ref = {"tanh": np.tanh(z), "logistic": 1 / (1 + np.exp(-z)), "identity": z}[op]
            else:
                ys = ysaved.astype(np.float64)
                ref = P * ({"tanh": 1 - ys * ys, "logistic": ys * (1 - ys)}[op])
            assert _rel(out.numpy(), ref) <= tol, (m, n, k, kind, op)
            if want_rowsum:
                assert _rel(rs.numpy(), 1.0 + ref.sum(axis=1)) <= tol * 10, (m, n, k, kind, op, "rowsum")
            # the emitted operand split must describe `out`: multiply it again on the tensor cores and compare
            if dtype == np.float32 and m * n >= 1 << 16 and m >= 128:
                E = ctx.asarray(np.eye(n, 128, dtype=np.float32))
                chk = ctx.empty((m, 128), np.float32)
                nb._lib.check(ctx.L.nb_gemm(ctx.h, b"N", b"N", m, 128, n, 1.0, out.ptr, m, E.ptr, n, 0.0, chk.ptr, m,
                                            chk.nb_dtype, nb.MATH_BF16X3), ctx.h)
                assert _rel(chk.numpy(), ref[:, :128]) <= 2e-5, "emitted split"


def test_gemv_ger_padded_leading_dimension(nb, ctx):
    """nb_gemv / nb_ger with lda > m (sub-matrix views), both dtypes, against NumPy."""
    rng = np.random.default_rng(13)
    for dtype, tol in ((np.float32, 2e-6), (np.float64, 1e-13)):
        m, n, lda = 1500, 700, 1504
        A = rng.standard_normal((m, n)).astype(dtype)
        x, xt = rng.standard_normal(n).astype(dtype), rng.standard_normal(m).astype(dtype)
        y0, y0t = rng.standard_normal(m).astype(dtype), rng.standard_normal(n).astype(dtype)
        Ad = ctx.asarray(_pad(A, lda))
        for tr, xin, yin in (("N", x, y0), ("T", xt, y0t)):
            yd = ctx.asarray(yin)
            nb._lib.check(ctx.L.nb_gemv(ctx.h, tr.encode(), m, n, 1.5, Ad.ptr, lda, ctx.asarray(xin).ptr, 0.5, yd.ptr,
                                        yd.nb_dtype), ctx.h)
            ref = 1.5 * ((A.T if tr == "T" else A).astype(np.float64) @ xin.astype(np.float64)) + 0.5 * yin.astype(np.float64)
            assert _rel(yd.numpy(), ref) <= tol
        G = ctx.asarray(_pad(A, lda))
        nb._lib.check(ctx.L.nb_ger(ctx.h, m, n, -2.0, ctx.asarray(xt).ptr, ctx.asarray(x).ptr, 1.0, G.ptr, lda, G.nb_dtype), ctx.h)
        ref = A.astype(np.float64) - 2.0 * np.outer(xt.astype(np.float64), x.astype(np.float64))
        assert _rel(G.numpy()[:m], ref) <= tol
