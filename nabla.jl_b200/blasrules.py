"""Symmetric / triangular BLAS sensitivities on the device: ``symm``, ``symv``, ``trmm``, ``trmv``,
``trsm``, ``trsv`` (src/sensitivities/linalg/blas.jl:60-328).

The structured operand is materialised once by ``nb_tri`` (symmetric matrix from a triangle, a
triangle with a unit / zeroed diagonal) and the products run on the dense GEMM / GEMV kernels; the
pullbacks follow the reference formulas line by line (``tmp + tmp' - Diagonal(tmp)`` then
``tril!/triu!``; transposed products; triangular solves with ``alpha = -1``).  Solves are ``nb_trsm``."""
from . import kernels as K
from ._lib import DimensionMismatch, check
from .core import Primitive, as_device
from .device import DevArray

_TRI = {"L": 0, "U": 1}
_SYM = {"L": 2, "U": 3}
_FOLD = {"L": 4, "U": 5}


def _U(c):
    return str(c).upper()


def _flip(ta):
    return "T" if _U(ta) == "N" else "N"


def _dev(x):
    x = K.materialize(x)
    return x if isinstance(x, DevArray) else as_device(x)


def _alpha(a):
    a = K.materialize(a) if not isinstance(a, (int, float)) else a
    return float(a.item()) if isinstance(a, DevArray) else float(a)


def _tri_op(op, diag, A):
    n = A.shape[0]
    if A.ndim != 2 or A.shape[1] != n:
        raise DimensionMismatch(6, f"matrix is not square: dimensions are {A.shape}")
    out = A.ctx.empty((n, n), A.dtype)
    check(A.ctx.L.nb_tri(A.ctx.h, op, diag, n, A.ptr, n, out.ptr, n, A.nb_dtype), A.ctx.h)
    return out


def _sym(ul, A):
    return _tri_op(_SYM[_U(ul)], 0, A)


def _tri(ul, dA, A):
    return _tri_op(_TRI[_U(ul)], 2 if _U(dA) == "U" else 0, A)


def _g(ul, M, dA="N"):
    return _tri_op(_TRI[_U(ul)], 1 if _U(dA) == "U" else 0, M)


def _fold(ul, M):
    return _tri_op(_FOLD[_U(ul)], 0, M)


def _as2d(x):
    return x.reshape((x.shape[0], 1)) if x.ndim == 1 else x


def _mm(tA, tB, alpha, A, B):
    A, B = _as2d(A), _as2d(B)
    m, k = (A.shape[1], A.shape[0]) if tA == "T" else A.shape
    k2, n = (B.shape[1], B.shape[0]) if tB == "T" else B.shape
    if k != k2:
        raise DimensionMismatch(6, f"matrix product of {A.shape} ({tA}) and {B.shape} ({tB})")
    out = A.ctx.empty((m, n), A.dtype)
    return K.gemm(tA, tB, alpha, A, B, 0.0, out, m, n, k)


def _like(out, ref):
    return out.reshape(ref.shape) if out.shape != ref.shape else out


def _dot_over_alpha(Yb, Y, alpha):
    out = Y.ctx.empty((), Y.dtype)
    K.bcast_reduce(_MUL, [_dev(Yb), Y], out, scale=1.0 / alpha)   # sum(Ȳ .* Y) / α
    return out


from .scalar import CATALOGUE, compile_scalar  # noqa: E402

_MUL = compile_scalar(CATALOGUE["mul"], 2)


def _scaled(alpha, x):
    if alpha == 1.0:
        return x
    out = x.ctx.empty(x.shape, x.dtype)
    K.bcast_reduce(None, [x], out, scale=alpha)
    return out


# ---- symm / symv ---------------------------------------------------------------------------------
def _split(rest):
    return (rest[0], rest[1], rest[2], True) if len(rest) == 3 else (1.0, rest[0], rest[1], False)


def _symm_raw(side, ul, alpha, A, B):
    S = _sym(ul, A)
    out = _mm("N", "N", alpha, S, B) if _U(side) == "L" else _mm("N", "N", alpha, B, S)
    return _like(out, B)


def _symm_forward(side, ul, *rest):
    alpha, A, B, _ = _split(rest)
    return _symm_raw(side, ul, _alpha(alpha), _dev(A), _dev(B))


_symm = Primitive("symm", _symm_forward)


def _symm_grad(n):
    def g(p, Y, Yb, side, ul, *rest):
        alpha, A, B, has_alpha = _split(rest)
        alpha, A, B, Yb = _alpha(alpha), _dev(A), _dev(B), _dev(Yb)
        if has_alpha and n == 3:
            return _dot_over_alpha(Yb, Y, alpha)                                  # blas.jl:65-71
        k = n - 3 - (1 if has_alpha else 0)
        if k == 0:                                                                # blas.jl:72-81
            tmp = _mm("N", "T", 1.0, Yb, B) if _U(side) == "L" else _mm("T", "N", 1.0, B, Yb)
            return _scaled(alpha, _fold(ul, tmp))
        return _symm_raw(side, ul, alpha, A, Yb)                                  # blas.jl:95-101
    return g


for _n in (3, 4, 5):
    _symm.grads[_n] = _symm_grad(_n)


def symm(side, ul, *rest):
    """``BLAS.symm(side, ul, [α,] A, B)``."""
    return _symm(side, ul, *rest)


def _symv_forward(ul, *rest):
    alpha, A, x, _ = _split(rest)
    A, x = _dev(A), _dev(x)
    n = A.shape[0]
    out = A.ctx.empty((n,), A.dtype)
    return K.gemv("N", _alpha(alpha), _sym(ul, A), x, 0.0, out, n, n)


_symv = Primitive("symv", _symv_forward)


def _symv_grad(n):
    def g(p, y, yb, ul, *rest):
        alpha, A, x, has_alpha = _split(rest)
        alpha, A, x, yb = _alpha(alpha), _dev(A), _dev(x), _dev(yb)
        if has_alpha and n == 2:
            return _dot_over_alpha(yb, y, alpha)                                  # blas.jl:146-151
        k = n - 2 - (1 if has_alpha else 0)
        if k == 0:                                                                # symm rule, side 'L'
            return _scaled(alpha, _fold(ul, _mm("N", "T", 1.0, yb, x)))
        return _symv_forward(ul, alpha, A, yb)
    return g


for _n in (2, 3, 4):
    _symv.grads[_n] = _symv_grad(_n)


def symv(ul, *rest):
    """``BLAS.symv(ul, [α,] A, x)``."""
    return _symv(ul, *rest)


# ---- trmm / trmv ---------------------------------------------------------------------------------
def _trmm_raw(side, ul, ta, dA, alpha, A, B):
    T = _tri(ul, dA, A)
    t = "T" if _U(ta) != "N" else "N"
    out = _mm(t, "N", alpha, T, B) if _U(side) == "L" else _mm("N", t, alpha, B, T)
    return _like(out, B)


_trmm = Primitive("trmm", lambda side, ul, ta, dA, alpha, A, B: _trmm_raw(side, ul, ta, dA, _alpha(alpha), _dev(A), _dev(B)))


def _trmm_grad(n):
    def g(p, Y, Yb, side, ul, ta, dA, alpha, A, B):
        alpha, A, B, Yb = _alpha(alpha), _dev(A), _dev(B), _dev(Yb)
        if n == 5:
            return _dot_over_alpha(Yb, Y, alpha)
        if n == 6:                                                                # blas.jl:217-231
            if _U(side) == "L":
                full = _mm("N", "T", alpha, Yb, B) if _U(ta) == "N" else _mm("N", "T", alpha, B, Yb)
            else:
                full = _mm("T", "N", alpha, B, Yb) if _U(ta) == "N" else _mm("T", "N", alpha, Yb, B)
            return _g(ul, full, dA)
        return _trmm_raw(side, ul, _flip(ta), dA, alpha, A, Yb)                   # blas.jl:232-244
    return g


for _n in (5, 6, 7):
    _trmm.grads[_n] = _trmm_grad(_n)


def trmm(side, ul, ta, dA, alpha, A, B):
    """``BLAS.trmm(side, ul, tA, dA, α, A, B)``."""
    return _trmm(side, ul, ta, dA, alpha, A, B)


def _trmv_raw(ul, ta, dA, A, b):
    n = A.shape[0]
    out = A.ctx.empty((n,), A.dtype)
    return K.gemv("T" if _U(ta) != "N" else "N", 1.0, _tri(ul, dA, A), b, 0.0, out, n, n)


_trmv = Primitive("trmv", lambda ul, ta, dA, A, b: _trmv_raw(ul, ta, dA, _dev(A), _dev(b)))


@_trmv.grad(4)
def _(p, y, yb, ul, ta, dA, A, b):                                                # blas.jl:252-260
    yb, b = _dev(yb), _dev(b)
    outer = _mm("N", "T", 1.0, yb, b) if _U(ta) == "N" else _mm("N", "T", 1.0, b, yb)
    return _g(ul, outer, dA)


@_trmv.grad(5)
def _(p, y, yb, ul, ta, dA, A, b):
    return _trmv_raw(ul, _flip(ta), dA, _dev(A), _dev(yb))


def trmv(ul, ta, dA, A, b):
    """``BLAS.trmv(ul, tA, dA, A, b)``."""
    return _trmv(ul, ta, dA, A, b)


# ---- trsm / trsv ---------------------------------------------------------------------------------
def _trsm_raw(side, ul, ta, dA, alpha, A, X):
    """α op(A)⁻¹ X (side 'L') or α X op(A)⁻¹ (side 'R'); X is copied, the solve runs in place."""
    X2 = _as2d(X)
    m, n = X2.shape
    na = m if _U(side) == "L" else n
    if A.shape != (na, na):
        raise DimensionMismatch(6, f"trsm: A has dimensions {A.shape}, X has dimensions {X.shape} (side {side})")
    out = X.copy()
    check(A.ctx.L.nb_trsm(A.ctx.h, _U(side).encode(), _U(ul).encode(), _U(ta).encode(), _U(dA).encode(), m, n,
                          float(alpha), A.ptr, na, out.ptr, m, A.nb_dtype), A.ctx.h)
    return out


_trsm = Primitive("trsm", lambda side, ul, ta, dA, alpha, A, X: _trsm_raw(side, ul, ta, dA, _alpha(alpha), _dev(A), _dev(X)))


def _trsm_abar(Y, Yb, side, ul, ta, dA, A):
    """blas.jl:280-294 — the sensitivity of the triangular factor."""
    Y, Yb = _as2d(Y), _as2d(Yb)
    if _U(side) == "L":
        full = (_trsm_raw("L", ul, "T", dA, -1.0, A, _mm("N", "T", 1.0, Yb, Y)) if _U(ta) == "N"
                else _trsm_raw("R", ul, "T", dA, -1.0, A, _mm("N", "T", 1.0, Y, Yb)))
    else:
        full = (_trsm_raw("R", ul, "T", dA, -1.0, A, _mm("T", "N", 1.0, Y, Yb)) if _U(ta) == "N"
                else _trsm_raw("L", ul, "T", dA, -1.0, A, _mm("T", "N", 1.0, Yb, Y)))
    return _g(ul, full, dA)


def _trsm_grad(n):
    def g(p, Y, Yb, side, ul, ta, dA, alpha, A, X):
        alpha, A, Yb = _alpha(alpha), _dev(A), _dev(Yb)
        if n == 5:
            return _dot_over_alpha(Yb, Y, alpha)
        if n == 6:
            return _trsm_abar(Y, Yb, side, ul, ta, dA, A)
        return _trsm_raw(side, ul, _flip(ta), dA, alpha, A, Yb)                   # blas.jl:295-307
    return g


for _n in (5, 6, 7):
    _trsm.grads[_n] = _trsm_grad(_n)


def trsm(side, ul, ta, dA, alpha, A, X):
    """``BLAS.trsm(side, ul, tA, dA, α, A, X)``."""
    return _trsm(side, ul, ta, dA, alpha, A, X)


_trsv = Primitive("trsv", lambda ul, ta, dA, A, x: _trsm_raw("L", ul, ta, dA, 1.0, _dev(A), _dev(x)))


@_trsv.grad(4)
def _(p, y, yb, ul, ta, dA, A, x):                                                # blas.jl:315-323
    return _trsm_abar(y, _dev(yb), "L", ul, ta, dA, _dev(A))


@_trsv.grad(5)
def _(p, y, yb, ul, ta, dA, A, x):
    return _trsm_raw("L", ul, _flip(ta), dA, 1.0, _dev(A), _dev(yb))


def trsv(ul, ta, dA, A, x):
    """``BLAS.trsv(ul, tA, dA, A, x)``."""
    return _trsv(ul, ta, dA, A, x)
