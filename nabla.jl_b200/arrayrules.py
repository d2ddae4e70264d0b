"""Array-structure rules on the device: ``hcat`` / ``vcat`` / ``getindex`` / ``reshape`` / ``fill`` and
gradient checkpointing.

Reference: src/sensitivities/chainrules_ignored.jl:12-40 (hcat / vcat pullbacks are
``copy(selectdim(ȳ, …))``), src/sensitivities/indexing.jl:3-11 + ChainRules ``rrule(getindex)``
(x̄ = zero(x); x̄[inds...] .+= ȳ, duplicates accumulate: test/sensitivities/indexing.jl #139),
ChainRules ``reshape`` / ``fill`` rules (test/sensitivities/array.jl:5-26), src/checkpointing.jl:10-54.
All data movement is ``nb_copy2d`` / ``nb_gather`` / ``nb_scatter_add`` (csrc/array.cu)."""
import ctypes as C
import numbers

import numpy as np

from . import kernels as K
from ._lib import DimensionMismatch, check
from .core import Leaf, Node, Primitive, Tape, as_device, nabla, unbox, unthunk, zeroslike
from .device import DevArray


class R:
    """Julia's ``lo:hi`` (1-based, inclusive)."""

    def __init__(self, lo, hi):
        self.lo, self.hi = int(lo), int(hi)


colon = R(1, -1)  # ``:``


def _copy2d(rows, cols, src, src_off, ld_src, dst, dst_off, ld_dst, accumulate=False):
    ctx = dst.ctx
    check(ctx.L.nb_copy2d(ctx.h, rows, cols, src.ptr, src_off, ld_src, dst.ptr, dst_off, ld_dst, dst.nb_dtype,
                          1 if accumulate else 0), ctx.h)


def _dev(x):
    x = K.materialize(x)
    return x if isinstance(x, DevArray) else as_device(x)


def _rc(a):
    """(rows, cols) of a vector or matrix viewed as a column-major matrix."""
    if a.ndim == 1:
        return a.shape[0], 1
    if a.ndim == 2:
        return a.shape
    raise TypeError("MethodError: hcat/vcat are implemented for vectors and matrices")


# ---- generic linalg: kron, A + λI, copy (src/sensitivities/linalg/generic.jl:7-51) -------------------
def _kron_views(A, B, Y):
    """kron as a 4-D broadcast: Y[(i-1)K + k, (j-1)L + l] = A[i,j] B[k,l] is, column-major, the array
    (K, I, L, J) = A viewed as (1, I, 1, J) .* B viewed as (K, 1, L, 1)."""
    (I, J), (Kk, L) = A.shape, B.shape
    return A.reshape((1, I, 1, J)), B.reshape((Kk, 1, L, 1)), Y.reshape((Kk, I, L, J))


def _kron_forward(A, B):
    A, B = _dev(A), _dev(B)
    if A.ndim != 2 or B.ndim != 2:
        raise TypeError("MethodError: the kron sensitivities destructure size(A), size(B) as matrices (generic.jl:19)")
    if A.dtype != B.dtype:
        raise TypeError("kron operands must share an element type")
    out = A.ctx.empty((A.shape[0] * B.shape[0], A.shape[1] * B.shape[1]), A.dtype)
    A4, B4, Y4 = _kron_views(A, B, out)
    K.bcast_fwd(_mul_prog(), [A4, B4], Y4)
    return out


def _mul_prog():
    from .rules import _MUL_PROG
    return _MUL_PROG


def _kron_fused(y, ybar, rvs):
    """Both sensitivities (generic.jl:13-40: ā_ij += Σ_kl Ȳ·B, B̄_kl += Σ_ij Ȳ·a_ij) in ONE launch of the broadcast
    pullback: the sums over the broadcast dimensions are its fused unbroadcast, `+=` its accumulate flag."""
    from .rules import _slot_target
    An, Bn = y.args
    A, B = _dev(unbox(An)), _dev(unbox(Bn))
    yb = _dev(unthunk(ybar))
    A4, B4, Y4 = _kron_views(A, B, yb)
    outs, acc = [None, None], [False, False]
    d_tape = rvs.tape
    if isinstance(An, Node):
        o, beta = _slot_target(d_tape, An.pos, A.shape, A.ctx, A.dtype)
        outs[0], acc[0] = o.reshape(A4.shape), beta == 1.0
    if isinstance(Bn, Node):
        if isinstance(An, Node) and Bn.pos == An.pos:
            raise TypeError("kron(x, x) with the same Node on both sides is not supported")
        o, beta = _slot_target(d_tape, Bn.pos, B.shape, B.ctx, B.dtype)
        outs[1], acc[1] = o.reshape(B4.shape), beta == 1.0
    K.bcast_pullback(_mul_prog(), Y4, None, [A4, B4], outs, acc)


_kron = Primitive("kron", _kron_forward, fused=_kron_fused)


def kron(A, B):
    """``kron(A, B)`` (generic.jl:7-40)."""
    return _kron(A, B)


def _plus_I_forward(A, lam):
    A = _dev(A)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise DimensionMismatch(6, "matrix is not square")
    n = A.shape[0]
    out = A.copy()
    if n:
        d = A.ctx.full((n,), float(lam), A.dtype)
        _copy2d(1, n, d, 0, 1, out, 0, n + 1, accumulate=True)   # the diagonal is a stride-(n+1) row
    return out


_plus_I = Primitive("+I", _plus_I_forward)


@_plus_I.grad(1)
def _(p, y, ybar, A, lam):
    return ybar   # generic.jl:43,46: the sensitivity is Ȳ itself (aliasing handled by the sweep)


def plus_I(A, lam):
    """``A + lam*I`` / ``lam*I + A`` with ``I`` the UniformScaling (generic.jl:42-46)."""
    return _plus_I(A, lam)


_copy = Primitive("copy", lambda A: _dev(A).copy())


@_copy.grad(1)
def _(p, y, ybar, A):
    return _dev(unthunk(ybar)).copy()   # generic.jl:51


def copy(A):
    """``copy(A)`` (generic.jl:48-51)."""
    return _copy(A)


# ---- hcat / vcat --------------------------------------------------------------------------------
def _hcat_forward(*A):
    A = [_dev(a) for a in A]
    n = _rc(A[0])[0]
    if any(_rc(a)[0] != n for a in A):
        raise DimensionMismatch(6, f"hcat: mismatched number of rows {[a.shape for a in A]}")
    out = A[0].ctx.empty((n, sum(_rc(a)[1] for a in A)), A[0].dtype)
    l = 0
    for a in A:
        c = _rc(a)[1]
        _copy2d(n, c, a, 0, n, out, l * n, n)
        l += c
    return out


_hcat = Primitive("hcat", _hcat_forward)


def _hcat_grad(i):
    def g(p, y, ybar, *A):
        ybar = _dev(ybar)
        n = y.shape[0]
        l = sum(_rc(A[j])[1] for j in range(i))
        c = _rc(A[i])[1]
        # a single column comes back as a Vector (selectdim with an integer index drops the dimension)
        out = ybar.ctx.empty((n,) if c == 1 else (n, c), ybar.dtype)
        _copy2d(n, c, ybar, l * n, n, out, 0, n)
        return out
    return g


def hcat(*A):
    for i in range(len(A)):
        _hcat.grads.setdefault(i + 1, _hcat_grad(i))
    return _hcat(*A)


def _vcat_forward(*A):
    A = [_dev(a) for a in A]
    cols = _rc(A[0])[1]
    nd = A[0].ndim
    if any(_rc(a)[1] != cols or a.ndim != nd for a in A):
        raise DimensionMismatch(6, f"vcat: mismatched shapes {[a.shape for a in A]}")
    rows = sum(_rc(a)[0] for a in A)
    out = A[0].ctx.empty((rows,) if nd == 1 else (rows, cols), A[0].dtype)
    l = 0
    for a in A:
        r = _rc(a)[0]
        _copy2d(r, cols, a, 0, r, out, l, rows)
        l += r
    return out


_vcat = Primitive("vcat", _vcat_forward)


def _vcat_grad(i):
    def g(p, y, ybar, *A):
        ybar = _dev(ybar)
        rows = y.shape[0]
        l = sum(_rc(A[j])[0] for j in range(i))
        r, cols = _rc(A[i])
        out = ybar.ctx.empty(A[i].shape, ybar.dtype)
        _copy2d(r, cols, ybar, l, rows, out, 0, r)
        return out
    return g


def vcat(*A):
    for i in range(len(A)):
        _vcat.grads.setdefault(i + 1, _vcat_grad(i))
    return _vcat(*A)


# ---- getindex -----------------------------------------------------------------------------------
class _IndexBuffer:
    """0-based int64 index vector in device memory."""

    def __init__(self, ctx, idx0):
        self.ctx = ctx
        a = np.ascontiguousarray(idx0, dtype=np.int64)
        self.n = a.size
        p = C.c_void_p()
        check(ctx.L.nb_alloc(ctx.h, max(a.nbytes, 8), C.byref(p)), ctx.h)
        self.ptr = p.value
        if a.nbytes:
            check(ctx.L.nb_h2d(ctx.h, self.ptr, a.ctypes.data, a.nbytes), ctx.h)
            ctx.sync()

    def __del__(self):
        try:
            if self.ptr and self.ctx.h:
                self.ctx.L.nb_release(self.ctx.h, self.ptr)
        except Exception:
            pass


def _span(ix, ext, shape):
    """-> (start0, length, keeps_dim) of an integer or range index along an extent."""
    if isinstance(ix, R):
        lo, hi = (1, ext) if ix is colon else (ix.lo, ix.hi)
        if lo < 1 or hi > ext:
            raise IndexError(f"BoundsError: attempt to access {shape} at index [{lo}:{hi}]")
        return lo - 1, max(hi - lo + 1, 0), True
    if isinstance(ix, numbers.Integral):
        if ix < 1 or ix > ext:
            raise IndexError(f"BoundsError: attempt to access {shape} at index [{ix}]")
        return int(ix) - 1, 1, False
    raise TypeError("MethodError: unsupported index type")


def _plan(x, idx):
    """Decode a Julia index tuple into either a sub-matrix copy or a gather."""
    if x.ndim == 1 and len(idx) == 1 and not isinstance(idx[0], (R, numbers.Integral)):
        iv = np.asarray(idx[0], dtype=np.int64)
        if iv.ndim != 1:
            raise TypeError("MethodError: index arrays must be vectors")
        if iv.size and (iv.min() < 1 or iv.max() > x.shape[0]):
            raise IndexError(f"BoundsError: attempt to access {x.shape} at index {list(iv)}")
        return ("gather", _IndexBuffer(x.ctx, iv - 1), (iv.size,))
    if x.ndim == 1 and len(idx) == 1:
        s, n, keep = _span(idx[0], x.shape[0], x.shape)
        return ("copy", n, 1, s, x.shape[0], (n,) if keep else ())
    if x.ndim == 2 and len(idx) == 2:
        rs, rn, rk = _span(idx[0], x.shape[0], x.shape)
        cs, cn, ck = _span(idx[1], x.shape[1], x.shape)
        shape = tuple(e for e, k in ((rn, rk), (cn, ck)) if k)
        return ("copy", rn, cn, rs + cs * x.shape[0], x.shape[0], shape)
    raise TypeError(f"MethodError: getindex(::Array{{{x.ndim}}}, {len(idx)} indices) is not intercepted")


def _getindex_forward(x, *idx):
    x = _dev(x)
    plan = _plan(x, idx)
    if plan[0] == "gather":
        _, ib, shape = plan
        out = x.ctx.empty(shape, x.dtype)
        check(x.ctx.L.nb_gather(x.ctx.h, ib.n, x.ptr, ib.ptr, out.ptr, x.nb_dtype), x.ctx.h)
        return out, plan
    _, rn, cn, off, ld, shape = plan
    out = x.ctx.empty(shape, x.dtype)
    _copy2d(rn, cn, x, off, ld, out, 0, max(rn, 1))
    return out, plan


def _getindex_fused(y, ybar, rvs):
    """x̄ = zero(x); x̄[inds...] .+= ȳ — accumulates in place when the slot already exists."""
    xn = y.args[0]
    if not isinstance(xn, Node):
        return
    from .core import UNDEF, own
    x = unbox(xn)
    ybar = _dev(unthunk(ybar))
    slot = rvs.tape[xn.pos - 1]
    xbar = zeroslike(x) if slot is UNDEF else own(slot)
    plan = y.saved
    if plan[0] == "gather":
        _, ib, _ = plan
        check(x.ctx.L.nb_scatter_add(x.ctx.h, ib.n, ybar.ptr, ib.ptr, xbar.ptr, x.nb_dtype), x.ctx.h)
    else:
        _, rn, cn, off, ld, _ = plan
        _copy2d(rn, cn, ybar, 0, max(rn, 1), xbar, off, ld, accumulate=True)
    rvs.tape[xn.pos - 1] = xbar


_getindex = Primitive("getindex", _getindex_forward, fused=_getindex_fused, returns_saved=True)


def getindex(x, *idx):
    """``x[idx...]`` with 1-based integers, ``R(lo, hi)`` ranges, ``colon`` and (for vectors) index
    vectors."""
    return _getindex(x, *idx)


# ---- reshape / fill -----------------------------------------------------------------------------
def _dims(dims):
    return tuple(int(d) for d in (dims[0] if len(dims) == 1 and isinstance(dims[0], (tuple, list)) else dims))


_reshape = Primitive("reshape", lambda x, *dims: _dev(x).reshape(_dims(dims)))


@_reshape.grad(1)
def _(p, y, ybar, x, *dims):
    return _dev(ybar).reshape(x.shape)


def reshape(x, *dims):
    return _reshape(x, *dims)


def _fill_forward(x, *dims):
    x = _dev(x)
    if x.ndim != 0:
        raise TypeError("MethodError: fill expects a scalar")
    out = x.ctx.empty(_dims(dims), x.dtype)
    K.bcast_reduce(None, [x], out)   # broadcast fill
    return out


_fill = Primitive("fill", _fill_forward)


@_fill.grad(1)
def _(p, y, ybar, x, *dims):
    ybar = _dev(ybar)
    out = ybar.ctx.empty((), ybar.dtype)
    K.bcast_reduce(None, [ybar], out)    # sum(ȳ)
    return out


def fill(x, *dims):
    return _fill(x, *dims)


# ---- checkpoint ---------------------------------------------------------------------------------
def _checkpoint_forward(f, *xs, **kw):
    return f(*xs, **kw)


def _checkpoint_fused(y, ybar, rvs):
    """checkpointing.jl:22-34: the forward pass kept only the inputs; re-run ``f`` on an inner tape
    inside the outer reverse pass and hand its sensitivities to the outer slots."""
    from .core import UNDEF, accumulate
    f, xs = y.args[0], y.args[1:]
    t = Tape()
    inner = [Leaf(t, unbox(a)) for a in xs]
    yi = f(*inner, **y.kwargs)
    if not isinstance(yi, Node):
        return
    df = nabla(yi, unthunk(ybar))
    for a, a_ in zip(xs, inner):
        if isinstance(a, Node) and df.isassigned(a_):
            g = df[a_]
            slot = rvs.tape[a.pos - 1]
            if slot is UNDEF:
                if isinstance(g, DevArray):
                    g.shared = True   # it also sits on the (dropped) inner reverse tape
                rvs.tape[a.pos - 1] = g
            else:
                rvs.tape[a.pos - 1] = accumulate(slot, g)
    t.tape.clear()


_checkpoint = Primitive("checkpoint", _checkpoint_forward, fused=_checkpoint_fused)


def checkpoint(f, args, kwargs=None):
    """``checkpoint(f, args::Tuple[, kwargs::NamedTuple])``: treat ``f`` as a primitive whose gradient is
    recomputed inside the outer reverse pass.  ``f`` must not be a closure (checkpointing.jl:11,15)."""
    if getattr(f, "__closure__", None):
        raise RuntimeError("f mustn't have fields. Can not checkpoint a closure or functor.")
    return _checkpoint(f, *args, **(kwargs or {}))
