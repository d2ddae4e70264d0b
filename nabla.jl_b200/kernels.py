"""Thin, typed wrappers over the C ABI entry points (include/nabla_cuda.h).  No arithmetic happens in
Python: every function here enqueues kernels of libnabla_cuda.so on the context stream."""
import ctypes as C

from . import _lib
from ._lib import NB_MAX_ARGS, check, nb_tensor
from .device import BroadcastView

_T = nb_tensor * NB_MAX_ARGS


def _pack(tensors):
    arr = _T()
    for i, t in enumerate(tensors):
        arr[i] = t
    return arr


def jl_broadcast_shape(*shapes):
    """``Broadcast.combine_axes`` (functional.jl:60): align from the leading dim."""
    nd = max((len(s) for s in shapes), default=0)
    out = []
    for d in range(nd):
        e = 1
        for s in shapes:
            x = s[d] if d < len(s) else 1
            if x != 1:
                if e != 1 and e != x:
                    raise _lib.DimensionMismatch(6, f"arrays could not be broadcast to a common size: {shapes}")
                e = x
        out.append(e)
    return tuple(out)


def _prog_handle(prog, ctx):
    return None if prog is None else prog.handle(ctx)


class _timed:
    """Bracket a launch with CUDA events when the context is profiling (see Context.profile_begin)."""

    __slots__ = ("ctx", "tag", "work", "unit", "ev")

    def __init__(self, ctx, tag, work, unit):
        self.ctx, self.tag, self.work, self.unit = ctx, tag, work, unit

    def __enter__(self):
        if self.ctx.profiler is not None:
            self.ev = self.ctx.event()
            self.ctx.record(self.ev)
        return self

    def __exit__(self, *exc):
        if self.ctx.profiler is not None and exc[0] is None:
            e1 = self.ctx.event()
            self.ctx.record(e1)
            self.ctx.profiler.append((self.tag, self.work, self.unit, self.ev, e1))
        return False


def bcast_fwd(prog, args, out):
    ctx = out.ctx
    check(ctx.L.nb_bcast_fwd(ctx.h, _prog_handle(prog, ctx), len(args), _pack([a.tensor() for a in args]),
                             C.byref(out.tensor())), ctx.h)
    return out


def bcast_reduce(prog, args, out, scale=1.0, accumulate=False):
    ctx = out.ctx
    check(ctx.L.nb_bcast_reduce(ctx.h, _prog_handle(prog, ctx), len(args),
                                _pack([a.tensor() for a in args]), float(scale), C.byref(out.tensor()),
                                1 if accumulate else 0), ctx.h)
    return out


def bcast_pullback(prog, ybar, y_saved, args, outs, accumulate):
    """outs: list (len(args)) of DevArray or None; accumulate: list of bool."""
    ctx = args[0].ctx
    want = acc = 0
    ts = []
    for k, o in enumerate(outs):
        if o is None:
            ts.append(nb_tensor())
        else:
            want |= 1 << k
            if accumulate[k]:
                acc |= 1 << k
            ts.append(o.tensor())
    yb = ybar.base if isinstance(ybar, BroadcastView) else ybar
    ybt = C.byref(yb.tensor()) if yb is not None else None
    yst = C.byref(y_saved.tensor()) if y_saved is not None else None
    with _timed(ctx, "bcast_pullback", 0, "B"):
        try:
            check(ctx.L.nb_bcast_pullback(ctx.h, _prog_handle(prog, ctx), ybt, yst, len(args),
                                          _pack([a.tensor() for a in args]), want, _pack(ts), acc), ctx.h)
        except _lib.NablaCudaError as e:   # say WHICH launch failed: program, operand and output shapes
            e.args = (f"{e.args[0]} [nb_bcast_pullback: ops={getattr(prog, 'ops', None)} a={getattr(prog, 'a', None)} "
                      f"b={getattr(prog, 'b', None)} args={[(tuple(a.shape), a.dtype.name) for a in args]} "
                      f"ybar={None if yb is None else tuple(yb.shape)} y_saved={None if y_saved is None else tuple(y_saved.shape)} "
                      f"outs={[None if o is None else tuple(o.shape) for o in outs]} acc={accumulate}]",)
            raise
    return outs


def gemm(tA, tB, alpha, A, B, beta, Cm, m, n, k, lda=None, ldb=None, ldc=None, math=None):
    """``math``: nb_math_mode of this product (default: the context's forward-product mode)."""
    ctx = Cm.ctx
    math = ctx.forward_math() if math is None else math
    lda = lda or (k if tA == "T" else m)
    ldb = ldb or (n if tB == "T" else k)
    ldc = ldc or m
    pa, pb, pc = A.ptr, B.ptr, Cm.ptr   # (forces deferred operands BEFORE the timed region starts)
    with _timed(ctx, f"gemm_{tA}{tB}_{m}x{n}x{k}", 2.0 * m * n * k, "FLOP"):
        check(ctx.L.nb_gemm(ctx.h, tA.encode(), tB.encode(), m, n, k, float(alpha), pa, lda, pb, ldb,
                            float(beta), pc, ldc, Cm.nb_dtype, math), ctx.h)
    return Cm


def gemv(tA, alpha, A, x, beta, y, m, n):
    ctx = y.ctx
    check(ctx.L.nb_gemv(ctx.h, tA.encode(), m, n, float(alpha), A.ptr, m, x.ptr, float(beta), y.ptr,
                        y.nb_dtype), ctx.h)
    return y


def ger(alpha, x, y, beta, A, m, n):
    ctx = A.ctx
    check(ctx.L.nb_ger(ctx.h, m, n, float(alpha), x.ptr, y.ptr, float(beta), A.ptr, m, A.nb_dtype), ctx.h)
    return A


def transpose(x):
    """Materialised adjoint of a matrix (m x n) -> (n x m)."""
    m, n = x.shape
    out = x.ctx.empty((n, m), x.dtype)
    check(x.ctx.L.nb_transpose(x.ctx.h, m, n, x.ptr, m, out.ptr, n, x.nb_dtype), x.ctx.h)
    return out


def materialize(x):
    return x.materialize() if isinstance(x, BroadcastView) else x
