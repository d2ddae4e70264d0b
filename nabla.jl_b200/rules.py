"""Sensitivity rules on the device: broadcast/map, reductions, dense linalg, scalar arithmetic.

Each rule cites the reference rule it replaces.  The arithmetic is in libnabla_cuda.so; this module
only decides WHICH launch to make, keeps the reference's semantics for who gets a sensitivity (only
Node operands, core.jl:151), and routes first-write vs accumulate (core.jl:152-154)."""
import numbers

import numpy as np

from . import ewfusion, fusion
from . import kernels as K
from ._lib import OPCODES, DimensionMismatch
from .core import UNDEF, Node, Primitive, accumulate, as_device, own, unbox
from .device import BroadcastView, DevArray, get_context
from .scalar import CATALOGUE, ScalarFn, compile_scalar

_ADD, _SUB, _ID, _MUL = OPCODES["add"], OPCODES["sub"], OPCODES["identity"], OPCODES["mul"]
_LAZY_SUM = __import__("os").environ.get("NB_LAZY_SUM", "1") != "0"


def _is_num(x):
    return isinstance(x, numbers.Real) and not isinstance(x, bool)


def _dev_operands(operands):
    """Split operands into device arrays and baked-in Python constants."""
    ctx = None
    for o in operands:
        if isinstance(o, (DevArray, BroadcastView)):
            ctx = o.ctx
            break
    ctx = ctx or get_context()
    dev, dev_pos, consts = [], [], {}
    dtype = None
    for o in operands:
        if isinstance(o, (DevArray, BroadcastView)):
            dtype = o.dtype if dtype is None else np.promote_types(dtype, o.dtype)
    for i, o in enumerate(operands):
        if _is_num(o):
            consts[i] = o
        else:
            if isinstance(o, Ref):
                o = o.x
                if _is_num(o):
                    consts[i] = o
                    continue
            d = K.materialize(o) if isinstance(o, BroadcastView) else as_device(o, ctx, dtype)
            if dtype is not None and d.dtype != dtype:
                raise TypeError(f"mixed element types in one broadcast: {d.dtype} vs {dtype}; convert "
                                "explicitly (device kernels do not promote)")
            dev.append(d)
            dev_pos.append(i)
    return ctx, dev, dev_pos, consts


class Ref:
    """``Ref(x)``: broadcasts as a scalar (test/sensitivities/functional/functional.jl:181)."""

    def __init__(self, x):
        self.x = x


# ---------------------------------------------------------------------------------------------
# map-reduce family: broadcast, map, sum(f, .), mean(f, .), mapreduce(f, +, .), dot, scalar rules
# ---------------------------------------------------------------------------------------------
def _reduced_shape(shape, dims):
    if dims is None:
        return ()
    if isinstance(dims, (int, np.integer)):
        dims = (dims,)
    ds = {int(d) for d in dims}
    return tuple(1 if (i + 1) in ds else s for i, s in enumerate(shape))


def _denom(shape, dims):
    """``_denom`` (functional/reducedim.jl:60-62)."""
    if dims is None:
        return int(np.prod(shape)) if shape else 1
    if isinstance(dims, (int, np.integer)):
        dims = (dims,)
    n = 1
    for d in {int(d) for d in dims}:
        if d <= len(shape):
            n *= shape[d - 1]
    return n


def _mr_forward(fn, *operands, reduce=False, dims=None, mean=False, same_shape=False):
    ctx, dev, dev_pos, consts = _dev_operands(operands)
    prog = compile_scalar(fn, len(operands), consts)
    if same_shape and len({d.shape for d in dev}) > 1:
        raise DimensionMismatch(6, f"map: arrays must have the same shape: {[d.shape for d in dev]}")
    full = K.jl_broadcast_shape(*[d.shape for d in dev])
    dtype = dev[0].dtype
    if not reduce:
        out = fusion.try_forward(prog, dev) if ctx.fuse else None  # `.+ bias` / `f.` after a deferred product
        if out is None:   # small arrays: stay lazy so that the next broadcast / reduction composes with this one
            out = ewfusion.try_forward(ctx, prog, dev, full, dtype)
        if out is None:
            out = ctx.empty(full, dtype)
            K.bcast_fwd(prog, dev, out)
        scale = 1.0
    else:
        out = ctx.empty(_reduced_shape(full, dims), dtype)
        scale = 1.0 / _denom(full, dims) if mean else 1.0
        c = ewfusion.try_reduce(ctx, prog, dev, full)   # sum(f, lazy chain): one launch over the chain's leaves
        if c is not None:
            K.bcast_reduce(c[0], c[1], out, scale=scale)
        else:
            K.bcast_reduce(prog, dev, out, scale=scale)
    return out, (prog, dev, dev_pos, scale, reduce)


def _alias_ok(prog, k, operand_shape, y_shape):
    """ȳ itself is the sensitivity when ∂f/∂x_k == 1 and no unbroadcast is needed (the reference
    materialises ȳ .* 1, functional.jl:65-67; `add`-like rules elsewhere return ȳ, generic.jl:43)."""
    if tuple(operand_shape) != tuple(y_shape):
        return False
    op = prog.single_op
    if op == _ID:
        return True
    if op == _ADD:
        return prog.a[0] != prog.b[0] and k in (prog.a[0], prog.b[0])
    if op == _SUB:
        return prog.a[0] == k and prog.b[0] != k
    return False


def _mr_fused(y, ybar, rvs):
    """Fused pullback for the whole map-reduce family: replaces ∇(broadcast, Arg{N}, ...) ->
    broadcastsum (functional.jl:92-95), ∇(map, ...) (:21-24), ∇(mapreduce/sum/mean, ...)
    (reducedim.jl:11-72, reduce.jl:4-12) and ChainRules' dot / scalar rules."""
    prog, dev, dev_pos, scale, reduce = y.saved
    d_tape = rvs.tape
    if scale != 1.0:  # mean: ȳ / n  (reducedim.jl:71)
        yb = K.materialize(ybar)
        scaled = yb.ctx.empty(yb.shape, yb.dtype)
        K.bcast_reduce(None, [yb], scaled, scale=scale)
        ybar = scaled
    # adjacent broadcast Branches whose intermediate value was never materialised: ONE fused pullback launch for the
    # whole chain, straight to the inner Branches' operands (ewfusion.py)
    if ewfusion.chain_pullback(y, ybar, rvs, own):
        return
    if reduce and prog.single_op == _ID and len(dev) == 1 and isinstance(y.args[dev_pos[0] + 1], Node) and _LAZY_SUM:
        # plain sum(x; dims) / mean(x; dims): x̄ is ȳ broadcast back up (ChainRules returns it as an
        # InplaceableThunk).  A fresh slot gets the lazy BroadcastView — the consumer's pullback kernel reads ȳ
        # in place as a broadcast operand — instead of a full-size fill (N·s written, then N·s re-read).
        xid = y.args[dev_pos[0] + 1].pos
        if d_tape[xid - 1] is UNDEF:
            base = K.materialize(ybar)
            base.shared = True
            d_tape[xid - 1] = BroadcastView(base, dev[0].shape)
            return
    y_saved = None if reduce else y.val
    y_dropped = False
    if y_saved is not None and ewfusion.lazy_of(y_saved) is not None:
        y_saved = None     # the value itself was never materialised (its consumers were fused): the kernel re-evaluates f
        y_dropped = True   # ... from the operands, so every operand is read
    kdev = dev
    if any(d.pending is not None for d in dev) or (y_saved is not None and y_saved.pending is not None):
        # hand the kernel only what the partials read: a deferred forward value (W*X, W*X .+ b) that nothing
        # reads is never launched
        wanted = [k for k, j in enumerate(dev_pos) if isinstance(y.args[j + 1], Node)]
        reads = fusion.pullback_reads(prog, wanted)
        if reads is not None and not (y_dropped and reads[1]):
            read_ops, read_y = reads
            if not read_y:
                y_saved = None
            kdev = [d if (d.pending is None or k in read_ops) else fusion.Shadow(d) for k, d in enumerate(dev)]
    outs = [None] * len(dev)
    acc = [False] * len(dev)
    seen = {}
    post = []
    for k, j in enumerate(dev_pos):
        a = y.args[j + 1]   # y.args = (f, operands...): operand j is argument j + 1
        if not isinstance(a, Node):
            continue
        xid = a.pos
        slot = d_tape[xid - 1]
        if xid in seen:  # the same Node twice in one call (σz .* σz, examples/vae.jl:85)
            tmp = dev[k].ctx.empty(dev[k].shape, dev[k].dtype)
            outs[k] = tmp
            post.append((xid, tmp))
        elif slot is UNDEF:
            if not reduce and _alias_ok(prog, k, dev[k].shape, y.val.shape):
                if isinstance(ybar, DevArray):
                    ybar.shared = True
                d_tape[xid - 1] = ybar
            else:
                outs[k] = dev[k].ctx.empty(dev[k].shape, dev[k].dtype)
                d_tape[xid - 1] = outs[k]
            seen[xid] = k
        else:
            o = own(slot)
            outs[k], acc[k] = o, True
            d_tape[xid - 1] = o
            seen[xid] = k
    if ctx_fuse(dev) and not reduce and scale == 1.0 and not post:
        _join_deferred(prog, ybar, y_saved, dev, outs, acc, d_tape, y, dev_pos)
    if any(o is not None for o in outs):
        K.bcast_pullback(prog, ybar, y_saved, kdev, outs, acc)
    for xid, tmp in post:
        d_tape[xid - 1] = accumulate(d_tape[xid - 1], tmp)


def ctx_fuse(dev):
    return bool(dev) and dev[0].ctx.fuse


def _join_deferred(prog, ybar, y_saved, dev, outs, acc, d_tape, y, dev_pos):
    """Reverse-side joins of fusion.py: a fresh `ȳ .* f'(y)` becomes the epilogue of the deferred product that
    produces ȳ, and a bias sensitivity becomes its row-sum output.  Joined outputs are removed from ``outs``."""
    if not isinstance(ybar, DevArray) or ybar.pending is None:
        return
    for k, o in enumerate(outs):
        if o is None:
            continue
        if len(dev) == 1 and not acc[k] and o.pending is None:
            lazy = fusion.try_pullback_act(prog, ybar, y_saved, dev[k])
            if lazy is not None:
                xid = y.args[dev_pos[k] + 1].pos
                d_tape[xid - 1] = lazy   # replaces the buffer allocated for the unfused launch
                outs[k] = None
        elif prog.single_op == _ADD and len(dev) == 2 and tuple(ybar.shape) != tuple(o.shape):
            if fusion.try_rowsum(ybar, o, acc[k]):
                outs[k] = None


def _make_mr(name, **fixed):
    def forward(fn, *operands, **kw):
        return _mr_forward(fn, *operands, **fixed, **kw)
    return Primitive(name, forward, fused=_mr_fused, returns_saved=True)


_broadcast = _make_mr("broadcast")
_map = _make_mr("map", same_shape=True)
_sum_f = _make_mr("sum", reduce=True)
_mean_f = _make_mr("mean", reduce=True, mean=True)
_mapreduce = _make_mr("mapreduce", reduce=True)
_dot = _make_mr("dot", reduce=True)


def broadcast(f, *args):
    """``broadcast(f, args...)`` / ``f.(args...)`` — one Branch per call, never fused across calls
    (functional.jl:48-51)."""
    return _broadcast(f, *args)


def map_(f, *args):
    """``map(f, As...)`` (functional.jl:10-24)."""
    return _map(f, *args)


def _callable_first(args):
    return len(args) == 2 and callable(args[0]) and not isinstance(args[0], (Node, DevArray, np.ndarray))


def _check_dims(x, dims):
    if dims is None:
        return None
    if isinstance(dims, (int, np.integer)):
        dims = (int(dims),)
    return tuple(sorted({int(d) for d in dims}))


def sum(*args, dims=None):  # noqa: A001
    """``sum(x; dims)`` (ChainRules rrule) and ``sum(f, x; dims)`` incl. ``sum(abs2, x)``
    (functional/reducedim.jl:23-51)."""
    dims = _check_dims(None, dims)
    if _callable_first(args):
        return _sum_f(args[0], args[1], dims=dims)
    return _sum_f(CATALOGUE["identity"], args[0], dims=dims)


def mean(*args, dims=None):
    """``mean(x; dims)`` (ChainRules) and ``mean(f, x)`` (reducedim.jl:53-72)."""
    dims = _check_dims(None, dims)
    if _callable_first(args):
        if dims is not None:
            raise TypeError("MethodError: mean(f, x; dims) is not intercepted (reducedim.jl:53-58)")
        return _mean_f(args[0], args[1], dims=None)
    return _mean_f(CATALOGUE["identity"], args[0], dims=dims)


def _check_plus(op):
    if not (op == "+" or (isinstance(op, ScalarFn) and op.name == "add")):
        raise TypeError("MethodError: only `+` reductions are intercepted (functional/reduce.jl:2)")


def mapreduce(f, op, A, dims=None):
    """``mapreduce(f, +, A; dims)`` (reducedim.jl:4-21)."""
    _check_plus(op)
    return _mapreduce(f, A, dims=_check_dims(None, dims))


def mapfoldl(f, op, A):
    """``mapfoldl(f, +, A)`` (reduce.jl:4-12)."""
    _check_plus(op)
    return _mapreduce(f, A, dims=None)


mapfoldr = mapfoldl


def dot(x, y):
    """``LinearAlgebra.dot`` (ChainRules): x̄ = ȳ·y, ȳ' = ȳ·x — one fused launch."""
    return _dot(CATALOGUE["mul"], x, y, dims=None)


def scalar_call(fn, args):
    """A catalogue primitive applied to scalar Nodes / device scalars (ChainRules @scalar_rule)."""
    for a in args:
        v = unbox(a)
        if isinstance(v, (DevArray, BroadcastView)) and len(v.shape) > 0:
            raise TypeError(f"MethodError: no method matching {fn.name}(::Array); use broadcast")
        if not isinstance(v, (DevArray, BroadcastView)) and not _is_num(v):
            raise TypeError(f"MethodError: no method matching {fn.name}(::{type(v).__name__})")
    if all(_is_num(unbox(a)) for a in args):
        raise TypeError(f"{fn.name} of plain Python numbers is host arithmetic; this package only "
                        "provides device rules (use `math`)")
    return _broadcast(fn, *args)


# ---------------------------------------------------------------------------------------------
# dense linalg: `*`, adjoint/transpose, BLAS.gemm / gemv
# ---------------------------------------------------------------------------------------------
def _slot_target(d_tape, xid, shape, ctx, dtype):
    """(buffer, beta) for a GEMM-style rule writing the sensitivity of tape position xid."""
    slot = d_tape[xid - 1]
    if slot is UNDEF:
        out = ctx.empty(shape, dtype)
        d_tape[xid - 1] = out
        return out, 0.0
    out = own(slot)
    d_tape[xid - 1] = out
    return out, 1.0


def _matmul_forward(A, B):
    A, B = K.materialize(A), K.materialize(B)
    if not isinstance(A, DevArray) or not isinstance(B, DevArray):
        ctx = A.ctx if isinstance(A, DevArray) else B.ctx if isinstance(B, DevArray) else get_context()
        dt = A.dtype if isinstance(A, DevArray) else B.dtype if isinstance(B, DevArray) else None
        A, B = as_device(A, ctx, dt), as_device(B, ctx, dt)
    if A.dtype != B.dtype:
        raise TypeError("matmul operands must share an element type")
    ctx = A.ctx
    if A.ndim == 2 and B.ndim == 2:
        m, k = A.shape
        k2, n = B.shape
        if k != k2:
            raise DimensionMismatch(6, f"A has dimensions {A.shape} but B has dimensions {B.shape}")
        out = fusion.defer_product("N", "N", 1.0, A, B, m, n, k)   # launched when somebody reads it
        if out is None:
            out = ctx.empty((m, n), A.dtype)
            K.gemm("N", "N", 1.0, A, B, 0.0, out, m, n, k)
        return out, ("mm", A, B)
    if A.ndim == 2 and B.ndim == 1:
        m, k = A.shape
        if k != B.shape[0]:
            raise DimensionMismatch(6, f"A has dimensions {A.shape} but x has length {B.shape[0]}")
        if A.adjvec:  # x' * y -> scalar
            out = ctx.empty((), A.dtype)
            K.bcast_reduce(_MUL_PROG, [A.reshape((k,)), B], out)
            return out, ("dot", A, B)
        out = ctx.empty((m,), A.dtype)
        K.gemv("N", 1.0, A, B, 0.0, out, m, k)
        return out, ("mv", A, B)
    raise TypeError(f"MethodError: no method matching *(::Array{{{A.ndim}}}, ::Array{{{B.ndim}}})")


_MUL_PROG = compile_scalar(CATALOGUE["mul"], 2)


def _matmul_fused(y, ybar, rvs):
    """ChainRules rrule(*): Ā = Ȳ·Bᵀ, B̄ = Aᵀ·Ȳ, each only if the operand is a Node (thunks of
    constants are never forced, core.jl:151), accumulated with beta = 1 when the slot exists
    (InplaceableThunk mul!(X̄, Ȳ, B', true, true))."""
    kind, A, B = y.saved
    An, Bn = y.args
    d_tape = rvs.tape
    ctx, dt = A.ctx, A.dtype
    ybar = K.materialize(ybar)
    adj = ctx.adjoint_math()
    if kind == "mm":
        m, k = A.shape
        n = B.shape[1]
        if isinstance(An, Node):
            out, beta = _slot_target(d_tape, An.pos, A.shape, ctx, dt)
            K.gemm("N", "T", 1.0, ybar, B, beta, out, m, k, n, math=adj)      # Ȳ (m×n) · Bᵀ (n×k)
        if isinstance(Bn, Node):
            lazy = (fusion.defer_product("T", "N", 1.0, A, ybar, k, n, m, math=adj)
                    if d_tape[Bn.pos - 1] is UNDEF else None)
            if lazy is not None:   # the pullback of the Branch that produced B may join this launch
                d_tape[Bn.pos - 1] = lazy
            else:
                out, beta = _slot_target(d_tape, Bn.pos, B.shape, ctx, dt)
                K.gemm("T", "N", 1.0, A, ybar, beta, out, k, n, m, math=adj)      # Aᵀ (k×m) · Ȳ (m×n)
    elif kind == "mv":
        m, k = A.shape
        if isinstance(An, Node):
            out, beta = _slot_target(d_tape, An.pos, A.shape, ctx, dt)
            K.ger(1.0, ybar, B, beta, out, m, k)                    # ȳ xᵀ
        if isinstance(Bn, Node):
            out, beta = _slot_target(d_tape, Bn.pos, B.shape, ctx, dt)
            K.gemv("T", 1.0, A, ybar, beta, out, m, k)              # Aᵀ ȳ
    else:  # "dot": x' * v
        k = B.shape[0]
        xa = A.reshape((k,))
        outs, acc = [None, None], [False, False]
        if isinstance(An, Node):
            o, beta = _slot_target(d_tape, An.pos, A.shape, ctx, dt)
            o.adjvec = True
            outs[0], acc[0] = (o.reshape((k,)) if o.shape != (k,) else o), beta == 1.0
        if isinstance(Bn, Node):
            if isinstance(An, Node) and Bn.pos == An.pos:
                raise TypeError("x' * x with the same Node on both sides: use dot(x, x)")
            o, beta = _slot_target(d_tape, Bn.pos, B.shape, ctx, dt)
            outs[1], acc[1] = o, beta == 1.0
        K.bcast_pullback(_MUL_PROG, ybar, None, [xa, B], outs, acc)


_matmul = Primitive("*", _matmul_forward, fused=_matmul_fused, returns_saved=True)


def matmul(A, B):
    """Julia ``A * B`` for matrix·matrix, matrix·vector and adjoint-vector·vector."""
    return _matmul(A, B)


def _adjoint_forward(x):
    x = K.materialize(x)
    if x.ndim == 1:
        v = x.reshape((1, x.shape[0]))
        v.adjvec = True
        return v
    if x.ndim == 2 and x.adjvec:
        return x.reshape((x.shape[1],))
    if x.ndim == 2:
        return K.transpose(x)
    raise TypeError("MethodError: adjoint is defined for vectors and matrices")


_adjoint = Primitive("adjoint", _adjoint_forward)


@_adjoint.grad(1)
def _adjoint_grad(p, y, ybar, x):
    """ChainRules rrule(adjoint): x̄ = ȳ', projected back to the operand's shape (a Vector operand
    gets a Vector sensitivity, not an n x 1 matrix)."""
    ybar = K.materialize(ybar)
    if x.ndim == 1:
        return ybar.reshape((x.shape[0],))
    if x.ndim == 2 and x.adjvec:
        v = ybar.reshape((1, x.shape[1]))
        v.adjvec = True
        return v
    return K.transpose(ybar)


def adjoint(x):
    return _adjoint(x)


transpose = adjoint


def _op_shape(t, M):
    return (M.shape[1], M.shape[0]) if t in ("T", "C") else M.shape


def _gemm_forward(tA, tB, *rest):
    if len(rest) == 3:
        alpha, A, B = rest
    else:
        (A, B), alpha = rest, 1.0
    alpha = float(alpha.item()) if isinstance(alpha, DevArray) else float(alpha)
    m, k = _op_shape(tA, A)
    k2, n = _op_shape(tB, B)
    if k != k2:
        raise DimensionMismatch(6, f"gemm: op(A) is {m}x{k} but op(B) is {k2}x{n}")
    out = A.ctx.empty((m, n), A.dtype)
    K.gemm("T" if tA in "TC" else "N", "T" if tB in "TC" else "N", alpha, A, B, 0.0, out, m, n, k)
    return out, (alpha, m, n, k)


def _gemm_fused(y, ybar, rvs):
    """Adjoints of ``BLAS.gemm(tA, tB, [α,] A, B)`` for all four transposes (pinned by
    test/sensitivities/linalg/blas.jl:36-47): with X = op(A), Z = op(B):
    X̄ = α Ȳ Zᵀ, Z̄ = α Xᵀ Ȳ, stored transposed when the operand was."""
    alpha, m, n, k = y.saved
    tA, tB = y.args[0], y.args[1]
    An, Bn = y.args[-2], y.args[-1]
    A, B = unbox(An), unbox(Bn)
    tAt, tBt = tA in "TC", tB in "TC"
    ybar = K.materialize(ybar)
    d_tape = rvs.tape
    ctx, dt = A.ctx, A.dtype
    adj = ctx.adjoint_math()
    if len(y.args) == 5 and isinstance(y.args[2], Node):
        aN = y.args[2]  # ᾱ = <Ȳ, op(A) op(B)> = <Ȳ, Y> / α
        tmp = ctx.empty((), dt)
        K.bcast_reduce(_MUL_PROG, [ybar, y.val], tmp, scale=1.0 / alpha)
        slot = d_tape[aN.pos - 1]
        d_tape[aN.pos - 1] = tmp if slot is UNDEF else accumulate(slot, tmp)
    if isinstance(An, Node):
        out, beta = _slot_target(d_tape, An.pos, A.shape, ctx, dt)
        if not tAt:   # Ā (m×k) = α Ȳ op(B)ᵀ
            K.gemm("N", "N" if tBt else "T", alpha, ybar, B, beta, out, m, k, n, math=adj)
        else:         # Ā (k×m) = α op(B) Ȳᵀ
            K.gemm("T" if tBt else "N", "T", alpha, B, ybar, beta, out, k, m, n, math=adj)
    if isinstance(Bn, Node):
        out, beta = _slot_target(d_tape, Bn.pos, B.shape, ctx, dt)
        if not tBt:   # B̄ (k×n) = α op(A)ᵀ Ȳ
            K.gemm("N" if tAt else "T", "N", alpha, A, ybar, beta, out, k, n, m, math=adj)
        else:         # B̄ (n×k) = α Ȳᵀ op(A)
            K.gemm("T", "T" if tAt else "N", alpha, ybar, A, beta, out, n, k, m, math=adj)


_gemm = Primitive("gemm", _gemm_forward, fused=_gemm_fused, returns_saved=True)


def gemm(tA, tB, *rest):
    """``BLAS.gemm(tA, tB, [α,] A, B)``."""
    return _gemm(tA, tB, *rest)


def _gemv_forward(tA, *rest):
    if len(rest) == 3:
        alpha, A, x = rest
    else:
        (A, x), alpha = rest, 1.0
    alpha = float(alpha.item()) if isinstance(alpha, DevArray) else float(alpha)
    m, n = A.shape
    t = tA in "TC"
    out = A.ctx.empty((n if t else m,), A.dtype)
    K.gemv("T" if t else "N", alpha, A, x, 0.0, out, m, n)
    return out, (alpha, t)


def _gemv_fused(y, ybar, rvs):
    """Adjoints of ``BLAS.gemv(tA, [α,] A, x)`` (test/sensitivities/linalg/blas.jl:49-61)."""
    alpha, t = y.saved
    An, xn = y.args[-2], y.args[-1]
    A, x = unbox(An), unbox(xn)
    m, n = A.shape
    ybar = K.materialize(ybar)
    d_tape = rvs.tape
    ctx, dt = A.ctx, A.dtype
    if len(y.args) == 4 and isinstance(y.args[1], Node):
        aN = y.args[1]
        tmp = ctx.empty((), dt)
        K.bcast_reduce(_MUL_PROG, [ybar, y.val], tmp, scale=1.0 / alpha)
        slot = d_tape[aN.pos - 1]
        d_tape[aN.pos - 1] = tmp if slot is UNDEF else accumulate(slot, tmp)
    if isinstance(An, Node):
        out, beta = _slot_target(d_tape, An.pos, A.shape, ctx, dt)
        if not t:
            K.ger(alpha, ybar, x, beta, out, m, n)   # α ȳ xᵀ
        else:
            K.ger(alpha, x, ybar, beta, out, m, n)   # α x ȳᵀ
    if isinstance(xn, Node):
        out, beta = _slot_target(d_tape, xn.pos, x.shape, ctx, dt)
        K.gemv("N" if t else "T", alpha, A, ybar, beta, out, m, n)


_gemv = Primitive("gemv", _gemv_forward, fused=_gemv_fused, returns_saved=True)


def gemv(tA, *rest):
    return _gemv(tA, *rest)


# ---------------------------------------------------------------------------------------------
# operator sugar on Node (Julia's +, -, *, /, \, ^, ')
# ---------------------------------------------------------------------------------------------
def _isarr(x):
    v = unbox(x)
    if isinstance(v, np.ndarray):
        return v.ndim > 0
    return isinstance(v, (DevArray, BroadcastView)) and len(v.shape) > 0


def _same_shape(a, b, opname):
    sa, sb = tuple(unbox(a).shape), tuple(unbox(b).shape)
    if sa != sb:
        raise DimensionMismatch(6, f"dimensions must match: a has dims {sa}, b has dims {sb} ({opname})")


def _plus(a, b):
    if _isarr(a) and _isarr(b):       # array + array -> broadcast rule (chainrules_ignored.jl:5-10)
        _same_shape(a, b, "+")
        return _broadcast(CATALOGUE["add"], a, b)
    if _isarr(a) or _isarr(b):
        raise TypeError("MethodError: no method matching +(::Array, ::Number); use broadcast")
    return _broadcast(CATALOGUE["add"], a, b)


def _minus(a, b):
    if _isarr(a) and _isarr(b):
        _same_shape(a, b, "-")
        return _broadcast(CATALOGUE["sub"], a, b)
    if _isarr(a) or _isarr(b):
        raise TypeError("MethodError: no method matching -(::Array, ::Number); use broadcast")
    return _broadcast(CATALOGUE["sub"], a, b)


def _times(a, b):
    if _isarr(a) and _isarr(b):
        return _matmul(a, b)
    return _broadcast(CATALOGUE["mul"], a, b)   # scalar*array (ChainRules) and scalar*scalar


def _divide(a, b):
    if _isarr(b):
        raise TypeError("MethodError: right-division by an array is out of scope (LAPACK)")
    return _broadcast(CATALOGUE["div"], a, b)   # array / scalar (functional.jl:98-103)


def ldivide(a, b):
    """Julia ``a \\ b`` for scalar ``a`` (functional.jl:105-111)."""
    if _isarr(a):
        raise TypeError("MethodError: left-division by an array is out of scope (LAPACK)")
    return _broadcast(CATALOGUE["ldiv"], a, b)


def _power(a, b):
    if _isarr(a) or _isarr(b):
        raise TypeError("MethodError: ^ on arrays is out of scope; use broadcast")
    return _broadcast(CATALOGUE["pow"], a, b)


def _install_operators():
    Node.__add__ = lambda s, o: _plus(s, o)
    Node.__radd__ = lambda s, o: _plus(o, s)
    Node.__sub__ = lambda s, o: _minus(s, o)
    Node.__rsub__ = lambda s, o: _minus(o, s)
    Node.__mul__ = lambda s, o: _times(s, o)
    Node.__rmul__ = lambda s, o: _times(o, s)
    Node.__matmul__ = lambda s, o: _matmul(s, o)
    Node.__rmatmul__ = lambda s, o: _matmul(o, s)
    Node.__truediv__ = lambda s, o: _divide(s, o)
    Node.__rtruediv__ = lambda s, o: _divide(o, s)
    Node.__pow__ = lambda s, o: _power(s, o)
    Node.__neg__ = lambda s: _broadcast(CATALOGUE["neg"], s)
    Node.T = property(lambda s: adjoint(s))


_install_operators()
