"""Tape-level fusion around dense products.

The reference records ``f.(W*X .+ b)`` as three Branches — ``*``, ``broadcast(+)``, ``broadcast(f)``
(functional.jl:48-51: no dot-fusion across calls) — and walks them one by one in reverse
(core.jl:140-158), each materialising a full-size array.  The tape keeps exactly that structure here
(same Branches, same positions, same slots), but the LAUNCH of a Float32 tensor-core product is deferred:
its result array carries a ``Deferred`` description instead of data, and

  forward   a following ``.+ bias`` (bias along the rows) and ``f.`` join the product's epilogue, so
            ``f.(W*X .+ b)`` is ONE launch that writes only ``f(...)``;
  reverse   ``X̄ = W'·Z̄`` is deferred the same way, the ``f.`` Branch's pullback ``X̄ .* f'(y)``
            (functional.jl:92-95, for f whose derivative is a function of the saved y) joins its epilogue, and
            the ``.+`` Branch's bias sensitivity (the unbroadcast row sum, functional.jl:61-63) becomes the
            epilogue's row-sum output.

Whoever reads ``DevArray.ptr`` of a deferred value forces the launch, so every other rule works unchanged;
intermediate values nobody reads (``W*X`` itself, ``W*X .+ b``) are never computed separately and never
allocated.  The kernel side is ``nb_gemm_fused`` (csrc/gemm_tc.cu); when it reports NB_ERR_UNSUPPORTED
(Float64, FP32 SIMT mode, tiny products) the same description is replayed as the unfused launches."""
import ctypes as C
import weakref

from . import _lib
from . import kernels as K
from ._lib import OPCODES, UnsupportedOp, check
from .device import DevArray

_ID, _ADD = OPCODES["identity"], OPCODES["add"]
_A, _B, _Y = 1, 2, 4

_NEEDS_NONE = {"identity", "neg", "deg2rad", "rad2deg", "sign", "step", "floor", "ceil", "round", "trunc",
               "gt", "lt", "ge", "le"}
_NEEDS_Y = {"tanh", "exp", "logistic", "sqrt", "inv", "tan", "exp2", "exp10", "cbrt", "cot", "coth", "tand",
            "cotd", "expm1"}
_NEEDS_Y |= {"erfinv", "erfcinv", "invdigamma"}
_NEEDS_AY = {"sec", "csc", "sech", "csch", "secd", "cscd", "gamma", "erfcx", "dawson"}


def _is_binary(op):
    return OPCODES["add"] <= op < OPCODES["binary_end"]
_NAME = {v: k for k, v in OPCODES.items()}


def partial_needs(op, wrt=0):
    """Which inputs the partial derivative of a catalogue op reads: bit0 operand a, bit1 operand b, bit2 the
    saved value y.  Mirror of ``nb_partial_needs`` (csrc/ops.cuh); tests/test_host_logic.py keeps them equal."""
    name = _NAME[op]
    if name in _NEEDS_NONE or name in ("add", "sub"):
        return 0
    if name in _NEEDS_Y:
        return _Y
    if name in _NEEDS_AY:
        return _A | _Y
    if name in ("mask", "maskn"):
        return 0 if wrt == 0 else _A
    if name == "mul":
        return _B if wrt == 0 else _A
    if name == "div":
        return _B if wrt == 0 else (_B | _Y)
    if name == "ldiv":
        return (_A | _Y) if wrt == 0 else _A
    if name == "pow":
        return (_A | _B) if wrt == 0 else (_A | _Y)
    if name == "hypot":
        return (_A | _Y) if wrt == 0 else (_B | _Y)
    if name in ("max", "min", "atan2"):
        return _A | _B
    if name == "beta":
        return _A | _B | _Y
    if name in ("besselj", "bessely", "besseli", "besselk", "polygamma"):
        return 0 if wrt == 0 else (_A | _B)   # the partial in the order is the constant NaN
    return _A


def pullback_reads(prog, wanted):
    """For a single-primitive program: (device-operand indices, reads_y) that the partials w.r.t. the operands in
    ``wanted`` read; None when unknown (composite program: everything is read)."""
    if prog.single_op is None:
        return None
    if not prog.ops:
        return set(), False
    op, a = prog.ops[0], prog.a[0]
    b = prog.b[0] if _is_binary(op) else None
    reads, ry = set(), False
    for k in wanted:
        for side, slot in ((0, a), (1, b)):
            if slot is not None and slot == k:
                n = partial_needs(op, side)
                if n & _A and a >= 0:
                    reads.add(a)
                if n & _B and b is not None and b >= 0:
                    reads.add(b)
                ry = ry or bool(n & _Y)
    return reads, ry


class Shadow:
    """Shape-only stand-in for a deferred operand the kernel does not read (its data pointer is a valid, aligned
    address that is never dereferenced: the kernels load only what nb_partial_needs reports)."""

    def __init__(self, arr):
        self.ctx, self.shape, self.dtype, self._arr = arr.ctx, arr.shape, arr.dtype, arr

    def tensor(self):
        t = _lib.nb_tensor()
        t.data = self.ctx.scratch_ptr()
        t.dtype = self._arr.nb_dtype
        t.ndim = len(self.shape)
        for i in range(_lib.NB_MAX_DIMS):
            t.shape[i] = self.shape[i] if i < len(self.shape) else 1
        return t


class nb_gemm_epilogue(C.Structure):
    _fields_ = [("kind", C.c_int32), ("op", C.c_int32), ("bias", C.c_void_p), ("y", C.c_void_p),
                ("ldy", C.c_int64), ("rowsum", C.c_void_p), ("rowsum_acc", C.c_int32),
                ("emit_split", C.c_int32)]


class Deferred:
    """One product ``alpha*op(A)*op(B)`` (m x n) with what has joined its epilogue, not launched yet.

    ``stage``: 0 product, 1 + bias, 2 activation applied (forward) / multiplied by f'(y) (pullback).
    The result array holds this object in ``.pending``; the object only holds a weak reference back, so an
    intermediate nobody keeps (and its operands) is released by reference counting."""

    def __init__(self, ctx, tA, tB, alpha, A, B, m, n, k, math=None):
        self.ctx = ctx
        self.math = ctx.forward_math() if math is None else math
        self.tA, self.tB, self.alpha, self.A, self.B = tA, tB, alpha, A, B
        self.m, self.n, self.k = m, n, k
        self.stage = 0
        self.kind = 0            # 0 plain, 1 forward epilogue, 2 pullback epilogue
        self.op = _ID
        self.bias = None
        self.ysaved = None
        self.rowsum_ref = None   # weakref to the bias-sensitivity array (side output)
        self.rowsum_acc = False
        self.out_ref = None
        self.done = False

    # -- construction --
    def attach(self, out):
        self.out_ref = weakref.ref(out)
        out.pending = self
        self.ctx._deferred.append(weakref.ref(self))
        return out

    def extended(self):
        d = Deferred(self.ctx, self.tA, self.tB, self.alpha, self.A, self.B, self.m, self.n, self.k, self.math)
        d.stage, d.kind, d.op, d.bias, d.ysaved = self.stage, self.kind, self.op, self.bias, self.ysaved
        return d

    def new_result(self):
        return self.attach(DevArray(self.ctx, (self.m, self.n), self.A.dtype, lazy=True))

    def has_side_outputs(self):
        return self.rowsum_ref is not None

    def operands(self):
        """The arrays the pending launch reads."""
        return [x for x in (self.A, self.B, self.bias, self.ysaved) if x is not None]

    # -- launch --
    def flush(self):
        if self.done:
            return
        self.done = True
        out = self.out_ref() if self.out_ref is not None else None
        if out is not None:
            out.pending = None
        else:  # only the side output is still wanted
            out = DevArray(self.ctx, (self.m, self.n), self.A.dtype)
        rowsum = self.rowsum_ref() if self.rowsum_ref is not None else None
        if rowsum is not None:
            rowsum.pending = None
        self._launch(out, rowsum)
        self.A = self.B = self.bias = self.ysaved = None  # operands may go back to the pool

    def _launch(self, out, rowsum):
        ctx, m, n, k = self.ctx, self.m, self.n, self.k
        if self.kind == 0 and rowsum is None:
            K.gemm(self.tA, self.tB, self.alpha, self.A, self.B, 0.0, out, m, n, k, math=self.math)
            return
        epi = nb_gemm_epilogue()
        epi.kind = self.kind or 1
        epi.op = self.op
        epi.bias = self.bias.ptr if self.bias is not None else None
        epi.y = self.ysaved.ptr if self.ysaved is not None else None
        epi.ldy = m
        epi.rowsum = rowsum.ptr if rowsum is not None else None
        epi.rowsum_acc = 1 if self.rowsum_acc else 0
        epi.emit_split = 1 if m * n >= (1 << 16) else 0
        if epi.emit_split and epi.kind == 1 and ctx.adjoint_math() == _lib.MATH_BF16X3 \
                and self.math in (_lib.MATH_FP16X3, _lib.MATH_DEFAULT):
            epi.emit_split |= 2   # a forward value: the adjoint products (bfloat16 splits) will read it as well
        lda = k if self.tA == "T" else m
        ldb = n if self.tB == "T" else k
        pa, pb, pc = self.A.ptr, self.B.ptr, out.ptr   # (forces deferred operands before the timed region)
        try:
            with K._timed(ctx, f"gemm_{self.tA}{self.tB}_{m}x{n}x{k}", 2.0 * m * n * k, "FLOP"):
                check(ctx.L.nb_gemm_fused(ctx.h, self.tA.encode(), self.tB.encode(), m, n, k, float(self.alpha),
                                          pa, lda, pb, ldb, pc, m, out.nb_dtype, self.math, C.byref(epi)),
                      ctx.h)
            return
        except UnsupportedOp:
            pass
        self._replay_unfused(out, rowsum)

    def _replay_unfused(self, out, rowsum):
        """The same arithmetic as separate launches (what the tape would have run without deferral)."""
        from .scalar import CATALOGUE, compile_scalar
        ctx, m, n, k = self.ctx, self.m, self.n, self.k
        z = ctx.empty((m, n), out.dtype)
        K.gemm(self.tA, self.tB, self.alpha, self.A, self.B, 0.0, z, m, n, k, math=self.math)
        f = CATALOGUE[_NAME[self.op]]
        if self.kind == 2:
            K.bcast_pullback(compile_scalar(f, 1), z, self.ysaved, [self.ysaved], [out], [False])
        else:
            if self.bias is not None:
                t = ctx.empty((m, n), out.dtype)
                K.bcast_fwd(compile_scalar(CATALOGUE["add"], 2), [z, self.bias.reshape((m,))], t)
                z = t
            K.bcast_fwd(compile_scalar(f, 1), [z], out)
        if rowsum is not None:
            K.bcast_reduce(None, [out], rowsum.reshape((m, 1)) if rowsum.shape != (m, 1) else rowsum,
                           accumulate=self.rowsum_acc)


def defer_product(tA, tB, alpha, A, B, m, n, k, math=None):
    """-> a lazy (m x n) result for ``alpha*op(A)*op(B)``, or None when the product should run now.

    Deferred: Float32 products on the tensor-core path (not the FP32 SIMT cross-check mode), also inside a
    CUDA-graph capture (everything a captured sweep returns is forced before the capture ends).  Float64 products
    are not: a fused DMMA epilogue was built and measured slower than the separate launches (the DMMA kernel is
    not persistent, so the epilogue is a serial tail of every CTA; see csrc/gemm_simt.cu nb_gemm_fused)."""
    ctx = A.ctx
    if not ctx.fuse or A.dtype.itemsize != 4 or m * n * k < (1 << 18) or min(m, n, k) < 1:
        return None
    math = ctx.forward_math() if math is None else math
    if math == _lib.MATH_FP32:
        return None
    return Deferred(ctx, tA, tB, alpha, A, B, m, n, k, math).new_result()


def _pending(x, kinds=(0,), max_stage=0):
    d = getattr(x, "pending", None) if isinstance(x, DevArray) else None
    if not isinstance(d, Deferred) or d.done or d.kind not in kinds or d.stage > max_stage or d.rowsum_ref is not None:
        return None
    return d


def try_forward(prog, dev):
    """Forward of a broadcast Branch whose operand is a deferred product: returns the (lazy) result or None."""
    op = prog.single_op
    if op is None:
        return None
    if op == _ADD and len(dev) == 2 and prog.a[0] != prog.b[0]:
        for zi in (0, 1):
            z, b = dev[zi], dev[1 - zi]
            d = _pending(z, kinds=(0,), max_stage=0)
            if d is None or b.pending is not None or b.dtype != z.dtype:
                continue
            if tuple(b.shape) not in ((d.m,), (d.m, 1)) or (d.n == 1 and tuple(b.shape) == (d.m, 1)):
                continue
            e = d.extended()
            e.kind, e.stage, e.bias = 1, 1, b
            return e.new_result()
        return None
    if len(dev) == 1 and not _is_binary(op) and (partial_needs(op) & (_A | _B)) == 0:
        d = _pending(dev[0], kinds=(0, 1), max_stage=1)
        if d is None or d.kind == 2:
            return None
        e = d.extended()
        e.kind, e.stage, e.op = 1, 2, op
        return e.new_result()
    return None


def try_pullback_act(prog, ybar, y_saved, operand):
    """``x̄ = ȳ .* f'(y)`` where ȳ is a deferred product: the lazy x̄, or None."""
    op = prog.single_op
    if op is None or _is_binary(op) or (partial_needs(op) & (_A | _B)) != 0:
        return None
    d = _pending(ybar, kinds=(0,), max_stage=0)
    if d is None or y_saved is None or tuple(y_saved.shape) != (d.m, d.n) or tuple(operand.shape) != (d.m, d.n):
        return None
    if op == _ID:
        return None  # plain alias: nothing to fuse
    e = d.extended()
    e.kind, e.stage, e.op, e.ysaved = 2, 2, op, y_saved
    return e.new_result()


def try_rowsum(ybar, out, acc):
    """The bias sensitivity ``sum(ȳ, dims=2)`` as a side output of the deferred launch that produces ȳ."""
    d = getattr(ybar, "pending", None) if isinstance(ybar, DevArray) else None
    if not isinstance(d, Deferred) or d.done or d.rowsum_ref is not None or out.pending is not None:
        return False
    if tuple(out.shape) not in ((d.m,), (d.m, 1)) or d.n == 1:
        return False
    if d.kind == 0:
        d.kind, d.op = 1, _ID  # a plain product with an identity epilogue
    d.rowsum_ref = weakref.ref(out)
    d.rowsum_acc = bool(acc)
    out.pending = d
    return True
