"""Device context and column-major device arrays (the `DevArray <: AbstractArray` of the Julia glue).

A ``Context`` wraps one ``nb_ctx``: one CUDA stream plus the pooled stream-ordered allocator that
replaces Julia's GC-managed ``Array`` storage (reference: src/core.jl:76,89 keeps every Branch value
alive on the tape; sensitivities are allocated at functional.jl:62,77,85)."""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import NB_F32, NB_F64, NB_MAX_DIMS, check, nb_tensor

_DT = {np.dtype(np.float32): NB_F32, np.dtype(np.float64): NB_F64}
_NP = {NB_F32: np.dtype(np.float32), NB_F64: np.dtype(np.float64)}


class Context:
    def __init__(self, device=None, pool_bytes=0):
        L = _lib.lib()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        check(L.nb_ctx_create(device, pool_bytes, C.byref(h)))
        self.h = h
        self.device = device
        self.L = L
        # Float32 GEMM arithmetic (include/nabla_cuda.h nb_math_mode).  MATH_DEFAULT is a policy, not one mode:
        # forward products run FP16X3 (22-bit operands: their error is amplified by saturating activations —
        # width-8192 MLP: BF16X3 forward products put 7e-4 on the gradient, outside rtol 1e-4), adjoint products
        # run BF16X3 (16-bit split of the sensitivities, any dynamic range; their error enters the gradient linearly)
        # re-using the cached 22-bit splits of the forward values.  Any other value applies to every product.
        self.math_mode = _lib.MATH_DEFAULT
        names = {"fp16x3": _lib.MATH_FP16X3, "bf16x3": _lib.MATH_BF16X3, "3xtf32": _lib.MATH_3XTF32,
                 "tf32": _lib.MATH_TF32, "fp32": _lib.MATH_FP32}
        self._fwd_math = names.get(os.environ.get("NB_F32_FWD_MATH", "").lower())
        self._adj_math = names.get(os.environ.get("NB_F32_ADJ_MATH", "").lower())
        self._progs = {}
        self.nranks, self.rank = 1, 0
        self._deferred = []   # weak references to launches that have been deferred (fusion.py)
        self.fuse = os.environ.get("NB_FUSE", "1") != "0"

    def close(self):
        if self.h:
            for p in self._progs.values():
                self.L.nb_prog_destroy(p)
            self._progs.clear()
            self.L.nb_ctx_destroy(self.h)
            self.h = None

    def forward_math(self):
        """nb_math_mode of a forward (primal) product."""
        if self.math_mode != _lib.MATH_DEFAULT:
            return self.math_mode
        return self._fwd_math if self._fwd_math is not None else _lib.MATH_FP16X3

    def adjoint_math(self):
        """nb_math_mode of an adjoint product (Ā = Ȳ·Bᵀ, B̄ = Aᵀ·Ȳ)."""
        if self.math_mode != _lib.MATH_DEFAULT:
            return self.math_mode
        return self._adj_math if self._adj_math is not None else _lib.MATH_BF16X3

    def sync(self):
        check(self.L.nb_sync(self.h), self.h)

    def scratch_ptr(self):
        """A small valid device address (256 B) for descriptors whose data is never read."""
        if getattr(self, "_scratch", None) is None:
            self._scratch = DevArray(self, (64,), np.float32)
        return self._scratch.ptr

    def flush_deferred(self, only_side_outputs=False):
        """Launch what fusion.py has deferred.  With ``only_side_outputs`` only the launches that owe a value to
        somebody besides their own result array (a fused bias sensitivity) are forced."""
        if not self._deferred:
            return
        live = []
        for r in self._deferred:
            d = r()
            if d is None or d.done:
                continue
            if only_side_outputs and not d.has_side_outputs():
                live.append(r)
                continue
            d.flush()
        self._deferred = live

    def flush_readers_of(self, arr):
        """Launch the deferred work that still READS ``arr`` (it is about to be overwritten).  Everything else stays
        deferred: a pending product nobody has asked for yet (an intermediate of the previous evaluation the caller
        still holds, a sensitivity it never read) must not be launched just because a new batch is copied in."""
        if not self._deferred:
            return
        root = arr._base or arr

        def reads(x, depth=0):
            if not isinstance(x, DevArray):
                return False
            if x is root or x._base is root:
                return True
            p = x.pending
            return p is not None and not p.done and depth < 64 and any(reads(o, depth + 1) for o in p.operands())

        for r in list(self._deferred):
            d = r()
            if d is not None and not d.done and any(reads(o) for o in d.operands()):
                d.flush()
        self._deferred = [r for r in self._deferred if r() is not None and not r().done]

    def stats(self):
        s = _lib.nb_stats()
        check(self.L.nb_ctx_stats(self.h, C.byref(s)), self.h)
        return {k: int(getattr(s, k)) for k, _ in s._fields_}

    def stream_ptr(self):
        p = C.c_void_p()
        check(self.L.nb_ctx_stream(self.h, C.byref(p)), self.h)
        return p.value

    # -- events (CUDA-event timing on the ctx stream) --
    def event(self):
        e = C.c_void_p()
        check(self.L.nb_event_create(C.byref(e)))
        return e

    def record(self, ev):
        check(self.L.nb_event_record(self.h, ev), self.h)

    def elapsed_ms(self, a, b):
        ms = C.c_float()
        check(self.L.nb_event_elapsed_ms(a, b, C.byref(ms)))
        return ms.value

    # -- per-launch device timing (bench.py roofline): when ``profiler`` is a list, the kernel
    #    wrappers bracket their launch with CUDA events on the ctx stream and append
    #    (tag, work, unit, ev_start, ev_stop) to it --
    profiler = None

    def profile_begin(self):
        self.profiler = []

    def profile_end(self):
        """-> list of (tag, work, unit, milliseconds); synchronises."""
        recs, self.profiler = self.profiler or [], None
        self.sync()
        out = []
        for tag, work, unit, a, b in recs:
            out.append((tag, work, unit, self.elapsed_ms(a, b)))
            self.L.nb_event_destroy(a)
            self.L.nb_event_destroy(b)
        return out

    # -- pinned host staging --
    def pinned_empty(self, shape, dtype):
        return PinnedArray(self, shape, dtype)

    # -- arrays --
    def empty(self, shape, dtype):
        return DevArray(self, tuple(int(s) for s in shape), np.dtype(dtype))

    def full(self, shape, value, dtype):
        a = self.empty(shape, dtype)
        check(self.L.nb_fill(self.h, a.ptr, a.nb_dtype, a.size, float(value)), self.h)
        return a

    def zeros(self, shape, dtype):
        return self.full(shape, 0.0, dtype)

    def ones(self, shape, dtype):
        return self.full(shape, 1.0, dtype)

    def scalar(self, value, dtype):
        return self.full((), value, dtype)

    def asarray(self, x, dtype=None):
        """Host -> device (column-major).  DevArrays pass through."""
        if isinstance(x, DevArray):
            return x
        a = np.asarray(x)
        if dtype is not None:
            a = a.astype(dtype, copy=False)
        elif a.dtype not in _DT:
            a = a.astype(np.float64)
        out = self.empty(a.shape, a.dtype)
        out.copy_from_host(a)
        return out


_default = None


def get_context():
    """Process-wide default context (device = LOCAL_RANK or 0).  Raises if no GPU is present."""
    global _default
    if _default is None:
        _default = Context()
    return _default


def set_context(ctx):
    global _default
    _default = ctx


class DevArray:
    """Dense column-major array in HBM, ``ndim <= 4`` (``ndim == 0`` is a device scalar).

    ``shared`` marks a buffer that is (or may be) referenced from more than one reverse-tape slot,
    e.g. when a rule returns ȳ itself; the accumulate path copies instead of mutating it
    (SURVEY.md §7 hard part 5; reference pin: test/core.jl:211-224)."""

    __slots__ = ("ctx", "_ptr", "shape", "dtype", "shared", "adjvec", "_base", "pending", "_tensor", "__weakref__")
    __array_priority__ = 3000.0

    def __init__(self, ctx, shape, dtype, ptr=None, base=None, lazy=False):
        if len(shape) > NB_MAX_DIMS:
            raise ValueError(f"DevArray supports at most {NB_MAX_DIMS} dimensions, got {shape}")
        if np.dtype(dtype) not in _DT:
            raise TypeError(f"DevArray eltype must be Float32 or Float64, got {dtype}")
        self.ctx = ctx
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.shared = False
        self.adjvec = False
        self._base = base
        self.pending = None  # fusion.Deferred: the launch that will produce this value (see `ptr`)
        self._tensor = None  # cached descriptor (shape and address are fixed once the buffer exists)
        if ptr is not None:
            self._ptr = ptr
        elif lazy:
            self._ptr = None  # allocated when the deferred launch is made (or never, if nobody reads it)
        else:
            self._alloc()

    def _alloc(self):
        p = C.c_void_p()
        check(self.ctx.L.nb_alloc(self.ctx.h, max(self.nbytes, 1), C.byref(p)), self.ctx.h)
        self._ptr = p.value

    @property
    def ptr(self):
        """Device address of the value.  Reading it is what forces a deferred product (fusion.py): every
        consumer that hands the buffer to a kernel, a copy or a collective goes through here."""
        if self.pending is not None:
            self.pending.flush()
        if self._ptr is None:
            self._alloc()
        return self._ptr

    def __del__(self):
        try:
            if self._base is None and self._ptr and self.ctx.h:
                self.ctx.L.nb_release(self.ctx.h, self._ptr)
        except Exception:
            pass

    # -- metadata --
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def nb_dtype(self):
        return _DT[self.dtype]

    def tensor(self, shape=None):
        if self.pending is not None:
            self.pending.flush()   # a descriptor is about to be handed to a kernel: the value must exist
        if shape is None:
            t = self._tensor
            if t is not None:
                return t
            t = self._make_tensor(self.shape)
            self._tensor = t
            return t
        return self._make_tensor(shape)

    def _make_tensor(self, shp):
        t = nb_tensor()
        t.data = self.ptr
        t.dtype = self.nb_dtype
        t.ndim = len(shp)
        for i in range(NB_MAX_DIMS):
            t.shape[i] = shp[i] if i < len(shp) else 1
        return t

    def reshape(self, shape):
        """Column-major reshape: a view on the same buffer (keeps it alive through ``_base``)."""
        shape = tuple(int(s) for s in shape)
        n = 1
        for s in shape:
            n *= s
        if n != self.size:
            raise ValueError(f"DimensionMismatch: cannot reshape {self.shape} to {shape}")
        v = DevArray(self.ctx, shape, self.dtype, ptr=self.ptr, base=self._base or self)
        v.shared = True
        return v

    # -- transfers --
    def copy_from_host(self, a):
        a = np.asarray(a, dtype=self.dtype)
        if a.ndim:   # (asfortranarray would promote a 0-dim scalar to shape (1,))
            a = np.asfortranarray(a)
        if a.shape != self.shape:
            raise ValueError(f"DimensionMismatch: {a.shape} vs {self.shape}")
        self.ctx.flush_readers_of(self)  # a deferred launch may still read this buffer
        if self.nbytes:
            # pageable source: the copy is complete (staged) when the call returns
            check(self.ctx.L.nb_h2d(self.ctx.h, self.ptr, a.ctypes.data, self.nbytes), self.ctx.h)
            self.ctx.sync()  # `a` may be a temporary: do not return before the DMA has read it
        return self

    def copy_from_pinned(self, pinned):
        """Asynchronous H2D from pinned host memory on the ctx stream (no host synchronisation)."""
        if pinned.nbytes != self.nbytes:
            raise ValueError(f"DimensionMismatch: {pinned.shape} vs {self.shape}")
        self.ctx.flush_readers_of(self)
        check(self.ctx.L.nb_h2d(self.ctx.h, self.ptr, pinned.ptr, self.nbytes), self.ctx.h)
        return self

    def copy_to_pinned(self, pinned):
        """Asynchronous D2H into pinned host memory; valid after ``ctx.sync()``."""
        if pinned.nbytes != self.nbytes:
            raise ValueError(f"DimensionMismatch: {pinned.shape} vs {self.shape}")
        check(self.ctx.L.nb_d2h_async(self.ctx.h, pinned.ptr, self.ptr, self.nbytes), self.ctx.h)
        return pinned

    def numpy(self):
        out = np.empty(self.shape, dtype=self.dtype, order="F")
        if self.nbytes:
            check(self.ctx.L.nb_d2h(self.ctx.h, out.ctypes.data, self.ptr, self.nbytes), self.ctx.h)
        if self.adjvec:
            return out
        return out if self.ndim else out[()]

    def item(self):
        return float(self.numpy())

    def __float__(self):
        return self.item()

    def copy(self):
        out = DevArray(self.ctx, self.shape, self.dtype)
        check(self.ctx.L.nb_d2d(self.ctx.h, out.ptr, self.ptr, self.nbytes), self.ctx.h)
        out.adjvec = self.adjvec
        return out

    def __repr__(self):
        return f"DevArray{{{self.dtype.name}}}{self.shape}"


class PinnedArray:
    """Page-locked host buffer (``nb_host_alloc``) viewed as a column-major NumPy array ``.a``."""

    def __init__(self, ctx, shape, dtype):
        self.ctx = ctx
        self.shape = tuple(int(x) for x in shape)
        self.dtype = np.dtype(dtype)
        n = 1
        for x in self.shape:
            n *= x
        self.nbytes = n * self.dtype.itemsize
        p = C.c_void_p()
        check(ctx.L.nb_host_alloc(max(self.nbytes, 1), C.byref(p)))
        self.ptr = p.value
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.a = np.frombuffer(buf, dtype=self.dtype, count=n).reshape(self.shape, order="F")

    def __del__(self):
        try:
            if self.ptr:
                self.ctx.L.nb_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class BroadcastView:
    """A lazily broadcast sensitivity: ``base`` (a DevArray whose extents are 1 or full) viewed at
    ``shape``.  This is the device analogue of the ``InplaceableThunk`` ChainRules returns from the
    ``sum``/``mean`` rules (x̄ = broadcast(last∘tuple, x, ȳ)): it is only materialised if a consumer
    cannot take a broadcast operand; the broadcast-pullback kernel reads it in place."""

    __slots__ = ("base", "shape")

    def __init__(self, base, shape):
        self.base = base
        self.shape = tuple(shape)

    @property
    def dtype(self):
        return self.base.dtype

    @property
    def ctx(self):
        return self.base.ctx

    def materialize(self):
        from .kernels import bcast_reduce
        out = self.base.ctx.empty(self.shape, self.base.dtype)
        bcast_reduce(None, [self.base], out)
        return out

    def numpy(self):
        return self.materialize().numpy()
