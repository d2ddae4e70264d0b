"""Graph core: ``Tape`` / ``Leaf`` / ``Branch`` / ``propagate`` / ``∇`` — the host side of the sweep.

This mirrors src/core.jl of the reference name for name (``nabla`` is ``∇``, positions are 1-based)
so the device backend is a drop-in behind the same API.  The host keeps tape construction, the
strictly descending sweep order (core.jl:160-166) and the first-write / accumulate slot semantics
(core.jl:149-156, 203-205); all array arithmetic is enqueued on the context's CUDA stream through
libnabla_cuda.so.  Nothing here computes on the CPU."""
import numbers

import numpy as np

from .device import BroadcastView, DevArray, get_context


class _Undef:
    def __repr__(self):
        return "#undef"


UNDEF = _Undef()


def unthunk(x):
    """Reading a slot forces lazy tangents (core.jl:27)."""
    return x.materialize() if isinstance(x, BroadcastView) else x


class Tape:
    """core.jl:10-35."""

    def __init__(self, n=None):
        self.tape = [] if n is None else [UNDEF] * n

    def __len__(self):
        return len(self.tape)

    def push(self, node):
        self.tape.append(node)
        return self

    def isassigned(self, n):
        if isinstance(n, Node):
            n = n.pos
        return 1 <= n <= len(self.tape) and self.tape[n - 1] is not UNDEF

    def __getitem__(self, n):
        if isinstance(n, Node):
            n = n.pos
        v = self.tape[n - 1]
        if v is UNDEF:
            raise IndexError(f"UndefRefError: access to undefined reference at tape[{n}]")
        return unthunk(v)

    def __setitem__(self, n, x):
        if isinstance(n, Node):
            n = n.pos
        self.tape[n - 1] = x

    def raw(self, n):
        if isinstance(n, Node):
            n = n.pos
        return self.tape[n - 1]

    def __repr__(self):
        lines = [f"Tape with {len(self)} element{'s' if len(self) != 1 else ''}:"]
        lines += [f"  [{i + 1}]: {v!r}" for i, v in enumerate(self.tape)]
        return "\n".join(lines)


class Node:
    """``abstract type Node{T}`` (core.jl:7).  Operator sugar is attached by ``rules``."""

    __array_priority__ = 4000.0
    val = None
    tape = None
    pos = -1

    @property
    def shape(self):
        return self.val.shape

    @property
    def dtype(self):
        return self.val.dtype

    def __len__(self):
        return self.val.size


class Leaf(Node):
    """``Leaf(tape, val)`` (core.jl:48-57).  Host values are moved to HBM once, here."""

    def __init__(self, tape, val, ctx=None):
        # the reference's Leaf wraps ANY value (core.jl:48-57); only arrays and real numbers move to HBM, anything
        # else (a callable, a tuple, a string flag) is kept as it is and simply never receives a sensitivity
        numeric = isinstance(val, (DevArray, BroadcastView, np.ndarray, numbers.Real, list)) and not isinstance(val, bool)
        self.val = as_device(val, ctx) if numeric else val
        self.tape = tape
        self.pos = len(tape) + 1
        tape.push(self)

    def __repr__(self):
        return f"Leaf{{{self.val!r}}}"


class Branch(Node):
    """``Branch(f, args, tape; kwargs...)`` (core.jl:75-95).  ``saved`` carries what the fused device
    pullback needs (compiled scalar program, operand mapping) — the analogue of ``Branch.pullback``
    set by the ChainRules bridge (chainrules.jl:183)."""

    def __init__(self, f, args, tape, **kwargs):
        unboxed = tuple(unbox(a) for a in args)
        res = f.forward(*unboxed, **kwargs)
        if f.returns_saved:
            self.val, self.saved = res
        else:
            self.val, self.saved = res, None
        self.f = f
        self.args = tuple(args)
        self.kwargs = kwargs
        self.tape = tape
        self.pos = len(tape) + 1
        tape.push(self)

    def __repr__(self):
        return f"Branch{{{self.val!r}}} f={self.f.name}"


class Output(Branch):
    """The live output Node ``∇(f; get_output=true)`` returns (a Branch, as in the reference: test/core.jl:171-173): same
    value, function, arguments and position as the Branch on the tape, and the only strong reference to that tape
    (dropping it releases every forward value at once)."""

    def __init__(self, y, tape):   # (not Branch.__init__: nothing is evaluated or recorded)
        self.val, self.pos, self.tape = y.val, y.pos, tape
        self.f, self.args, self.kwargs = getattr(y, "f", None), getattr(y, "args", ()), getattr(y, "kwargs", {})
        self.saved = getattr(y, "saved", None)

    def __repr__(self):
        return f"Output{{{self.val!r}}}"


class Arg:
    """``Arg{N}`` (core.jl:177)."""

    def __init__(self, n):
        self.n = n


def unbox(x):
    return x.val if isinstance(x, Node) else x


def pos(x):
    return x.pos if isinstance(x, Node) else -1


def tape(x):
    return x.tape


def as_device(x, ctx=None, dtype=None):
    """Host value -> DevArray (numbers become 0-dim device scalars)."""
    if isinstance(x, (DevArray, BroadcastView)):
        return x
    ctx = ctx or get_context()
    if isinstance(x, numbers.Real):
        if dtype is None and isinstance(x, np.floating) and x.dtype in (np.float32, np.float64):
            dtype = x.dtype  # a Float32 scalar stays Float32 (Julia does not widen it either)
        return ctx.asarray(np.asarray(float(x), dtype=dtype or np.float64))
    return ctx.asarray(x, dtype)


def oneslike(x):
    x = unbox(x)
    return x.ctx.ones(x.shape, x.dtype)


def zeroslike(x):
    x = unbox(x)
    if not isinstance(x, (DevArray, BroadcastView)):
        return None      # a non-numeric argument (kept as it is by Leaf) has no sensitivity
    return x.ctx.zeros(x.shape, x.dtype)


class Primitive:
    """One intercepted function: the registration surface of src/sensitivity.jl.

    * ``forward(*unboxed, **kw)`` — primal on device values (core.jl:89).
    * ``grads[N](p, y, ȳ, *xs, **kw)`` — ``∇(f, Arg{N}, p, y, ȳ, xs...)`` (core.jl:189-193).
    * ``preprocess(y, ȳ, *xs)`` — run once per Branch (sensitivity.jl:206-221).
    * ``fused(branch, ȳ, rvs_tape)`` — device fast path: ONE launch produces every Node operand's
      sensitivity and fuses the accumulate (replaces the per-argument loop of core.jl:149-156)."""

    def __init__(self, name, forward, grads=None, preprocess=None, fused=None, returns_saved=False):
        self.name = name
        self.forward = forward
        self.grads = dict(grads or {})
        self._preprocess = preprocess
        self.fused = fused
        self.returns_saved = returns_saved

    def __repr__(self):
        return f"<primitive {self.name}>"

    def grad(self, n):
        """Decorator: register ``∇(f, Arg{n}, p, y, ȳ, xs...)`` (1-based n)."""
        def deco(fn):
            self.grads[n] = fn
            return fn
        return deco

    def preprocess(self, y, ybar, *args):
        if self._preprocess is None:
            return ()
        return self._preprocess(y.val, ybar, *[unbox(a) for a in args])

    def __call__(self, *args, **kwargs):
        nodes = [a for a in args if isinstance(a, Node)]
        if not nodes:
            res = self.forward(*args, **kwargs)
            return res[0] if self.returns_saved else res
        for nd in nodes[1:]:
            if nd.tape is not nodes[0].tape:   # the reference "fails silently" here (docs/src/index.md:88): be loud
                raise ValueError("Nodes of different Tapes in one call: every Leaf of a computation must share one Tape")
        return Branch(self, args, nodes[0].tape, **kwargs)


def explicit_intercepts(name=None):
    """``@explicit_intercepts f Tuple{...}`` (sensitivity.jl:39-49): turn a device function into a
    tape-aware primitive; attach sensitivities with ``@prim.grad(N)``."""
    def deco(fn):
        return Primitive(name or fn.__name__, fn)
    return deco


def accumulate(xbar, fresh):
    """Default ``∇(x̄, f, Arg{N}, ...) = add!!(x̄, fresh)`` (core.jl:203-205): in place when the slot
    owns its buffer, copy-on-accumulate when it is shared or lazy."""
    from .kernels import bcast_reduce
    xbar = own(xbar)
    fresh_base = fresh.base if isinstance(fresh, BroadcastView) else fresh
    bcast_reduce(None, [fresh_base], xbar, accumulate=True)
    return xbar


def own(x):
    """Return a buffer that may be mutated in place."""
    if isinstance(x, BroadcastView):
        return x.materialize()
    if x.shared:
        return x.copy()
    return x


def propagate_branch(y, rvs_tape):
    """``propagate(y::Branch, rvs_tape)`` (core.jl:140-158)."""
    f = y.f
    if f.fused is not None:
        f.fused(y, rvs_tape.raw(y), rvs_tape)
        return
    ybar = rvs_tape[y]
    d_tape = rvs_tape.tape
    xs = tuple(unbox(a) for a in y.args)
    p = f.preprocess(y, ybar, *y.args)
    for j, a in enumerate(y.args):
        xid = pos(a)
        if xid > 0:
            g = f.grads.get(j + 1)
            if g is None:
                raise TypeError(f"MethodError: no method matching ∇(::typeof({f.name}), Arg{{{j + 1}}}, ...)")
            fresh = g(p, y.val, ybar, *xs, **y.kwargs)
            if d_tape[xid - 1] is not UNDEF:
                d_tape[xid - 1] = accumulate(d_tape[xid - 1], fresh)
            else:
                if fresh is ybar and isinstance(fresh, DevArray):
                    fresh.shared = True  # the rule returned ȳ itself (aliasing)
                d_tape[xid - 1] = fresh


def last_contribution_positions(fwd_tape, upto):
    """For every Leaf: the SMALLEST tape position among the Branches (<= upto) that take it as an
    argument.  The sweep runs in descending order (core.jl:160-166) and a Leaf slot only receives
    contributions from its direct consumers (core.jl:149-156), so once the sweep has processed that
    position the Leaf's sensitivity is final."""
    first_use = {}
    for node in fwd_tape.tape[:upto]:
        if isinstance(node, Branch):
            for a in node.args:
                if isinstance(a, Leaf) and a.pos not in first_use:
                    first_use[a.pos] = node.pos
    return first_use


def propagate(fwd_tape, rvs_tape, on_final=None):
    """``propagate(fwd_tape, rvs_tape)`` (core.jl:160-166): strictly descending tape index.

    ``on_final(leaf_pos, slot)`` (optional, not in the reference) is called as soon as a Leaf's
    sensitivity can no longer change — used to start its allreduce while the sweep continues."""
    rvs_tape._fwd = fwd_tape      # (rules that look across Branches — ewfusion.py — find the forward tape here)
    finals = {}
    if on_final is not None:
        for leaf_pos, p in last_contribution_positions(fwd_tape, len(rvs_tape)).items():
            finals.setdefault(p, []).append(leaf_pos)
    for delta in range(len(rvs_tape), 0, -1):
        if rvs_tape.tape[delta - 1] is not UNDEF:
            node = fwd_tape.tape[delta - 1]
            if isinstance(node, Branch):
                propagate_branch(node, rvs_tape)
        for leaf_pos in finals.get(delta, ()):
            if rvs_tape.tape[leaf_pos - 1] is not UNDEF:
                on_final(leaf_pos, rvs_tape)
    # launches deferred by fusion.py that owe a value to a second array (a fused bias sensitivity) are made
    # before the sweep returns: every slot of the reverse tape is then final
    for node in fwd_tape.tape[:1]:
        v = unbox(node)
        if isinstance(v, DevArray):
            v.ctx.flush_deferred(only_side_outputs=True)
    return rvs_tape


def reverse_tape(y, ybar):
    """core.jl:170-174."""
    t = Tape(y.pos)
    if isinstance(ybar, DevArray):
        ybar.shared = True  # the caller's seed is never mutated
    t[len(t)] = ybar
    return t


def nabla(f_or_node, ybar=None, *, get_output=False):
    """``∇(y::Node[, ȳ])`` -> reverse tape (core.jl:200-201), or ``∇(f; get_output)`` -> gradient
    function (core.jl:214-229).  Gradients come back as DevArrays (``.numpy()`` / ``.item()``)."""
    if isinstance(f_or_node, Node):
        y = f_or_node
        if ybar is None:
            if y.val.ndim != 0:
                raise TypeError("MethodError: no method matching ∇(::Node{<:AbstractArray}); array "
                                "outputs need an explicit ȳ (core.jl:200-201)")
            ybar = oneslike(y)
        else:
            ybar = as_device(ybar, y.val.ctx, y.val.dtype)
            if tuple(ybar.shape) != tuple(y.val.shape):
                raise ValueError("DimensionMismatch: ȳ must have the shape of y")
        return propagate(y.tape, reverse_tape(y, ybar))
    f = f_or_node

    def gradient_fn(*args, **kwargs):
        t = Tape()
        args_ = [Leaf(t, a) for a in args]
        y = f(*args_, **kwargs)
        if isinstance(y, Node):
            df = nabla(y)
            out = tuple(df[a_] if df.isassigned(a_) else zeroslike(a_) for a_ in args_)
        else:
            out = tuple(zeroslike(a_) for a_ in args_)
        # Tape -> Node -> Tape is a reference cycle: dropping the node list here lets reference
        # counting return every forward value and intermediate sensitivity to the device pool as
        # soon as this call returns, instead of whenever Python's cyclic collector runs (which made
        # successive evaluations overlap in memory and hit cudaMalloc).  `y` stays a valid Node.
        # Tape -> Node -> Tape is a reference cycle: it is broken here so that reference counting returns every forward
        # value and intermediate sensitivity to the device pool as soon as the caller lets go, instead of whenever
        # Python's cyclic collector runs (successive evaluations would overlap in memory and hit cudaMalloc).  With
        # get_output the caller receives y as a LIVE Node (core.jl:214-229: it can be passed to ∇(y, ȳ) again): a
        # detached twin of y that owns the tape, while the nodes on the tape drop their back references.
        if not get_output or not isinstance(y, Node):
            t.tape.clear()
            return (y, out) if get_output else out
        for node in t.tape:
            node.tape = None
        yo = Output(y, t)
        return yo, out

    return gradient_fn


grad = nabla
