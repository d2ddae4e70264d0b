"""Tape-level fusion of ADJACENT broadcast Branches (second half).

The reference records every dot call as its own Branch with its own full-size array (functional.jl:48-51: no
dot-fusion across calls) and walks them one by one in reverse (core.jl:140-158): the log-likelihood of
examples/mlp.jl:68-73 is eight elementwise Branches and two reductions, each a launch here and each a pass over HBM.
The tape keeps exactly that structure (same Branches, same positions, same slots); what changes is WHEN and HOW MANY
launches are made:

  forward   the value of a broadcast Branch is LAZY (a ``LazyEw``: a scalar program over already materialised
            arrays).  A following broadcast / map / sum(f, .) that consumes it composes the two programs
            (``scalar.symbolic`` + ``scalar.linearize``) and runs — or stays lazy — as ONE launch over the leaves; the
            intermediate is never written unless something else reads it (reading ``DevArray.ptr`` forces it).
  reverse   when the sweep reaches a Branch whose operand is such a never-materialised value produced by a broadcast
            Branch with no other consumer, the pullback of BOTH Branches is one fused launch of the composed program
            (exact reverse accumulation over the program DAG, the derivative fmad's duals give, core.jl:285-292): the
            sensitivities go straight to the inner Branch's operands, its own slot stays unassigned and the sweep
            skips it (core.jl:163).  Chains of any length collapse the same way.

Only small arrays are treated this way (``NB_EW_FUSE_MAX`` elements, default 4 Mi): there the sweep is launch-bound
(C1 / C4: ~100 launches of a few microseconds).  Large arrays stay on the single-primitive lean kernels, which run at
80-90 % of the HBM peak while composed programs are interpreter-bound."""
import os
import weakref

from . import kernels as K
from ._lib import NB_MAX_ARGS, UnsupportedOp
from .device import BroadcastView, DevArray
from .scalar import Sym, linearize, symbolic

MAX_ELEMS = int(os.environ.get("NB_EW_FUSE_MAX", str(1 << 22)))
MAX_LEAVES = min(6, NB_MAX_ARGS)   # composed programs: more operands than this no longer fit the shared-memory interpreter
ENABLED = os.environ.get("NB_EW_FUSE", "1") != "0"


class LazyEw:
    """``prog`` over ``leaves`` (materialised or independently deferred DevArrays), not launched yet."""

    __slots__ = ("ctx", "prog", "leaves", "out_ref", "done", "__weakref__")

    def __init__(self, ctx, prog, leaves):
        self.ctx, self.prog, self.leaves = ctx, prog, list(leaves)
        self.out_ref = None
        self.done = False

    def attach(self, out):
        self.out_ref = weakref.ref(out)
        out.pending = self
        self.ctx._deferred.append(weakref.ref(self))
        return out

    def has_side_outputs(self):
        return False

    def operands(self):
        return list(self.leaves or ())

    def flush(self):
        if self.done:
            return
        self.done = True
        out = self.out_ref() if self.out_ref is not None else None
        leaves, self.leaves = self.leaves, None
        if out is None:
            return      # nobody holds the value any more: nothing to compute
        out.pending = None
        K.bcast_fwd(self.prog, leaves, out)


def lazy_of(x):
    p = getattr(x, "pending", None) if isinstance(x, DevArray) else None
    return p if isinstance(p, LazyEw) and not p.done else None


def small(shape):
    n = 1
    for s in shape:
        n *= s
    return n <= MAX_ELEMS


def compose(prog, dev):
    """``prog`` applied to ``dev`` where some operands are lazy -> (composed program, leaves), or None when the
    composition does not fit one device program (operand / instruction / constant limits of include/nabla_cuda.h)."""
    leaves, index = [], {}

    def leaf(d):
        k = index.get(id(d))
        if k is None:
            k = index[id(d)] = len(leaves)
            leaves.append(d)
        return Sym.arg(k)

    syms = []
    for d in dev:
        L = lazy_of(d)
        if L is None:
            syms.append(leaf(d))
        else:
            syms.append(symbolic(L.prog, [leaf(x) for x in L.leaves]))
    if len(leaves) > MAX_LEAVES:
        return None
    try:
        return linearize(symbolic(prog, syms), len(leaves), "fused"), leaves
    except UnsupportedOp:
        return None


def try_forward(ctx, prog, dev, full, dtype):
    """Lazy result of a broadcast Branch (composed with its lazy operands), or None."""
    if not (ENABLED and ctx.fuse and small(full)):
        return None
    c = compose(prog, dev)
    if c is None:
        return None
    out = DevArray(ctx, full, dtype, lazy=True)
    return LazyEw(ctx, c[0], c[1]).attach(out)


def try_reduce(ctx, prog, dev, full):
    """(program, operands) of a reduction over lazy operands composed into one launch, or None."""
    if not (ENABLED and ctx.fuse and small(full)) or not any(lazy_of(d) is not None for d in dev):
        return None
    return compose(prog, dev)


# ---------------------------------------------------------------------------------------------------------------
# reverse
# ---------------------------------------------------------------------------------------------------------------
def _consumers(y, rvs):
    """How many Branch arguments refer to each tape position (positions <= len(rvs)); cached on the reverse tape."""
    c = getattr(rvs, "_ew_consumers", None)
    if c is None:
        from .core import Branch, Node
        c = {}
        fwd = getattr(rvs, "_fwd", None) or y.tape
        for node in fwd.tape[:len(rvs)]:
            if isinstance(node, Branch):
                for a in node.args:
                    if isinstance(a, Node):
                        c[a.pos] = c.get(a.pos, 0) + 1
        rvs._ew_consumers = c
    return c


def _chainable(a, d, rvs, cons):
    """Operand Node ``a`` with device value ``d`` can be differentiated THROUGH: a broadcast / map Branch whose value
    was never materialised, with this one consumer and no sensitivity of its own yet."""
    from .core import UNDEF, Branch
    return (isinstance(a, Branch) and a.f.name in ("broadcast", "map") and a.saved is not None and not a.saved[4]
            and lazy_of(d) is not None and cons.get(a.pos, 0) == 1 and rvs.tape[a.pos - 1] is UNDEF)


def _leaf(d, a, leaves, nodes, index):
    from .core import Node
    k = index.get(id(d))
    if k is None:
        k = index[id(d)] = len(leaves)
        leaves.append(d)
        nodes.append(a if isinstance(a, Node) else None)
    return Sym.arg(k)


def _expr(branch, rvs, cons, leaves, nodes, index):
    """Expression of ``branch``'s value over the leaves: chainable operands are expanded recursively."""
    p, dv, dp = branch.saved[0], branch.saved[1], branch.saved[2]
    syms = []
    for k, j in enumerate(dp):
        a = branch.args[j + 1]
        if _chainable(a, dv[k], rvs, cons):
            syms.append(_expr(a, rvs, cons, leaves, nodes, index))
        else:
            syms.append(_leaf(dv[k], a, leaves, nodes, index))
    return symbolic(p, syms)


def _full_shape(arrays):
    """Broadcast shape of the operands (trailing singleton dimensions dropped like Julia's)."""
    nd = max((len(a.shape) for a in arrays), default=0)
    full = [1] * nd
    for a in arrays:
        for i, e in enumerate(a.shape):
            if e != 1:
                full[i] = e
    return tuple(full)


def chain_pullback(y, ybar, rvs, own, alias_ok=None):
    """Fused pullback of Branch ``y`` together with every chainable Branch below it.  Returns False when nothing is
    chainable or the composition does not fit (the caller then runs the ordinary per-Branch pullback)."""
    from .core import UNDEF, Node
    if not ENABLED:
        return False
    prog, dev, dev_pos, scale, reduce = y.saved
    cons = _consumers(y, rvs)
    if not any(_chainable(y.args[j + 1], dev[k], rvs, cons) for k, j in enumerate(dev_pos)):
        return False
    leaves, nodes, index = [], [], {}
    # (module-level recursion, not a nested closure calling itself: a self-referencing closure is a reference cycle
    #  that keeps the reverse tape — and every deferred launch hanging off its slots — alive until the cyclic collector
    #  runs, long after the sweep)
    root = _expr(y, rvs, cons, leaves, nodes, index)
    if len(leaves) > MAX_LEAVES:
        return False
    try:
        pc = linearize(root, len(leaves), "fused")
    except UnsupportedOp:
        return False
    d_tape = rvs.tape
    outs, acc = [None] * len(leaves), [False] * len(leaves)
    for k, (d, a) in enumerate(zip(leaves, nodes)):
        if a is None:
            continue
        slot = d_tape[a.pos - 1]
        if slot is UNDEF:
            outs[k] = d.ctx.empty(d.shape, d.dtype)
        else:
            outs[k], acc[k] = own(slot), True
        d_tape[a.pos - 1] = outs[k]
    # the root Branch's own forward value, when it was materialised, is the saved y of the composed program: the kernels
    # take f' from it where f' is a function of y alone (tanh.(A .+ b): 1 - y^2 instead of a second tanh)
    y_saved = None
    if not reduce and scale == 1.0:
        v = y.val
        if isinstance(v, DevArray) and v.pending is None and tuple(v.shape) == tuple(_full_shape(leaves)):
            y_saved = v
    if any(o is not None for o in outs):
        K.bcast_pullback(pc, ybar, y_saved, leaves, outs, acc)
    return True
