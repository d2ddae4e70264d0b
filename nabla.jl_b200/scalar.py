"""Scalar functions for ``broadcast``/``map``/``sum(f, .)``: the primitive catalogue and the tracer that
lowers an arbitrary Python scalar function to a device program.

Reference: Nabla differentiates *any* Julia scalar function inside ``broadcast`` by evaluating it on a
``ForwardDiff.Dual`` per element (src/core.jl:285-292).  With no device code generation allowed, a
scalar function is traced ONCE on symbolic scalars into a short register program over the catalogue
(include/nabla_cuda.h ``nb_op``; the list pinned by test/runtests.jl:28-52), which the kernels
differentiate exactly.  A primitive outside the catalogue raises ``UnsupportedOp`` (the Julia glue
raises MethodError) — there is no CPU fallback."""
import ctypes as C
import math
import numbers
import types

from ._lib import NB_MAX_ARGS, NB_MAX_CONSTS, NB_MAX_INSTR, OPCODES, UnsupportedOp, check

UNARY = [n for n, v in OPCODES.items() if not n.endswith(("_end", "_begin")) and
         (v < OPCODES["unary_end"] or OPCODES["sf_begin"] <= v < OPCODES["sf_end"])]
BINARY = [n for n, v in OPCODES.items() if OPCODES["add"] <= v < OPCODES["binary_end"] and not n.endswith("_end")]


class Sym:
    """A node of the traced expression DAG: an operand, a constant or an op application."""

    __slots__ = ("kind", "op", "a", "b", "value", "index")
    __array_priority__ = 5000.0

    def __init__(self, kind, op=None, a=None, b=None, value=None, index=None):
        self.kind, self.op, self.a, self.b, self.value, self.index = kind, op, a, b, value, index

    @staticmethod
    def arg(i):
        return Sym("arg", index=i)

    @staticmethod
    def lift(x):
        if isinstance(x, Sym):
            return x
        if isinstance(x, numbers.Real):
            return Sym("const", value=float(x))
        raise UnsupportedOp(4, f"cannot trace a {type(x).__name__} inside a scalar function")

    def _bin(self, op, o, swap=False):
        o = Sym.lift(o)
        a, b = (o, self) if swap else (self, o)
        return Sym("op", op=op, a=a, b=b)

    def __add__(self, o): return self._bin("add", o)
    def __radd__(self, o): return self._bin("add", o, True)
    def __sub__(self, o): return self._bin("sub", o)
    def __rsub__(self, o): return self._bin("sub", o, True)
    def __mul__(self, o): return self._bin("mul", o)
    def __rmul__(self, o): return self._bin("mul", o, True)
    def __truediv__(self, o): return self._bin("div", o)
    def __rtruediv__(self, o): return self._bin("div", o, True)
    def __pow__(self, o): return self._bin("pow", o)
    def __rpow__(self, o): return self._bin("pow", o, True)
    def __neg__(self): return Sym("op", op="neg", a=self)
    def __pos__(self): return self
    def __abs__(self): return Sym("op", op="abs", a=self)

    def __bool__(self):
        raise UnsupportedOp(4, "data-dependent control flow (`if x > 0`, `a if c else b`, `and`/`or`) cannot be traced "
                               "into a device program: write the selection as ifelse(x > 0, a, b), max_, min_, abs or "
                               "sign — both branches are traced, the derivative is that of the branch taken")

    # comparisons are VALUES (1.0 / 0.0, derivative zero) that ifelse / arithmetic consume
    def __gt__(self, o): return self._bin("gt", o)
    def __lt__(self, o): return self._bin("lt", o)
    def __ge__(self, o): return self._bin("ge", o)
    def __le__(self, o): return self._bin("le", o)
    __hash__ = object.__hash__


class ScalarFn:
    """A catalogue primitive.  Callable on Syms (tracing), Python numbers (host math is NOT provided:
    raises) and scalar Nodes (device scalar rule through the broadcast kernels)."""

    def __init__(self, name):
        self.name = name
        self.__name__ = name
        self.opcode = OPCODES[name]
        self.arity = 2 if OPCODES["add"] <= self.opcode < OPCODES["binary_end"] else 1

    def __repr__(self):
        return f"<nabla.jl_b200 scalar fn {self.name}>"

    def __call__(self, *args):
        if len(args) != self.arity:
            raise TypeError(f"MethodError: {self.name} takes {self.arity} argument(s)")
        if any(isinstance(a, Sym) for a in args):
            a = Sym.lift(args[0])
            b = Sym.lift(args[1]) if self.arity == 2 else None
            return Sym("op", op=self.name, a=a, b=b)
        from . import rules  # scalar Nodes / device scalars: a 0-dim broadcast
        return rules.scalar_call(self, args)


def ifelse(cond, a, b):
    """``ifelse(cond, a, b)`` inside a traced scalar function: ``cond`` is a comparison of traced values (``x > 0``),
    both branches are traced, and per element the value AND the derivative are those of the branch taken — what
    ForwardDiff's duals give for ``x > 0 ? a : b`` under ``fmad`` (src/core.jl:285-292).  The branch not taken
    contributes exactly zero (no ``0 * Inf``): lowered to ``MASK(c, a) + MASKN(c, b)`` (include/nabla_cuda.h)."""
    if not isinstance(cond, Sym):
        if isinstance(cond, (bool, numbers.Real)):
            return a if cond else b
        raise UnsupportedOp(4, "ifelse: the condition must be a comparison of traced values")
    return Sym("op", op="add", a=Sym("op", op="mask", a=cond, b=Sym.lift(a)), b=Sym("op", op="maskn", a=cond, b=Sym.lift(b)))


def _is_const(s, v=None):
    return s.kind == "const" and (v is None or s.value == v)


def _simplify(s):
    """Peephole rewrites to fused catalogue ops.  logistic: 1/(1+exp(-x)) in any operand order
    (examples/mlp.jl:42, examples/vae.jl:32)."""
    if s.kind != "op":
        return s
    a = _simplify(s.a)
    b = _simplify(s.b) if s.b is not None else None
    s = Sym("op", op=s.op, a=a, b=b)
    den = None
    if s.op == "div" and _is_const(a, 1.0):
        den = b
    elif s.op == "inv":
        den = a
    if den is not None and den.kind == "op" and den.op == "add":
        e = den.b if _is_const(den.a, 1.0) else den.a if _is_const(den.b, 1.0) else None
        if e is not None and e.kind == "op" and e.op == "exp" and e.a.kind == "op" and e.a.op == "neg":
            return Sym("op", op="logistic", a=e.a.a)
    # x ^ const folded forms used by literal_pow
    if s.op == "pow" and _is_const(b, 2.0):
        return Sym("op", op="abs2", a=a)
    if s.op == "pow" and _is_const(b, 1.0):
        return a
    return s


class Program:
    """Linearised scalar program + its device handle (cached per context)."""

    __slots__ = ("nargs", "ops", "a", "b", "consts", "key", "name")

    def __init__(self, nargs, ops, a, b, consts, name="f"):
        self.nargs, self.ops, self.a, self.b, self.consts = nargs, tuple(ops), tuple(a), tuple(b), tuple(consts)
        self.key = (nargs, self.ops, self.a, self.b, self.consts)
        self.name = name

    def handle(self, ctx):
        h = ctx._progs.get(self.key)
        if h is None:
            n = len(self.ops)
            I = C.c_int32 * max(n, 1)
            D = C.c_double * max(len(self.consts), 1)
            out = C.c_void_p()
            check(ctx.L.nb_prog_create(self.nargs, n, I(*self.ops), I(*self.a), I(*self.b),
                                       len(self.consts), D(*self.consts), C.byref(out)))
            h = out
            ctx._progs[self.key] = h
        return h

    @property
    def single_op(self):
        return self.ops[0] if len(self.ops) == 1 else (OPCODES["identity"] if not self.ops else None)


_TRACE_CACHE = {}


_CODE_NAMES = {}
_KEY_DEPTH = 4


def _code_names(code):
    """Global / attribute names a code object (and the code objects nested in it) refers to."""
    names = _CODE_NAMES.get(code)
    if names is None:
        acc = list(code.co_names)
        for c in code.co_consts:
            if isinstance(c, types.CodeType):
                acc.extend(_code_names(c))
        names = _CODE_NAMES[code] = tuple(dict.fromkeys(acc))
    return names


def _value_key(v, depth):
    """Key component of a value a traced function captures (closure cell or global), or None when the value cannot
    be keyed SAFELY.  Numbers go in by value, catalogue primitives by name, modules / builtins / classes as the
    objects themselves (the key then keeps them alive, so no identity is ever reused), plain functions structurally
    (recursively).  Anything mutable or opaque — dicts, lists, arrays, instances — makes the function uncacheable:
    an identity key would go stale when the object is mutated or when its id is reused after it is freed."""
    if isinstance(v, bool):
        return ("b", v)
    if isinstance(v, numbers.Real):
        return ("n", float(v))
    if isinstance(v, ScalarFn):
        return ("prim", v.name)
    if isinstance(v, (types.ModuleType, types.BuiltinFunctionType, type)):
        return ("o", v)
    if isinstance(v, types.FunctionType) and depth < _KEY_DEPTH:
        k = _function_key(v, depth + 1)
        return None if k is None else ("f", k)
    return None


def _function_key(f, depth=0):
    code = getattr(f, "__code__", None)
    if code is None or getattr(f, "__defaults__", None) or getattr(f, "__kwdefaults__", None):
        return None
    parts = [code]
    for c in (getattr(f, "__closure__", None) or ()):
        try:
            k = _value_key(c.cell_contents, depth)
        except ValueError:   # empty cell
            return None
        if k is None:
            return None
        parts.append(k)
    g = getattr(f, "__globals__", None) or {}
    for name in _code_names(code):
        if name in g:      # (other names are attributes or builtins)
            k = _value_key(g[name], depth)
            if k is None:
                return None
            parts.append((name, k))
    return tuple(parts)


def _trace_key(f, nargs, const_args):
    """Cache key of a traced function: catalogue primitives by name; Python functions by code object plus the
    VALUES of everything the code can see — closure cells and referenced globals (`_value_key`) — so that a lambda
    re-created by the same source line with the same captured values hits the cache, while a rebound cell, a changed
    global or any captured mutable object never yields a stale program (such functions are simply re-traced)."""
    consts = tuple(sorted((i, float(v)) for i, v in const_args.items()))
    if isinstance(f, ScalarFn):
        return ("prim", f.name, nargs, consts)
    k = _function_key(f)
    if k is None:
        return None
    return ("fn", k, nargs, consts)


def compile_scalar(f, nargs, const_args=None):
    """Trace ``f`` on ``nargs`` symbolic operands -> Program.  ``const_args`` maps operand positions
    to Python numbers that are baked in as constants (non-differentiable literals such as ϵ); the
    remaining operands are renumbered 0..k-1 in order.  Traces are cached (`_trace_key`)."""
    const_args = const_args or {}
    key = _trace_key(f, nargs, const_args)
    if key is not None:
        hit = _TRACE_CACHE.get(key)
        if hit is not None:
            return hit[0]
    prog = _compile_scalar(f, nargs, const_args)
    if key is not None:
        if len(_TRACE_CACHE) > 4096:
            _TRACE_CACHE.clear()
        _TRACE_CACHE[key] = (prog,)
    return prog


def _compile_scalar(f, nargs, const_args):
    syms, k = [], 0
    for i in range(nargs):
        if i in const_args:
            syms.append(Sym("const", value=float(const_args[i])))
        else:
            syms.append(Sym.arg(k))
            k += 1
    if k == 0:
        raise TypeError("broadcast needs at least one array/Node operand")
    if k > NB_MAX_ARGS:
        raise UnsupportedOp(4, f"at most {NB_MAX_ARGS} array operands per broadcast (got {k})")
    if isinstance(f, ScalarFn):
        if f.arity != nargs:
            raise TypeError(f"MethodError: {f.name} takes {f.arity} argument(s), got {nargs}")
        root = Sym("op", op=f.name, a=syms[0], b=syms[1] if nargs == 2 else None)
    else:
        root = Sym.lift(f(*syms))
    return linearize(root, k, getattr(f, "__name__", "f"))


def linearize(root, k, name="f"):
    """Expression DAG over ``k`` operands -> Program (peephole rewrites, structural CSE, catalogue / size checks)."""
    root = _simplify(root)
    ops, ia, ib, consts = [], [], [], []
    memo = {}

    def const_index(v):
        for i, c in enumerate(consts):
            if c == v or (math.isnan(c) and math.isnan(v)):
                return -(i + 1)
        if len(consts) >= NB_MAX_CONSTS:
            raise UnsupportedOp(4, "scalar function uses too many distinct constants")
        consts.append(v)
        return -len(consts)

    def emit(s):
        if s.kind == "arg":
            return s.index
        if s.kind == "const":
            return const_index(s.value)
        a = emit(s.a)
        b = emit(s.b) if s.b is not None else 0
        key = (s.op, a, b)  # structural CSE: identical sub-expressions are emitted once
        if key in memo:
            return memo[key]
        if s.op not in OPCODES:
            raise UnsupportedOp(4, f"scalar primitive `{s.op}` is not in the device catalogue")
        if len(ops) >= NB_MAX_INSTR:
            raise UnsupportedOp(4, f"scalar function is longer than {NB_MAX_INSTR} instructions")
        ops.append(OPCODES[s.op])
        ia.append(a)
        ib.append(b)
        memo[key] = k + len(ops) - 1
        return memo[key]

    r = emit(root)
    if root.kind != "op":  # f(x) = x or f(x) = c : route through identity / add-zero
        if root.kind == "arg":
            if r != 0:
                ops.append(OPCODES["identity"]); ia.append(r); ib.append(0)
        else:
            ops.append(OPCODES["add"]); ia.append(r); ib.append(const_index(0.0))
            # a constant function still needs operand 0 to define the shape
    return Program(k, ops, ia, ib, consts, name)


_OPNAME = {v: n for n, v in OPCODES.items() if not n.endswith(("_end", "_begin"))}


def symbolic(prog, syms):
    """The value of ``prog`` on symbolic operands ``syms`` (one Sym per program operand) as an expression DAG — used to
    compose the programs of adjacent broadcast Branches into one (ewfusion.py)."""
    vals = list(syms)
    if len(vals) != prog.nargs:
        raise ValueError("symbolic: operand count mismatch")

    def slot(i):
        return vals[i] if i >= 0 else Sym("const", value=float(prog.consts[-i - 1]))
    for op, a, b in zip(prog.ops, prog.a, prog.b):
        name = _OPNAME[op]
        binary = OPCODES["add"] <= op < OPCODES["binary_end"]
        vals.append(Sym("op", op=name, a=slot(a), b=slot(b) if binary else None))
    return vals[-1] if prog.ops else vals[0]


def _make_catalogue():
    out = {}
    for n in UNARY + BINARY:
        out[n] = ScalarFn(n)
    return out


CATALOGUE = _make_catalogue()
