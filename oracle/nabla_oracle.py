"""CPU oracle: a NumPy restatement of Nabla.jl's reverse sweep for array sensitivities.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``nabla.jl_b200/``) imports this
module; it is imported by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` as the checker and the timed CPU baseline.

PARITY STATUS.  The reference is pure Julia and Julia is not installed in the build container or
on the GPU box, so this file cannot be pinned against outputs of the reference itself.  It IS
pinned (tests/test_oracle_golden.py) against every golden vector / known-answer case the reference
holds for this path (docs/src/pages/autodiff.md:203-286, test/core.jl:104-173,
test/sensitivities/functional/functional.jl:6-12,208-227, test/sensitivities/functional/
reducedim.jl:4-61, docs/src/index.md:25-39) and against the reference's own test *method*
(5-point central finite differences, src/finite_differencing.jl:18-25,106-120, fresh-slot and
pre-filled-slot variants).  Whole-model gradients (examples/mlp.jl, examples/vae.jl) have no
reference fixture anywhere: **parity unpinned** for those; they are checked by FD only.

Third-party arithmetic that is NOT in /root/reference and is restated here from its published
definitions: ChainRules.jl =1.35.2 (Project.toml:22) rrules for ``*``, ``sum``, ``mean``,
``dot``, ``adjoint``/``transpose``, ``-``, ``BLAS.gemm/gemv`` and the ``@scalar_rule`` table;
ChainRulesCore (Thunk / InplaceableThunk / add!! / unthunk); ForwardDiff >=0.10.12 + DiffRules
(dual-number arithmetic behind ``fmad``, src/core.jl:285-301).

All citations are relative to /root/reference.  Arrays follow Julia semantics: shapes broadcast
from the LEADING dimension (a (m,) vector broadcasts against (m, n) as a column), reductions
with ``dims`` keep the reduced dimensions with extent 1, and positions on the tape are 1-based.
"""
from __future__ import annotations

import math
import numpy as np

__all__ = [
    "Tape", "Node", "Leaf", "Branch", "Arg", "unbox", "pos", "tape", "nabla", "grad",
    "propagate", "reverse_tape", "broadcast", "map_", "mapreduce", "mapfoldl", "mapfoldr",
    "sum", "mean", "dot", "matmul", "transpose", "adjoint", "gemm", "gemv", "fmad",
    "broadcastsum", "broadcastsum_inplace", "Dual", "Ref", "Thunk", "InplaceableThunk",
    "unthunk", "add_bang_bang", "check_errs", "oneslike", "zeroslike",
    "hcat", "vcat", "getindex", "reshape", "fill", "checkpoint", "R", "colon",
    "symm", "symv", "trmm", "trmv", "trsm", "trsv",
    "kron", "plus_I", "copy", "ifelse",
]

_pysum = sum


# --------------------------------------------------------------------------------------------
# Julia-style broadcasting helpers
# --------------------------------------------------------------------------------------------
class Ref:
    """``Ref(x)``: a 0-dim container that broadcasts as a scalar (test/.../functional.jl:181)."""

    def __init__(self, x):
        self.x = x


def _is_array(x):
    return isinstance(x, np.ndarray)


def _is_scalar(x):
    return isinstance(x, (int, float, np.floating, np.integer)) and not isinstance(x, bool)


def _jl_shape(x):
    if isinstance(x, Ref):
        return ()
    return tuple(np.shape(x))


def jl_broadcast_shape(*shapes):
    """``Broadcast.combine_axes``: align from the leading dim, pad with trailing 1s."""
    nd = max((len(s) for s in shapes), default=0)
    out = []
    for d in range(nd):
        ext = 1
        for s in shapes:
            e = s[d] if d < len(s) else 1
            if e != 1:
                if ext != 1 and ext != e:
                    raise ValueError(f"DimensionMismatch: arrays could not be broadcast: {shapes}")
                ext = e
        out.append(ext)
    return tuple(out)


def _pad(x, nd):
    """View ``x`` with trailing singleton dims so NumPy's broadcasting equals Julia's."""
    if isinstance(x, Ref):
        x = x.x
    if _is_array(x):
        if x.ndim < nd:
            return x.reshape(x.shape + (1,) * (nd - x.ndim))
        return x
    return x


def _sum_to_shape(tmp, zshape):
    """``sum!(z, tmp)``: sum over every dim where ``z`` has extent 1 and ``tmp`` does not."""
    zpad = tuple(zshape) + (1,) * (tmp.ndim - len(zshape))
    axes = tuple(d for d in range(tmp.ndim) if zpad[d] == 1 and tmp.shape[d] != 1)
    out = tmp.sum(axis=axes, keepdims=True) if axes else tmp
    return out.reshape(zshape)


# --------------------------------------------------------------------------------------------
# Dual numbers (ForwardDiff.Dual with one partial), vectorised over NumPy arrays
# --------------------------------------------------------------------------------------------
class Dual:
    """value + one partial.  Follows ForwardDiff/DiffRules formulas (src/core.jl:285-292)."""

    __slots__ = ("v", "d")
    __array_priority__ = 1000.0

    def __init__(self, v, d):
        self.v = v
        self.d = d

    @staticmethod
    def _split(o):
        return (o.v, o.d) if isinstance(o, Dual) else (o, None)

    def __add__(self, o):
        ov, od = Dual._split(o)
        return Dual(self.v + ov, self.d if od is None else self.d + od)

    __radd__ = __add__

    def __sub__(self, o):
        ov, od = Dual._split(o)
        return Dual(self.v - ov, self.d if od is None else self.d - od)

    def __rsub__(self, o):
        return Dual(o - self.v, -self.d)

    def __neg__(self):
        return Dual(-self.v, -self.d)

    def __pos__(self):
        return self

    def __mul__(self, o):
        ov, od = Dual._split(o)
        if od is None:
            return Dual(self.v * ov, self.d * ov)
        return Dual(self.v * ov, self.d * ov + self.v * od)

    __rmul__ = __mul__

    def __truediv__(self, o):
        ov, od = Dual._split(o)
        if od is None:
            return Dual(self.v / ov, self.d / ov)
        q = self.v / ov
        return Dual(q, (self.d - q * od) / ov)

    def __rtruediv__(self, o):
        q = o / self.v
        return Dual(q, -q / self.v * self.d)

    def __pow__(self, o):
        ov, od = Dual._split(o)
        if od is None:
            if isinstance(ov, (int, np.integer)):  # Base.literal_pow path: x .^ 2
                return Dual(self.v ** ov, ov * self.v ** (ov - 1) * self.d)
            return Dual(self.v ** ov, ov * self.v ** (ov - 1) * self.d)
        val = self.v ** ov
        return Dual(val, ov * self.v ** (ov - 1) * self.d + val * np.log(self.v) * od)

    def __rpow__(self, o):
        val = o ** self.v
        return Dual(val, val * math.log(o) * self.d)

    # comparisons act on values (ForwardDiff semantics) so data-dependent branches work
    def __lt__(self, o):
        return self.v < Dual._split(o)[0]

    def __le__(self, o):
        return self.v <= Dual._split(o)[0]

    def __gt__(self, o):
        return self.v > Dual._split(o)[0]

    def __ge__(self, o):
        return self.v >= Dual._split(o)[0]


def fmad(f, xs, n):
    """∂f/∂x_n with the other operands held constant: ``fmad(f, x, Val{n})`` (core.jl:285-292).

    ``n`` is 0-based here.  Returns an array/scalar broadcastable against the operands."""
    args = list(xs)
    args[n] = Dual(xs[n], 1.0)
    out = f(*args)
    if isinstance(out, Dual):
        return out.d
    return 0.0  # f does not depend on x_n


# --------------------------------------------------------------------------------------------
# Scalar function catalogue (DiffRules / ChainRules @scalar_rule formulas; list pinned at
# test/runtests.jl:28-52).  Each entry works on floats, ndarrays, Duals and scalar Nodes.
# --------------------------------------------------------------------------------------------
_D2R = math.pi / 180.0
_R2D = 180.0 / math.pi
_LN2 = math.log(2.0)
_LN10 = math.log(10.0)


def _sec(x):
    return 1.0 / np.cos(x)


def _csc(x):
    return 1.0 / np.sin(x)


def _cot(x):
    return 1.0 / np.tan(x)


def _sech(x):
    return 1.0 / np.cosh(x)


def _csch(x):
    return 1.0 / np.sinh(x)


def _coth(x):
    return 1.0 / np.tanh(x)


def _sinpi(x):
    return np.sin(np.pi * x)


def _cospi(x):
    return np.cos(np.pi * x)


import scipy.special as _sp  # noqa: E402  (special functions: values only, the rules are written out below)

_2_SQRTPI = 2.0 / math.sqrt(math.pi)


def _polygamma_any(n, x):
    """polygamma(n, x) for any real x (scipy's is defined for x > 0): reflection for x <= 0,
    psi1(x) = pi^2/sin^2(pi x) - psi1(1-x),  psi2(x) = psi2(1-x) - 2 pi^3 cos(pi x)/sin^3(pi x)."""
    x = np.asarray(x, dtype=np.float64)
    pos = x > 0
    xp = np.where(pos, x, 1.0 - x)
    v = _sp.polygamma(n, xp)
    with np.errstate(all="ignore"):
        s, c = np.sin(np.pi * x), np.cos(np.pi * x)
        if n == 1:
            r = np.pi ** 2 / (s * s) - v
        else:
            r = v - 2.0 * np.pi ** 3 * c / (s ** 3)
    out = np.where(pos, v, r)
    return out if out.ndim else float(out)


def _invdigamma(y):
    """Inverse of digamma on x > 0: Newton from Minka's closed-form start ("Estimating a Dirichlet distribution",
    2000), the scheme SpecialFunctions.jl's invdigamma uses (its loop stops at |dx| <= 1e-12; here it runs on to
    the last bit, the results agree to that tolerance)."""
    y = np.asarray(y, dtype=np.float64)
    with np.errstate(all="ignore"):
        x = np.where(y >= -2.22, np.exp(y) + 0.5, -1.0 / (y - _sp.digamma(1.0)))
        for _ in range(40):
            xn = x - (_sp.digamma(x) - y) / _sp.polygamma(1, x)
            xn = np.where(xn > 0.0, xn, 0.5 * x)
            if np.all(np.abs(xn - x) <= 4e-16 * xn):
                x = xn
                break
            x = xn
    return x if x.ndim else float(x)


_ZERO = lambda x: 0.0 * x  # noqa: E731   (derivative of a piecewise-constant function, as ForwardDiff gives)


def ifelse(cond, a, b):
    """Julia ``ifelse(cond, a, b)`` / ``cond ? a : b`` element-wise: value and partial of the branch taken — comparisons
    of ``Dual``s act on the values (ForwardDiff semantics, src/core.jl:285-292)."""
    if isinstance(cond, Dual):
        cond = cond.v
    cond = np.asarray(cond) != 0
    av, ad = Dual._split(a)
    bv, bd = Dual._split(b)
    val = np.where(cond, av, bv)
    if ad is None and bd is None:
        return val if val.ndim else val[()]
    d = np.where(cond, 0.0 if ad is None else ad, 0.0 if bd is None else bd)
    return Dual(val, d)


_UNARY = {
    # name: (f, f')  -- SURVEY.md §8(c) catalogue
    "sign": (np.sign, _ZERO),
    "step": (lambda x: (x > 0) * 1.0, _ZERO),
    "floor": (np.floor, _ZERO),
    "ceil": (np.ceil, _ZERO),
    "round": (np.rint, _ZERO),
    "trunc": (np.trunc, _ZERO),
    "identity": (lambda x: x, lambda x: 1.0 + 0.0 * x),
    "neg": (lambda x: -x, lambda x: -1.0 + 0.0 * x),
    "abs": (np.abs, np.sign),
    "abs2": (lambda x: x * x, lambda x: 2.0 * x),
    "sqrt": (np.sqrt, lambda x: 0.5 / np.sqrt(x)),
    "cbrt": (np.cbrt, lambda x: 1.0 / (3.0 * np.cbrt(x) ** 2)),
    "inv": (lambda x: 1.0 / x, lambda x: -1.0 / (x * x)),
    "exp": (np.exp, np.exp),
    "exp2": (np.exp2, lambda x: _LN2 * np.exp2(x)),
    "exp10": (lambda x: np.power(10.0, x), lambda x: _LN10 * np.power(10.0, x)),
    "expm1": (np.expm1, np.exp),
    "log": (np.log, lambda x: 1.0 / x),
    "log2": (np.log2, lambda x: 1.0 / (x * _LN2)),
    "log10": (np.log10, lambda x: 1.0 / (x * _LN10)),
    "log1p": (np.log1p, lambda x: 1.0 / (1.0 + x)),
    "sin": (np.sin, np.cos),
    "cos": (np.cos, lambda x: -np.sin(x)),
    "tan": (np.tan, lambda x: 1.0 + np.tan(x) ** 2),
    "sec": (_sec, lambda x: _sec(x) * np.tan(x)),
    "csc": (_csc, lambda x: -_csc(x) * _cot(x)),
    "cot": (_cot, lambda x: -(1.0 + _cot(x) ** 2)),
    "sinh": (np.sinh, np.cosh),
    "cosh": (np.cosh, np.sinh),
    "tanh": (np.tanh, lambda x: 1.0 - np.tanh(x) ** 2),
    "sech": (_sech, lambda x: -np.tanh(x) * _sech(x)),
    "csch": (_csch, lambda x: -_coth(x) * _csch(x)),
    "coth": (_coth, lambda x: -(_csch(x) ** 2)),
    "asin": (np.arcsin, lambda x: 1.0 / np.sqrt(1.0 - x * x)),
    "acos": (np.arccos, lambda x: -1.0 / np.sqrt(1.0 - x * x)),
    "atan": (np.arctan, lambda x: 1.0 / (1.0 + x * x)),
    "asec": (lambda x: np.arccos(1.0 / x), lambda x: 1.0 / (np.abs(x) * np.sqrt(x * x - 1.0))),
    "acsc": (lambda x: np.arcsin(1.0 / x), lambda x: -1.0 / (np.abs(x) * np.sqrt(x * x - 1.0))),
    "acot": (lambda x: np.arctan(1.0 / x), lambda x: -1.0 / (1.0 + x * x)),
    "asinh": (np.arcsinh, lambda x: 1.0 / np.sqrt(x * x + 1.0)),
    "acosh": (np.arccosh, lambda x: 1.0 / np.sqrt(x * x - 1.0)),
    "atanh": (np.arctanh, lambda x: 1.0 / (1.0 - x * x)),
    "asech": (lambda x: np.arccosh(1.0 / x), lambda x: -1.0 / (x * np.sqrt(1.0 - x * x))),
    "acsch": (lambda x: np.arcsinh(1.0 / x), lambda x: -1.0 / (np.abs(x) * np.sqrt(1.0 + x * x))),
    "acoth": (lambda x: np.arctanh(1.0 / x), lambda x: 1.0 / (1.0 - x * x)),
    "sind": (lambda x: np.sin(_D2R * x), lambda x: _D2R * np.cos(_D2R * x)),
    "cosd": (lambda x: np.cos(_D2R * x), lambda x: -_D2R * np.sin(_D2R * x)),
    "tand": (lambda x: np.tan(_D2R * x), lambda x: _D2R * (1.0 + np.tan(_D2R * x) ** 2)),
    "secd": (lambda x: _sec(_D2R * x), lambda x: _D2R * _sec(_D2R * x) * np.tan(_D2R * x)),
    "cscd": (lambda x: _csc(_D2R * x), lambda x: -_D2R * _csc(_D2R * x) * _cot(_D2R * x)),
    "cotd": (lambda x: _cot(_D2R * x), lambda x: -_D2R * (1.0 + _cot(_D2R * x) ** 2)),
    "asind": (lambda x: _R2D * np.arcsin(x), lambda x: _R2D / np.sqrt(1.0 - x * x)),
    "acosd": (lambda x: _R2D * np.arccos(x), lambda x: -_R2D / np.sqrt(1.0 - x * x)),
    "atand": (lambda x: _R2D * np.arctan(x), lambda x: _R2D / (1.0 + x * x)),
    "asecd": (lambda x: _R2D * np.arccos(1.0 / x),
              lambda x: _R2D / (np.abs(x) * np.sqrt(x * x - 1.0))),
    "acscd": (lambda x: _R2D * np.arcsin(1.0 / x),
              lambda x: -_R2D / (np.abs(x) * np.sqrt(x * x - 1.0))),
    "acotd": (lambda x: _R2D * np.arctan(1.0 / x), lambda x: -_R2D / (1.0 + x * x)),
    "sinpi": (_sinpi, lambda x: np.pi * _cospi(x)),
    "cospi": (_cospi, lambda x: -np.pi * _sinpi(x)),
    "deg2rad": (lambda x: _D2R * x, lambda x: _D2R + 0.0 * x),
    "rad2deg": (lambda x: _R2D * x, lambda x: _R2D + 0.0 * x),
    # SpecialFunctions.jl (test/runtests.jl:34-36), ChainRules @scalar_rule formulas; values from scipy.special
    "erf": (lambda x: _sp.erf(x), lambda x: _2_SQRTPI * np.exp(-x * x)),
    "erfc": (lambda x: _sp.erfc(x), lambda x: -_2_SQRTPI * np.exp(-x * x)),
    "erfcx": (lambda x: _sp.erfcx(x), lambda x: 2.0 * x * _sp.erfcx(x) - _2_SQRTPI),
    "erfinv": (lambda x: _sp.erfinv(x), lambda x: np.exp(_sp.erfinv(x) ** 2) / _2_SQRTPI),
    "erfcinv": (lambda x: _sp.erfcinv(x), lambda x: -np.exp(_sp.erfcinv(x) ** 2) / _2_SQRTPI),
    "gamma": (lambda x: _sp.gamma(x), lambda x: _sp.gamma(x) * _sp.digamma(x)),
    "digamma": (lambda x: _sp.digamma(x), lambda x: _polygamma_any(1, x)),
    "trigamma": (lambda x: _polygamma_any(1, x), lambda x: _polygamma_any(2, x)),
    "besselj0": (lambda x: _sp.j0(x), lambda x: -_sp.j1(x)),
    "besselj1": (lambda x: _sp.j1(x), lambda x: (_sp.j0(x) - _sp.jv(2, x)) / 2.0),
    "bessely0": (lambda x: _sp.y0(x), lambda x: -_sp.y1(x)),
    "bessely1": (lambda x: _sp.y1(x), lambda x: (_sp.y0(x) - _sp.yv(2, x)) / 2.0),
    "airyai": (lambda x: _sp.airy(x)[0], lambda x: _sp.airy(x)[1]),
    "airyaiprime": (lambda x: _sp.airy(x)[1], lambda x: x * _sp.airy(x)[0]),          # Ai'' = x Ai
    "airybi": (lambda x: _sp.airy(x)[2], lambda x: _sp.airy(x)[3]),
    "airybiprime": (lambda x: _sp.airy(x)[3], lambda x: x * _sp.airy(x)[2]),
    "dawson": (lambda x: _sp.dawsn(x), lambda x: 1.0 - 2.0 * x * _sp.dawsn(x)),
    "erfi": (lambda x: _sp.erfi(x), lambda x: _2_SQRTPI * np.exp(x * x)),
    "invdigamma": (_invdigamma, lambda x: 1.0 / _sp.polygamma(1, _invdigamma(x))),
}

_BINARY = {
    # name: (f, ∂x, ∂y)
    "add": (lambda x, y: x + y, lambda x, y: 1.0, lambda x, y: 1.0),
    "sub": (lambda x, y: x - y, lambda x, y: 1.0, lambda x, y: -1.0),
    "mul": (lambda x, y: x * y, lambda x, y: y, lambda x, y: x),
    "div": (lambda x, y: x / y, lambda x, y: 1.0 / y, lambda x, y: -x / (y * y)),
    "ldiv": (lambda x, y: y / x, lambda x, y: -y / (x * x), lambda x, y: 1.0 / x),
    "pow": (lambda x, y: np.power(x, y), lambda x, y: y * np.power(x, y - 1.0),
            lambda x, y: np.power(x, y) * np.log(x)),
    "hypot": (np.hypot, lambda x, y: x / np.hypot(x, y), lambda x, y: y / np.hypot(x, y)),
    "max": (np.maximum, lambda x, y: (x > y) * 1.0, lambda x, y: (x <= y) * 1.0),
    "min": (np.minimum, lambda x, y: (x < y) * 1.0, lambda x, y: (x >= y) * 1.0),
    "atan2": (np.arctan2, lambda x, y: y / (x * x + y * y), lambda x, y: -x / (x * x + y * y)),
    "gt": (lambda x, y: (x > y) * 1.0, lambda x, y: 0.0 * (x + y), lambda x, y: 0.0 * (x + y)),
    "lt": (lambda x, y: (x < y) * 1.0, lambda x, y: 0.0 * (x + y), lambda x, y: 0.0 * (x + y)),
    "ge": (lambda x, y: (x >= y) * 1.0, lambda x, y: 0.0 * (x + y), lambda x, y: 0.0 * (x + y)),
    "le": (lambda x, y: (x <= y) * 1.0, lambda x, y: 0.0 * (x + y), lambda x, y: 0.0 * (x + y)),
    "mask": (lambda c, v: np.where(np.asarray(c) != 0, v, 0.0), lambda c, v: 0.0 * (c + v),
             lambda c, v: (np.asarray(c) != 0) * 1.0 + 0.0 * v),
    "maskn": (lambda c, v: np.where(np.asarray(c) == 0, v, 0.0), lambda c, v: 0.0 * (c + v),
              lambda c, v: (np.asarray(c) == 0) * 1.0 + 0.0 * v),
    "beta": (lambda x, y: _sp.beta(x, y), lambda x, y: _sp.beta(x, y) * (_sp.digamma(x) - _sp.digamma(x + y)),
             lambda x, y: _sp.beta(x, y) * (_sp.digamma(y) - _sp.digamma(x + y))),
}

_NAN = lambda x, y: np.nan + 0.0 * (x + y)  # noqa: E731


def _polygamma_int(n, x):
    """polygamma(n, x) for integer n >= 0 and any real x that is not a pole: the defining sum
    (-1)^(n+1) n! sum_k (x+k)^-(n+1) shifts x to the positive axis, where scipy's polygamma applies."""
    n, x = np.broadcast_arrays(np.asarray(n, dtype=np.float64), np.asarray(x, dtype=np.float64))
    out = np.empty(x.shape)
    for idx in np.ndindex(x.shape):
        ni, xi, acc = int(n[idx]), float(x[idx]), 0.0
        while xi <= 0.0:
            acc += (1.0 / xi if ni == 0 else xi ** -(ni + 1.0))
            xi += 1.0
        fact = math.factorial(ni)
        out[idx] = (_sp.digamma(xi) - acc) if ni == 0 else (_sp.polygamma(ni, xi) + (-1.0) ** (ni + 1) * fact * acc)
    return out if out.ndim else float(out)


# Differentiable in the second argument only (test/runtests.jl:43-47; DiffRules gives NaN for the order):
# name: (f, d/d(order) = NaN, d/dx)
_SECOND_ARG_ONLY = {
    "besselj": (lambda n, x: _sp.jv(n, x), _NAN, lambda n, x: (_sp.jv(n - 1, x) - _sp.jv(n + 1, x)) / 2.0),
    "bessely": (lambda n, x: _sp.yv(n, x), _NAN, lambda n, x: (_sp.yv(n - 1, x) - _sp.yv(n + 1, x)) / 2.0),
    "besseli": (lambda n, x: _sp.iv(n, x), _NAN, lambda n, x: (_sp.iv(n - 1, x) + _sp.iv(n + 1, x)) / 2.0),
    "besselk": (lambda n, x: _sp.kv(n, x), _NAN, lambda n, x: -(_sp.kv(n - 1, x) + _sp.kv(n + 1, x)) / 2.0),
    "polygamma": (_polygamma_int, _NAN, lambda n, x: _polygamma_int(np.asarray(n) + 1, x)),
}


class ScalarFn:
    """A named scalar function that accepts floats, ndarrays, Duals and scalar Nodes."""

    def __init__(self, name, arity):
        self.name = name
        self.arity = arity
        self.__name__ = name
        if arity == 1:
            self.f, self.df = _UNARY[name]
        else:
            self.f, self.dfx, self.dfy = (_BINARY.get(name) or _SECOND_ARG_ONLY[name])

    def __repr__(self):
        return f"<oracle scalar fn {self.name}>"

    def __call__(self, *args):
        if any(isinstance(a, Node) for a in args):
            return _scalar_rule_call(self, args)
        if self.arity == 1:
            (x,) = args
            if isinstance(x, Dual):
                return Dual(self.f(x.v), self.df(x.v) * x.d)
            return self.f(x)
        x, y = args
        xd, yd = isinstance(x, Dual), isinstance(y, Dual)
        if not (xd or yd):
            return self.f(x, y)
        xv = x.v if xd else x
        yv = y.v if yd else y
        d = 0.0
        if xd:
            d = d + self.dfx(xv, yv) * x.d
        if yd:
            d = d + self.dfy(xv, yv) * y.d
        return Dual(self.f(xv, yv), d)


for _n in _UNARY:
    globals()["abs_" if _n == "abs" else _n] = ScalarFn(_n, 1)
for _n in _BINARY:
    globals()["max_" if _n == "max" else "min_" if _n == "min" else "pow_" if _n == "pow" else _n] = \
        ScalarFn(_n, 2)
for _n in _SECOND_ARG_ONLY:
    globals()[_n] = ScalarFn(_n, 2)
UNARY_NAMES = tuple(_UNARY)
BINARY_NAMES = tuple(_BINARY)
SECOND_ARG_ONLY_NAMES = tuple(_SECOND_ARG_ONLY)


def scalar_fn(name):
    return ScalarFn(name, 1 if name in _UNARY else 2)  # (binary: _BINARY or _SECOND_ARG_ONLY)


# --------------------------------------------------------------------------------------------
# ChainRulesCore tangents: Thunk / InplaceableThunk / unthunk / add!!
# --------------------------------------------------------------------------------------------
class Thunk:
    """``@thunk expr``: recomputed on every ``unthunk`` (ChainRulesCore semantics)."""

    def __init__(self, fn):
        self.fn = fn

    def unthunk(self):
        return self.fn()


class InplaceableThunk(Thunk):
    """``InplaceableThunk(add!, val)``: ``add!!(x̄, t)`` runs ``add!(x̄)`` on mutable ``x̄``."""

    def __init__(self, add, fn):
        super().__init__(fn)
        self.add = add


def unthunk(x):
    while isinstance(x, Thunk):
        x = x.unthunk()
    return x


def add_bang_bang(xbar, fresh):
    """``ChainRulesCore.add!!`` as used by the default accumulate rule (core.jl:203-205)."""
    if isinstance(xbar, Thunk):
        return unthunk(xbar) + unthunk(fresh)
    if _is_array(xbar) and xbar.flags.writeable:
        if isinstance(fresh, InplaceableThunk):
            fresh.add(xbar)
            return xbar
        xbar += unthunk(fresh)  # in place: pinned by test/core.jl:211-224
        return xbar
    return xbar + unthunk(fresh)


# --------------------------------------------------------------------------------------------
# Graph core (src/core.jl:6-229)
# --------------------------------------------------------------------------------------------
class _Undef:
    def __repr__(self):
        return "#undef"


_UNDEF = _Undef()


class Tape:
    """``Tape`` (core.jl:10-35): a 1-based vector of Any; reads ``unthunk`` (core.jl:27)."""

    def __init__(self, n=None):
        self.tape = [] if n is None else [_UNDEF] * n

    def __len__(self):
        return len(self.tape)

    def push(self, node):
        self.tape.append(node)
        return self

    def isassigned(self, n):
        if isinstance(n, Node):
            n = n.pos
        return 1 <= n <= len(self.tape) and self.tape[n - 1] is not _UNDEF

    def __getitem__(self, n):
        if isinstance(n, Node):
            n = n.pos
        v = self.tape[n - 1]
        if v is _UNDEF:
            raise IndexError(f"UndefRefError: access to undefined reference at tape[{n}]")
        return unthunk(v)

    def __setitem__(self, n, x):
        if isinstance(n, Node):
            n = n.pos
        self.tape[n - 1] = x


class Node:
    """``abstract type Node{T}`` (core.jl:7) + operator overloads generated by the rule surface."""

    __array_priority__ = 2000.0
    val = None
    tape = None
    pos = -1

    @property
    def shape(self):
        return _jl_shape(self.val)

    # arithmetic sugar (Julia's `+`, `-`, `*`, `/`, `^`, adjoint `'`)
    def __add__(self, o):
        return _plus(self, o)

    def __radd__(self, o):
        return _plus(o, self)

    def __sub__(self, o):
        return _minus(self, o)

    def __rsub__(self, o):
        return _minus(o, self)

    def __mul__(self, o):
        return _times(self, o)

    def __rmul__(self, o):
        return _times(o, self)

    def __matmul__(self, o):
        return matmul(self, o)

    def __rmatmul__(self, o):
        return matmul(o, self)

    def __truediv__(self, o):
        return _divide(self, o)

    def __rtruediv__(self, o):
        return _divide(o, self)

    def __pow__(self, o):
        return _power(self, o)

    def __neg__(self):
        return _negate(self)

    @property
    def T(self):
        return adjoint(self)


class Leaf(Node):
    """``Leaf(tape, val)`` (core.jl:48-57)."""

    def __init__(self, tape, val):
        self.val = val
        self.tape = tape
        self.pos = len(tape) + 1
        tape.push(self)


class Branch(Node):
    """``Branch(f, args, tape; kwargs...)`` (core.jl:75-95, chainrules.jl:179-191)."""

    def __init__(self, f, args, tape, **kwargs):
        unboxed = tuple(unbox(a) for a in args)
        if f.rrule is not None:
            self.val, self.pullback = f.rrule(*unboxed, **kwargs)
        else:
            self.val = f.forward(*unboxed, **kwargs)
            self.pullback = None
        self.f = f
        self.args = tuple(args)
        self.kwargs = kwargs
        self.tape = tape
        self.pos = len(tape) + 1
        tape.push(self)


class Arg:
    """``Arg{N}`` (core.jl:177); N is 1-based as in the reference."""

    def __init__(self, n):
        self.n = n


def unbox(x):
    return x.val if isinstance(x, Node) else x


def pos(x):
    return x.pos if isinstance(x, Node) else -1


def tape(x):
    return x.tape


def oneslike(x):
    return np.ones_like(x) if _is_array(x) else type(x)(1)


def zeroslike(x):
    return np.zeros_like(x) if _is_array(x) else type(x)(0)


def _zero(x):
    if _is_array(x):
        return np.zeros_like(x)
    if isinstance(x, Ref):
        return 0.0
    return type(x)(0)


def propagate_branch(y, rvs_tape):
    """``propagate(y::Branch, rvs_tape)`` (core.jl:140-158)."""
    ybar = rvs_tape[y]
    d_tape = rvs_tape.tape
    f, args, kwargs = y.f, y.args, y.kwargs
    xs = tuple(unbox(a) for a in args)
    xids = tuple(pos(a) for a in args)
    p = f.preprocess(y, ybar, *args)
    for j, xid in enumerate(xids):
        if xid > 0:
            if d_tape[xid - 1] is not _UNDEF:
                d_tape[xid - 1] = f.grad_acc(d_tape[xid - 1], j + 1, p, y.val, ybar, *xs, **kwargs)
            else:
                d_tape[xid - 1] = f.grad(j + 1, p, y.val, ybar, *xs, **kwargs)
    return None


def propagate(fwd_tape, rvs_tape):
    """``propagate(fwd_tape, rvs_tape)`` (core.jl:160-166): strictly descending index."""
    n = len(rvs_tape)
    for delta in range(n, 0, -1):
        if rvs_tape.tape[delta - 1] is not _UNDEF:
            node = fwd_tape.tape[delta - 1]
            if isinstance(node, Branch):
                propagate_branch(node, rvs_tape)
    return rvs_tape


def reverse_tape(y, ybar):
    """core.jl:170-174."""
    t = Tape(y.pos)
    t[len(t)] = ybar
    return t


def nabla(f_or_node, ybar=None, *, get_output=False):
    """``∇`` in both of the reference's forms.

    ``nabla(y::Node[, ȳ])``  -> reverse tape (core.jl:200-201; scalar ``y`` is seeded with one).
    ``nabla(f; get_output)`` -> function returning the gradient tuple (core.jl:214-229)."""
    if isinstance(f_or_node, Node):
        y = f_or_node
        if ybar is None:
            if not (_is_scalar(y.val) or (_is_array(y.val) and y.val.ndim == 0)):
                raise TypeError("MethodError: no method matching ∇(::Node{<:AbstractArray}); "
                                "array outputs need an explicit ȳ (core.jl:200-201)")
            ybar = oneslike(y.val)
        return propagate(y.tape, reverse_tape(y, ybar))
    f = f_or_node

    def gradient_fn(*args, **kwargs):
        t = Tape()
        args_ = [Leaf(t, a) for a in args]
        y = f(*args_, **kwargs)
        if isinstance(y, Node):
            df = nabla(y)
            out = tuple(df[a_] if df.isassigned(a_) else _zero(a) for a_, a in zip(args_, args))
        else:
            out = tuple(_zero(a) for a in args)
        out = tuple(unthunk(o) for o in out)
        return (y, out) if get_output else out

    return gradient_fn


grad = nabla


# --------------------------------------------------------------------------------------------
# Rule surface: Nabla-native rules (forward + ∇ per arg) and the ChainRules bridge (rrule)
# --------------------------------------------------------------------------------------------
class Primitive:
    """One intercepted function.  ``grads[N]`` is ``∇(f, Arg{N}, p, y, ȳ, xs...)``; ``rrule``
    marks a ChainRules-bridged op whose ``preprocess`` is ``pullback(ȳ)`` and whose ``∇`` is
    ``p[N+1]`` (chainrules.jl:233-290)."""

    def __init__(self, name, forward=None, grads=None, rrule=None, preprocess=None):
        self.name = name
        self.forward = forward
        self.grads = grads or {}
        self.rrule = rrule
        self._preprocess = preprocess

    def __repr__(self):
        return f"<oracle primitive {self.name}>"

    def preprocess(self, y, ybar, *args):
        if self.rrule is not None:
            return y.pullback(ybar)
        if self._preprocess is not None:
            return self._preprocess(y.val, ybar, *[unbox(a) for a in args])
        return ()  # sensitivity.jl:220

    def grad(self, n, p, y, ybar, *xs, **kwargs):
        if self.rrule is not None:
            return p[n]  # p[N+1] with 1-based tuples == p[n] with 0-based (skip ∂self)
        g = self.grads.get(n, self.grads.get("any"))
        if g is None:
            raise TypeError(f"MethodError: no method matching ∇(::typeof({self.name}), Arg{{{n}}}, ...)")
        return g(n, p, y, ybar, *xs, **kwargs)

    def grad_acc(self, xbar, n, p, y, ybar, *xs, **kwargs):
        return add_bang_bang(xbar, self.grad(n, p, y, ybar, *xs, **kwargs))  # core.jl:203-205

    def __call__(self, *args, **kwargs):
        nodes = [a for a in args if isinstance(a, Node)]
        if not nodes:
            if self.rrule is not None:
                return self.rrule(*args, **kwargs)[0]
            return self.forward(*args, **kwargs)
        return Branch(self, args, nodes[0].tape, **kwargs)


# ---- broadcast / map (src/sensitivities/functional/functional.jl) -------------------------------
def _broadcast_forward(f, *As):
    nd = max((len(_jl_shape(a)) for a in As), default=0)
    out = f(*[_pad(a, nd) for a in As])
    shape = jl_broadcast_shape(*[_jl_shape(a) for a in As])
    if _is_array(out) and out.shape != shape:
        out = np.broadcast_to(out, shape).copy()
    elif not _is_array(out) and shape != ():
        out = np.full(shape, out, dtype=np.result_type(*[np.asarray(unbox_ref(a)).dtype for a in As]))
    return out


def unbox_ref(a):
    return a.x if isinstance(a, Ref) else a


def broadcastsum_inplace(f, add, z, *As):
    """``broadcastsum!(f, add, z, As...)`` (functional.jl:59-69)."""
    tmp_shape = jl_broadcast_shape(*[_jl_shape(a) for a in As])
    nd = len(tmp_shape)
    if z.shape != tmp_shape:
        tmp = np.empty(tmp_shape, dtype=z.dtype)
        tmp[...] = f(*[_pad(a, nd) for a in As])          # broadcast!(f, tmp, As...)
        red = _sum_to_shape(tmp, z.shape)                  # sum!(z, tmp, init=!add)
        if add:
            z += red
        else:
            z[...] = red
        return z
    if add:
        z += f(*[_pad(a, nd) for a in As])
    else:
        z[...] = f(*[_pad(a, nd) for a in As])
    return z


def broadcastsum(f, add, z, *As):
    """``broadcastsum`` for array (functional.jl:76-77) and Number/Ref (``:84-89``) targets."""
    if isinstance(z, Ref):
        z = z.x
    if _is_array(z):
        return broadcastsum_inplace(f, add, np.empty(z.shape, dtype=z.dtype), *As)
    tmp_shape = jl_broadcast_shape(*[_jl_shape(a) for a in As])
    nd = len(tmp_shape)
    tmp = np.empty(tmp_shape, dtype=np.result_type(z))
    tmp[...] = f(*[_pad(a, nd) for a in As])
    return tmp.sum() + (z if add else 0.0)


def _grad_broadcast(n, p, y, ybar, f, *A):
    # ∇(broadcast, Arg{N}, ...) -> _∇(broadcast, Arg{N-1}, ...) (functional.jl:92-95)
    k = n - 2  # 0-based operand index among A
    return broadcastsum(lambda yb, *xn: yb * fmad(f, xn, k), False, A[k], ybar, *A)


def _fwd_map(f, *As):
    shapes = {np.shape(a) for a in As}
    if len(shapes) != 1:
        raise ValueError("DimensionMismatch in map")
    return f(*As)


def _grad_map(n, p, y, ybar, f, *A):
    k = n - 2  # functional.jl:21-24
    out = ybar * fmad(f, A, k)
    return np.asarray(out, dtype=ybar.dtype) if _is_array(ybar) else out


_broadcast = Primitive("broadcast", forward=_broadcast_forward, grads={"any": _grad_broadcast})
_map = Primitive("map", forward=_fwd_map, grads={"any": _grad_map})


def broadcast(f, *args):
    """``broadcast(f, args...)`` / ``f.(args...)``: ONE Branch per call, never fused
    (functional.jl:48-51)."""
    return _broadcast(f, *args)


def map_(f, *args):
    return _map(f, *args)


# ---- reductions (functional/reduce.jl, functional/reducedim.jl) ---------------------------------
def _norm_dims(dims, nd):
    if dims is None:
        return None
    if isinstance(dims, (int, np.integer)):
        dims = (dims,)
    return tuple(sorted({int(d) for d in dims if int(d) <= nd}))  # 1-based; dims > ndims are no-ops


def _jl_reduce(a, dims, fn=None, mean_=False):
    """Base ``sum``/``mean``/``mapreduce(f, +, A; dims)``: keeps reduced dims with extent 1."""
    v = a if fn is None else fn(a)
    if dims is None:
        return v.mean() if mean_ else v.sum()
    axes = tuple(d - 1 for d in _norm_dims(dims, np.ndim(a)))
    return v.mean(axis=axes, keepdims=True) if mean_ else v.sum(axis=axes, keepdims=True)


def _denom(x, dims):
    """``_denom`` (reducedim.jl:60-62)."""
    if dims is None:
        return x.size
    n = 1
    for d in _norm_dims(dims, x.ndim):
        n *= x.shape[d - 1]
    return n


def _fwd_mapreduce(f, op, A, dims=None):
    return _jl_reduce(A, dims, fn=f)


def _grad_mapreduce(n, p, y, ybar, f, op, A, dims=None):
    # broadcast((An, ȳn)->ȳn * ForwardDiff.derivative(f, An), A, ȳ)   (reducedim.jl:11-20)
    return np.asarray(ybar * fmad(f, (A,), 0) + 0.0 * A, dtype=A.dtype)


def _fwd_sum_f(f, A, dims=None):
    return _jl_reduce(A, dims, fn=f)


def _grad_sum_f(n, p, y, ybar, f, A, dims=None):
    if isinstance(f, ScalarFn) and f.name == "abs2":
        return 2 * ybar * A                                   # reducedim.jl:42-51
    return _grad_mapreduce(3, p, y, ybar, f, None, A, dims=dims)  # reducedim.jl:29-39


def _fwd_mean_f(f, A):
    return _jl_reduce(A, None, fn=f, mean_=True)


def _grad_mean_f(n, p, y, ybar, f, A):
    return _grad_sum_f(2, p, y, ybar, f, A, dims=None) / _denom(A, None)  # reducedim.jl:64-72


def _fwd_mapfold(f, op, A):
    return _jl_reduce(A, None, fn=f) if _is_array(A) else f(A)


def _grad_mapfold(n, p, y, ybar, f, op, A):
    return ybar * fmad(f, (A,), 0) + 0.0 * A                   # reduce.jl:4-12


_mapreduce = Primitive("mapreduce", forward=_fwd_mapreduce, grads={3: _grad_mapreduce})
_mapfoldl = Primitive("mapfoldl", forward=_fwd_mapfold, grads={3: _grad_mapfold})
_mapfoldr = Primitive("mapfoldr", forward=_fwd_mapfold, grads={3: _grad_mapfold})
_sum_f = Primitive("sum", forward=_fwd_sum_f, grads={2: _grad_sum_f})
_mean_f = Primitive("mean", forward=_fwd_mean_f, grads={2: _grad_mean_f})


def _rrule_sum(x, dims=None):
    # ChainRules 1.35.2 rrule(sum, x; dims): x̄ = broadcast(last∘tuple, x, ȳ); InplaceableThunk
    y = _jl_reduce(x, dims)

    def sum_pullback(dy):
        dy = unthunk(dy)

        def val():
            out = np.empty(x.shape, dtype=x.dtype)
            out[...] = dy
            return out

        def add(dx):
            dx += dy
            return dx

        return (None, InplaceableThunk(add, val))

    return y, sum_pullback


def _rrule_mean(x, dims=None):
    y_sum, sum_pb = _rrule_sum(x, dims=dims)
    n = _denom(x, dims)

    def mean_pullback(dy):
        return sum_pb(unthunk(dy) / n)

    return y_sum / n, mean_pullback


def _rrule_dot(x, y):
    def dot_pullback(d):
        d = unthunk(d)
        return (None, Thunk(lambda: d * y), Thunk(lambda: d * x))

    return float(np.vdot(x, y)) if x.dtype == np.float64 else np.vdot(x, y), dot_pullback


_sum = Primitive("sum", rrule=_rrule_sum)
_mean = Primitive("mean", rrule=_rrule_mean)
_dot = Primitive("dot", rrule=_rrule_dot)


def _callable_first(args):
    return len(args) == 2 and callable(args[0]) and not isinstance(args[0], (Node, np.ndarray))


def sum(*args, dims=None):  # noqa: A001  (mirrors Base.sum)
    """``sum(x; dims)`` (ChainRules) or ``sum(f, x; dims)`` (reducedim.jl:23-51)."""
    if _callable_first(args):
        return _sum_f(args[0], args[1], dims=dims) if dims is not None else _sum_f(*args)
    return _sum(args[0], dims=dims) if dims is not None else _sum(args[0])


def mean(*args, dims=None):
    if _callable_first(args):
        if dims is not None:
            raise TypeError("MethodError: mean(f, x; dims) is not intercepted (reducedim.jl:53-58)")
        return _mean_f(*args)
    return _mean(args[0], dims=dims) if dims is not None else _mean(args[0])


def mapreduce(f, op, A, dims=None):
    _check_plus(op)
    return _mapreduce(f, op, A, dims=dims) if dims is not None else _mapreduce(f, op, A)


def mapfoldl(f, op, A):
    _check_plus(op)
    return _mapfoldl(f, op, A)


def mapfoldr(f, op, A):
    _check_plus(op)
    return _mapfoldr(f, op, A)


def _check_plus(op):
    if not (op == "+" or (isinstance(op, ScalarFn) and op.name == "add")):
        raise TypeError("MethodError: only `+` reductions are intercepted (reduce.jl:2)")


def dot(x, y):
    return _dot(x, y)


# ---- dense linalg (ChainRules `*`, adjoint/transpose, BLAS.gemm/gemv) ---------------------------
class AdjVec(np.ndarray):
    """``Adjoint{T,Vector}``: a (1, m) row whose product with a vector is a scalar."""


def _adj_value(x):
    if x.ndim == 1:
        return x.reshape(1, -1).view(AdjVec)
    if isinstance(x, AdjVec):
        return np.asarray(x).reshape(-1)
    return np.asarray(x).T


def _rrule_adjoint(x):
    def adjoint_pullback(d):
        d = unthunk(d)
        if x.ndim == 1:  # ProjectTo(x): the tangent of a Vector is a Vector
            return (None, Thunk(lambda: np.ascontiguousarray(np.asarray(d).reshape(-1))))
        return (None, Thunk(lambda: np.ascontiguousarray(_adj_value(d))))

    return _adj_value(x), adjoint_pullback


_adjoint = Primitive("adjoint", rrule=_rrule_adjoint)


def adjoint(x):
    return _adjoint(x)


transpose = adjoint  # real eltypes only


def _mm(a, b):
    if isinstance(a, AdjVec) and b.ndim == 1:
        return np.asarray(a).reshape(-1) @ b  # x' * y -> scalar
    return np.asarray(a) @ np.asarray(b)


def _rrule_matmul(A, B):
    # ChainRules 1.35.2 rrule(*, A::AbstractVecOrMat, B::AbstractVecOrMat):
    #   Ā = InplaceableThunk(X̄ -> mul!(X̄, Ȳ, B', true, true), @thunk(Ȳ * B')),  B̄ likewise with A'Ȳ
    Y = _mm(A, B)

    def times_pullback(dY):
        dY = unthunk(dY)

        def dA():
            if isinstance(A, AdjVec) and B.ndim == 1:
                return (dY * B).reshape(1, -1).view(AdjVec)
            if B.ndim == 1:
                d = np.asarray(dY)
                return np.outer(d, B) if A.ndim == 2 else d * B
            return np.asarray(dY) @ np.asarray(B).T

        def dB():
            if isinstance(A, AdjVec) and B.ndim == 1:
                return dY * np.asarray(A).reshape(-1)
            return np.asarray(A).T @ np.asarray(dY)

        def addA(X):
            X += dA()
            return X

        def addB(X):
            X += dB()
            return X

        return (None, InplaceableThunk(addA, dA), InplaceableThunk(addB, dB))

    return Y, times_pullback


def _rrule_scale(a, B):
    # rrule(*, a::Number, B::AbstractArray): ā = dot(B, Ȳ), B̄ = a * Ȳ  (either order)
    if _is_array(a):
        a, B, swapped = B, a, True
    else:
        swapped = False

    def pb(dY):
        dY = unthunk(dY)
        da = Thunk(lambda: np.vdot(B, dY))
        dB = Thunk(lambda: a * dY)
        return (None, dB, da) if swapped else (None, da, dB)

    return a * B, pb


_matmul = Primitive("*", rrule=_rrule_matmul)
_scale = Primitive("*", rrule=_rrule_scale)


def matmul(A, B):
    """Julia ``A * B`` on arrays."""
    return _matmul(A, B)


def _op(t, M):
    return M.T if t in ("T", "C") else M


def _rrule_gemm(tA, tB, *rest):
    # BLAS.gemm(tA, tB, [α,] A, B); adjoints pinned at test/sensitivities/linalg/blas.jl:36-47
    if len(rest) == 3:
        alpha, A, B = rest
        has_alpha = True
    else:
        A, B = rest
        alpha, has_alpha = 1.0, False
    AB = _op(tA, A) @ _op(tB, B)
    Y = alpha * AB

    def pb(dY):
        dY = unthunk(dY)

        def dA():
            g = alpha * (dY @ _op(tB, B).T)
            return g.T.copy() if tA in ("T", "C") else g

        def dB():
            g = alpha * (_op(tA, A).T @ dY)
            return g.T.copy() if tB in ("T", "C") else g

        out = [None, None, None]
        if has_alpha:
            out.append(Thunk(lambda: float(np.vdot(dY, AB))))
        out += [Thunk(dA), Thunk(dB)]
        return tuple(out)

    return Y, pb


def _rrule_gemv(tA, *rest):
    if len(rest) == 3:
        alpha, A, x = rest
        has_alpha = True
    else:
        A, x = rest
        alpha, has_alpha = 1.0, False
    Ax = _op(tA, A) @ x
    y = alpha * Ax

    def pb(dy):
        dy = unthunk(dy)

        def dA():
            g = alpha * np.outer(dy, x)
            return g.T.copy() if tA in ("T", "C") else g

        out = [None, None]
        if has_alpha:
            out.append(Thunk(lambda: float(np.vdot(dy, Ax))))
        out += [Thunk(dA), Thunk(lambda: alpha * (_op(tA, A).T @ dy))]
        return tuple(out)

    return y, pb


_gemm = Primitive("gemm", rrule=_rrule_gemm)
_gemv = Primitive("gemv", rrule=_rrule_gemv)


def gemm(tA, tB, *rest):
    return _gemm(tA, tB, *rest)


def gemv(tA, *rest):
    return _gemv(tA, *rest)


# ---- scalar rules on Node{<:Real} (ChainRules @scalar_rule via the bridge) -----------------------
_scalar_prims = {}


def _scalar_rule_call(fn, args):
    prim = _scalar_prims.get(fn.name)
    if prim is None:
        if fn.arity == 1:
            def rrule(x, _fn=fn):
                d = _fn.df(x)
                return _fn.f(x), (lambda dy: (None, unthunk(dy) * d))
        else:
            def rrule(x, y, _fn=fn):
                dx, dy_ = _fn.dfx(x, y), _fn.dfy(x, y)
                return _fn.f(x, y), (lambda dz: (None, unthunk(dz) * dx, unthunk(dz) * dy_))
        prim = Primitive(fn.name, rrule=rrule)
        _scalar_prims[fn.name] = prim
    for a in args:
        v = unbox(a)
        if _is_array(v) and v.ndim > 0:
            raise TypeError(f"MethodError: no method matching {fn.name}(::Array); use broadcast")
    return prim(*args)


# array `+` routed to the broadcast rule (chainrules_ignored.jl:5-10)
_array_plus = Primitive(
    "+", forward=lambda x, y: _same_shape_add(x, y),
    grads={1: lambda n, p, z, zb, x, y: _grad_broadcast(2, p, z, zb, add, x, y),   # noqa: F821
           2: lambda n, p, z, zb, x, y: _grad_broadcast(3, p, z, zb, add, x, y)})  # noqa: F821
# array / scalar and scalar \ array (functional.jl:98-111)
_array_div = Primitive(
    "/", forward=lambda x, y: x / y,
    grads={1: lambda n, p, z, zb, x, y: _grad_broadcast(2, p, z, zb, div, x, y),   # noqa: F821
           2: lambda n, p, z, zb, x, y: _grad_broadcast(3, p, z, zb, div, x, y)})  # noqa: F821
_array_ldiv = Primitive(
    "\\", forward=lambda x, y: y / x,
    grads={1: lambda n, p, z, zb, x, y: _grad_broadcast(2, p, z, zb, ldiv, x, y),   # noqa: F821
           2: lambda n, p, z, zb, x, y: _grad_broadcast(3, p, z, zb, ldiv, x, y)})  # noqa: F821


def _same_shape_add(x, y):
    if np.shape(x) != np.shape(y):
        raise ValueError(f"DimensionMismatch: {np.shape(x)} + {np.shape(y)}")
    return x + y


def _rrule_array_minus(x, y):
    def pb(d):
        d = unthunk(d)
        return (None, d, Thunk(lambda: -d))
    if np.shape(x) != np.shape(y):
        raise ValueError("DimensionMismatch")
    return x - y, pb


def _rrule_array_neg(x):
    return -x, (lambda d: (None, Thunk(lambda: -unthunk(d))))


_array_minus = Primitive("-", rrule=_rrule_array_minus)
_array_neg = Primitive("-", rrule=_rrule_array_neg)


def _isarr(x):
    v = unbox(x)
    return _is_array(v) and v.ndim > 0


def _plus(a, b):
    if _isarr(a) and _isarr(b):
        return _array_plus(a, b)
    if _isarr(a) or _isarr(b):
        raise TypeError("MethodError: no method matching +(::Array, ::Number); use broadcast")
    return add(a, b)  # noqa: F821


def _minus(a, b):
    if _isarr(a) and _isarr(b):
        return _array_minus(a, b)
    if _isarr(a) or _isarr(b):
        raise TypeError("MethodError: no method matching -(::Array, ::Number); use broadcast")
    return sub(a, b)  # noqa: F821


def _negate(a):
    return _array_neg(a) if _isarr(a) else neg(a)  # noqa: F821


def _times(a, b):
    if _isarr(a) and _isarr(b):
        return _matmul(a, b)
    if _isarr(a) or _isarr(b):
        return _scale(a, b)
    return mul(a, b)  # noqa: F821


def _divide(a, b):
    if _isarr(a) and not _isarr(b):
        return _array_div(a, b)
    if _isarr(b):
        raise TypeError("MethodError: array right-division is out of scope (LAPACK)")
    return div(a, b)  # noqa: F821


def ldivide(a, b):
    """Julia ``a \\ b`` for scalar ``a`` and array ``b``."""
    if not _isarr(a) and _isarr(b):
        return _array_ldiv(a, b)
    return ldiv(a, b)  # noqa: F821


def _power(a, b):
    if _isarr(a) or _isarr(b):
        raise TypeError("MethodError: ^ on arrays is out of scope; use broadcast")
    return pow_(a, b)  # noqa: F821


# --------------------------------------------------------------------------------------------
# Array structure: hcat / vcat / getindex / reshape / fill, and gradient checkpointing
# 
# --------------------------------------------------------------------------------------------
class R:
    """Julia's ``lo:hi`` (1-based, inclusive)."""

    def __init__(self, lo, hi):
        self.lo, self.hi = int(lo), int(hi)

    def __len__(self):
        return max(self.hi - self.lo + 1, 0)


colon = R(1, -1)  # ``:`` — resolved against the indexed extent


def _col2(a):
    a = np.asarray(a)
    return a.reshape(a.shape[0], 1) if a.ndim == 1 else a


def _grad_hcat(n, p, y, ybar, *A):
    # chainrules_ignored.jl:14-26: copy(u > l+1 ? selectdim(ȳ, 2, l+1:u) : selectdim(ȳ, 2, u))
    i = n - 1
    l = _pysum(_col2(A[j]).shape[1] for j in range(i))
    u = l + _col2(A[i]).shape[1]
    yb = np.asarray(unthunk(ybar))
    return yb[:, l:u].copy() if u > l + 1 else yb[:, u - 1].copy()


def _grad_vcat(n, p, y, ybar, *A):
    # chainrules_ignored.jl:29-40: copy(selectdim(ȳ, 1, l+1:u))
    i = n - 1
    l = _pysum(np.shape(A[j])[0] for j in range(i))
    u = l + np.shape(A[i])[0]
    return np.asarray(unthunk(ybar))[l:u].copy()


_hcat = Primitive("hcat", forward=lambda *A: np.hstack([_col2(a) for a in A]), grads={"any": _grad_hcat})
_vcat = Primitive("vcat", forward=lambda *A: np.concatenate([np.asarray(a) for a in A], axis=0),
                  grads={"any": _grad_vcat})


def hcat(*A):
    return _hcat(*A)


def vcat(*A):
    return _vcat(*A)


def _np_index(shape, idx):
    """Julia index tuple -> NumPy index (0-based); ints drop their dimension like Julia."""
    if len(idx) == 1 and len(shape) > 1:  # linear indexing into a column-major array
        raise TypeError("MethodError: linear indexing of matrices is not intercepted here")
    out = []
    for d, ix in enumerate(idx):
        ext = shape[d]
        if isinstance(ix, R):
            hi = ext if ix is colon else ix.hi
            lo = 1 if ix is colon else ix.lo
            if lo < 1 or hi > ext:
                raise IndexError(f"BoundsError: attempt to access {shape} at index [{lo}:{hi}]")
            out.append(slice(lo - 1, hi))
        elif isinstance(ix, (int, np.integer)):
            if ix < 1 or ix > ext:
                raise IndexError(f"BoundsError: attempt to access {shape} at index [{ix}]")
            out.append(int(ix) - 1)
        else:
            iv = np.asarray(ix, dtype=np.int64)
            if iv.size and (iv.min() < 1 or iv.max() > ext):
                raise IndexError(f"BoundsError: attempt to access {shape} at index {list(iv)}")
            out.append(iv - 1)
    return tuple(out)


def _rrule_getindex(x, *idx):
    # ChainRules rrule(getindex, x::Array, inds...): x̄ = zero(x); x̄[inds...] .+= ȳ (duplicates add, #139)
    x = np.asarray(x)
    ni = _np_index(x.shape, idx)
    y = x[ni]

    def pb(ybar):
        def thunk():
            xb = np.zeros_like(x, dtype=np.result_type(x.dtype, np.float64) if x.dtype.kind in "iu" else x.dtype)
            np.add.at(xb, ni, unthunk(ybar))
            return xb
        return (None, Thunk(thunk)) + (None,) * len(idx)
    return (y.copy() if isinstance(y, np.ndarray) else y), pb


_getindex = Primitive("getindex", rrule=_rrule_getindex)


def getindex(x, *idx):
    return _getindex(x, *idx)


def _rrule_reshape(x, *dims):
    if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
        dims = tuple(dims[0])
    shp = np.shape(x)
    return np.reshape(x, dims, order="F"), (lambda ybar: (None, np.reshape(unthunk(ybar), shp, order="F")) + (None,) * len(dims))


_reshape = Primitive("reshape", rrule=_rrule_reshape)


def reshape(x, *dims):
    return _reshape(x, *dims)


def _rrule_fill(x, *dims):
    if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
        dims = tuple(dims[0])
    return np.full(dims, x, dtype=np.result_type(x, np.float64)), (lambda ybar: (None, np.sum(unthunk(ybar))) + (None,) * len(dims))


_fill = Primitive("fill", rrule=_rrule_fill)


def fill(x, *dims):
    return _fill(x, *dims)


# ---- generic linalg: kron, A + λI, copy (src/sensitivities/linalg/generic.jl:7-51) -------------
def _kron_check(A, B):
    A, B = np.asarray(A), np.asarray(B)
    if A.ndim != 2 or B.ndim != 2:
        raise TypeError("MethodError: the kron sensitivities destructure size(A), size(B) as matrices (generic.jl:19)")
    return A, B


def _grad_kron(n, p, y, ybar, A, B):
    """generic.jl:13-40 restated: the reference walks the column-major linear index m of Ȳ downwards over
    (j, l, i, k) — Y[(i-1)K + k, (j-1)L + l] = A[i,j] B[k,l] — accumulating ā_ij += Ȳ[m] B[k,l] (Arg 1) or
    B̄[k,l] += Ȳ[m] a_ij (Arg 2) into a zero matrix; as contractions of Ȳ viewed as (I, K, J, L)."""
    A, B = _kron_check(A, B)
    (I, J), (Kk, L) = A.shape, B.shape
    Y4 = np.asarray(unthunk(ybar)).reshape(I, Kk, J, L)
    if n == 1:
        return np.einsum("ikjl,kl->ij", Y4, B)
    return np.einsum("ikjl,ij->kl", Y4, A)


_kron = Primitive("kron", forward=lambda A, B: np.kron(*_kron_check(A, B)), grads={1: _grad_kron, 2: _grad_kron})


def kron(A, B):
    return _kron(A, B)


def _plus_I_forward(A, lam):
    A = np.asarray(A)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("DimensionMismatch: matrix is not square")
    return A + float(lam) * np.eye(A.shape[0], dtype=A.dtype)


# `A + λI` / `λI + A` (generic.jl:42-46): the sensitivity of A is Ȳ itself
_plus_I = Primitive("+I", forward=_plus_I_forward, grads={1: lambda n, p, y, ybar, A, lam: ybar})


def plus_I(A, lam):
    """``A + lam*I`` with ``I`` the UniformScaling (either operand order: the rule is the same)."""
    return _plus_I(A, lam)


# `copy` (generic.jl:48-51): Ā = copy(Ȳ)
_copy = Primitive("copy", forward=lambda A: np.array(A, copy=True),
                  grads={1: lambda n, p, y, ybar, A: np.array(unthunk(ybar), copy=True)})


def copy(A):
    return _copy(A)


def _checkpoint_forward(f, *xs, **kw):
    return f(*xs, **kw)


def _grad_checkpoint(n, p, y, ybar, f, *xs, **kw):
    # checkpointing.jl:22-34: re-run f on an inner tape inside the outer reverse pass
    t = Tape()
    xs_ = [Leaf(t, x) for x in xs]
    yi = f(*xs_, **kw)
    if isinstance(yi, Node):
        df = nabla(yi, unthunk(ybar))
        return df[xs_[n - 2]] if df.isassigned(xs_[n - 2]) else _zero(xs[n - 2])
    return _zero(xs[n - 2])


_checkpoint = Primitive("checkpoint", forward=_checkpoint_forward, grads={"any": _grad_checkpoint})


def checkpoint(f, args, kwargs=None):
    """``checkpoint(f, args::Tuple[, kwargs::NamedTuple])`` (checkpointing.jl:10-20): ``f`` must not be a
    closure."""
    if getattr(f, "__closure__", None):
        raise RuntimeError("f mustn't have fields. Can not checkpoint a closure or functor.")
    return _checkpoint(f, *args, **(kwargs or {}))


# --------------------------------------------------------------------------------------------
# Symmetric / triangular BLAS family (src/sensitivities/linalg/blas.jl:60-328)
# --------------------------------------------------------------------------------------------
def _U(c):
    return str(c).upper()


def _sym(ul, A):
    A = np.asarray(A)
    return np.tril(A) + np.tril(A, -1).T if _U(ul) == "L" else np.triu(A) + np.triu(A, 1).T


def _tri(ul, dA, A):
    A = np.asarray(A)
    T = np.tril(A) if _U(ul) == "L" else np.triu(A)
    if _U(dA) == "U":
        T = T - np.diag(np.diag(T)) + np.eye(A.shape[0], dtype=A.dtype)
    return T


def _g(ul, M, dA="N"):
    """``tril!``/``triu!`` (and a zeroed diagonal for unit-triangular operands)."""
    M = np.tril(M) if _U(ul) == "L" else np.triu(M)
    if _U(dA) == "U":
        M = M - np.diag(np.diag(M))
    return M


def _op(ta, M):
    return M if _U(ta) == "N" else M.T


def _flip(ta):
    return "T" if _U(ta) == "N" else "N"


def _symm_raw(side, ul, alpha, A, B):
    S = _sym(ul, A)
    return alpha * (S @ B) if _U(side) == "L" else alpha * (B @ S)


def _symm_split(rest):
    return rest if len(rest) == 3 else (1.0,) + tuple(rest)


def _symm_fwd(side, ul, *rest):
    return _symm_raw(side, ul, *_symm_split(rest))


def _symm_grad(n, p, Y, Yb, side, ul, *rest):
    has_alpha = len(rest) == 3
    alpha, A, B = _symm_split(rest)
    Yb = unthunk(Yb)
    k = n - 3 - (1 if has_alpha else 0)   # 0: A, 1: B
    if has_alpha and n == 3:
        return np.sum(Yb * Y) / alpha                                        # blas.jl:65-71
    if k == 0:
        B2, Yb2 = np.atleast_2d(B.T).T, np.atleast_2d(Yb.T).T
        tmp = Yb2 @ B2.T if _U(side) == "L" else B2.T @ Yb2                   # blas.jl:78-80
        return alpha * _g(ul, tmp + tmp.T - np.diag(np.diag(tmp)))
    return _symm_raw(side, ul, alpha, A, Yb)                                 # blas.jl:95-101


_symm = Primitive("symm", forward=_symm_fwd, grads={"any": _symm_grad})


def symm(side, ul, *rest):
    return _symm(side, ul, *rest)


def _symv_fwd(ul, *rest):
    alpha, A, x = _symm_split(rest)
    return alpha * (_sym(ul, A) @ x)


def _symv_grad(n, p, y, yb, ul, *rest):
    has_alpha = len(rest) == 3
    alpha, A, x = _symm_split(rest)
    yb = unthunk(yb)
    if has_alpha and n == 2:
        return np.dot(yb, y) / alpha                                         # blas.jl:146-151
    k = n - 2 - (1 if has_alpha else 0)
    if k == 0:
        tmp = np.outer(yb, x)                                                # symm rule with side 'L'
        return alpha * _g(ul, tmp + tmp.T - np.diag(np.diag(tmp)))
    return alpha * (_sym(ul, A) @ yb)


_symv = Primitive("symv", forward=_symv_fwd, grads={"any": _symv_grad})


def symv(ul, *rest):
    return _symv(ul, *rest)


def _trmm_raw(side, ul, ta, dA, alpha, A, B):
    T = _op(ta, _tri(ul, dA, A))
    return alpha * (T @ B) if _U(side) == "L" else alpha * (B @ T)


def _trmm_grad(n, p, Y, Yb, side, ul, ta, dA, alpha, A, B):
    Yb = unthunk(Yb)
    if n == 5:
        return np.sum(Yb * Y) / alpha
    if n == 6:                                                               # blas.jl:217-231
        if _U(side) == "L":
            full = alpha * (Yb @ B.T) if _U(ta) == "N" else alpha * (B @ Yb.T)
        else:
            full = alpha * (B.T @ Yb) if _U(ta) == "N" else alpha * (Yb.T @ B)
        return _g(ul, full, dA)
    return _trmm_raw(side, ul, _flip(ta), dA, alpha, A, Yb)                  # blas.jl:232-244


_trmm = Primitive("trmm", forward=_trmm_raw, grads={"any": _trmm_grad})


def trmm(side, ul, ta, dA, alpha, A, B):
    return _trmm(side, ul, ta, dA, alpha, A, B)


def _trmv_raw(ul, ta, dA, A, b):
    return _op(ta, _tri(ul, dA, A)) @ b


def _trmv_grad(n, p, y, yb, ul, ta, dA, A, b):
    yb = unthunk(yb)
    if n == 4:                                                               # blas.jl:252-260
        return _g(ul, np.outer(yb, b) if _U(ta) == "N" else np.outer(b, yb), dA)
    return _trmv_raw(ul, _flip(ta), dA, A, yb)


_trmv = Primitive("trmv", forward=_trmv_raw, grads={"any": _trmv_grad})


def trmv(ul, ta, dA, A, b):
    return _trmv(ul, ta, dA, A, b)


def _trsm_raw(side, ul, ta, dA, alpha, A, X):
    T = _op(ta, _tri(ul, dA, A))
    X2 = np.atleast_2d(np.asarray(X).T).T
    out = alpha * np.linalg.solve(T, X2) if _U(side) == "L" else alpha * np.linalg.solve(T.T, X2.T).T
    return out.reshape(np.shape(X))


def _trsm_grad(n, p, Y, Yb, side, ul, ta, dA, alpha, A, X):
    Yb = unthunk(Yb)
    if n == 5:
        return np.sum(Yb * Y) / alpha
    if n == 6:                                                               # blas.jl:280-294
        Y2, Yb2 = np.atleast_2d(np.asarray(Y).T).T, np.atleast_2d(np.asarray(Yb).T).T
        if _U(side) == "L":
            full = (_trsm_raw("L", ul, "T", dA, -1.0, A, Yb2 @ Y2.T) if _U(ta) == "N"
                    else _trsm_raw("R", ul, "T", dA, -1.0, A, Y2 @ Yb2.T))
        else:
            full = (_trsm_raw("R", ul, "T", dA, -1.0, A, Y2.T @ Yb2) if _U(ta) == "N"
                    else _trsm_raw("L", ul, "T", dA, -1.0, A, Yb2.T @ Y2))
        return _g(ul, full, dA)
    return _trsm_raw(side, ul, _flip(ta), dA, alpha, A, Yb)                  # blas.jl:295-307


_trsm = Primitive("trsm", forward=_trsm_raw, grads={"any": _trsm_grad})


def trsm(side, ul, ta, dA, alpha, A, X):
    return _trsm(side, ul, ta, dA, alpha, A, X)


def _trsv_raw(ul, ta, dA, A, x):
    return _trsm_raw("L", ul, ta, dA, 1.0, A, x)


def _trsv_grad(n, p, y, yb, ul, ta, dA, A, x):
    if n == 4:                                                               # blas.jl:315-323
        return _trsm_grad(6, p, y, yb, "L", ul, ta, dA, 1.0, A, x)
    return _trsv_raw(ul, _flip(ta), dA, A, unthunk(yb))


_trsv = Primitive("trsv", forward=_trsv_raw, grads={"any": _trsv_grad})


def trsv(ul, ta, dA, A, x):
    return _trsv(ul, ta, dA, A, x)


# --------------------------------------------------------------------------------------------
# The reference's own test method (src/finite_differencing.jl)
# --------------------------------------------------------------------------------------------
def _central5(g, h=None):
    """``FDM.Central(5, 1)``: 5-point central first derivative at 0 (weights 1,-8,0,8,-1 / 12h).
    FDM 0.6.1 picks an adaptive step; a fixed step near eps^(1/5) has the same O(h^4) order."""
    h = 1e-3 if h is None else h
    return (g(-2 * h) - 8 * g(-h) + 8 * g(h) - g(2 * h)) / (12 * h)


def _tuplify(x):
    return x if isinstance(x, tuple) else (x,)


def approximate_Dv(f, ybar, x, v):
    x, v = _tuplify(x), _tuplify(v)
    return _central5(lambda e: np.sum(ybar * unbox(f(*[xi + e * vi for xi, vi in zip(x, v)]))))


def compute_Dv(f, ybar, x, v):
    x, v = _tuplify(x), _tuplify(v)
    t = Tape()
    x_ = [Leaf(t, xi) for xi in x]
    df = nabla(f(*x_), ybar)
    return _pysum(np.sum(df[xi_] * vi) for xi_, vi in zip(x_, v))


def compute_Dv_update(f, ybar, x, v, rng=None):
    """Pre-filled leaf slots exercise the accumulate path (finite_differencing.jl:57-85)."""
    rng = np.random.default_rng(0) if rng is None else rng
    x, v = _tuplify(x), _tuplify(v)
    t = Tape()
    x_ = [Leaf(t, xi) for xi in x]
    y = f(*x_)
    rt = reverse_tape(y, ybar)
    inits = {}
    for i in range(1, len(rt) + 1):
        node = y.tape.tape[i - 1]
        if isinstance(node, Leaf):
            val = node.val
            init = rng.standard_normal(np.shape(val)) if _is_array(val) else float(rng.standard_normal())
            inits[i] = init
            rt[i] = np.array(init, copy=True) if _is_array(init) else init
    df = propagate(y.tape, rt)
    tot = 0.0
    for xi_, vi in zip(x_, v):
        tot += np.sum((df[xi_] - inits[xi_.pos]) * vi)
    return tot


def check_errs(f, ybar, x, v, eps_abs=1e-10, eps_rel=1e-7):
    """``check_errs`` (finite_differencing.jl:106-120) with the reference's default tolerances."""
    a = compute_Dv(f, ybar, x, v)
    b = compute_Dv_update(f, ybar, x, v)
    c = approximate_Dv(f, ybar, x, v)
    for name, val in (("allocated", a), ("in-place", b)):
        err_abs = abs(val - c)
        err_rel = err_abs / (abs(val) + 1e-20)
        if err_abs > eps_abs and err_rel > eps_rel:
            raise AssertionError(f"<{getattr(f, '__name__', f)}> {name}: AD {val} vs FD {c} "
                                 f"(abs {err_abs:.3e}, rel {err_rel:.3e})")
    return True
