"""Generates tests/golden/scalar_catalogue_mpmath.json: every function of the reference's scalar catalogue
(test/runtests.jl:28-52) and its derivative at the reference's own test points (src/finite_differencing.jl:175-179:
-pi+.1, -pi/2+.1, -.9, -.1, .1, .9, pi/2-.1, pi-.1), computed with mpmath at 50 digits — a source independent of
both the oracle (NumPy / SciPy) and the device code.  Points where a function is not real and finite are left out.
The reference itself stores no values for these functions (its tests are finite-difference property tests), so this
is the known-answer table the oracle and the kernels are pinned to.   Run: python tests/golden/make_scalar_catalogue.py"""
import json
import os

import mpmath as mp

mp.mp.dps = 50
PI = mp.pi
PTS = [-PI + mp.mpf("0.1"), -PI / 2 + mp.mpf("0.1"), mp.mpf("-0.9"), mp.mpf("-0.1"), mp.mpf("0.1"), mp.mpf("0.9"),
       PI / 2 - mp.mpf("0.1"), PI - mp.mpf("0.1")]
# the points as the IEEE doubles the tests will feed in
PTS = [mp.mpf(float(p)) for p in PTS]
D2R, R2D = PI / 180, 180 / PI


def invdigamma(y):
    x0 = mp.exp(y) + mp.mpf("0.5") if y >= mp.mpf("-2.22") else -1 / (y - mp.digamma(1))
    return mp.findroot(lambda t: mp.digamma(t) - y, x0)


UNARY = {
    "identity": lambda x: x, "neg": lambda x: -x, "abs": abs, "abs2": lambda x: x * x, "sqrt": mp.sqrt, "cbrt": mp.cbrt,
    "inv": lambda x: 1 / x, "exp": mp.exp, "exp2": lambda x: mp.mpf(2) ** x, "exp10": lambda x: mp.mpf(10) ** x,
    "expm1": mp.expm1, "log": mp.log, "log2": lambda x: mp.log(x, 2), "log10": mp.log10, "log1p": mp.log1p,
    "sin": mp.sin, "cos": mp.cos, "tan": mp.tan, "sec": mp.sec, "csc": mp.csc, "cot": mp.cot,
    "sinh": mp.sinh, "cosh": mp.cosh, "tanh": mp.tanh, "sech": mp.sech, "csch": mp.csch, "coth": mp.coth,
    "asin": mp.asin, "acos": mp.acos, "atan": mp.atan, "asec": mp.asec, "acsc": mp.acsc, "acot": lambda x: mp.atan(1 / x),
    "asinh": mp.asinh, "acosh": mp.acosh, "atanh": mp.atanh, "asech": mp.asech, "acsch": mp.acsch, "acoth": mp.acoth,
    "sind": lambda x: mp.sin(D2R * x), "cosd": lambda x: mp.cos(D2R * x), "tand": lambda x: mp.tan(D2R * x),
    "secd": lambda x: mp.sec(D2R * x), "cscd": lambda x: mp.csc(D2R * x), "cotd": lambda x: mp.cot(D2R * x),
    "asind": lambda x: R2D * mp.asin(x), "acosd": lambda x: R2D * mp.acos(x), "atand": lambda x: R2D * mp.atan(x),
    "asecd": lambda x: R2D * mp.asec(x), "acscd": lambda x: R2D * mp.acsc(x), "acotd": lambda x: R2D * mp.atan(1 / x),
    "sinpi": mp.sinpi, "cospi": mp.cospi, "deg2rad": lambda x: D2R * x, "rad2deg": lambda x: R2D * x,
    "logistic": lambda x: 1 / (1 + mp.exp(-x)),
    "erf": mp.erf, "erfc": mp.erfc, "erfcx": lambda x: mp.exp(x * x) * mp.erfc(x), "erfinv": mp.erfinv,
    "erfcinv": lambda x: mp.erfinv(1 - x), "gamma": mp.gamma, "digamma": mp.digamma,
    "trigamma": lambda x: mp.polygamma(1, x), "besselj0": lambda x: mp.besselj(0, x), "besselj1": lambda x: mp.besselj(1, x),
    "bessely0": lambda x: mp.bessely(0, x), "bessely1": lambda x: mp.bessely(1, x),
    "airyai": mp.airyai, "airyaiprime": lambda x: mp.airyai(x, derivative=1), "airybi": mp.airybi,
    "airybiprime": lambda x: mp.airybi(x, derivative=1),
    "dawson": lambda x: mp.sqrt(mp.pi) / 2 * mp.exp(-x * x) * mp.erfi(x), "erfi": mp.erfi, "invdigamma": invdigamma,
    # piecewise-constant functions (none of the test points is a breakpoint; mp.diff gives exactly 0)
    "sign": mp.sign, "step": lambda x: mp.mpf(1) if x > 0 else mp.mpf(0), "floor": mp.floor, "ceil": mp.ceil,
    "round": mp.nint, "trunc": lambda x: mp.floor(x) if x >= 0 else mp.ceil(x),
}
SECOND = {"besselj": mp.besselj, "bessely": mp.bessely, "besseli": mp.besseli, "besselk": mp.besselk,
          "polygamma": mp.polygamma}


def real_finite(v):
    try:
        v = mp.mpmathify(v)
    except Exception:
        return False
    if isinstance(v, mp.mpc):
        if abs(v.imag) > mp.mpf(10) ** -40:
            return False
    return mp.isfinite(v)


def sample(f, x):
    """(f(x), f'(x)) as doubles, or None outside the real domain."""
    try:
        v = f(x)
        if not real_finite(v):
            return None
        d = mp.diff(f, x)
        if not real_finite(d):
            return None
        return float(mp.re(v)), float(mp.re(d))
    except (ValueError, ZeroDivisionError, OverflowError, TypeError):
        return None


out = {"source": "mpmath %s, 50 digits; points of src/finite_differencing.jl:175-179" % mp.__version__, "unary": {},
       "second_arg_only": {}}
for name, f in UNARY.items():
    xs, fs, ds = [], [], []
    for p in PTS:
        if name == "abs" and p == 0:
            continue
        r = sample(f, p)
        if r is not None:
            xs.append(float(p)); fs.append(r[0]); ds.append(r[1])
    out["unary"][name] = {"x": xs, "f": fs, "df": ds}
for name, f in SECOND.items():
    rows = []
    for n in range(6):     # test/sensitivities/scalar.jl:43: integer first argument 0..5
        xs, fs, ds = [], [], []
        for p in PTS:
            r = sample(lambda t: f(n, t), p)
            if r is not None:
                xs.append(float(p)); fs.append(r[0]); ds.append(r[1])
        rows.append({"n": n, "x": xs, "f": fs, "df": ds})
    out["second_arg_only"][name] = rows
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scalar_catalogue_mpmath.json")
with open(path, "w") as fh:
    json.dump(out, fh, indent=0)
print("wrote", path, {k: len(v["x"]) for k, v in out["unary"].items()})
