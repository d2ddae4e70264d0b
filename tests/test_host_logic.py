"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the opcode table matches, the scalar tracer lowers functions correctly, and the product
path fails loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import subprocess

import pytest

import nabla_jl_b200 as nb
from nabla_jl_b200 import _lib, scalar

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _lib.declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/nabla_cuda.h but not exported"
    assert set(_lib._SIGS) == set(declared)
    assert L.nb_abi_version() == 1


def test_integration_doc_lists_every_entry_point():
    """INTEGRATION.md's appendix (tools/gen_ccall_stubs.py) holds a Julia ccall stub for every function the header
    declares — the binding a Nabla maintainer would load."""
    import re
    hdr = open(os.path.join(ROOT, "include", "nabla_cuda.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    names = set(re.findall(r"\b(nb_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)))
    assert len(names) >= 50
    for n in names:
        assert f"ccall((:{n}, LIB)" in doc, f"{n} has no ccall stub in INTEGRATION.md (run tools/gen_ccall_stubs.py)"


def test_julia_glue_opcode_table_matches_the_header():
    """julia/NablaCUDA.jl cannot be executed here (no Julia in the image); its function => nb_op tables are at least
    kept equal to the header's enum, and cover the whole catalogue."""
    import re
    src = open(os.path.join(ROOT, "julia", "NablaCUDA.jl")).read()
    ops = _lib.OPCODES
    alias = {"+": "add", "-": "sub", "*": "mul", "/": "div", "\\": "ldiv", "^": "pow"}
    seen = {}
    for table in ("UNARY_OPS", "BINARY_OPS"):
        body = src[src.index(f"const {table} = Dict"):]
        body = body[:body.index(")\nconst ") if ")\nconst " in body else len(body)]
        for name, val in re.findall(r"(?:SF\.)?\(?([A-Za-z0-9_]+|[-+*/^\\])\)?\s*=>\s*(\d+)", body):
            key = alias.get(name, name)
            assert key in ops, f"{table}: {name} is not a catalogue op"
            assert ops[key] == int(val), f"{table}: {name} => {val}, header says {ops[key]}"
            seen[key] = int(val)
    m = re.search(r"OP_NEG, OP_LOGISTIC, OP_ATAN2 = Int32\((\d+)\), Int32\((\d+)\), Int32\((\d+)\)", src)
    consts = dict(zip(("neg", "logistic", "atan2"), map(int, m.groups())))
    for k, v in consts.items():
        assert ops[k] == v
        seen[k] = v
    m = re.search(r"OP_GT, OP_LT, OP_GE, OP_LE, OP_MASK, OP_MASKN = " + ", ".join([r"Int32\((\d+)\)"] * 6), src)
    for k, v in zip(("gt", "lt", "ge", "le", "mask", "maskn"), map(int, m.groups())):
        assert ops[k] == v
        seen[k] = v
    missing = [n for n in scalar.UNARY + scalar.BINARY if n not in seen]
    assert not missing, f"catalogue ops without a Julia mapping: {missing}"


def test_julia_glue_ccall_signatures_match_the_header():
    """Every hand-written `ccall` in julia/NablaCUDA.jl passes the argument list the header declares (same count,
    same scalar widths, pointers where the header has pointers)."""
    import re
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_ccall_stubs
    want = {n: (r, [t for t, _ in al]) for n, r, al in gen_ccall_stubs.prototypes()}
    src = open(os.path.join(ROOT, "julia", "NablaCUDA.jl")).read()

    def cls(t):
        t = t.strip()
        return "ptr" if t.startswith(("Ptr", "Ref")) or t == "Cstring" else t

    def split_top(text):
        out, depth, cur = [], 0, ""
        for ch in text:
            depth += ch in "{(" 
            depth -= ch in "})"
            if ch == "," and depth == 0:
                out.append(cur)
                cur = ""
            else:
                cur += ch
        return [x.strip() for x in out + [cur] if x.strip()]

    checked = 0
    for m in re.finditer(r"ccall\(\(:(nb_[a-z0-9_]+), LIB\),\s*([A-Za-z0-9{}]+),\s*\(", src):
        name, ret = m.group(1), m.group(2)
        assert name in want, f"{name} is not declared in include/nabla_cuda.h"
        j, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        got = [cls(t) for t in split_top(src[m.end():j - 1])]
        assert cls(ret) == cls(want[name][0]) and got == [cls(t) for t in want[name][1]], f"ccall of {name}: {got}"
        checked += 1
    assert checked >= 15


def test_ctypes_signatures_match_the_header():
    """nabla.jl_b200/_lib.py declares argtypes for every entry point of the header with the same argument count,
    scalar widths and pointer positions."""
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_ccall_stubs
    want = {n: [t for t, _ in al] for n, _, al in gen_ccall_stubs.prototypes()}
    assert set(want) == set(_lib._SIGS)
    scalars = {C.c_int32: "Int32", C.c_int64: "Int64", C.c_uint32: "UInt32", C.c_uint64: "UInt64", C.c_double: "Float64",
               C.c_float: "Float32", C.c_char: "Cchar"}

    def cls(t):
        return scalars.get(t, "ptr")    # c_void_p, c_char_p, POINTER(...) and arrays are all pointers at the ABI

    for name, (argtypes, _) in _lib._SIGS.items():
        exp = ["ptr" if t.startswith(("Ptr", "Ref")) or t == "Cstring" else t.replace("Csize_t", "UInt64") for t in want[name]]
        assert [cls(t) for t in argtypes] == exp, name


def test_header_is_plain_c_and_struct_layouts_match_ctypes(tmp_path):
    """include/nabla_cuda.h compiles as strict C99 (the boundary is a C ABI: no C++ in the signatures), and the
    struct layouts the C compiler sees are the ones the ctypes mirror uses."""
    import ctypes as C
    from nabla_jl_b200 import fusion
    csrc = tmp_path / "abi.c"
    csrc.write_text("""
#include <stdio.h>
#include <stddef.h>
#include "nabla_cuda.h"
int main(void) {
  printf("%zu %zu %zu\\n", sizeof(nb_tensor), offsetof(nb_tensor, ndim), offsetof(nb_tensor, shape));
  printf("%zu\\n", sizeof(nb_stats));
  printf("%zu %zu %zu %zu\\n", sizeof(nb_gemm_epilogue), offsetof(nb_gemm_epilogue, bias), offsetof(nb_gemm_epilogue, ldy),
         offsetof(nb_gemm_epilogue, rowsum));
  return 0;
}
""")
    exe = str(tmp_path / "abi")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-o", exe, str(csrc)])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split("\n")
    T, S, E = _lib.nb_tensor, _lib.nb_stats, fusion.nb_gemm_epilogue
    assert [int(v) for v in out[0].split()] == [C.sizeof(T), T.ndim.offset, T.shape.offset]
    assert int(out[1]) == C.sizeof(S)
    assert [int(v) for v in out[2].split()] == [C.sizeof(E), E.bias.offset, E.ldy.offset, E.rowsum.offset]


def test_library_is_sm100a_and_has_no_oracle_dependency():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    # the oracle is test infrastructure: only tests/, smoke() and bench.py's CPU legs may import it
    for top in ("nabla.jl_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "nabla_oracle" not in src.replace("nb = nabla_oracle", ""), f"{top}/{f} must not import the oracle"


def test_opcodes_match_header():
    ops = _lib.OPCODES
    assert ops["identity"] == 0 and ops["tanh"] == 23 and ops["logistic"] == 55
    assert ops["add"] == 64 and ops["atan2"] == 73
    assert ops["beta"] == 74 and ops["gamma"] == 96 and ops["erfcx"] == 105
    assert ops["airyai"] == 106 and ops["invdigamma"] == 112 and ops["sf_end"] == 113
    assert ops["besselj"] == 75 and ops["polygamma"] == 79 and ops["gt"] == 80 and ops["maskn"] == 85 and ops["binary_end"] == 86
    assert len(scalar.UNARY) == 64 + 17 and len(scalar.BINARY) == 22   # + every SpecialFunctions entry the reference lists
    assert "sf_begin" not in scalar.UNARY and "binary_end" not in scalar.BINARY


def test_special_functions_source_on_the_host(tmp_path):
    """csrc/special.cuh (airy*, dawson, erfi, invdigamma: the SpecialFunctions.jl entries of test/runtests.jl:34-36
    the CUDA math library lacks) is plain C++ as well as device code: build it with g++ and check every function
    against 40-digit mpmath values over the whole real line.  The device kernels compile the same source."""
    import ctypes as C
    import subprocess
    import mpmath as mp
    import numpy as np
    lib = str(tmp_path / "libsf.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", lib,
                           os.path.join(ROOT, "tests", "special_host_shim.cpp")])
    L = C.CDLL(lib)
    L.sf_airy.restype, L.sf_airy.argtypes = C.c_double, [C.c_double, C.c_int]
    for n in ("sf_dawson", "sf_erfi", "sf_invdigamma"):
        getattr(L, n).restype, getattr(L, n).argtypes = C.c_double, [C.c_double]
    mp.mp.dps = 40
    rng = np.random.default_rng(0)
    airy = [lambda x: mp.airyai(x), lambda x: mp.airyai(x, derivative=1),
            lambda x: mp.airybi(x), lambda x: mp.airybi(x, derivative=1)]
    xs = np.concatenate([rng.uniform(-16, 16, 400), rng.uniform(-40, -16, 60), rng.uniform(16, 40, 60),
                         np.linspace(-16, 16, 65), np.linspace(-16, 16, 65)[:-1] + 0.25, [0.0, 1e-300, 16.0001, -16.0001]])
    for w in range(4):
        for x in xs:
            X = mp.mpf(float(x))
            ref = airy[w](X)
            # oscillatory side: error relative to the local amplitude (values cross zero); else to the value
            amp = mp.sqrt(airy[w & 1](X) ** 2 + airy[2 + (w & 1)](X) ** 2) if x < 0 else abs(ref)
            assert abs(L.sf_airy(float(x), w) - ref) <= 1e-13 * amp, (w, x)
    assert L.sf_airy(200.0, 0) == 0.0 and L.sf_airy(200.0, 2) == float("inf") and np.isnan(L.sf_airy(float("nan"), 1))
    xs = np.concatenate([rng.uniform(-30, 30, 500), rng.uniform(-1, 1, 100), [0.5, -0.5, 25.0, 25.0001, 1e-8, 1e5, -1e8]])
    for x in xs:
        X = mp.mpf(float(x))
        ref = mp.sqrt(mp.pi) / 2 * mp.exp(-X * X) * mp.erfi(X)
        assert abs(L.sf_dawson(float(x)) - ref) <= 4e-15 * abs(ref), x
        if abs(x) < 26:
            assert abs(L.sf_erfi(float(x)) - mp.erfi(X)) <= 4e-15 * abs(mp.erfi(X)), x
    assert L.sf_dawson(0.0) == 0.0 and L.sf_erfi(0.0) == 0.0 and L.sf_erfi(27.0) == float("inf")
    for y in np.concatenate([rng.uniform(-50, 10, 300), rng.uniform(-3.1, 3.1, 100), [-2.22, -2.2200001, 0.0, 100.0, -1e6]]):
        G = mp.mpf(L.sf_invdigamma(float(y)))
        assert abs(mp.digamma(G) - mp.mpf(float(y))) <= 1e-14 * mp.polygamma(1, G) * G, y   # relative error of x
    # integer-order Bessel functions (all regimes: series, trapezoid rules, large-argument expansions) and polygamma
    L.sf_bessel.restype, L.sf_bessel.argtypes = C.c_double, [C.c_int, C.c_double, C.c_double]
    L.sf_bessel_dx.restype, L.sf_bessel_dx.argtypes = C.c_double, [C.c_int, C.c_double, C.c_double]
    L.sf_polygamma.restype, L.sf_polygamma.argtypes = C.c_double, [C.c_double, C.c_double]
    mp.mp.dps = 30
    ref = [mp.besselj, mp.bessely, mp.besseli, mp.besselk]
    xs = np.concatenate([rng.uniform(1e-3, 3, 12), rng.uniform(3, 50, 20), rng.uniform(50, 300, 8)])
    for w in range(4):
        for n in (-1, 0, 1, 2, 5, 7):
            for x in xs:
                for X in ((float(x), -float(x)) if w in (0, 2) else (float(x),)):
                    r = ref[w](n, mp.mpf(X))
                    amp = abs(r)
                    if w <= 1 and abs(X) > abs(n):   # oscillatory: relative to the modulus
                        amp = mp.sqrt(mp.besselj(n, abs(X)) ** 2 + mp.bessely(n, abs(X)) ** 2)
                    assert abs(L.sf_bessel(w, float(n), X) - r) <= 2e-13 * amp, (w, n, X)
        for n in range(6):
            for x in rng.uniform(0.1, 3.1, 5):
                r = mp.diff(lambda t: ref[w](n, t), mp.mpf(float(x)))
                assert abs(L.sf_bessel_dx(w, float(n), float(x)) - r) <= 1e-13 * max(abs(r), 1e-3), (w, n, x)
    assert np.isnan(L.sf_bessel(0, 0.5, 1.0)) and L.sf_bessel(3, 2.0, 0.0) == float("inf") and np.isnan(L.sf_bessel(1, 0.0, -1.0))
    for n in (0, 1, 3):          # K_n next to the origin (leading-term branch below 1e-8 and just above it)
        for x in (1e-300, 1e-100, 9.9e-9, 1.01e-8, 1e-5):
            r = mp.besselk(n, mp.mpf(x))
            got = L.sf_bessel(3, float(n), x)
            assert (got == float("inf") and r > mp.mpf(1.7e308)) or abs(got - r) <= 1e-13 * r, (n, x)
    for n in range(9):
        for x in np.concatenate([rng.uniform(0.05, 30, 25), rng.uniform(-6, 0, 15), [1e-3, 100.0, 1e4]]):
            r = mp.polygamma(n, mp.mpf(float(x)))
            cond = min(1.0, 10 * abs(x - round(x))) if x < 0 else 1.0    # conditioning next to the poles
            assert abs(L.sf_polygamma(float(n), float(x)) - r) * cond <= 1e-13 * abs(r), (n, x)
    assert np.isnan(L.sf_polygamma(2.0, -3.0)) and np.isnan(L.sf_polygamma(1.5, 1.0))
    # infinite arguments give the limits, never NaN from inf * 0
    inf = float("inf")
    assert [L.sf_airy(inf, w) for w in range(4)] == [0.0, 0.0, inf, inf] and L.sf_airy(-inf, 0) == 0.0 == L.sf_airy(-inf, 2)
    assert L.sf_dawson(inf) == 0.0 and L.sf_erfi(inf) == inf and L.sf_erfi(-inf) == -inf and L.sf_invdigamma(inf) == inf
    assert [L.sf_bessel(w, 2.0, inf) for w in range(4)] == [0.0, 0.0, inf, 0.0] and L.sf_bessel(2, 3.0, -inf) == -inf
    assert L.sf_polygamma(0.0, inf) == inf and L.sf_polygamma(2.0, inf) == 0.0 and np.isnan(L.sf_polygamma(2.0, -inf))


def test_tracer_single_primitive_and_constants():
    p = nb.compile_scalar(nb.tanh, 1)
    assert p.key == (1, (23,), (0,), (0,), ())
    p = nb.compile_scalar(nb.sub, 2, {0: 1.5})          # (1+ϵ) .- f : constant baked in
    assert p.nargs == 1 and p.ops == (65,) and p.a == (-1,) and p.b == (0,) and p.consts == (1.5,)
    p = nb.compile_scalar(lambda x: x, 1)
    assert p.ops == () and p.single_op == 0


def test_tracer_recognises_logistic_and_literal_pow():
    for f in (lambda x: 1.0 / (1.0 + nb.exp(-x)), lambda x: 1 / (nb.exp(-x) + 1), lambda x: nb.inv(1 + nb.exp(-x))):
        assert nb.compile_scalar(f, 1).ops == (_lib.OPCODES["logistic"],)
    assert nb.compile_scalar(lambda x: x ** 2, 1).ops == (_lib.OPCODES["abs2"],)


def test_tracer_composite_program_with_cse():
    f = lambda x, y, z: x * y + y * z + x * z
    p = nb.compile_scalar(f, 3)
    assert p.nargs == 3 and p.ops == (66, 66, 64, 66, 64)
    g = lambda x: nb.sin(x) * nb.sin(x)   # structural CSE: sin(x) emitted once
    assert len(nb.compile_scalar(g, 1).ops) == 2
    def h(x):
        s = nb.sin(x)
        return s * s
    assert len(nb.compile_scalar(h, 1).ops) == 2    # shared sub-expression emitted once


def _run_program(orc, P, xs):
    """Python restatement of nb_prog_eval / nb_prog_grad (csrc/ops.cuh): value of the program and d(value)/d(operand k),
    with the catalogue's f / partials taken from the oracle's tables."""
    import numpy as np
    names = {v: k for k, v in _lib.OPCODES.items()}
    slot = lambda vals, i: vals[i] if i >= 0 else np.float64(P.consts[-i - 1])
    vals = [np.asarray(x, dtype=np.float64) for x in xs]
    for op, a, b in zip(P.ops, P.a, P.b):
        n = names[op]
        if n == "logistic":
            vals.append(1.0 / (1.0 + np.exp(-slot(vals, a))))
        elif n in orc._UNARY:
            vals.append(orc._UNARY[n][0](slot(vals, a)))
        else:
            vals.append(orc._BINARY[n][0](slot(vals, a), slot(vals, b)))
    y = vals[-1] if P.ops else vals[0]
    adj = [np.zeros_like(y) for _ in vals]
    adj[-1 if P.ops else 0] = np.ones_like(y)
    for i in range(len(P.ops) - 1, -1, -1):
        n, a, b, g = names[P.ops[i]], P.a[i], P.b[i], adj[P.nargs + i]
        if n == "logistic":
            yy = vals[P.nargs + i]
            adj[a] = adj[a] + g * yy * (1 - yy) if a >= 0 else adj[a]
        elif n in orc._UNARY:
            if a >= 0:
                adj[a] = adj[a] + g * orc._UNARY[n][1](slot(vals, a))
        else:
            _, dx, dy = orc._BINARY[n]
            if a >= 0:
                adj[a] = adj[a] + g * dx(slot(vals, a), slot(vals, b))
            if b >= 0:
                adj[b] = adj[b] + g * dy(slot(vals, a), slot(vals, b))
    return y, adj[:P.nargs]


def test_tracer_random_expressions_evaluate_and_differentiate_correctly(orc):
    """Property test of the tracer (operator overloads, operand swaps, constant pool, structural CSE, peephole
    rewrites): 150 random scalar expressions are traced to programs, the programs are run by a Python restatement
    of the device interpreter, and value / gradient are compared with the expression evaluated directly in NumPy
    and differentiated by central differences."""
    import numpy as np
    rng = np.random.default_rng(2024)
    un = ["tanh", "sin", "cos", "exp", "abs2", "atan", "logistic", "neg", "sinh"]

    def gen(depth, nargs):
        """-> (callable over a namespace ns and operands) ; bounded values so that nothing overflows"""
        r = rng.random()
        if depth == 0 or r < 0.15:
            k = int(rng.integers(nargs))
            return lambda ns, xs: xs[k]
        if r < 0.25:
            c = float(np.round(rng.uniform(-2, 2), 2))
            return lambda ns, xs: c
        if r < 0.55:
            name = un[int(rng.integers(len(un)))]
            sub = gen(depth - 1, nargs)

            def u(ns, xs):
                v = sub(ns, xs)
                if name == "neg":
                    return -v
                if isinstance(v, float):          # a constant subtree: tie it to an operand so both namespaces
                    v = v + 0 * xs[0]             # see a non-float from here on
                if name == "logistic":
                    return 1.0 / (1.0 + ns.exp(-v))
                if name in ("exp", "sinh"):
                    v = ns.tanh(v)                # keep the argument bounded
                return getattr(ns, name)(v)
            return u
        l, rr = gen(depth - 1, nargs), gen(depth - 1, nargs)
        kind = int(rng.integers(5))

        def b(ns, xs):
            p, q = l(ns, xs), rr(ns, xs)
            if kind == 0:
                return p + q
            if kind == 1:
                return p - q
            if kind == 2:
                return p * q
            if kind == 3:
                return p / (1.5 + q * q)
            return (p * p + 0.5) ** 1.5 if not isinstance(p, float) else p + q
        return b

    class NP:   # the same expression evaluated directly
        tanh, sin, cos, exp, atan, sinh = np.tanh, np.sin, np.cos, np.exp, np.arctan, np.sinh
        abs2 = staticmethod(lambda v: v * v)

    done = 0
    for trial in range(400):
        nargs = int(rng.integers(1, 4))
        expr = gen(4, nargs)
        f = lambda *xs: expr(nb, list(xs))
        try:
            P = nb.compile_scalar(f, nargs)
        except nb.UnsupportedOp:      # more than 32 instructions / 16 constants: legitimately refused
            continue
        except TypeError:
            continue                  # the expression collapsed to a constant without operands
        xs = [rng.uniform(-1.5, 1.5, 7) for _ in range(nargs)]
        y, grads = _run_program(orc, P, xs)
        ref = expr(NP, xs) + 0 * xs[0]
        np.testing.assert_allclose(y + 0 * xs[0], ref, rtol=1e-12, atol=1e-12, err_msg=f"trial {trial}")
        h = 1e-6
        for k in range(nargs):
            xp = [x.copy() for x in xs]; xm = [x.copy() for x in xs]
            xp[k] += h; xm[k] -= h
            fd = (expr(NP, xp) - expr(NP, xm)) / (2 * h) + 0 * xs[0]
            np.testing.assert_allclose(grads[k] + 0 * xs[0], fd, rtol=2e-5, atol=2e-7, err_msg=f"trial {trial} arg {k}")
        done += 1
        if done == 150:
            break
    assert done == 150


_SCALE = 2.0   # a module global read by a traced lambda below


def test_trace_cache_never_serves_a_stale_program():
    """The trace cache keys a Python function by its code object and the VALUES of everything the code can see.
    Regression for identity-based keys: a rebound closure cell (the freed object's id gets reused), a changed global
    and a mutated captured container must all re-trace; re-creating the same lambda with the same captured values
    must hit."""
    global _SCALE
    consts = lambda p: set(p.consts)

    def make(c):
        return lambda x: c * nb.tanh(x)
    p1, p2, p3 = nb.compile_scalar(make(0.5), 1), nb.compile_scalar(make(0.5), 1), nb.compile_scalar(make(0.75), 1)
    assert p1 is p2 and p3 is not p1 and consts(p3) == {0.75}                 # value-keyed: hit / miss
    for c in (0.1, 0.2, 0.3, 0.4):                                            # one cell, rebound every iteration
        inner = make(c)
        assert consts(nb.compile_scalar(lambda x: inner(x) + 1.0, 1)) == {c, 1.0}
    f = lambda x: _SCALE * x
    assert consts(nb.compile_scalar(f, 1)) == {2.0}
    _SCALE = 3.0
    try:
        assert consts(nb.compile_scalar(f, 1)) == {3.0}                       # the global changed: re-traced
    finally:
        _SCALE = 2.0
    box = {"a": 1.5}
    g = lambda x: box["a"] * x
    assert consts(nb.compile_scalar(g, 1)) == {1.5}
    box["a"] = 2.5
    assert consts(nb.compile_scalar(g, 1)) == {2.5}                           # captured container: never cached
    assert scalar._trace_key(g, 1, {}) is None and scalar._trace_key(f, 1, {}) is not None
    assert nb.compile_scalar(nb.tanh, 1) is nb.compile_scalar(nb.tanh, 1)


def test_tracer_rejects_unsupported():
    with pytest.raises(nb.UnsupportedOp):
        nb.compile_scalar(lambda x: 1.0 if x > 0 else 2.0, 1)       # data-dependent branch
    def too_long(x):
        for _ in range(40):
            x = nb.sin(x)
        return x
    with pytest.raises(nb.UnsupportedOp):
        nb.compile_scalar(too_long, 1)


def test_no_cpu_fallback_without_gpu():
    n = ctypes.c_int32()
    _lib.lib().nb_device_count(ctypes.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(nb.NablaCudaError, match="no CUDA device"):
        nb.Context()


def test_julia_broadcast_shape_rule():
    from nabla_jl_b200.kernels import jl_broadcast_shape
    assert jl_broadcast_shape((5,), (5, 7)) == (5, 7)          # vector is a column
    assert jl_broadcast_shape((1, 7), (5,)) == (5, 7)
    assert jl_broadcast_shape((), (3, 1, 2)) == (3, 1, 2)
    with pytest.raises(nb.DimensionMismatch):
        jl_broadcast_shape((5,), (7,))


def test_partial_needs_table_matches_the_kernels():
    """fusion.partial_needs (host) decides which operands a pullback launch is handed; it must equal the
    table the kernels use (csrc/ops.cuh nb_partial_needs, exported as nb_op_partial_needs)."""
    import nabla_jl_b200 as nb
    from nabla_jl_b200 import fusion
    L = nb._lib.lib()
    for name, op in nb._lib.OPCODES.items():
        if name.endswith(("_end", "_begin")):
            continue
        for wrt in (0, 1):
            assert fusion.partial_needs(op, wrt) == L.nb_op_partial_needs(op, wrt), (name, wrt)


def test_pullback_reads():
    import nabla_jl_b200 as nb
    from nabla_jl_b200 import fusion
    add = nb.compile_scalar(nb.CATALOGUE["add"], 2)
    assert fusion.pullback_reads(add, [0, 1]) == (set(), False)
    mul = nb.compile_scalar(nb.CATALOGUE["mul"], 2)
    assert fusion.pullback_reads(mul, [0]) == ({1}, False)
    assert fusion.pullback_reads(mul, [0, 1]) == ({0, 1}, False)
    tanh = nb.compile_scalar(nb.CATALOGUE["tanh"], 1)
    assert fusion.pullback_reads(tanh, [0]) == (set(), True)
    sin = nb.compile_scalar(nb.CATALOGUE["sin"], 1)
    assert fusion.pullback_reads(sin, [0]) == ({0}, False)
    comp = nb.compile_scalar(lambda a, b: nb.tanh(a) * b, 2)
    assert fusion.pullback_reads(comp, [0]) is None


def _pullback_plan(P, wanted):
    """C side: nb_prog_pullback_plan of the traced program -> (live_instr, fwd_skip, arg_live) bit masks."""
    L = _lib.lib()
    n = len(P.ops)
    I = ctypes.c_int32 * max(n, 1)
    D = ctypes.c_double * max(len(P.consts), 1)
    h = ctypes.c_void_p()
    assert L.nb_prog_create(P.nargs, n, I(*P.ops), I(*P.a), I(*P.b), len(P.consts), D(*P.consts), ctypes.byref(h)) == 0
    live, skip, alive = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    assert L.nb_prog_pullback_plan(h, wanted, ctypes.byref(live), ctypes.byref(skip), ctypes.byref(alive)) == 0
    L.nb_prog_destroy(h)
    return live.value, skip.value, alive.value


def _pullback_plan_py(P, wanted):
    """Independent restatement: which instructions a pullback for the operands in `wanted` runs in reverse, and which
    forward values it has to compute (everything a computed partial reads, plus what those values are computed from)."""
    from nabla_jl_b200 import fusion
    na, n = P.nargs, len(P.ops)
    binary = lambda op: _lib.OPCODES["add"] <= op < _lib.OPCODES["binary_end"]
    operands = lambda i: [P.a[i]] + ([P.b[i]] if binary(P.ops[i]) else [])
    reach = [bool(wanted >> k & 1) for k in range(na)]
    for i in range(n):
        reach.append(any(s >= 0 and reach[s] for s in operands(i)))
    live, need, written = [False] * n, [False] * (na + n), set()
    if n:
        live[n - 1] = True
    for i in range(n - 1, -1, -1):
        if not (live[i] and reach[na + i]):
            live[i] = False
            continue
        for side, s in enumerate(operands(i)):
            if s < 0 or not reach[s]:
                continue
            bits = fusion.partial_needs(P.ops[i], side)
            if bits & 1 and P.a[i] >= 0:
                need[P.a[i]] = True
            if bits & 2 and binary(P.ops[i]) and P.b[i] >= 0:
                need[P.b[i]] = True
            if bits & 4:
                need[na + i] = True
            written.add(s)
            if s >= na:
                live[s - na] = True
    skip = 0
    for i in range(n - 1, -1, -1):
        if not need[na + i]:
            skip |= 1 << i
            continue
        for s in operands(i):
            if s >= 0:
                need[s] = True
    live_mask = sum(1 << i for i in range(n) if live[i])
    alive = sum(1 << k for k in range(na) if k in written or n == 0)
    return live_mask, skip, alive


def test_pullback_plan_skips_what_no_partial_reads():
    """examples/vae.jl:79 / mlp.jl:73: the log-likelihood term `x * log(f + eps) + (1 - x) * log(1 + eps - f)` with `x`
    data.  The sensitivity of `f` needs x and the two arguments of the logarithms — never the logarithms: the pullback
    launch for f alone skips them (fmad's duals evaluate everything, core.jl:285-301); asked for x as well it cannot."""
    eps = 1e-12
    P = nb.compile_scalar(lambda x, f: x * nb.log(f + eps) + (1.0 - x) * nb.log(1.0 + eps - f), 2)
    logs = [i for i, op in enumerate(P.ops) if op == _lib.OPCODES["log"]]
    assert len(logs) == 2
    live, skip, alive = _pullback_plan(P, 0b10)          # wanted: f (operand 1)
    assert all(skip >> i & 1 for i in logs) and alive == 0b10
    assert all(live >> i & 1 for i in logs)               # the adjoint still flows THROUGH log (its partial reads a)
    live_x, skip_x, alive_x = _pullback_plan(P, 0b01)    # wanted: x — d/dx = log(..) - log(..): the values are needed
    assert not any(skip_x >> i & 1 for i in logs) and alive_x == 0b01
    assert not any(live_x >> i & 1 for i in logs)         # ... but no adjoint goes through the logarithms towards f
    _, skip_all, alive_all = _pullback_plan(P, 0b11)
    assert not any(skip_all >> i & 1 for i in logs) and alive_all == 0b11
    assert skip_all & ~skip == 0 and skip_all != skip             # wanting more never skips more
    # an operand that does not reach the result has no sensitivity
    Q = nb.compile_scalar(lambda a, b: nb.tanh(a) + 0.0 * b, 2)
    assert _pullback_plan(Q, 0b11)[2] in (0b01, 0b11)


def test_pullback_plan_matches_its_restatement_on_random_programs():
    """nb_prog_derive (csrc/ops.cuh) against an independent Python restatement: for random traced programs and every
    subset of wanted operands the reverse instructions, the skipped forward values and the live operands are equal."""
    import numpy as np
    rng = np.random.default_rng(77)
    un = [nb.tanh, nb.sin, nb.exp, nb.log, nb.sqrt, nb.abs2, lambda v: -v, lambda v: 1.0 / (1.0 + nb.exp(-v))]

    def gen(depth, nargs):
        r = rng.random()
        if depth == 0 or r < 0.2:
            k = int(rng.integers(nargs))
            return lambda xs: xs[k]
        if r < 0.3:
            c = float(np.round(rng.uniform(0.5, 2), 2))
            return lambda xs: c
        if r < 0.6:
            f, sub = un[int(rng.integers(len(un)))], gen(depth - 1, nargs)
            return lambda xs: f(sub(xs) + 0 * xs[0])
        l, rr, kind = gen(depth - 1, nargs), gen(depth - 1, nargs), int(rng.integers(4))
        return lambda xs: (l(xs) + rr(xs), l(xs) - rr(xs), l(xs) * rr(xs), l(xs) / (2.0 + rr(xs)))[kind]

    done = 0
    for _ in range(300):
        nargs = int(rng.integers(1, 5))
        expr = gen(4, nargs)
        try:
            P = nb.compile_scalar(lambda *xs: expr(list(xs)), nargs)
        except (nb.UnsupportedOp, TypeError):
            continue
        if len(P.ops) < 2:
            continue
        for wanted in range(1, 1 << nargs):
            assert _pullback_plan(P, wanted) == _pullback_plan_py(P, wanted), (P.ops, P.a, P.b, wanted)
        done += 1
        if done == 120:
            break
    assert done == 120


def test_copy_into_an_array_flushes_only_the_launches_that_read_it():
    """Host logic behind DevArray.copy_from_pinned / copy_from_host (device.Context.flush_readers_of): a deferred launch
    must run before a buffer it READS is overwritten — directly, through a view, or through an operand that is itself
    still pending — and nothing else may be launched (a pending product nobody asked for would otherwise run on every
    new batch: the end-to-end step once took twice the resident one for exactly this reason).  No GPU: stand-in
    objects with the attributes the method touches."""
    import weakref
    import numpy as np
    from nabla_jl_b200.device import Context, DevArray

    ctx = Context.__new__(Context)
    ctx._deferred = []

    def arr(base=None):
        a = DevArray.__new__(DevArray)
        a.ctx, a._ptr, a.shape, a.dtype = None, None, (2, 2), np.dtype(np.float64)
        a.shared, a.adjvec, a._base, a.pending, a._tensor = False, False, base, None, None
        return a

    class Pending:
        def __init__(self, reads, out):
            self.reads, self.done, self.flushed = list(reads), False, 0
            out.pending = self
            ctx._deferred.append(weakref.ref(self))

        def operands(self):
            return list(self.reads)

        def flush(self):
            self.done = True
            self.flushed += 1

    X, Y, W = arr(), arr(), arr()
    Xview = arr(base=X)
    z1, z2, z3, z4 = arr(), arr(), arr(), arr()
    p1 = Pending([W, X], z1)          # reads X directly
    p2 = Pending([W, Xview], z2)      # reads a view of X
    p3 = Pending([W, z1], z3)         # reads z1: once p1 has run, z1 is an ordinary buffer and p3 no longer depends on X
    p4 = Pending([W, Y], z4)          # does not read X
    ctx.flush_readers_of(X)
    assert (p1.flushed, p2.flushed, p3.flushed, p4.flushed) == (1, 1, 0, 0)
    assert [r() for r in ctx._deferred] == [p3, p4]
    ctx.flush_readers_of(Xview)       # writing through a view is writing the base: nothing of X is pending any more
    assert (p3.flushed, p4.flushed) == (0, 0)
    ctx.flush_readers_of(Y)
    assert p4.flushed == 1 and [r() for r in ctx._deferred] == [p3]
    # a launch that reads X only THROUGH a value that is still pending must run (it would force that value later,
    # after X has changed)
    ctx._deferred = []
    u1, u2 = arr(), arr()
    q2 = Pending([W, u1], u2)         # registered first, reads u1 ...
    q1 = Pending([W, X], u1)          # ... which is pending and reads X
    ctx.flush_readers_of(X)
    assert (q1.flushed, q2.flushed) == (1, 1) and ctx._deferred == []


def test_composed_programs_equal_the_branches_they_replace(orc):
    """Tape-level fusion of adjacent broadcast Branches (ewfusion.py) composes the programs of consecutive dot calls
    with scalar.symbolic + scalar.linearize.  Property: for random chains `g.(f.(x, y), z)` the composed program,
    run by the Python restatement of the device interpreter, gives the value and the operand sensitivities of the two
    Branches evaluated one after the other (functional.jl:48-51 tapes them separately; the chain rule joins them)."""
    import numpy as np
    from nabla_jl_b200.scalar import Sym, linearize, symbolic
    rng = np.random.default_rng(99)
    inner = [lambda a, b: nb.tanh(a + b), lambda a, b: a * b + 0.5, lambda a, b: nb.exp(-a * a) - b,
             lambda a, b: a / (2.0 + b * b), lambda a, b: nb.sin(a) * nb.cos(b)]
    outer = [lambda u, c: nb.log(u * u + 1.5) * c, lambda u, c: u * c + u, lambda u, c: nb.tanh(u) - c * c,
             lambda u, c: 1.0 / (1.0 + nb.exp(-u)) + c, lambda u, c: nb.sqrt(u * u + c * c + 0.1)]
    for fi in inner:
        for go in outer:
            Pf, Pg = nb.compile_scalar(fi, 2), nb.compile_scalar(go, 2)
            # leaves (x, y, z): g's first operand is the lazy value f(x, y), its second the leaf z
            root = symbolic(Pg, [symbolic(Pf, [Sym.arg(0), Sym.arg(1)]), Sym.arg(2)])
            Pc = linearize(root, 3, "fused")
            x, y, z = (rng.uniform(-1.2, 1.2, 9) for _ in range(3))
            u, (ux, uy) = _run_program(orc, Pf, [x, y])
            v, (vu, vz) = _run_program(orc, Pg, [u, z])
            w, (wx, wy, wz) = _run_program(orc, Pc, [x, y, z])
            np.testing.assert_allclose(w, v, rtol=1e-13, atol=1e-13)
            np.testing.assert_allclose(wx, vu * ux, rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(wy, vu * uy, rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(wz, vz, rtol=1e-12, atol=1e-13)
            # the same operand on both levels (sigma .* sigma style reuse): one leaf, contributions add up
            root2 = symbolic(Pg, [symbolic(Pf, [Sym.arg(0), Sym.arg(1)]), Sym.arg(0)])
            Pd = linearize(root2, 2, "fused")
            w2, (w2x, w2y) = _run_program(orc, Pd, [x, y])
            v2, (v2u, v2c) = _run_program(orc, Pg, [u, x])
            np.testing.assert_allclose(w2, v2, rtol=1e-13, atol=1e-13)
            np.testing.assert_allclose(w2x, v2u * ux + v2c, rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(w2y, v2u * uy, rtol=1e-12, atol=1e-13)


def test_ozaki_split_scheme_in_numpy():
    """The Float64-on-int8 scheme of csrc/ozaki.cu restated in NumPy (no GPU): power-of-two row/column scales,
    S signed 7-bit slices by q = rint(128 x), x <- 128 x - q, slices concatenated along K (A in order, B
    reversed) so that level l is ONE integer product over K' = (l+1) K, exact in int32, combined with
    2^(-7(l+2)).  Pins the slice ranges, the exactness bound and the accuracy the device path is tested against,
    for the default seven slices and the looser NB_OZAKI_SLICES=6."""
    import numpy as np
    for S, rec_tol, norm_tol, elem_tol in ((7, 2.0 ** -50, 2e-13, 1e-12), (6, 2.0 ** -43, 2e-11, 1e-10)):
        _ozaki_numpy(np, S, rec_tol, norm_tol, elem_tol)


def _ozaki_numpy(np, S, rec_tol, norm_tol, elem_tol):
    rng = np.random.default_rng(5)
    m, n, k = 37, 29, 200
    A = rng.standard_normal((m, k)) * np.exp(rng.uniform(-4, 4, (m, 1)))
    B = rng.standard_normal((k, n)) * np.exp(rng.uniform(-4, 4, (1, n)))

    def scales(X, axis):
        mx = np.max(np.abs(X), axis=axis, keepdims=True)
        return np.ldexp(1.0, np.frexp(mx)[1] + 1)      # 2^(ilogb(mx) + 2): |x| / scale < 1/2

    def slices(X):
        out = []
        for _ in range(S):
            t = X * 128.0
            q = np.rint(t)
            out.append(q.astype(np.int64))
            X = t - q                                    # exact
        return out

    ra, cb = scales(A, 1), scales(B, 0)
    assert np.all(np.abs(A / ra) < 0.5) and np.all(np.abs(B / cb) < 0.5)
    sa, sb = slices(A / ra), slices(B / cb)
    assert all(np.abs(q).max() <= 64 for q in sa + sb)
    # the slices reconstruct the scaled operand to 2^-(7 S + 1)
    rec = sum(q * 2.0 ** (-7 * (p + 1)) for p, q in enumerate(sa))
    assert np.max(np.abs(rec - A / ra)) <= rec_tol
    Acat = np.concatenate(sa, axis=1)                   # m x (S k): slices 0 .. S-1
    Bcat = np.concatenate(sb[::-1], axis=0)             # (S k) x n: slices S-1 .. 0
    C = np.zeros((m, n))
    for l in range(S):
        acc = Acat[:, :(l + 1) * k] @ Bcat[(S - 1 - l) * k:, :]        # = sum_{p+q=l} A_p B_q, one product
        direct = sum(sa[p] @ sb[l - p] for p in range(l + 1))
        assert np.array_equal(acc, direct)
        assert np.abs(acc).max() < 2 ** 31 and (l + 1) * k * 4096 < 2 ** 31     # exact in int32
        C += np.ldexp(acc.astype(np.float64), -7 * (l + 2))
    C *= ra * cb
    ref = A @ B
    assert np.linalg.norm(C - ref) / np.linalg.norm(ref) < norm_tol
    assert np.max(np.abs(C - ref) / (np.abs(A).max(1, keepdims=True) * np.abs(B).max(0, keepdims=True) * np.sqrt(k))) < elem_tol


def test_tracer_lowers_selection_and_refuses_control_flow():
    """ifelse / comparisons trace to MASK / MASKN / GT ... (include/nabla_cuda.h); Python control flow on a traced value
    cannot be traced and raises UnsupportedOp (the Julia glue: TypeError for a non-Bool condition)."""
    import nabla_jl_b200 as nb
    ops = _lib.OPCODES
    p = nb.compile_scalar(lambda x: nb.ifelse(x > 0.0, x, 0.1 * x), 1)
    assert p.ops == (ops["gt"], ops["mask"], ops["mul"], ops["maskn"], ops["add"])
    assert nb.compile_scalar(lambda x, y: (x <= y) * x, 2).ops == (ops["le"], ops["mul"])
    for bad in (lambda x: x if x > 0 else 0.1 * x, lambda x: max(x, 0.0), lambda x: (x > 0 and x) or 0.0):
        with pytest.raises(nb.UnsupportedOp):
            nb.compile_scalar(bad, 1)
    # a host-side constant condition is just Python
    assert nb.ifelse(True, 1.0, 2.0) == 1.0
