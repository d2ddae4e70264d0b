"""nabla.jl_b200 — B200-native backend for the array-valued reverse sweep of invenia/Nabla.jl.

Public surface mirrors the reference (``∇`` is spelled ``nabla``): ``Tape``, ``Leaf``, ``Branch``,
``nabla(f; get_output)`` / ``nabla(y[, ȳ])``, ``broadcast`` / ``map_`` / ``sum`` / ``mean`` /
``mapreduce`` / ``dot`` / ``matmul`` / ``adjoint`` / ``gemm`` / ``gemv``, the scalar catalogue
(``tanh``, ``exp``, ``log``, ...), and the registration surface (``explicit_intercepts``,
``Primitive``).  All array arithmetic runs in libnabla_cuda.so (hand-written sm_100a kernels); the
package raises if the library or a GPU is missing — there is no CPU path (see oracle/ for the
test-only CPU restatement)."""
from . import _lib
from ._lib import (DimensionMismatch, NablaCudaError, UnsupportedOp, MATH_DEFAULT, MATH_TF32,
                   MATH_3XTF32, MATH_FP32, MATH_BF16X3, MATH_FP16X3)
from .device import BroadcastView, Context, DevArray, PinnedArray, get_context, set_context
from .core import (Arg, Branch, Leaf, Node, Primitive, Tape, explicit_intercepts, grad, nabla,
                   oneslike, pos, propagate, reverse_tape, tape, unbox, unthunk, zeroslike, as_device)
from .scalar import CATALOGUE, Program, ScalarFn, Sym, compile_scalar, ifelse
from .rules import (Ref, adjoint, broadcast, dot, gemm, gemv, ldivide, map_, mapfoldl, mapfoldr,
                    mapreduce, matmul, mean, sum, transpose)

from .blasrules import symm, symv, trmm, trmv, trsm, trsv
from .arrayrules import R, checkpoint, colon, copy, fill, getindex, hcat, kron, plus_I, reshape, vcat
from .sweep import (CapturedGradient, Communicator, OverlappedShardedGradient, ShardedGradient,
                    shard_bounds)

for _n, _f in CATALOGUE.items():
    globals()[{"abs": "abs_", "max": "max_", "min": "min_", "pow": "pow_"}.get(_n, _n)] = _f
UNARY_NAMES = tuple(n for n, f in CATALOGUE.items() if f.arity == 1)
BINARY_NAMES = tuple(n for n, f in CATALOGUE.items() if f.arity == 2)


def scalar_fn(name):
    return CATALOGUE[name]


def to_host(x):
    """Device value -> NumPy (test / user convenience)."""
    x = unthunk(unbox(x))
    return x.numpy() if isinstance(x, DevArray) else x
