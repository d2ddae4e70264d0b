"""ctypes binding of libnabla_cuda.so (include/nabla_cuda.h).

This is the Python twin of the `ccall` stubs in julia/NablaCUDA.jl.  There is NO CPU fallback: if
the shared library is missing, or no CUDA device is present when a context is created, the error is
raised to the caller (BASELINE.json north_star: "no CPU fallback")."""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NABLA_CUDA_LIB") or os.path.join(_HERE, "libnabla_cuda.so")  # (same override as julia/NablaCUDA.jl)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "nabla_cuda.h")

NB_MAX_DIMS = 4
NB_MAX_ARGS = 10
NB_MAX_INSTR = 32
NB_MAX_CONSTS = 16
NB_F32, NB_F64 = 0, 1
MATH_DEFAULT, MATH_TF32, MATH_3XTF32, MATH_FP32, MATH_BF16X3, MATH_FP16X3 = 0, 1, 2, 3, 4, 5
STATUS_NAMES = {0: "NB_OK", 1: "NB_ERR_INVALID", 2: "NB_ERR_CUDA", 3: "NB_ERR_OOM",
                4: "NB_ERR_UNSUPPORTED", 5: "NB_ERR_COMM", 6: "NB_ERR_SHAPE"}


class NablaCudaError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {msg}")
        self.status = status


class DimensionMismatch(NablaCudaError, ValueError):
    pass


class UnsupportedOp(NablaCudaError, TypeError):
    """The device has no rule for this op (the Julia glue raises MethodError)."""


class nb_tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * NB_MAX_DIMS)]


class nb_stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("pool_bytes_in_use", C.c_uint64),
                ("pool_bytes_reserved", C.c_uint64), ("pool_high_water", C.c_uint64),
                ("cuda_mallocs", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


def parse_opcodes(header_path=HEADER_PATH):
    """The opcode values are part of the ABI; read them from the header so the two never drift."""
    text = open(header_path).read()
    body = text[text.index("typedef enum nb_op"):text.index("} nb_op;")]
    return {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"NB_OP_([A-Z0-9_]+)\s*=\s*(\d+)", body)}


def declared_symbols(header_path=HEADER_PATH):
    text = open(header_path).read()
    return sorted(set(re.findall(r"\b(nb_[a-z0-9_]+)\s*\(", text)))


OPCODES = parse_opcodes()

_p = C.c_void_p
_pp = C.POINTER(C.c_void_p)
_i32, _i64, _u64, _u32, _dbl = C.c_int32, C.c_int64, C.c_uint64, C.c_uint32, C.c_double
_tp = C.POINTER(nb_tensor)

_SIGS = {
    "nb_abi_version": ([], _i32),
    "nb_last_error": ([_p], C.c_char_p),
    "nb_ctx_create": ([_i32, _u64, _pp], _i32),
    "nb_ctx_destroy": ([_p], _i32),
    "nb_sweep_begin": ([_p], _i32),
    "nb_sweep_end": ([_p], _i32),
    "nb_sync": ([_p], _i32),
    "nb_ctx_stream": ([_p, _pp], _i32),
    "nb_ctx_stats": ([_p, C.POINTER(nb_stats)], _i32),
    "nb_device_count": ([C.POINTER(_i32)], _i32),
    "nb_alloc": ([_p, _u64, _pp], _i32),
    "nb_retain": ([_p, _p], _i32),
    "nb_release": ([_p, _p], _i32),
    "nb_refcount": ([_p, _p, C.POINTER(_i32)], _i32),
    "nb_h2d": ([_p, _p, _p, _u64], _i32),
    "nb_d2h": ([_p, _p, _p, _u64], _i32),
    "nb_d2h_async": ([_p, _p, _p, _u64], _i32),
    "nb_d2d": ([_p, _p, _p, _u64], _i32),
    "nb_fill": ([_p, _p, _i32, _u64, _dbl], _i32),
    "nb_host_alloc": ([_u64, _pp], _i32),
    "nb_host_free": ([_p], _i32),
    "nb_buffer_written": ([_p, _p, _u64], _i32),
    "nb_prog_create": ([_i32, _i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), _i32,
                        C.POINTER(_dbl), _pp], _i32),
    "nb_prog_destroy": ([_p], _i32),
    "nb_prog_pullback_plan": ([_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)], _i32),
    "nb_bcast_fwd": ([_p, _p, _i32, _tp, _tp], _i32),
    "nb_bcast_pullback": ([_p, _p, _tp, _tp, _i32, _tp, _u32, _tp, _u32], _i32),
    "nb_bcast_reduce": ([_p, _p, _i32, _tp, _dbl, _tp, _i32], _i32),
    "nb_gemm": ([_p, C.c_char, C.c_char, _i64, _i64, _i64, _dbl, _p, _i64, _p, _i64, _dbl, _p, _i64,
                 _i32, _i32], _i32),
    "nb_op_partial_needs": ([_i32, _i32], _i32),
    "nb_gemm_fused": ([_p, C.c_char, C.c_char, _i64, _i64, _i64, _dbl, _p, _i64, _p, _i64, _p, _i64, _i32, _i32,
                       _p], _i32),
    "nb_gemv": ([_p, C.c_char, _i64, _i64, _dbl, _p, _i64, _p, _dbl, _p, _i32], _i32),
    "nb_ger": ([_p, _i64, _i64, _dbl, _p, _p, _dbl, _p, _i64, _i32], _i32),
    "nb_transpose": ([_p, _i64, _i64, _p, _i64, _p, _i64, _i32], _i32),
    "nb_tri": ([_p, _i32, _i32, _i64, _p, _i64, _p, _i64, _i32], _i32),
    "nb_trsm": ([_p, C.c_char, C.c_char, C.c_char, C.c_char, _i64, _i64, _dbl, _p, _i64, _p, _i64, _i32], _i32),
    "nb_copy2d": ([_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i32, _i32], _i32),
    "nb_gather": ([_p, _i64, _p, _p, _p, _i32], _i32),
    "nb_scatter_add": ([_p, _i64, _p, _p, _p, _i32], _i32),
    "nb_comm_unique_id": ([C.POINTER(C.c_uint8)], _i32),
    "nb_comm_init": ([_p, _i32, _i32, C.POINTER(C.c_uint8)], _i32),
    "nb_allreduce_sum": ([_p, _i32, _pp, C.POINTER(_i64), _i32], _i32),
    "nb_allreduce_sum_async": ([_p, _i32, _pp, C.POINTER(_i64), _i32], _i32),
    "nb_comm_join": ([_p], _i32),
    "nb_comm_destroy": ([_p], _i32),
    "nb_graph_begin": ([_p], _i32),
    "nb_graph_end": ([_p, _pp], _i32),
    "nb_graph_launch": ([_p, _p], _i32),
    "nb_graph_destroy": ([_p], _i32),
    "nb_event_create": ([_pp], _i32),
    "nb_event_record": ([_p, _p], _i32),
    "nb_event_elapsed_ms": ([_p, _p, C.POINTER(C.c_float)], _i32),
    "nb_event_destroy": ([_p], _i32),
}

_lib = None


def lib():
    """Load libnabla_cuda.so (once).  Fails loudly when the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  nabla.jl_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (argtypes, restype) in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = L
    return _lib


def check(status, ctx_handle=None):
    if status == 0:
        return
    msg = lib().nb_last_error(ctx_handle)
    msg = msg.decode() if msg else ""
    if status == 6:
        raise DimensionMismatch(status, msg)
    if status == 4:
        raise UnsupportedOp(status, msg)
    raise NablaCudaError(status, msg)
