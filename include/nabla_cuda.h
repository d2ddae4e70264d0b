/*
 * nabla_cuda.h — C ABI of libnabla_cuda.so, the B200 (sm_100a) backend for the array-valued
 * reverse sweep of invenia/Nabla.jl.
 *
 * Callable identically from Julia `ccall` (julia/NablaCUDA.jl) and Python `ctypes`
 * (nabla.jl_b200/_lib.py).  Plain pointers and sizes only.  Every function returns an int32 status
 * (NB_OK == 0), never throws or aborts; `nb_last_error(ctx)` gives the message.  One `nb_ctx` per
 * host thread / reverse sweep; all device work is enqueued asynchronously on the ctx stream and is
 * ordered by it (the reference sweep is strictly sequential: src/core.jl:160-166).
 *
 * Arrays follow Julia: COLUMN-MAJOR, dense, shape[0] is the contiguous dimension, ndim <= 4,
 * ndim == 0 is a device-resident scalar.  An operand whose extent in a dimension is 1 (or whose
 * ndim is smaller) while the result's is larger is broadcast along it, exactly like
 * `Broadcast.combine_axes` (src/sensitivities/functional/functional.jl:60).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference
 * checkout).
 */
#ifndef NABLA_CUDA_H
#define NABLA_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB_ABI_VERSION 1
#define NB_MAX_DIMS 4
#define NB_MAX_ARGS 10     /* operands of one broadcast / map call (the reference generates map for up to ten: functional.jl:10-18) */
#define NB_MAX_INSTR 32    /* instructions of one scalar program */
#define NB_MAX_CONSTS 16

typedef struct nb_ctx nb_ctx;
typedef struct nb_prog nb_prog;
typedef struct nb_graph nb_graph;

typedef enum nb_status {
  NB_OK = 0,
  NB_ERR_INVALID = 1,      /* bad argument (Julia side raises ArgumentError) */
  NB_ERR_CUDA = 2,         /* CUDA runtime/driver error */
  NB_ERR_OOM = 3,
  NB_ERR_UNSUPPORTED = 4,  /* no rule for this op on the device (Julia side raises MethodError) */
  NB_ERR_COMM = 5,         /* NCCL error or NCCL not loadable */
  NB_ERR_SHAPE = 6         /* DimensionMismatch */
} nb_status;

typedef enum nb_dtype { NB_F32 = 0, NB_F64 = 1 } nb_dtype;

/* GEMM arithmetic.
 * Float32 on the tensor cores: operands are split x = hi + lo into 16-bit (or tf32) halves and the product is
 * issued as lo*hi + hi*lo + hi*hi into one FP32 accumulator.  The split copies are the library's own temporaries,
 * cached per source array until that array is written through an nb_* entry point or released (a caller that writes
 * an array with its own kernels on nb_ctx_stream must call nb_buffer_written).
 * Float64: exact FP64 FMA on the DMMA path for small products, exact int8 slice products on tcgen05 (Ozaki split)
 * for m*n*k >= 3e9 — its error is ~2^-49 relative to max|row of A| * max|col of B| (4e-14 norm-wise measured), NOT
 * relative to |A||B| element-wise like dgemm: rows or columns with a very large dynamic range lose low-order bits of
 * their small entries.  A range guard measures, in the pass that finds the row / column scales, how far the largest
 * entry of a line stands above the others (peak = max * (K-1) / (sum - max)); when peak_A * peak_B exceeds
 * NB_OZAKI_PEAK_LIMIT (default 2^24, i.e. ordinary entries of a line would keep fewer than 25 bits; 4096 = strict,
 * 0 = off) the product runs on DMMA instead — decided on the device, no host synchronisation.
 * NB_MATH_FP32 (or NB_F64_OZAKI=0) keeps every Float64 product on DMMA. */
typedef enum nb_math_mode {
  NB_MATH_DEFAULT = 0,  /* f32: NB_MATH_FP16X3; f64: DMMA / Ozaki split by size */
  NB_MATH_TF32 = 1,     /* f32: single-pass kind::tf32 on tcgen05 (~8e-4 relative error) */
  NB_MATH_3XTF32 = 2,   /* f32: x = hi + lo in tf32 (21 bits), three kind::tf32 passes (~4e-7) */
  NB_MATH_FP32 = 3,     /* f32: FP32 FMA on CUDA cores (no tensor cores); f64: DMMA only */
  NB_MATH_BF16X3 = 4,   /* f32: x = hi + lo in bfloat16 (16 bits, any dynamic range), three kind::f16 passes (~5e-6 per
                           product).  Enough for the adjoint products, whose error enters the gradient linearly; NOT
                           for forward products that feed saturating activations (width 8192: 7e-4 on the gradient).
                           A cached binary16 split of an operand (a forward value) is re-used when there is one. */
  NB_MATH_FP16X3 = 5    /* f32: x * 2^s = hi + lo in binary16 with one power-of-two scale per array (22 bits for
                           every element within 2^18 of the array's largest magnitude, absolute floor 2^-40 of it
                           below), three kind::f16 passes at the BF16X3 cost (~3e-7 per product): FP32-class */
} nb_math_mode;

typedef struct nb_tensor {
  void* data;                  /* device pointer returned by nb_alloc (may be offset into it) */
  int32_t dtype;               /* nb_dtype */
  int32_t ndim;                /* 0..NB_MAX_DIMS */
  int64_t shape[NB_MAX_DIMS];  /* column-major contiguous; entries >= ndim are ignored */
} nb_tensor;

/* Scalar-function catalogue: the derivative of each is the DiffRules / ChainRules @scalar_rule
 * formula that ForwardDiff applies inside `fmad` (src/core.jl:285-292); the list is the one pinned
 * by test/runtests.jl:28-52 (Base part).  Values are part of the ABI. */
typedef enum nb_op {
  /* unary */
  NB_OP_IDENTITY = 0, NB_OP_NEG = 1, NB_OP_ABS = 2, NB_OP_ABS2 = 3, NB_OP_SQRT = 4, NB_OP_CBRT = 5,
  NB_OP_INV = 6, NB_OP_EXP = 7, NB_OP_EXP2 = 8, NB_OP_EXP10 = 9, NB_OP_EXPM1 = 10, NB_OP_LOG = 11,
  NB_OP_LOG2 = 12, NB_OP_LOG10 = 13, NB_OP_LOG1P = 14, NB_OP_SIN = 15, NB_OP_COS = 16,
  NB_OP_TAN = 17, NB_OP_SEC = 18, NB_OP_CSC = 19, NB_OP_COT = 20, NB_OP_SINH = 21,
  NB_OP_COSH = 22, NB_OP_TANH = 23, NB_OP_SECH = 24, NB_OP_CSCH = 25, NB_OP_COTH = 26,
  NB_OP_ASIN = 27, NB_OP_ACOS = 28, NB_OP_ATAN = 29, NB_OP_ASEC = 30, NB_OP_ACSC = 31,
  NB_OP_ACOT = 32, NB_OP_ASINH = 33, NB_OP_ACOSH = 34, NB_OP_ATANH = 35, NB_OP_ASECH = 36,
  NB_OP_ACSCH = 37, NB_OP_ACOTH = 38, NB_OP_SIND = 39, NB_OP_COSD = 40, NB_OP_TAND = 41,
  NB_OP_SECD = 42, NB_OP_CSCD = 43, NB_OP_COTD = 44, NB_OP_ASIND = 45, NB_OP_ACOSD = 46,
  NB_OP_ATAND = 47, NB_OP_ASECD = 48, NB_OP_ACSCD = 49, NB_OP_ACOTD = 50, NB_OP_SINPI = 51,
  NB_OP_COSPI = 52, NB_OP_DEG2RAD = 53, NB_OP_RAD2DEG = 54,
  NB_OP_LOGISTIC = 55,  /* 1/(1+exp(-x)): examples/mlp.jl:42, examples/vae.jl:32, fused */
  NB_OP_ERF = 56, NB_OP_ERFC = 57,
  /* piecewise-constant functions (derivative zero, as ForwardDiff gives): the building blocks of value-dependent
   * scalar functions such as `x > 0 ? x : 0.1x`, which fmad differentiates along the taken branch (core.jl:285-292) */
  NB_OP_SIGN = 58,
  NB_OP_STEP = 59,      /* x > 0 ? 1 : 0 */
  NB_OP_FLOOR = 60, NB_OP_CEIL = 61, NB_OP_ROUND = 62, NB_OP_TRUNC = 63,
  NB_OP_UNARY_END = 64,
  /* binary */
  NB_OP_ADD = 64, NB_OP_SUB = 65, NB_OP_MUL = 66, NB_OP_DIV = 67,
  NB_OP_LDIV = 68,      /* x \ y == y / x */
  NB_OP_POW = 69, NB_OP_HYPOT = 70, NB_OP_MAX = 71, NB_OP_MIN = 72,
  NB_OP_ATAN2 = 73,     /* atan(y_, x_) with operand a = y_, b = x_ */
  NB_OP_BETA = 74,      /* SpecialFunctions.beta(a, b) */
  /* differentiable in the SECOND operand only (test/runtests.jl:43-47): operand a is the integer order, its partial is
   * NaN as in DiffRules; a non-integer order gives NaN (csrc/special.cuh) */
  NB_OP_BESSELJ = 75, NB_OP_BESSELY = 76, NB_OP_BESSELI = 77, NB_OP_BESSELK = 78, NB_OP_POLYGAMMA = 79,
  /* comparisons: 1.0 / 0.0, both partials zero */
  NB_OP_GT = 80, NB_OP_LT = 81, NB_OP_GE = 82, NB_OP_LE = 83,
  /* selection: MASK(c, v) = c != 0 ? v : 0, MASKN(c, v) = c == 0 ? v : 0; d/dv = the indicator, d/dc = 0.
   * ifelse(c, a, b) = MASK(c, a) + MASKN(c, b): the branch not taken contributes exactly 0 to value and derivative
   * (no 0 * Inf), i.e. the derivative of the taken branch, which is what ForwardDiff's duals give */
  NB_OP_MASK = 84, NB_OP_MASKN = 85,
  NB_OP_BINARY_END = 86,
  /* SpecialFunctions.jl unary functions of test/runtests.jl:34-36 (second unary range) */
  NB_OP_SF_BEGIN = 96,
  NB_OP_GAMMA = 96, NB_OP_DIGAMMA = 97, NB_OP_TRIGAMMA = 98, NB_OP_BESSELJ0 = 99, NB_OP_BESSELJ1 = 100,
  NB_OP_BESSELY0 = 101, NB_OP_BESSELY1 = 102, NB_OP_ERFINV = 103, NB_OP_ERFCINV = 104, NB_OP_ERFCX = 105,
  NB_OP_AIRYAI = 106, NB_OP_AIRYAIPRIME = 107, NB_OP_AIRYBI = 108, NB_OP_AIRYBIPRIME = 109,
  NB_OP_DAWSON = 110, NB_OP_ERFI = 111, NB_OP_INVDIGAMMA = 112,   /* csrc/special.cuh */
  NB_OP_SF_END = 113
} nb_op;

typedef struct nb_stats {
  uint64_t kernel_launches;   /* kernels of this library launched on the ctx stream */
  uint64_t pool_bytes_in_use;
  uint64_t pool_bytes_reserved;
  uint64_t pool_high_water;
  uint64_t cuda_mallocs;      /* pool misses that went to the driver */
  uint64_t h2d_bytes, d2h_bytes;
} nb_stats;

/* ---- library / context ---------------------------------------------------------------------
 * No function of the reference corresponds to these: in Nabla the Julia runtime owns array storage, threads and
 * error propagation (exceptions: MethodError for a missing rule test/core.jl:159, ErrorException
 * src/checkpointing.jl:11,15).  Here a status code + message replaces the exception, a ctx replaces the runtime. */
int32_t nb_abi_version(void);
/* Message of the last failing call on this ctx (ctx may be NULL for create failures). */
const char* nb_last_error(nb_ctx* ctx);
/* Replaces Julia's GC-owned `Array` storage + implicit single thread: owns one CUDA stream and a
 * pooled, stream-ordered allocator (pool_bytes is pre-reserved; the pool grows on demand). */
int32_t nb_ctx_create(int32_t device, uint64_t pool_bytes, nb_ctx** out);
int32_t nb_ctx_destroy(nb_ctx* ctx);
/* One reverse sweep == one stream epoch (src/core.jl:160-166, 200).  begin/end are cheap markers;
 * end synchronises the stream so results may be read by the host. */
int32_t nb_sweep_begin(nb_ctx* ctx);
int32_t nb_sweep_end(nb_ctx* ctx);
int32_t nb_sync(nb_ctx* ctx);
int32_t nb_ctx_stream(nb_ctx* ctx, void** cuda_stream_out);
int32_t nb_ctx_stats(nb_ctx* ctx, nb_stats* out);
int32_t nb_device_count(int32_t* out);

/* ---- buffers (replace Array allocation: functional.jl:62,77,85) ---------------------------- */
int32_t nb_alloc(nb_ctx* ctx, uint64_t nbytes, void** out);
int32_t nb_retain(nb_ctx* ctx, void* buf);
int32_t nb_release(nb_ctx* ctx, void* buf);  /* safe from a GC finalizer thread */
int32_t nb_refcount(nb_ctx* ctx, void* buf, int32_t* out);
int32_t nb_h2d(nb_ctx* ctx, void* dst_dev, const void* src_host, uint64_t nbytes);
int32_t nb_d2h(nb_ctx* ctx, void* dst_host, const void* src_dev, uint64_t nbytes); /* syncs */
int32_t nb_d2h_async(nb_ctx* ctx, void* dst_host, const void* src_dev, uint64_t nbytes);
int32_t nb_d2d(nb_ctx* ctx, void* dst_dev, const void* src_dev, uint64_t nbytes);
int32_t nb_fill(nb_ctx* ctx, void* dst_dev, int32_t dtype, uint64_t count, double value);
/* The caller wrote [ptr, ptr + nbytes) with its own kernels or copies enqueued on nb_ctx_stream (every nb_* entry
 * point does this bookkeeping itself).  The library caches tensor-core operand splits of Float32 arrays it has
 * multiplied (see nb_math_mode); this drops the cached splits of every array overlapping the range.  No reference
 * counterpart: Julia arrays have no derived copies. */
int32_t nb_buffer_written(nb_ctx* ctx, const void* ptr, uint64_t nbytes);
int32_t nb_host_alloc(uint64_t nbytes, void** out); /* pinned host memory */
int32_t nb_host_free(void* p);

/* ---- scalar programs (replace "any Julia scalar function under ForwardDiff duals",
 *      src/core.jl:285-301).  Value slots: 0..nargs-1 are the operands, nargs+i is the result of
 *      instruction i; a negative operand index -(c+1) names consts[c].  The program's value is the
 *      last instruction's result (or operand 0 when n_instr == 0). */
int32_t nb_prog_create(int32_t nargs, int32_t n_instr, const int32_t* ops, const int32_t* a,
                       const int32_t* b, int32_t n_consts, const double* consts, nb_prog** out);
int32_t nb_prog_destroy(nb_prog* prog);
/* What a pullback launch of `prog` computes when it differentiates the operands in `wanted_mask` (bit k = operand k;
 * host only, no device work).  fmad evaluates the whole function on duals once per operand (core.jl:285-301); the
 * device propagates adjoints only along paths that reach a wanted operand and skips forward values no remaining
 * partial reads.  Bit i of *live_instr: instruction i takes part in the reverse pass; bit i of *fwd_skip: its value is
 * not computed; bit k of *arg_live: operand k reaches the result (otherwise its sensitivity is zero). */
int32_t nb_prog_pullback_plan(const nb_prog* prog, uint32_t wanted_mask, uint32_t* live_instr, uint32_t* fwd_skip,
                              uint32_t* arg_live);

/* ---- family (1): broadcast forward / pullback fused with unbroadcast ----------------------- */
/* out = f.(args...)   — primal of Branch(broadcast, (f, args...)): functional.jl:48-51, core.jl:89 */
int32_t nb_bcast_fwd(nb_ctx* ctx, const nb_prog* prog, int32_t nargs, const nb_tensor* args,
                     const nb_tensor* out);
/* For every k in want_mask:  outs[k] (+)= unbroadcast_to(shape(outs[k]), ybar .* df/dx_k(args...))
 * — replaces ∇(broadcast, Arg{N}, ...) -> broadcastsum/broadcastsum! (functional.jl:59-95), one
 * fused pass for all wanted operands instead of one pass per operand (core.jl:149-156); bit k of
 * accumulate_mask fuses the `add!!` of core.jl:203-205.  `ybar` may itself be broadcast (e.g. the
 * scalar seed of `sum`).  `y_saved` (nullable) is the Branch's forward value (core.jl:76): used
 * instead of re-evaluating f where df/dx is a function of y (tanh, exp, logistic, sqrt, ...). */
int32_t nb_bcast_pullback(nb_ctx* ctx, const nb_prog* prog, const nb_tensor* ybar,
                          const nb_tensor* y_saved, int32_t nargs, const nb_tensor* args,
                          uint32_t want_mask, const nb_tensor* outs, uint32_t accumulate_mask);

/* ---- family (2): reductions ----------------------------------------------------------------- */
/* out (+)= scale * sum over the dims where `out` has extent 1 of f.(args...)
 * — replaces sum / sum(f, .) / sum(abs2, .) / mean / mapreduce(f, +, .; dims) / dot forward
 * (functional/reducedim.jl:4-72, functional/reduce.jl:4-12, ChainRules sum/mean/dot) and
 * broadcastsum! itself (functional.jl:59-69).  prog == NULL means identity on args[0]. */
int32_t nb_bcast_reduce(nb_ctx* ctx, const nb_prog* prog, int32_t nargs, const nb_tensor* args,
                        double scale, const nb_tensor* out, int32_t accumulate);

/* ---- family (3): dense linalg (BLAS column-major semantics) -------------------------------- */
/* C = alpha*op(A)*op(B) + beta*C — replaces `*` and BLAS.gemm forward and the adjoints
 * Abar = Ybar*B' ('N','T'), Bbar = A'*Ybar ('T','N') of ChainRules rrule(*) reached through
 * src/sensitivities/chainrules.jl:179-191,286 and src/sensitivities/linalg/blas.jl:224-228;
 * beta == 1 is the in-place accumulate (InplaceableThunk mul!(X, Y, B', true, true)). */
int32_t nb_gemm(nb_ctx* ctx, char tA, char tB, int64_t m, int64_t n, int64_t k, double alpha,
                const void* A, int64_t lda, const void* B, int64_t ldb, double beta, void* C,
                int64_t ldc, int32_t dtype, int32_t math_mode);
/* The product with the Branches around it fused into the GEMM epilogue (tape-level
 * fusion; the reference tapes `f.(W*X .+ b)` as three Branches with three full-size arrays,
 * functional.jl:48-51, and walks them one by one in reverse, core.jl:140-158):
 *   kind 1 (forward)   C = f.(alpha*op(A)*op(B) .+ bias)          bias: length m, nullable
 *   kind 2 (pullback)  C = (alpha*op(A)*op(B)) .* f'(y)            y: the saved forward value f(...) (m x n, ldy);
 *                      f must be a catalogue function whose derivative is a function of y alone
 *                      (identity, tanh, logistic, exp, sqrt, inv, ...: DiffRules forms used by fmad, core.jl:285-292)
 *   rowsum (nullable)  rowsum[i] (+)= sum_j C[i,j]                 the bias sensitivity: unbroadcast of the
 *                      `.+` pullback (functional.jl:61-63), deterministic two-phase sum
 *   emit_split         bit 0: also keep the tensor-core operand split of C (in this product's own 16-bit format) for
 *                      the products that consume it next; bit 1: a forward value emitted as binary16 is ALSO emitted
 *                      as bfloat16, for the adjoint products that run in NB_MATH_BF16X3
 * Only the Float32 tensor-core path implements it: NB_ERR_UNSUPPORTED (nothing launched) otherwise, and the
 * caller runs the unfused sequence nb_gemm + nb_bcast_fwd / nb_bcast_pullback. */
typedef struct nb_gemm_epilogue {
  int32_t kind;        /* 0 none, 1 forward, 2 pullback */
  int32_t op;          /* unary nb_op code */
  const void* bias;
  const void* y;
  int64_t ldy;
  void* rowsum;
  int32_t rowsum_acc;
  int32_t emit_split;
} nb_gemm_epilogue;
/* Which inputs the partial derivative of catalogue op `op` w.r.t. operand `wrt` (0 = a, 1 = b) reads:
 * bit 0 operand a, bit 1 operand b, bit 2 the saved forward value y (host logic uses it to decide what a
 * pullback launch must be handed; no GPU needed). */
int32_t nb_op_partial_needs(int32_t op, int32_t wrt);
int32_t nb_gemm_fused(nb_ctx* ctx, char tA, char tB, int64_t m, int64_t n, int64_t k, double alpha,
                      const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                      int32_t dtype, int32_t math_mode, const nb_gemm_epilogue* epi);
/* y = alpha*op(A)*x + beta*y   (A is m x n) — matrix*vector and BLAS.gemv (blas.jl tests :49-61) */
int32_t nb_gemv(nb_ctx* ctx, char tA, int64_t m, int64_t n, double alpha, const void* A,
                int64_t lda, const void* x, double beta, void* y, int32_t dtype);
/* A = alpha*x*y' + beta*A  (A is m x n) — the outer-product adjoint of matrix*vector */
int32_t nb_ger(nb_ctx* ctx, int64_t m, int64_t n, double alpha, const void* x, const void* y,
               double beta, void* A, int64_t lda, int32_t dtype);
/* out(n x m) = in(m x n)' — materialised adjoint/transpose (ChainRules rrule(adjoint)) */
int32_t nb_transpose(nb_ctx* ctx, int64_t m, int64_t n, const void* in, int64_t ld_in, void* out,
                     int64_t ld_out, int32_t dtype);

/* ---- symmetric / triangular BLAS family (blas.jl:60-328) -------------------
 * nb_tri materialises the structured operand or finishes a pullback:
 *   op 0/1  dst = tril(src) / triu(src); diag 0 keeps, 1 zeroes ('U' pullbacks), 2 sets the diagonal to one
 *   op 2/3  dst = the symmetric matrix stored in the lower / upper triangle of src (symm, symv operands)
 *   op 4/5  dst = tril / triu of (src + src' - Diagonal(src))            (symm pullback, blas.jl:78-80)
 * nb_trsm overwrites B with alpha*op(A)^-1*B (side 'L') or alpha*B*op(A)^-1 (side 'R'), A triangular:
 * BLAS.trsm; with n == 1 it is BLAS.trsv. */
int32_t nb_tri(nb_ctx* ctx, int32_t op, int32_t diag, int64_t n, const void* src, int64_t ld_src, void* dst,
               int64_t ld_dst, int32_t dtype);
int32_t nb_trsm(nb_ctx* ctx, char side, char ul, char ta, char diag, int64_t m, int64_t n, double alpha,
                const void* A, int64_t lda, void* B, int64_t ldb, int32_t dtype);

/* ---- array structure --------------------------------------------------- */
/* dst[dst_off + r + c*ld_dst] (+)= src[src_off + r + c*ld_src] for r < rows, c < cols (offsets in elements):
 * hcat / vcat forward and their pullbacks `copy(selectdim(ybar, ...))` (chainrules_ignored.jl:12-40),
 * range getindex forward and its pullback (indexing.jl:3-11; ChainRules rrule(getindex)). */
int32_t nb_copy2d(nb_ctx* ctx, int64_t rows, int64_t cols, const void* src, int64_t src_off, int64_t ld_src,
                  void* dst, int64_t dst_off, int64_t ld_dst, int32_t dtype, int32_t accumulate);
/* dst[i] = src[idx[i]] and dst[idx[i]] += src[i] (0-based device index vector; duplicates accumulate,
 * test/sensitivities/indexing.jl "Overlapping indices (#139)") — getindex with an index vector. */
int32_t nb_gather(nb_ctx* ctx, int64_t n, const void* src, const int64_t* idx_dev, void* dst, int32_t dtype);
int32_t nb_scatter_add(nb_ctx* ctx, int64_t n, const void* src, const int64_t* idx_dev, void* dst, int32_t dtype);

/* ---- batch-sharded gradients (new capability) ----------------------------- */
int32_t nb_comm_unique_id(uint8_t id_out[128]);
int32_t nb_comm_init(nb_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t id[128]);
/* In-place sum of each buffer over all ranks (one grouped NCCL launch on the ctx stream). */
int32_t nb_allreduce_sum(nb_ctx* ctx, int32_t n_bufs, void* const* bufs, const int64_t* counts,
                         int32_t dtype);
/* Overlapped exchange: the reduction runs on the communicator's own stream, ordered after everything
 * enqueued so far on the ctx stream, so the sweep keeps launching backward kernels while parameter
 * sensitivities that are already final travel over NVLink.  nb_comm_join makes the ctx stream wait
 * for every reduction started this way.  The buffers must stay allocated until the join. */
int32_t nb_allreduce_sum_async(nb_ctx* ctx, int32_t n_bufs, void* const* bufs, const int64_t* counts,
                               int32_t dtype);
int32_t nb_comm_join(nb_ctx* ctx);
int32_t nb_comm_destroy(nb_ctx* ctx);

/* ---- CUDA-graph capture of a whole forward + reverse sweep ---------------------------------
 * The reference re-runs the host loop of propagate (src/core.jl:160-166) for every gradient evaluation; a tape with
 * a fixed signature (same Branches, same shapes: examples/mlp.jl, examples/vae.jl inside an optimiser loop) is
 * captured once between begin/end and replayed with one launch. */
int32_t nb_graph_begin(nb_ctx* ctx);
int32_t nb_graph_end(nb_ctx* ctx, nb_graph** out);
int32_t nb_graph_launch(nb_ctx* ctx, nb_graph* g);
int32_t nb_graph_destroy(nb_graph* g);

/* ---- device timing on the ctx stream --------------------------------------------------------
 * (measurement only: the reference times with @elapsed / BenchmarkTools on the host; no rule depends on these) */
int32_t nb_event_create(void** out);
int32_t nb_event_record(nb_ctx* ctx, void* ev);
int32_t nb_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms_out); /* syncs ev_stop */
int32_t nb_event_destroy(void* ev);

#ifdef __cplusplus
}
#endif
#endif /* NABLA_CUDA_H */
