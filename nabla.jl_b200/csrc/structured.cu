// Structured-matrix helpers behind the symm / symv / trmm / trmv / trsm / trsv family
// (src/sensitivities/linalg/blas.jl:60-328).  The products themselves run on
// the dense GEMM / GEMV kernels after the symmetric or triangular operand has been materialised; the
// pullbacks need `tril!/triu!`, a zeroed or unit diagonal, and `tmp + tmp' - Diagonal(tmp)`.
// Triangular solves (`trsm`, `trsv`) are substitution kernels, one CTA per right-hand side: they make
// the rules available and exact, they are not tuned (no BASELINE configuration uses them).
#include "common.cuh"

namespace {

enum { TRI_L = 0, TRI_U = 1, SYM_L = 2, SYM_U = 3, FOLD_L = 4, FOLD_U = 5 };

template <typename T>
__global__ void __launch_bounds__(256) k_tri(int op, int diag, int64_t n, const T* __restrict__ src, int64_t lds,
                                             T* __restrict__ dst, int64_t ldd) {
  const int64_t total = n * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t i = e % n, j = e / n;
    const bool lower = i > j, upper = i < j;
    T v;
    if (op == TRI_L || op == TRI_U) {
      const bool keep = op == TRI_L ? !upper : !lower;
      v = keep ? src[i + j * lds] : T(0);
      if (i == j) v = diag == 1 ? T(0) : diag == 2 ? T(1) : v;
    } else if (op == SYM_L || op == SYM_U) {
      const bool from_here = op == SYM_L ? !upper : !lower;
      v = from_here ? src[i + j * lds] : src[j + i * lds];
    } else {
      const bool keep = op == FOLD_L ? !upper : !lower;
      v = !keep ? T(0) : (i == j ? src[i + j * lds] : src[i + j * lds] + src[j + i * lds]);
    }
    dst[i + j * ldd] = v;
  }
}

// One CTA solves op(T) x = alpha * b for one right-hand side held in shared memory.
//   el(i, j) = element (i, j) of op(T);  forward substitution when op(T) is lower triangular.
// trans == 0: column sweeps (T[i, j] contiguous in i);  trans == 1: row dot products (contiguous in j).
template <typename T>
__global__ void __launch_bounds__(256) k_trsolve(int lower_eff, int trans, int unit, int64_t n, T alpha,
                                                 const T* __restrict__ A, int64_t lda, T* __restrict__ B,
                                                 int64_t rhs_stride, int64_t elem_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* x = reinterpret_cast<T*>(smem_raw);
  __shared__ T red[32];
  T* b = B + (int64_t)blockIdx.x * rhs_stride;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) x[i] = alpha * b[i * elem_stride];
  __syncthreads();
  for (int64_t s = 0; s < n; ++s) {
    const int64_t j = lower_eff ? s : n - 1 - s;  // unknown finalised in this step
    if (!trans) {
      if (threadIdx.x == 0 && !unit) x[j] /= A[j + j * lda];
      __syncthreads();
      const T xj = x[j];
      if (lower_eff) {
        for (int64_t i = j + 1 + threadIdx.x; i < n; i += blockDim.x) x[i] -= A[i + j * lda] * xj;
      } else {
        for (int64_t i = threadIdx.x; i < j; i += blockDim.x) x[i] -= A[i + j * lda] * xj;
      }
      __syncthreads();
    } else {
      // op(T)[j, k] = A[k + j*lda]: dot product over the already known unknowns
      T acc = T(0);
      if (lower_eff) {
        for (int64_t k = threadIdx.x; k < j; k += blockDim.x) acc += A[k + j * lda] * x[k];
      } else {
        for (int64_t k = j + 1 + threadIdx.x; k < n; k += blockDim.x) acc += A[k + j * lda] * x[k];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
      __syncthreads();
      if (threadIdx.x == 0) {
        T t = T(0);
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        const T v = x[j] - t;
        x[j] = unit ? v : v / A[j + j * lda];
      }
      __syncthreads();
    }
  }
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) b[i * elem_stride] = x[i];
}

inline bool is_t(char c) { return c == 'T' || c == 't' || c == 'C' || c == 'c'; }

}  // namespace

extern "C" {

int32_t nb_tri(nb_ctx* ctx, int32_t op, int32_t diag, int64_t n, const void* src, int64_t ld_src, void* dst,
               int64_t ld_dst, int32_t dtype) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  NB_REQUIRE(ctx, op >= 0 && op <= 5 && diag >= 0 && diag <= 2, NB_ERR_INVALID, "nb_tri: bad op/diag");
  NB_REQUIRE(ctx, dtype == NB_F32 || dtype == NB_F64, NB_ERR_INVALID, "bad dtype");
  NB_REQUIRE(ctx, n >= 0 && ld_src >= n && ld_dst >= n, NB_ERR_SHAPE, "nb_tri: leading dimension too small");
  if (n == 0) return NB_OK;
  NB_REQUIRE(ctx, src && dst && src != dst, NB_ERR_INVALID, "nb_tri: NULL or aliased buffers");
  nb_note_write(ctx, dst);
  int64_t blocks = (n * n + 255) / 256;
  if (blocks > ctx->num_sms * 8) blocks = ctx->num_sms * 8;
  if (dtype == NB_F32) k_tri<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(op, diag, n, (const float*)src, ld_src, (float*)dst, ld_dst);
  else k_tri<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(op, diag, n, (const double*)src, ld_src, (double*)dst, ld_dst);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int32_t nb_trsm(nb_ctx* ctx, char side, char ul, char ta, char diag, int64_t m, int64_t n, double alpha,
                const void* A, int64_t lda, void* B, int64_t ldb, int32_t dtype) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  NB_REQUIRE(ctx, dtype == NB_F32 || dtype == NB_F64, NB_ERR_INVALID, "bad dtype");
  const bool left = side == 'L' || side == 'l', right = side == 'R' || side == 'r';
  const bool lower = ul == 'L' || ul == 'l', upper = ul == 'U' || ul == 'u';
  NB_REQUIRE(ctx, (left || right) && (lower || upper), NB_ERR_INVALID, "nb_trsm: bad side/uplo flag");
  NB_REQUIRE(ctx, m >= 0 && n >= 0 && ldb >= m, NB_ERR_SHAPE, "nb_trsm: bad dimensions");
  const int64_t na = left ? m : n;  // order of A
  NB_REQUIRE(ctx, lda >= na, NB_ERR_SHAPE, "nb_trsm: lda too small");
  if (m == 0 || n == 0) return NB_OK;
  NB_REQUIRE(ctx, A && B, NB_ERR_INVALID, "nb_trsm: NULL matrix");
  const size_t es = nb_dtype_size(dtype);
  NB_REQUIRE(ctx, (size_t)na * es <= 200 * 1024, NB_ERR_UNSUPPORTED, "nb_trsm: triangular order too large");
  nb_note_write(ctx, B);
  // side L: op(A) X = alpha B, one RHS per column of B.
  // side R: X op(A) = alpha B  <=>  op(A)' x_r = alpha b_r for every row r of B.
  bool trans = is_t(ta);
  if (right) trans = !trans;
  const int lower_eff = (lower != trans) ? 1 : 0;  // op(T) (after the side-R transposition) is lower
  const int unit = (diag == 'U' || diag == 'u') ? 1 : 0;
  const int64_t nrhs = left ? n : m;
  const int64_t rhs_stride = left ? ldb : 1, elem_stride = left ? 1 : ldb;
  const size_t smem = (size_t)na * es;
  if (dtype == NB_F32) {
    if (smem > 48 * 1024) NB_CUDA_OK(ctx, cudaFuncSetAttribute(k_trsolve<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_trsolve<float><<<(unsigned)nrhs, 256, smem, ctx->stream>>>(lower_eff, trans, unit, na, (float)alpha, (const float*)A, lda, (float*)B, rhs_stride, elem_stride);
  } else {
    if (smem > 48 * 1024) NB_CUDA_OK(ctx, cudaFuncSetAttribute(k_trsolve<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_trsolve<double><<<(unsigned)nrhs, 256, smem, ctx->stream>>>(lower_eff, trans, unit, na, alpha, (const double*)A, lda, (double*)B, rhs_stride, elem_stride);
  }
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

}  // extern "C"
