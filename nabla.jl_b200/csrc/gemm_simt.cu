// Family (3), CUDA-core part: the nb_gemm dispatcher, a tiled SIMT GEMM that accepts every shape,
// stride and transpose (the correctness floor and the path for shapes TMA cannot describe, e.g. the
// 10-row output layer whose leading dimension is not a multiple of 16 bytes), and the transpose
// kernel behind `adjoint`.
//
// Reference: ChainRules rrule(*) / BLAS.gemm adjoints reached through
// src/sensitivities/chainrules.jl:179-191,286 and src/sensitivities/linalg/blas.jl:224-228.
#include "ops.cuh"

// implemented in gemm_f64.cu / gemm_tf32.cu; return NB_ERR_UNSUPPORTED when the shape does not fit
int32_t nb_gemm_dmma(nb_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, double alpha,
                     const double* A, int64_t lda, const double* B, int64_t ldb, double beta,
                     double* C, int64_t ldc, const int32_t* run_if);
int32_t nb_gemm_ozaki(nb_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                      int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc);
int32_t nb_gemm_tcgen05(nb_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, float alpha,
                        const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
                        float* C, int64_t ldc, int32_t math_mode, const nb_gemm_epilogue* epi);

int32_t nb_gemm_skinny(nb_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, double alpha, const void* A,
                       int64_t lda, const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int32_t dtype,
                       const nb_gemm_epilogue* epi, int32_t math_mode);  // gemm_skinny.cu

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

template <typename T>
__global__ void __launch_bounds__(256) k_gemm_simt(bool tA, bool tB, int64_t m, int64_t n, int64_t k,
                                                   T alpha, const T* __restrict__ A, int64_t lda,
                                                   const T* __restrict__ B, int64_t ldb, T beta,
                                                   T* __restrict__ C, int64_t ldc) {
  __shared__ T As[BK][BM + 4];
  __shared__ T Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // tx along rows of C (contiguous), ty along columns
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

  for (int64_t k0 = 0; k0 < k; k0 += BK) {
    // A tile: op(A)[m0+i, k0+kk]
#pragma unroll
    for (int e = tid; e < BM * BK; e += 256) {
      int i, kk;
      if (!tA) { i = e % BM; kk = e / BM; } else { kk = e % BK; i = e / BK; }
      const int64_t gi = m0 + i, gk = k0 + kk;
      T v = T(0);
      if (gi < m && gk < k) v = tA ? A[gk + gi * lda] : A[gi + gk * lda];
      As[kk][i] = v;
    }
    // B tile: op(B)[k0+kk, n0+j]
#pragma unroll
    for (int e = tid; e < BN * BK; e += 256) {
      int j, kk;
      if (!tB) { kk = e % BK; j = e / BK; } else { j = e % BN; kk = e / BN; }
      const int64_t gj = n0 + j, gk = k0 + kk;
      T v = T(0);
      if (gj < n && gk < k) v = tB ? B[gj + gk * ldb] : B[gk + gj * ldb];
      Bs[kk][j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][tx * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][ty * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int64_t gj = n0 + ty * TN + j;
    if (gj >= n) continue;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int64_t gi = m0 + tx * TM + i;
      if (gi >= m) continue;
      T* c = C + gi + gj * ldc;
      const T r = alpha * acc[i][j];
      *c = beta == T(0) ? r : r + beta * *c;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_transpose(int64_t m, int64_t n, const T* __restrict__ in,
                                                   int64_t ldi, T* __restrict__ out, int64_t ldo) {
  __shared__ T tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;  // 32 x 8
  for (int jj = ty; jj < 32; jj += 8) {
    const int64_t i = i0 + tx, j = j0 + jj;
    if (i < m && j < n) tile[jj][tx] = in[i + j * ldi];
  }
  __syncthreads();
  for (int ii = ty; ii < 32; ii += 8) {
    const int64_t j = j0 + tx, i = i0 + ii;
    if (i < m && j < n) out[j + i * ldo] = tile[tx][ii];
  }
}

template <typename T>
__global__ void k_scale_mat(int64_t m, int64_t n, T beta, T* C, int64_t ldc) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * n) return;
  T* c = C + idx % m + (idx / m) * ldc;
  *c = beta == T(0) ? T(0) : beta * *c;
}

template <typename T>
int32_t launch_simt(nb_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, double alpha,
                    const void* A, int64_t lda, const void* B, int64_t ldb, double beta, void* C,
                    int64_t ldc) {
  dim3 grid((unsigned)((m + BM - 1) / BM), (unsigned)((n + BN - 1) / BN));
  k_gemm_simt<T><<<grid, 256, 0, ctx->stream>>>(tA, tB, m, n, k, (T)alpha, (const T*)A, lda,
                                                (const T*)B, ldb, (T)beta, (T*)C, ldc);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

inline bool is_t(char c) { return c == 'T' || c == 't' || c == 'C' || c == 'c'; }
inline bool is_n(char c) { return c == 'N' || c == 'n'; }

}  // namespace

extern "C" {

int32_t nb_gemm(nb_ctx* ctx, char tAc, char tBc, int64_t m, int64_t n, int64_t k, double alpha,
                const void* A, int64_t lda, const void* B, int64_t ldb, double beta, void* C,
                int64_t ldc, int32_t dtype, int32_t math_mode) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  NB_REQUIRE(ctx, (is_t(tAc) || is_n(tAc)) && (is_t(tBc) || is_n(tBc)), NB_ERR_INVALID,
             "nb_gemm: transpose flags must be 'N', 'T' or 'C'");
  NB_REQUIRE(ctx, m >= 0 && n >= 0 && k >= 0, NB_ERR_INVALID, "nb_gemm: negative dimension");
  NB_REQUIRE(ctx, dtype == NB_F32 || dtype == NB_F64, NB_ERR_INVALID, "nb_gemm: bad dtype");
  const bool tA = is_t(tAc), tB = is_t(tBc);
  NB_REQUIRE(ctx, lda >= (tA ? k : m) && ldb >= (tB ? n : k) && ldc >= m, NB_ERR_SHAPE,
             "nb_gemm: leading dimension too small (DimensionMismatch)");
  if (m == 0 || n == 0) return NB_OK;
  NB_REQUIRE(ctx, A && B && C, NB_ERR_INVALID, "nb_gemm: NULL matrix");
  nb_note_write(ctx, C);
  if (k == 0 || alpha == 0.0) {  // C = beta*C
    const int64_t tot = m * n;
    if (dtype == NB_F32) k_scale_mat<float><<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(m, n, (float)beta, (float*)C, ldc);
    else k_scale_mat<double><<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(m, n, beta, (double*)C, ldc);
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
  }
  {  // one dimension <= 16: HBM-bound streaming kernels, plain FMA in the element type (gemm_skinny.cu)
    int32_t rc = nb_gemm_skinny(ctx, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, dtype, nullptr, math_mode);
    if (rc != NB_ERR_UNSUPPORTED) return rc;
  }
  if (dtype == NB_F64) {
    // large products: exact integer slice products on the tensor cores (ozaki.cu); NB_F64_OZAKI=0 or
    // math_mode == NB_MATH_FP32 (the "no tensor-core tricks" cross-check mode) keeps everything on DMMA
    static const bool ozaki_on = !(getenv("NB_F64_OZAKI") && atoi(getenv("NB_F64_OZAKI")) == 0);
    if (ozaki_on && math_mode != NB_MATH_FP32) {
      int32_t rc = nb_gemm_ozaki(ctx, tA, tB, m, n, k, alpha, (const double*)A, lda, (const double*)B, ldb, beta,
                                 (double*)C, ldc);
      if (rc != NB_ERR_UNSUPPORTED) return rc;
    }
    int32_t rc = nb_gemm_dmma(ctx, tA, tB, m, n, k, alpha, (const double*)A, lda, (const double*)B,
                              ldb, beta, (double*)C, ldc, nullptr);
    if (rc != NB_ERR_UNSUPPORTED) return rc;
    return launch_simt<double>(ctx, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  }
  if (math_mode != NB_MATH_FP32) {
    int32_t rc = nb_gemm_tcgen05(ctx, tA, tB, m, n, k, (float)alpha, (const float*)A, lda,
                                 (const float*)B, ldb, (float)beta, (float*)C, ldc, math_mode, nullptr);
    if (rc != NB_ERR_UNSUPPORTED) return rc;
  }
  return launch_simt<float>(ctx, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}

int32_t nb_gemm_fused(nb_ctx* ctx, char tAc, char tBc, int64_t m, int64_t n, int64_t k, double alpha,
                      const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                      int32_t dtype, int32_t math_mode, const nb_gemm_epilogue* epi) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  NB_REQUIRE(ctx, (is_t(tAc) || is_n(tAc)) && (is_t(tBc) || is_n(tBc)), NB_ERR_INVALID,
             "nb_gemm_fused: transpose flags must be 'N', 'T' or 'C'");
  NB_REQUIRE(ctx, epi && (epi->kind == 1 || epi->kind == 2), NB_ERR_INVALID, "nb_gemm_fused: bad epilogue kind");
  NB_REQUIRE(ctx, nb_op_is_unary(epi->op), NB_ERR_INVALID, "nb_gemm_fused: op must be unary");
  NB_REQUIRE(ctx, epi->kind != 2 || (epi->y && epi->ldy >= m), NB_ERR_INVALID, "nb_gemm_fused: pullback needs y");
  NB_REQUIRE(ctx, epi->kind != 2 || (nb_partial_needs(epi->op, 0) & 3) == 0, NB_ERR_INVALID,
             "nb_gemm_fused: f' must be a function of y alone");
  const bool tA = is_t(tAc), tB = is_t(tBc);
  NB_REQUIRE(ctx, m > 0 && n > 0 && k > 0, NB_ERR_UNSUPPORTED, "nb_gemm_fused: empty product");
  NB_REQUIRE(ctx, lda >= (tA ? k : m) && ldb >= (tB ? n : k) && ldc >= m, NB_ERR_SHAPE,
             "nb_gemm_fused: leading dimension too small (DimensionMismatch)");
  NB_REQUIRE(ctx, A && B && C, NB_ERR_INVALID, "nb_gemm_fused: NULL matrix");
  // Float64: not fused.  A fused DMMA epilogue was built and measured (same box, graph replay of the Float64
  // examples): MLP 0.93 ms vs 0.81 ms, VAE 1.60 vs 1.42 — the DMMA kernel is not persistent, so the epilogue work
  // becomes a serial tail of every CTA instead of overlapping the next tile, and it slowed the plain products too.
  if (alpha == 0.0) return NB_ERR_UNSUPPORTED;
  if (math_mode != NB_MATH_FP32 || dtype == NB_F64) {
    nb_note_write(ctx, C);
    if (epi->rowsum) nb_note_write(ctx, epi->rowsum);
    int32_t rc = nb_gemm_skinny(ctx, tA, tB, m, n, k, alpha, A, lda, B, ldb, 0.0, C, ldc, dtype, epi, math_mode);
    if (rc != NB_ERR_UNSUPPORTED) return rc;
  }
  if (dtype != NB_F32 || math_mode == NB_MATH_FP32) return NB_ERR_UNSUPPORTED;
  nb_note_write(ctx, C);
  if (epi->rowsum) nb_note_write(ctx, epi->rowsum);
  return nb_gemm_tcgen05(ctx, tA, tB, m, n, k, (float)alpha, (const float*)A, lda, (const float*)B, ldb, 0.f,
                         (float*)C, ldc, math_mode, epi);
}

int32_t nb_op_partial_needs(int32_t op, int32_t wrt) { return nb_partial_needs(op, wrt); }

int32_t nb_transpose(nb_ctx* ctx, int64_t m, int64_t n, const void* in, int64_t ld_in, void* out,
                     int64_t ld_out, int32_t dtype) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && (in || !(m * n)) && (out || !(m * n)), NB_ERR_INVALID, "NULL argument");
  NB_REQUIRE(ctx, ld_in >= m && ld_out >= n, NB_ERR_SHAPE, "nb_transpose: leading dimension too small");
  if (m == 0 || n == 0) return NB_OK;
  nb_note_write(ctx, out);
  dim3 grid((unsigned)((m + 31) / 32), (unsigned)((n + 31) / 32));
  if (dtype == NB_F32) k_transpose<float><<<grid, 256, 0, ctx->stream>>>(m, n, (const float*)in, ld_in, (float*)out, ld_out);
  else if (dtype == NB_F64) k_transpose<double><<<grid, 256, 0, ctx->stream>>>(m, n, (const double*)in, ld_in, (double*)out, ld_out);
  else return nb_fail(ctx, NB_ERR_INVALID, "bad dtype");
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

}  // extern "C"
