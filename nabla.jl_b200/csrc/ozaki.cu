// Float64 products on the 5th-generation tensor cores by an Ozaki split.
//
// Every reference example is Float64 (rtol 1e-10), tcgen05 has no f64 kind, and the DMMA path (gemm_f64.cu) tops
// out at ~28 TFLOP/s.  Large products C = alpha*op(A)*op(B) + beta*C are therefore computed from EXACT integer
// products (replaces OpenBLAS dgemm behind ChainRules' rrule(*), chainrules.jl:179-191, and BLAS.gemm adjoints,
// blas.jl:224-228):
//   1. every row i of op(A) and column j of op(B) is scaled by a power of two 2^-e so that |x| < 1/2
//      (k_absmax_* -> row_scale / col_scale = 2^e);
//   2. the scaled values are cut into S = 7 signed 7-bit slices  x = sum_p q_p 2^(-7(p+1)),  |q_p| <= 64
//      (k_slice: q = rint(128 x), x <- 128 x - q, exact in Float64); the slices of one operand are stored K-major
//      and concatenated along K — A in order 0..S-1, B in REVERSE order — each padded to a multiple of 128;
//   3. level l = p + q of  sum_{p+q<S} A_p B_q 2^(-7(p+q+2))  is then ONE kind::i8 GEMM over K' = (l+1)*Kp
//      (A slices 0..l against B slices l..0, which are contiguous in that layout), accumulated exactly in int32 TMEM
//      (K' * 2^12 < 2^31), and its epilogue adds  alpha * 2^(-7(l+2)) * row_scale * col_scale * acc  into the Float64 C.
// Dropped: the levels p + q >= S (~2^-51 of max|row| * max|col| per product term) and the slice residual (2^-50):
// measured norm-wise error 4e-14 per product (DMMA: 2e-15), and a five-layer MLP gradient built from a dozen chained
// 4096^3 split products agrees with the CPU oracle to 9e-13 (DMMA: 8e-14) — a hundred times inside rtol 1e-10; the
// elementwise error is relative to max|row| * max|col|, like any scaled fixed-point scheme.  A row / column holding
// a NaN or Inf makes its row / column of C NaN.
// RANGE GUARD.  dgemm's error is relative to sum_k |a_ik||b_kj|; the split's is relative to max|a_i.| * max|b_.j|.  The
// two differ when a line is dominated by an outlier: the pass that finds the line maxima also sums |x|, and
//   peak(line) = max|x| * (K - 1) / (sum|x| - max|x|)          (how far the largest entry stands above the others)
// is reduced to its maximum over the lines of op(A) and of op(B).  A line with peak p keeps 49 - log2(p) bits of its
// ordinary entries.  When peak_A * peak_B exceeds NB_OZAKI_PEAK_LIMIT (default 2^24: the bulk of a line would keep fewer
// than 25 bits — a mis-scaled feature, a sentinel value) the product is computed by DMMA instead.  Measured peaks:
// Gaussian data ~8 per operand; the sensitivities of the saturated wide MLP 1e4 - 1e6 per line (a few samples dominate a
// unit's row) — there the dominant entries also dominate the RESULT, and the whole Float64 gradient agrees with the
// oracle to 2e-12 on the split; NB_OZAKI_PEAK_LIMIT=4096 is the strict setting (error <= ~5e-12 of dgemm's scale in
// every case; it sends those weight-sensitivity products to DMMA, 1.2 instead of 2.1 evaluations/s on that workload),
// 0 disables the guard.  The decision is taken ON THE DEVICE (no host
// synchronisation, valid inside a CUDA-graph capture): the int8 launches start with their tile counter past the last
// tile when the flag is set (k_counter_init, gemm_tc.cu), and a DMMA launch that exits at once unless it is set follows.
// 28 int8 passes at twice the bf16 rate = 14 bf16 passes per product.  NB_OZAKI_SLICES=6 is the looser, faster
// setting (21 passes, 26 in the one-launch form; 20-25 % less time): 5e-12 per product, which the same chain grows
// to 8e-11 — too close to the 1e-10 bar to be the default (tools/check_f64_split_depth.py).
#include "common.cuh"

int32_t nb_gemm_i8(nb_ctx* ctx, int64_t m, int64_t n, int64_t kp, const int8_t* A8, int64_t lda8, const int8_t* B8,
                   int64_t ldb8, double scale, const double* row_scale, const double* col_scale, double beta,
                   double* C, int64_t ldc, const int32_t* skip_if);
int32_t nb_gemm_dmma(nb_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                     int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
                     const int32_t* run_if);

int32_t nb_gemm_i8all(nb_ctx* ctx, int64_t m, int64_t n, int64_t kp, int32_t S, int32_t extra_level, const int8_t* A8,
                      int64_t ld8, const int8_t* B8, double alpha, const double* row_scale, const double* col_scale,
                      double beta, double* C, int64_t ldc, const int32_t* skip_if);

namespace {

constexpr int MAX_S = 7;  // slices per operand (<= 8 levels fit nb_gemm_i8all's table)

// scale[i] = 2^e with |x| * 2^-e < 1/2 for every element of line i of op(X) (MN index i, reduced over K)
//   contiguous-K form  (element (i, kk) at X[kk + i*ld]): one warp per line
// non-negative doubles order like their bit patterns: max through an integer atomic (order-independent, deterministic)
__device__ __forceinline__ void atomic_max_nonneg(double* dst, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)__double_as_longlong(v));
}
// peak = max|x| * (K - 1) / (sum|x| - max|x|); 0 for an all-zero line and for a line with a single non-zero entry
// (both exact in the split: zeros cost nothing), NaN-free for finite input
__device__ __forceinline__ double line_peak(double mx, double sum, int64_t k) {
  if (!(mx > 0.0) || !isfinite(mx)) return 0.0;
  const double rest = sum - mx;
  return rest > 0.0 ? mx * (double)(k - 1) / rest : 0.0;
}

__global__ void __launch_bounds__(256) k_absmax_kcontig(const double* __restrict__ X, int64_t ld, int64_t mn, int64_t k,
                                                        double* __restrict__ scale, double* __restrict__ peak) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= mn) return;
  double mx = 0.0, sum = 0.0;
  for (int64_t kk = lane; kk < k; kk += 32) {
    const double v = fabs(X[kk + i * ld]);
    mx = (v == v) ? fmax(mx, v) : INFINITY;  // fmax drops NaNs: poison instead
    sum += v;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    sum += __shfl_xor_sync(0xffffffffu, sum, s);
  }
  // (a line with a NaN or Inf gets a NaN scale: its row / column of C becomes NaN instead of finite garbage)
  if (lane == 0) {
    scale[i] = mx == 0.0 ? 1.0 : isfinite(mx) ? ldexp(1.0, ilogb(mx) + 2) : nan("");
    atomic_max_nonneg(peak, line_peak(mx, sum, k));
  }
}
//   contiguous-MN form (element (i, kk) at X[i + kk*ld]): CTAs of 32 lines x 8 k-lanes over a k range; the partial
//   maxima meet in scale[] through atomicMax on the bit pattern, the partial sums of |x| go to sums[blockIdx.y][i];
//   k_absmax_finish adds them in order and turns the maximum into the power of two
__global__ void __launch_bounds__(256) k_absmax_mncontig(const double* __restrict__ X, int64_t ld, int64_t mn, int64_t k,
                                                         int64_t k_per_cta, double* __restrict__ scale,
                                                         double* __restrict__ sums) {
  __shared__ double red[8][33], reds[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + tx;
  const int64_t kb = (int64_t)blockIdx.y * k_per_cta, ke = min(k, kb + k_per_cta);
  double mx = 0.0, sum = 0.0;
  if (i < mn)
    for (int64_t kk = kb + ty; kk < ke; kk += 8) {
      const double v = fabs(X[i + kk * ld]);
      mx = (v == v) ? fmax(mx, v) : INFINITY;
      sum += v;
    }
  red[ty][tx] = mx;
  reds[ty][tx] = sum;
  __syncthreads();
  if (ty == 0 && i < mn) {
#pragma unroll
    for (int y = 1; y < 8; ++y) {
      mx = fmax(mx, red[y][tx]);
      sum += reds[y][tx];
    }
    if (!(mx == mx)) mx = INFINITY;  // NaN -> poison
    atomic_max_nonneg(scale + i, mx);
    sums[(int64_t)blockIdx.y * mn + i] = sum;
  }
}
__global__ void __launch_bounds__(256) k_absmax_finish(int64_t mn, int64_t k, double* __restrict__ scale,
                                                       const double* __restrict__ sums, int nsums,
                                                       double* __restrict__ peak) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mn) return;
  const double mx = scale[i];
  double sum = 0.0;
  for (int s = 0; s < nsums; ++s) sum += sums[(int64_t)s * mn + i];
  scale[i] = mx == 0.0 ? 1.0 : isfinite(mx) ? ldexp(1.0, ilogb(mx) + 2) : nan("");
  atomic_max_nonneg(peak, line_peak(mx, sum, k));
}
// flag = 1: the split would lose too much against dgemm's error model, use DMMA (see RANGE GUARD above)
__global__ void k_ozaki_decide(const double* __restrict__ peaks, double limit, int32_t* __restrict__ flag) {
  *flag = (limit > 0.0 && peaks[0] * peaks[1] > limit) ? 1 : 0;
}

// out[(off_p + kk) + i*ldo] = slice p of X(i, kk) / scale[i], p = 0..S-1, off_p = (reverse ? S-1-p : p) * kp;
// kk in [k, kp) is zero padding.  A 32 x 32 shared-memory tile turns the MN-contiguous form around so that both the
// Float64 reads and the int8 writes are coalesced.
template <bool MN_CONTIG>
__global__ void __launch_bounds__(256) k_slice(const double* __restrict__ X, int64_t ld, int64_t mn, int64_t k, int64_t kp,
                                               const double* __restrict__ scale, int8_t* __restrict__ out, int64_t ldo,
                                               int reverse, int S) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t k0 = (int64_t)blockIdx.x * 32, i0 = (int64_t)blockIdx.y * 32;
  // load a 32 (k) x 32 (mn) tile: tile[kk][i]
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    if (MN_CONTIG) {  // X[i + kk*ld]: threads along i
      const int64_t kk = k0 + r, i = i0 + tx;
      tile[r][tx] = (kk < k && i < mn) ? X[i + kk * ld] : 0.0;
    } else {          // X[kk + i*ld]: threads along kk
      const int64_t kk = k0 + tx, i = i0 + r;
      tile[tx][r] = (kk < k && i < mn) ? X[kk + i * ld] : 0.0;
    }
  }
  __syncthreads();
  // write: threads along kk (the contiguous direction of the output)
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t kk = k0 + tx, i = i0 + r;
    if (kk >= kp || i >= mn) continue;
    double x = tile[tx][r] / scale[i];  // power-of-two divide: exact
    int8_t* o = out + kk + i * ldo;
#pragma unroll
    for (int p = 0; p < MAX_S; ++p) {
      if (p >= S) break;
      const double t = x * 128.0;
      const double q = rint(t);
      o[(int64_t)(reverse ? S - 1 - p : p) * kp] = (int8_t)(int)q;
      x = t - q;
    }
  }
}

}  // namespace

// Returns NB_ERR_UNSUPPORTED (nothing launched) when the product is not worth / not fit for the split path; the
// caller (nb_gemm) then uses DMMA.
int32_t nb_gemm_ozaki(nb_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                      int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
  if (m < 128 || n < 128 || k < 256) return NB_ERR_UNSUPPORTED;
  static const double min_work = getenv("NB_OZAKI_MIN_MNK") ? atof(getenv("NB_OZAKI_MIN_MNK")) : 3.0e9;  // crossover measured at ~3e9 (tools/ozaki_crossover.py)
  if ((double)m * (double)n * (double)k < min_work) return NB_ERR_UNSUPPORTED;  // small products: DMMA wins
  static const int S = [] {
    const char* e = getenv("NB_OZAKI_SLICES");
    const int v = e ? atoi(e) : MAX_S;
    return v < 2 ? 2 : v > MAX_S ? MAX_S : v;
  }();
  const int64_t kp = (k + 127) / 128 * 128;
  if ((double)S * (double)kp * 4096.0 >= 2147483648.0) return NB_ERR_UNSUPPORTED;  // int32 accumulator must stay exact
  const int64_t ld8 = (int64_t)S * kp;
  static const double peak_limit = getenv("NB_OZAKI_PEAK_LIMIT") ? atof(getenv("NB_OZAKI_PEAK_LIMIT")) : 16777216.0;
  void *sa = nullptr, *sb = nullptr, *a8 = nullptr, *b8 = nullptr, *gd = nullptr, *sums = nullptr;
  // guard words: peaks[2] (doubles) + the decision flag
  int32_t rc = nb_pool_alloc(ctx, 32, &gd);
  if (rc == NB_OK) rc = nb_pool_alloc(ctx, (size_t)m * sizeof(double), &sa);
  if (rc == NB_OK) rc = nb_pool_alloc(ctx, (size_t)n * sizeof(double), &sb);
  if (rc == NB_OK) rc = nb_pool_alloc(ctx, (size_t)m * ld8, &a8);
  if (rc == NB_OK) rc = nb_pool_alloc(ctx, (size_t)n * ld8, &b8);
  auto cleanup = [&]() {
    if (sa) nb_pool_release(ctx, sa);
    if (sb) nb_pool_release(ctx, sb);
    if (a8) nb_pool_release(ctx, a8);
    if (b8) nb_pool_release(ctx, b8);
    if (gd) nb_pool_release(ctx, gd);
    if (sums) nb_pool_release(ctx, sums);
  };
  if (rc != NB_OK) { cleanup(); return rc; }
  cudaStream_t st = ctx->stream;
  double* peaks = (double*)gd;
  int32_t* flag = (int32_t*)(peaks + 2);
  cudaMemsetAsync(gd, 0, 32, st);
  // MN-contiguous passes: CTAs of 32 lines over K ranges; (K ranges, K per range) for an operand with mn lines
  auto mn_plan = [&](int64_t mn, int64_t* kpc) -> int64_t {
    const int64_t lines = (mn + 31) / 32;
    int64_t ksplit = ((int64_t)ctx->num_sms * 8 + lines - 1) / lines;
    if (ksplit > (k + 255) / 256) ksplit = (k + 255) / 256;
    if (ksplit < 1) ksplit = 1;
    *kpc = (k + ksplit - 1) / ksplit;
    return (k + *kpc - 1) / *kpc;
  };
  // op(A) line i over kk: not transposed -> A[i + kk*lda] (MN-contiguous); transposed -> A[kk + i*lda]
  // op(B) line j over kk: not transposed -> B[kk + j*ldb] (K-contiguous);  transposed -> B[j + kk*ldb]
  const bool a_mnc = !tA, b_mnc = tB;
  {  // partial sums of |x| of those passes ([K range][line]), added in order by k_absmax_finish
    int64_t kpc = 0, need = 1;
    if (a_mnc) need = mn_plan(m, &kpc) * m;
    if (b_mnc) { const int64_t nb_ = mn_plan(n, &kpc) * n; if (nb_ > need) need = nb_; }
    rc = nb_pool_alloc(ctx, (size_t)need * sizeof(double), &sums);
    if (rc != NB_OK) { cleanup(); return rc; }
  }
  auto absmax = [&](bool mnc, const double* X, int64_t ld, int64_t mn, double* sc, double* pk) {
    if (mnc) {
      cudaMemsetAsync(sc, 0, (size_t)mn * sizeof(double), st);
      int64_t kpc = 0;
      const int64_t ny = mn_plan(mn, &kpc);
      k_absmax_mncontig<<<dim3((unsigned)((mn + 31) / 32), (unsigned)ny), 256, 0, st>>>(X, ld, mn, k, kpc, sc, (double*)sums);
      k_absmax_finish<<<(unsigned)((mn + 255) / 256), 256, 0, st>>>(mn, k, sc, (const double*)sums, (int)ny, pk);
      ctx->stats.kernel_launches += 2;
    } else {
      k_absmax_kcontig<<<(unsigned)((mn * 32 + 255) / 256), 256, 0, st>>>(X, ld, mn, k, sc, pk);
      ctx->stats.kernel_launches++;
    }
  };
  auto slice = [&](bool mnc, const double* X, int64_t ld, int64_t mn, const double* sc, int8_t* out, int rev) {
    const dim3 grid((unsigned)(kp / 32), (unsigned)((mn + 31) / 32));
    if (mnc) k_slice<true><<<grid, 256, 0, st>>>(X, ld, mn, k, kp, sc, out, ld8, rev, S);
    else k_slice<false><<<grid, 256, 0, st>>>(X, ld, mn, k, kp, sc, out, ld8, rev, S);
    ctx->stats.kernel_launches++;
  };
  absmax(a_mnc, A, lda, m, (double*)sa, peaks);
  absmax(b_mnc, B, ldb, n, (double*)sb, peaks + 1);
  k_ozaki_decide<<<1, 1, 0, st>>>(peaks, peak_limit, flag);
  ctx->stats.kernel_launches++;
  if (getenv("NB_OZAKI_DEBUG") && !ctx->capturing) {   // (synchronises: diagnostics only)
    double hp[2] = {0, 0};
    cudaStreamSynchronize(st);
    cudaMemcpy(hp, peaks, sizeof(hp), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[nb_gemm_ozaki] %c%c %lld x %lld x %lld: peak_A %.3g peak_B %.3g product %.3g limit %.3g -> %s\n",
            tA ? 'T' : 'N', tB ? 'T' : 'N', (long long)m, (long long)n, (long long)k, hp[0], hp[1], hp[0] * hp[1],
            peak_limit, (peak_limit > 0 && hp[0] * hp[1] > peak_limit) ? "DMMA" : "split");
  }
  slice(a_mnc, A, lda, m, (const double*)sa, (int8_t*)a8, 0);
  slice(b_mnc, B, ldb, n, (const double*)sb, (int8_t*)b8, 1);
  if (cudaGetLastError() != cudaSuccess) { cleanup(); return nb_fail(ctx, NB_ERR_CUDA, "nb_gemm_ozaki: split kernels failed to launch"); }
  // mid-size products: all levels in ONE launch with a Float64 running sum in the epilogue registers (with fewer
  // than seven slices, the first dropped level rides along); large ones: one CTA-pair launch per level (higher int8
  // pipe utilisation, S passes over C)
  static const double one_launch_below = getenv("NB_OZAKI_ONE_LAUNCH_MNK") ? atof(getenv("NB_OZAKI_ONE_LAUNCH_MNK")) : 2.0e10;
  // (small K: the S read-modify-writes of C outweigh the pair kernel's better pipe utilisation at any m x n)
  if (((double)m * (double)n * (double)k < one_launch_below || k <= 2048) && (double)S * (double)kp < 2147483648.0 / 4096.0) {
    rc = nb_gemm_i8all(ctx, m, n, kp, S, S < MAX_S ? 1 : 0, (const int8_t*)a8, ld8, (const int8_t*)b8, alpha, (const double*)sa,
                       (const double*)sb, beta, C, ldc, flag);
    if (rc != NB_ERR_UNSUPPORTED) {
      if (rc == NB_OK) rc = nb_gemm_dmma(ctx, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, flag);
      cleanup();
      return rc;
    }
    rc = NB_OK;
  }
  for (int l = 0; l < S && rc == NB_OK; ++l) {
    // A slices 0..l (columns [0, (l+1)kp) of the K axis) against B slices l..0 (rows [(S-1-l)kp, S*kp))
    rc = nb_gemm_i8(ctx, m, n, (int64_t)(l + 1) * kp, (const int8_t*)a8, ld8, (const int8_t*)b8 + (int64_t)(S - 1 - l) * kp,
                    ld8, alpha * ldexp(1.0, -7 * (l + 2)), (const double*)sa, (const double*)sb, l == 0 ? beta : 1.0, C, ldc,
                    flag);
  }
  if (rc == NB_OK) rc = nb_gemm_dmma(ctx, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, flag);
  cleanup();
  return rc;
}
