// Family (3), Float64: C = alpha*op(A)*op(B) + beta*C on the FP64 tensor path (DMMA).
//
// tcgen05 has no f64 kind, and Float64 is the element type of every reference example (rtol 1e-10), so
// the Float64 GEMM sensitivities run on the legacy warp-level `mma.sync.m8n8k4.f64` (SASS DMMA): exact
// FP64 FMA semantics, same rounding class as OpenBLAS dgemm behind ChainRules' rrule(*)
// (src/sensitivities/chainrules.jl:179-191,286; src/sensitivities/linalg/blas.jl:224-228).
//
// CTA tile 128 x 128 x 16, 8 warps (2 x 4), warp tile 64 x 32 = 8 x 4 m8n8 accumulator tiles
// (128 accumulator registers per lane).  Operands of all four transposes are normalised while being
// staged: shared memory always holds As[k][m] and Bs[k][n] with a row pitch of 132 doubles, which makes
// the fragment loads (lane -> (row = lane/4, k = lane%4)) bank-conflict free.  Global loads of tile i+1
// are issued into registers before the DMMAs of tile i (register double buffering).
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, LD = 132;

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}

// Loads the (BK x 128) staged slice of one operand into 8 registers per thread.
//   contiguous-along-MN operands (A not transposed / B transposed): element (mn, k) at P[mn + k*ld]
//   contiguous-along-K  operands (A transposed / B not transposed): element (mn, k) at P[k + mn*ld]
// r[2*j], r[2*j+1] are two neighbours along the contiguous direction.
template <bool MN_CONTIG>
__device__ __forceinline__ void load_tile(const double* __restrict__ P, int64_t ld, int64_t mn0, int64_t k0,
                                          int64_t MN, int64_t K, bool vec, int tid, double (&r)[8]) {
  if (MN_CONTIG) {
    const int mn = (tid & 63) * 2;
    const int kb = tid >> 6;  // 0..3
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gk = k0 + kb + 4 * j, gmn = mn0 + mn;
      const double* p = P + gmn + gk * ld;
      if (vec && gk < K && gmn + 1 < MN) {
        const double2 v = *reinterpret_cast<const double2*>(p);
        r[2 * j] = v.x; r[2 * j + 1] = v.y;
      } else {
        r[2 * j] = (gk < K && gmn < MN) ? p[0] : 0.0;
        r[2 * j + 1] = (gk < K && gmn + 1 < MN) ? p[1] : 0.0;
      }
    }
  } else {
    const int k = (tid & 7) * 2;
    const int mb = tid >> 3;  // 0..31
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gmn = mn0 + mb + 32 * j, gk = k0 + k;
      const double* p = P + gk + gmn * ld;
      if (vec && gmn < MN && gk + 1 < K) {
        const double2 v = *reinterpret_cast<const double2*>(p);
        r[2 * j] = v.x; r[2 * j + 1] = v.y;
      } else {
        r[2 * j] = (gmn < MN && gk < K) ? p[0] : 0.0;
        r[2 * j + 1] = (gmn < MN && gk + 1 < K) ? p[1] : 0.0;
      }
    }
  }
}

template <bool MN_CONTIG>
__device__ __forceinline__ void store_tile(double (*S)[LD], int tid, const double (&r)[8]) {
  if (MN_CONTIG) {
    const int mn = (tid & 63) * 2, kb = tid >> 6;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<double2*>(&S[kb + 4 * j][mn]) = make_double2(r[2 * j], r[2 * j + 1]);
  } else {
    const int k = (tid & 7) * 2, mb = tid >> 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      S[k][mb + 32 * j] = r[2 * j];
      S[k + 1][mb + 32 * j] = r[2 * j + 1];
    }
  }
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(256, 1)
k_gemm_dmma(int64_t M, int64_t N, int64_t K, double alpha, const double* __restrict__ A, int64_t lda,
            const double* __restrict__ B, int64_t ldb, double beta, double* __restrict__ C, int64_t ldc,
            int vecA, int vecB, int tiles_m, int64_t k_per_split, double* __restrict__ ws,
            const int32_t* __restrict__ run_if) {
  if (run_if && !*run_if) return;  // a product decided on the device (ozaki.cu range guard): not this path
  // two shared-memory stages: slice i+1 is stored while slice i is multiplied, one barrier per slice (ncu on the
  // single-stage version: Tensor (DP) pipe 76.8 % active, no eligible warp 84 % of the cycles — the 8 warps
  // drained the pipe at both barriers of every slice)
  extern __shared__ __align__(16) unsigned char smem_dyn[];
  double (*As)[BK][LD] = reinterpret_cast<double (*)[BK][LD]>(smem_dyn);
  double (*Bs)[BK][LD] = reinterpret_cast<double (*)[BK][LD]>(smem_dyn + 2 * sizeof(double) * BK * LD);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;  // 2 x 4 warps
  const int64_t m0 = (int64_t)(blockIdx.x % tiles_m) * BM, n0 = (int64_t)(blockIdx.x / tiles_m) * BN;
  const int g = lane >> 2, t = lane & 3;

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // split-K: blockIdx.y owns the K range [kbeg, kend); partial products go to the workspace slice
  const int64_t kbeg = (int64_t)blockIdx.y * k_per_split;
  const int64_t kend = min(K, kbeg + k_per_split);
  double ra[8], rb[8];
  load_tile<!TA>(A, lda, m0, kbeg, M, kend, vecA, tid, ra);
  load_tile<TB>(B, ldb, n0, kbeg, N, kend, vecB, tid, rb);
  store_tile<!TA>(As[0], tid, ra);
  store_tile<TB>(Bs[0], tid, rb);
  if (kbeg + BK < kend) {
    load_tile<!TA>(A, lda, m0, kbeg + BK, M, kend, vecA, tid, ra);
    load_tile<TB>(B, ldb, n0, kbeg + BK, N, kend, vecB, tid, rb);
  }
  __syncthreads();
  int st = 0;
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK, st ^= 1) {
    const bool more = k0 + BK < kend;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[8], bf[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) af[i] = As[st][kk + t][wm * 64 + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[st][kk + t][wn * 32 + j * 8 + g];
      if (kk == 4 && more) {
        // the registers hold slice i+1 (loaded during the previous slice): park it in the other stage, then
        // start the global loads of slice i+2 — both overlap the remaining DMMAs of this slice
        store_tile<!TA>(As[st ^ 1], tid, ra);
        store_tile<TB>(Bs[st ^ 1], tid, rb);
        if (k0 + 2 * BK < kend) {
          load_tile<!TA>(A, lda, m0, k0 + 2 * BK, M, kend, vecA, tid, ra);
          load_tile<TB>(B, ldb, n0, k0 + 2 * BK, N, kend, vecB, tid, rb);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j], af[i], bf[j]);
    }
    __syncthreads();  // stage st^1 is complete and stage st is free for slice i+2
  }
  // epilogue: lane holds C[row = g, cols = 2t, 2t+1] of every 8x8 tile
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gm = m0 + wm * 64 + i * 8 + g;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t gn = n0 + wn * 32 + j * 8 + 2 * t + e;
        if (gn >= N) continue;
        if (ws) {
          ws[(int64_t)blockIdx.y * M * N + gm + gn * M] = acc[i][j][e];
        } else {
          double* c = C + gm + gn * ldc;
          const double v = alpha * acc[i][j][e];
          *c = beta == 0.0 ? v : v + beta * *c;
        }
      }
    }
  }
}

// C = alpha * sum_s ws[s] + beta * C   (deterministic second phase of split-K)
__global__ void __launch_bounds__(256) k_splitk_reduce(const double* __restrict__ ws, int splits, int64_t M,
                                                       int64_t N, double alpha, double beta,
                                                       double* __restrict__ C, int64_t ldc,
                                                       const int32_t* __restrict__ run_if) {
  if (run_if && !*run_if) return;
  const int64_t total = M * N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    double s = 0.0;
    for (int z = 0; z < splits; ++z) s += ws[(int64_t)z * total + i];
    double* c = C + i % M + (i / M) * ldc;
    *c = beta == 0.0 ? alpha * s : alpha * s + beta * *c;
  }
}

}  // namespace

int32_t nb_gemm_dmma(nb_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, double alpha,
                     const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
                     int64_t ldc, const int32_t* run_if) {
  const int64_t tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN;
  if (tiles_m * tiles_n > 0x7fffffffLL) return NB_ERR_UNSUPPORTED;
  const int vecA = ((uintptr_t)A % 16 == 0) && (lda % 2 == 0);
  const int vecB = ((uintptr_t)B % 16 == 0) && (ldb % 2 == 0);
  const int64_t tiles = tiles_m * tiles_n;
  cudaStream_t st = ctx->stream;
  // split-K when the output has too few tiles to occupy the SMs (weight gradients: small m x n, k = batch)
  int splits = 1;
  if (tiles < ctx->num_sms && k >= 8 * BK) {
    // one CTA per SM: pick the split count (up to ~2 waves) whose grid fills whole waves best — 28 tiles x 11 splits
    // = 308 CTAs ran as three waves on 148 SMs, 28 x 10 = 280 as two.  With more than half a wave of tiles the split
    // must pay for its workspace pass: taken only for a quarter more of the part (96 tiles -> 3 x 96 = 288 CTAs)
    int64_t max_splits = k / (4 * BK);
    if (max_splits > 64) max_splits = 64;
    const int64_t hi = (2 * (int64_t)ctx->num_sms + tiles - 1) / tiles;
    const double eff1 = (double)tiles / (double)ctx->num_sms;
    double best = tiles * 2 <= ctx->num_sms ? 0.0 : 1.25 * eff1;
    for (int64_t s = 2; s <= hi && s <= max_splits; ++s) {
      const int64_t ctas = tiles * s, waves = (ctas + ctx->num_sms - 1) / ctx->num_sms;
      const double eff = (double)ctas / (double)(waves * ctx->num_sms);
      if (eff >= best - 1e-9) { best = eff; splits = (int)s; }   // ties: the larger split (shorter CTAs)
    }
  }
  int64_t kps = k;
  double* ws = nullptr;
  if (splits > 1) {
    kps = ((k + splits - 1) / splits + BK - 1) / BK * BK;  // multiple of BK keeps the vector loads aligned
    splits = (int)((k + kps - 1) / kps);
    void* p = nullptr;
    int32_t rc = nb_pool_alloc(ctx, (size_t)splits * m * n * sizeof(double), &p);
    if (rc != NB_OK) return rc;
    ws = (double*)p;
  }
  const dim3 grid((unsigned)tiles, (unsigned)splits);
  constexpr int DMMA_SMEM = 4 * (int)sizeof(double) * BK * LD;  // two stages of As and Bs
#define NB_DMMA(TA_, TB_)                                                                                       \
  do {                                                                                                          \
    static bool attr_done[64] = {};  /* per device: cudaFuncSetAttribute applies to the current device only */                                                                              \
    if (!attr_done[ctx->device & 63]) {                                                                                           \
      NB_CUDA_OK(ctx, cudaFuncSetAttribute(k_gemm_dmma<TA_, TB_>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           DMMA_SMEM));                                                         \
      attr_done[ctx->device & 63] = true;                                                                                         \
    }                                                                                                           \
    k_gemm_dmma<TA_, TB_><<<grid, 256, DMMA_SMEM, st>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, vecA,     \
                                                        vecB, (int)tiles_m, kps, ws, run_if);                   \
  } while (0)
  if (!tA && !tB) NB_DMMA(false, false);
  else if (!tA && tB) NB_DMMA(false, true);
  else if (tA && !tB) NB_DMMA(true, false);
  else NB_DMMA(true, true);
#undef NB_DMMA
  NB_LAUNCH_CHECK(ctx);
  if (ws) {
    int64_t blocks = (m * n + 255) / 256;
    if (blocks > ctx->num_sms * 8) blocks = ctx->num_sms * 8;
    k_splitk_reduce<<<(unsigned)blocks, 256, 0, st>>>(ws, splits, m, n, alpha, beta, C, ldc, run_if);
    NB_LAUNCH_CHECK(ctx);
    nb_pool_release(ctx, ws);
  }
  return NB_OK;
}
