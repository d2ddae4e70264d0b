// Family (3), skinny shapes: products with one dimension <= 16 are not tensor-core work — they stream one big
// operand (or one big result) once and are bound by HBM, so they run as FP32 / FP64 FMA kernels built for bytes in
// flight instead of on a 128-row tensor tile:
//
//   small M / small N   C(m x n) = op(A) op(B) with m <= 16 or n <= 16: the output layer of the MLP
//                       (`logits = W5*H4`, 10 x B x 8192) and its weight sensitivity (`W5bar = Z5bar*H4'`,
//                       10 x 8192 x B), matrix*vector and vector'*matrix (ChainRules rrule(*) with a vector operand,
//                       BLAS.gemv: examples/symmetric_quadratic.jl:15).  The big operand is streamed exactly once with
//                       16-byte loads, split over enough CTAs (split-K, deterministic two-phase sum) to fill the part.
//   small K             C(m x n) = op(A) op(B) with k <= 16: `H4bar = W5'*Z5bar` (8192 x B x 10) and the outer product
//                       ybar*x' (BLAS.ger, the adjoint of matrix*vector).  One pass over the result with the fused
//                       epilogues of nb_gemm_fused (`.* f'(y)`, bias + activation, row sums, operand split emission).
//
// Reference: ChainRules rrule(*) / BLAS.gemm / BLAS.gemv adjoints reached through
// src/sensitivities/chainrules.jl:179-191,286 and src/sensitivities/linalg/blas.jl:224-228; OpenBLAS sgemm/dgemm/gemv
// do this work in the reference.  Arithmetic is plain FMA in the element type (no operand splitting): at least as
// accurate as every tensor-core mode.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ops.cuh"

// gemm_tc.cu: operand-split emission for the product that consumes C next
int32_t nb_split_emit_begin(nb_ctx* ctx, int32_t math_mode, int32_t epi_kind, int32_t epi_op, int64_t m, int64_t n,
                            void** hi, void** lo, int64_t* ld, int32_t* fp16, float* scale, float** sc);
void nb_split_emit_end(nb_ctx* ctx, bool ok, const void* C, int64_t m, int64_t n, int64_t ldc, int64_t ld, void* hi,
                       void* lo, float* sc);

namespace {

template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };

template <typename T>
__device__ __forceinline__ void ldvec(const T* p, T (&v)[Vec<T>::N]) {
  using VT = typename Vec<T>::type;
  const VT q = *reinterpret_cast<const VT*>(p);
  if constexpr (Vec<T>::N == 4) { v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; } else { v[0] = q.x; v[1] = q.y; }
}
template <typename T>
__device__ __forceinline__ void stvec(T* p, const T (&v)[Vec<T>::N]) {
  using VT = typename Vec<T>::type;
  VT q;
  if constexpr (Vec<T>::N == 4) { q.x = v[0]; q.y = v[1]; q.z = v[2]; q.w = v[3]; } else { q.x = v[0]; q.y = v[1]; }
  *reinterpret_cast<VT*>(p) = q;
}

// the streamed operand is read exactly once: keep it out of L1 so the small operand (re-read by every warp) stays there
template <typename T>
__device__ __forceinline__ void ldstream(const T* p, T (&v)[Vec<T>::N]) {
  if constexpr (sizeof(T) == 4) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
  } else {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
  }
}

constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

// The small operand S (R <= 16 rows) is re-laid out once per call into the form each streaming kernel reads with
// aligned 16-byte loads and no shared-memory staging (it stays L1/L2 resident: R x K elements, <= 4 MB):
//   k_skinny_cc reads Sp[k][MRP]  (the R values of one k contiguous, padded to a multiple of 4)
//   k_skinny_kc reads St[r][Kp]   (row r contiguous along k, Kp a multiple of 128, zero padded)
template <typename T>
__global__ void __launch_bounds__(256) k_pack_s(const T* __restrict__ S, int64_t sr, int64_t sk, int R, int64_t K,
                                                T* __restrict__ out, int64_t o_r, int64_t o_k, int RP, int64_t Kp) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)RP * Kp) return;
  // thread order follows the OUTPUT's contiguous index so the writes coalesce
  int64_t k;
  int r;
  if (o_r == 1) { r = (int)(e % RP); k = e / RP; } else { k = e % Kp; r = (int)(e / Kp); }
  out[(int64_t)r * o_r + k * o_k] = (r < R && k < K) ? S[(int64_t)r * sr + k * sk] : T(0);
}

// ------------------------------------------------------------------------------------------------------------
// out[r, c] = sum_k S[r, k] * G[c + k * ldg]      the big operand G has the OUTPUT index c contiguous
// (small M: C = A * B' with B stored n x k;  small N: C' = B' * A' with A stored m x k).  Threads own V consecutive
// c, CTAs own 256 * V of them and one K range (split-K over gridDim.z); partial sums go to ws[z][r][c].
// No barriers: eight 16-byte loads of G in flight per thread, S read through L1 as warp-uniform 16-byte loads.
// ------------------------------------------------------------------------------------------------------------
template <typename T, int MR>
__global__ void __launch_bounds__(256, 3) k_skinny_cc(const T* __restrict__ G, int64_t ldg, int64_t ncols, int64_t K,
                                                      const T* __restrict__ Sp, int R, T* __restrict__ ws, int64_t kper,
                                                      int vec_ok, int ncg, int nitems) {
  constexpr int V = Vec<T>::N;
  constexpr int MRP = round_up(MR, 4);
  constexpr int UK = 8;
  const int tid = threadIdx.x;
  // work item = (column group, K range); the grid is exactly one resident wave and strides over the items
  for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
  const int cg = w % ncg, z = w / ncg;
  const int64_t c0 = ((int64_t)cg * 256 + tid) * V;
  if (c0 >= ncols) continue;
  const int64_t kb = (int64_t)z * kper;
  const int64_t ke = kb + kper < K ? kb + kper : K;
  const bool fast = vec_ok && c0 + V <= ncols;
  T acc[MR][V];
#pragma unroll
  for (int r = 0; r < MR; ++r)
#pragma unroll
    for (int v = 0; v < V; ++v) acc[r][v] = T(0);
  const T* gp = G + c0;
  int64_t k = kb;
  if (fast) {
    for (; k + UK <= ke; k += UK) {
      T b[UK][V];
#pragma unroll
      for (int u = 0; u < UK; ++u) ldstream<T>(gp + (k + u) * ldg, b[u]);
#pragma unroll
      for (int u = 0; u < UK; ++u) {
        T s[MRP];
#pragma unroll
        for (int q = 0; q < MRP / V; ++q) ldvec<T>(Sp + (k + u) * MRP + q * V, *reinterpret_cast<T(*)[V]>(&s[q * V]));
#pragma unroll
        for (int r = 0; r < MR; ++r)
#pragma unroll
          for (int v = 0; v < V; ++v) acc[r][v] = fma(s[r], b[u][v], acc[r][v]);
      }
    }
  }
  for (; k < ke; ++k) {
    T b[V];
#pragma unroll
    for (int v = 0; v < V; ++v) b[v] = c0 + v < ncols ? gp[k * ldg + v] : T(0);
#pragma unroll
    for (int r = 0; r < MR; ++r) {
      const T s = Sp[k * MRP + r];
#pragma unroll
      for (int v = 0; v < V; ++v) acc[r][v] = fma(s, b[v], acc[r][v]);
    }
  }
#pragma unroll
  for (int r = 0; r < MR; ++r) {
    if (r >= R) break;
    T* wp = ws + ((int64_t)z * R + r) * ncols + c0;
    if (c0 + V <= ncols && (ncols % V) == 0) {
      stvec<T>(wp, acc[r]);
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (c0 + v < ncols) wp[v] = acc[r][v];
    }
  }
  }
}

// ------------------------------------------------------------------------------------------------------------
// out[r, c] = sum_k S[r, k] * G[k + c * ldg]      the big operand G has the REDUCTION index k contiguous
// (small M: C = A * B with B stored k x n;  small N: C' = B' * A with A stored k x m).  A warp owns NC columns of G
// and walks K with 16-byte loads per lane (two batches = 2 * NC loads in flight per lane); S comes from the packed
// k-contiguous copy through L1; warp-shuffle reduction at the end; split-K partials go to ws[z][r][c].
// ------------------------------------------------------------------------------------------------------------
template <typename T, int MR>
__global__ void __launch_bounds__(256, 2) k_skinny_kc(const T* __restrict__ G, int64_t ldg, int64_t ncols, int64_t K,
                                                      const T* __restrict__ St, int64_t Kp, int R, T* __restrict__ ws,
                                                      int64_t kper, int vec_ok, int ncg, int nitems) {
  constexpr int V = Vec<T>::N;
  constexpr int NC = sizeof(T) == 8 ? 2 : 4;   // columns per warp (Float64 accumulators take two registers each)
  constexpr int STEP = 32 * V;   // reduction indices per warp iteration
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
  const int cg = w % ncg, z = w / ncg;
  const int64_t cb = ((int64_t)cg * 8 + warp) * NC;
  if (cb >= ncols) continue;
  const int64_t kb = (int64_t)z * kper;
  const int64_t ke = kb + kper < K ? kb + kper : K;
  T acc[MR][NC];
#pragma unroll
  for (int r = 0; r < MR; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[r][c] = T(0);
  const bool colfull = cb + NC <= ncols;
  int64_t k0 = kb;
  if (vec_ok && colfull) {
    for (; k0 + 2 * STEP <= ke; k0 += 2 * STEP) {
      const int64_t kg = k0 + lane * V;
      T b0[NC][V], b1[NC][V];
#pragma unroll
      for (int c = 0; c < NC; ++c) ldstream<T>(G + (cb + c) * ldg + kg, b0[c]);
#pragma unroll
      for (int c = 0; c < NC; ++c) ldstream<T>(G + (cb + c) * ldg + kg + STEP, b1[c]);
#pragma unroll
      for (int r = 0; r < MR; ++r) {
        T s0[V], s1[V];
        ldvec<T>(St + (int64_t)r * Kp + kg, s0);
        ldvec<T>(St + (int64_t)r * Kp + kg + STEP, s1);
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int v = 0; v < V; ++v) acc[r][c] = fma(s0[v], b0[c][v], acc[r][c]);
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int v = 0; v < V; ++v) acc[r][c] = fma(s1[v], b1[c][v], acc[r][c]);
      }
    }
  }
  for (; k0 < ke; k0 += STEP) {   // tails, ragged columns, unaligned operands
    const int64_t kg = k0 + lane * V;
    T b[NC][V];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int v = 0; v < V; ++v) b[c][v] = (cb + c < ncols && kg + v < ke) ? G[(cb + c) * ldg + kg + v] : T(0);
#pragma unroll
    for (int r = 0; r < MR; ++r) {
      T s[V];
      ldvec<T>(St + (int64_t)r * Kp + kg, s);   // (Kp is padded with zeros to a multiple of STEP)
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int v = 0; v < V; ++v) acc[r][c] = fma(s[v], b[c][v], acc[r][c]);
    }
  }
  // butterfly: every lane ends with the warp's total; lane (r * NC + c) % 32 stores element (r, c)
#pragma unroll
  for (int r = 0; r < MR; ++r) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      T v = acc[r][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == ((r * NC + c) & 31) && r < R && cb + c < ncols)
        ws[((int64_t)z * R + r) * ncols + cb + c] = v;
    }
  }
  }
}

// Second phase of the two kernels above: sums the split-K partials in order (deterministic) and applies
// alpha / beta and the fused epilogue.  Element (r, c) of the skinny result is C[i, j] with (i, j) = (r, c) (small M)
// or (c, r) (small N, `swapped`).
template <typename T>
__global__ void __launch_bounds__(256) k_skinny_finish(const T* __restrict__ ws, int nsplit, int R, int64_t ncols,
                                                       T alpha, T beta, T* __restrict__ C, int64_t ldc, int swapped,
                                                       int ekind, int eop, const T* __restrict__ bias,
                                                       const T* __restrict__ y, int64_t ldy) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;   // one thread per element: the partial sums of an element are summed in order
  if (c >= ncols) return;
  const T* wp = ws + (int64_t)r * ncols + c;
  const int64_t zs = (int64_t)R * ncols;
  T s = T(0);
  int z = 0;
  for (; z + 8 <= nsplit; z += 8) {   // eight independent loads in flight, added in order
    T t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = wp[(int64_t)(z + u) * zs];
#pragma unroll
    for (int u = 0; u < 8; ++u) s += t[u];
  }
  for (; z < nsplit; ++z) s += wp[(int64_t)z * zs];
  const int64_t i = swapped ? c : r, j = swapped ? r : c;
  T v = alpha * s;
  T* cp = C + i + j * ldc;
  if (beta != T(0)) v += beta * *cp;
  if (ekind == 1) v = nb_eval<T>(eop, v + (bias ? bias[i] : T(0)), T(0));
  else if (ekind == 2) v *= nb_dunary<T>(eop, T(0), y[i + j * ldy]);
  *cp = v;
}

// ------------------------------------------------------------------------------------------------------------
// small K:  C[i, j] = epilogue(alpha * sum_{k < KK} opA[i, k] * opB[k, j] + beta * C[i, j])
// Threads own V consecutive rows i (16-byte accesses along the columns of C), a CTA owns 256 * V rows x NCB columns;
// the KK x V operand-A values live in registers for the whole CTA, opB is staged in shared memory and broadcast.
// ------------------------------------------------------------------------------------------------------------
template <typename T>
struct SmallK {
  int64_t m, n;
  int k;
  const T* A; int64_t a_si, a_sk;
  const T* B; int64_t b_sk, b_sj;
  T* C; int64_t ldc;
  T alpha, beta;
  int ekind, eop;
  const T* bias;
  const T* y; int64_t ldy;
  T* rowsum_partial;        // [gridDim.y][m]
  int ncb;                  // columns per CTA (<= 128; fewer when 128 would leave most SMs without a CTA)
  void* s_hi; void* s_lo; int64_t ld_s; int s_fp16; float s_scale;   // Float32 only
  int vec;
};

__device__ __forceinline__ uint16_t f2bf(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint16_t f2h(float x) { return __half_as_ushort(__float2half_rn(x)); }

template <typename T, int KK>
__global__ void __launch_bounds__(256, 2) k_gemm_smallk(const SmallK<T> g) {
  constexpr int V = Vec<T>::N;
  constexpr int NCB = 128, U = 4;
  constexpr int KKP = round_up(KK, V);
  __shared__ __align__(16) T Bs[NCB][KKP];
  const int tid = threadIdx.x;
  const int64_t i0 = ((int64_t)blockIdx.x * 256 + tid) * V;
  const int64_t j0 = (int64_t)blockIdx.y * g.ncb;
  const int ncb = (int)(g.n - j0 < g.ncb ? g.n - j0 : g.ncb);
  for (int e = tid; e < g.ncb * KKP; e += 256) {
    const int jj = e / KKP, kk = e % KKP;
    Bs[jj][kk] = (jj < ncb && kk < g.k) ? g.B[(int64_t)kk * g.b_sk + (j0 + jj) * g.b_sj] : T(0);
  }
  T a[KK][V];
#pragma unroll
  for (int kk = 0; kk < KK; ++kk)
#pragma unroll
    for (int v = 0; v < V; ++v)
      a[kk][v] = (kk < g.k && i0 + v < g.m) ? g.A[(i0 + v) * g.a_si + (int64_t)kk * g.a_sk] : T(0);
  __syncthreads();
  T rs[V];
#pragma unroll
  for (int v = 0; v < V; ++v) rs[v] = T(0);
  if (i0 < g.m) {
    const bool fast = g.vec && i0 + V <= g.m;
    T bias[V];
#pragma unroll
    for (int v = 0; v < V; ++v) bias[v] = (g.ekind == 1 && g.bias && i0 + v < g.m) ? g.bias[i0 + v] : T(0);
    // the one full-size array the epilogue READS (the saved y of a pullback, or C when beta != 0) is prefetched one
    // column group ahead, so its loads are in flight while the current group is computed and stored
    const T* auxp = g.ekind == 2 ? g.y : (g.beta != T(0) ? g.C : nullptr);
    const int64_t auxld = g.ekind == 2 ? g.ldy : g.ldc;
    const bool need_aux = fast && auxp != nullptr;
    T auxn[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int v = 0; v < V; ++v) auxn[u][v] = T(0);
    if (need_aux) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (u < ncb) ldvec<T>(auxp + i0 + (j0 + u) * auxld, auxn[u]);
    }
    for (int jj = 0; jj < ncb; jj += U) {
      T x[U][V];
      T aux[U][V];
      if (need_aux) {
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int v = 0; v < V; ++v) aux[u][v] = auxn[u][v];
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (jj + U + u < ncb) ldvec<T>(auxp + i0 + (j0 + jj + U + u) * auxld, auxn[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int v = 0; v < V; ++v) x[u][v] = T(0);
        if (jj + u < ncb) {
#pragma unroll
          for (int kk = 0; kk < KK; ++kk) {
            const T b = Bs[jj + u][kk];
#pragma unroll
            for (int v = 0; v < V; ++v) x[u][v] = fma(a[kk][v], b, x[u][v]);
          }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) x[u][v] *= g.alpha;
      }
      if (fast) {
        if (g.ekind == 2) {
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (jj + u < ncb)
#pragma unroll
              for (int v = 0; v < V; ++v) x[u][v] *= nb_dunary<T>(g.eop, T(0), aux[u][v]);
        } else {
          if (g.beta != T(0)) {
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (jj + u < ncb)
#pragma unroll
                for (int v = 0; v < V; ++v) x[u][v] = fma(g.beta, aux[u][v], x[u][v]);
          }
          if (g.ekind == 1) {
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (jj + u < ncb)
#pragma unroll
                for (int v = 0; v < V; ++v) x[u][v] = nb_eval<T>(g.eop, x[u][v] + bias[v], T(0));
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (jj + u >= ncb) continue;
          const int64_t j = j0 + jj + u;
          stvec<T>(g.C + i0 + j * g.ldc, x[u]);
#pragma unroll
          for (int v = 0; v < V; ++v) rs[v] += x[u][v];
          if constexpr (sizeof(T) == 4) {
            if (g.s_hi) {
              __align__(8) uint16_t h[4], l[4];
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                if (g.s_fp16) {
                  const float xs = (float)x[u][v] * g.s_scale;
                  h[v] = f2h(xs);
                  l[v] = f2h(xs - __half2float(__ushort_as_half(h[v])));
                } else {
                  h[v] = f2bf((float)x[u][v]);
                  l[v] = f2bf((float)x[u][v] - __bfloat162float(__ushort_as_bfloat16(h[v])));
                }
              }
              *reinterpret_cast<uint2*>((uint16_t*)g.s_hi + i0 + j * g.ld_s) = *reinterpret_cast<const uint2*>(h);
              *reinterpret_cast<uint2*>((uint16_t*)g.s_lo + i0 + j * g.ld_s) = *reinterpret_cast<const uint2*>(l);
            }
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (jj + u >= ncb) continue;
          const int64_t j = j0 + jj + u;
#pragma unroll
          for (int v = 0; v < V; ++v) {
            if (i0 + v >= g.m) continue;
            T xv = x[u][v];
            T* cp = g.C + i0 + v + j * g.ldc;
            if (g.beta != T(0)) xv = fma(g.beta, *cp, xv);
            if (g.ekind == 2) xv *= nb_dunary<T>(g.eop, T(0), g.y[i0 + v + j * g.ldy]);
            else if (g.ekind == 1) xv = nb_eval<T>(g.eop, xv + bias[v], T(0));
            *cp = xv;
            rs[v] += xv;
            if constexpr (sizeof(T) == 4) {
              if (g.s_hi) {
                uint16_t h, l;
                if (g.s_fp16) {
                  const float xs = (float)xv * g.s_scale;
                  h = f2h(xs);
                  l = f2h(xs - __half2float(__ushort_as_half(h)));
                } else {
                  h = f2bf((float)xv);
                  l = f2bf((float)xv - __bfloat162float(__ushort_as_bfloat16(h)));
                }
                ((uint16_t*)g.s_hi)[i0 + v + j * g.ld_s] = h;
                ((uint16_t*)g.s_lo)[i0 + v + j * g.ld_s] = l;
              }
            }
          }
        }
      }
    }
    if (g.rowsum_partial) {
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (i0 + v < g.m) g.rowsum_partial[(int64_t)blockIdx.y * g.m + i0 + v] = rs[v];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_rowsum_finish_t(const T* __restrict__ partial, int nparts, int64_t M,
                                                         T* __restrict__ out, int acc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  T s = T(0);
  for (int p = 0; p < nparts; ++p) s += partial[(int64_t)p * M + i];
  out[i] = acc ? out[i] + s : s;
}

// ---- host ---------------------------------------------------------------------------------------------------
inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <typename T, template <typename, int> class Launcher, typename... Args>
int32_t dispatch_mr(int R, Args&... args) {
  switch (R) {
    case 1: return Launcher<T, 1>::run(args...);
    case 2: return Launcher<T, 2>::run(args...);
    case 3: return Launcher<T, 3>::run(args...);
    case 4: return Launcher<T, 4>::run(args...);
    case 5: case 6: return Launcher<T, 6>::run(args...);
    case 7: case 8: return Launcher<T, 8>::run(args...);
    case 9: case 10: return Launcher<T, 10>::run(args...);
    case 11: case 12: return Launcher<T, 12>::run(args...);
    default: return Launcher<T, 16>::run(args...);
  }
}

struct StreamArgs {
  nb_ctx* ctx;
  const void* G; int64_t ldg, ncols, K;
  const void* Sp; int64_t Kp;
  int R;
  void* ws; int64_t kper; int nsplit; int vec_ok;
  int ncg;        // column groups (256 * V columns for cc, 32 for kc)
  int query;      // 1: only report the resident CTA capacity in `capacity` (no launch)
  int capacity;
};

template <typename T, int MR>
struct LaunchCC {
  static int32_t run(StreamArgs& a) {
    auto kern = k_skinny_cc<T, MR>;
    if (a.query) {
      int occ = 1;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0);
      a.capacity = (occ > 0 ? occ : 1) * a.ctx->num_sms;
      return NB_OK;
    }
    const int nitems = a.ncg * a.nsplit;
    const int grid = nitems < a.capacity ? nitems : a.capacity;
    kern<<<grid, 256, 0, a.ctx->stream>>>((const T*)a.G, a.ldg, a.ncols, a.K, (const T*)a.Sp, a.R, (T*)a.ws, a.kper,
                                          a.vec_ok, a.ncg, nitems);
    NB_LAUNCH_CHECK(a.ctx);
    return NB_OK;
  }
};
template <typename T, int MR>
struct LaunchKC {
  static int32_t run(StreamArgs& a) {
    auto kern = k_skinny_kc<T, MR>;
    if (a.query) {
      int occ = 1;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0);
      a.capacity = (occ > 0 ? occ : 1) * a.ctx->num_sms;
      return NB_OK;
    }
    const int nitems = a.ncg * a.nsplit;
    const int grid = nitems < a.capacity ? nitems : a.capacity;
    kern<<<grid, 256, 0, a.ctx->stream>>>((const T*)a.G, a.ldg, a.ncols, a.K, (const T*)a.Sp, a.Kp, a.R, (T*)a.ws,
                                          a.kper, a.vec_ok, a.ncg, nitems);
    NB_LAUNCH_CHECK(a.ctx);
    return NB_OK;
  }
};

inline int padded_mr(int R) {   // the MR the dispatcher instantiates for R rows, rounded up to a multiple of 4
  const int mr = R <= 4 ? R : R <= 6 ? 6 : R <= 8 ? 8 : R <= 10 ? 10 : R <= 12 ? 12 : 16;
  return (mr + 3) / 4 * 4;
}

// out[r, c] = sum_k S[r, k] * G(k, c): G is the streamed operand; g_c_contig says which of its indices is contiguous
template <typename T>
int32_t run_stream(nb_ctx* ctx, bool g_c_contig, const T* G, int64_t ldg, int64_t ncols, int64_t K, const T* S,
                   int64_t sr, int64_t sk, int R, double alpha, double beta, T* C, int64_t ldc, bool swapped,
                   const nb_gemm_epilogue* epi) {
  constexpr int V = Vec<T>::N;
  StreamArgs a{};
  a.ctx = ctx; a.G = G; a.ldg = ldg; a.ncols = ncols; a.K = K; a.R = R;
  a.vec_ok = al16(G) && (ldg % V == 0);
  a.query = 1;
  int32_t rc = g_c_contig ? dispatch_mr<T, LaunchCC>(R, a) : dispatch_mr<T, LaunchKC>(R, a);
  if (rc != NB_OK) return rc;
  a.query = 0;
  // Work items = column groups x K ranges, run by ONE resident wave of CTAs striding over them: split K so that the
  // items fill whole rounds of that wave (a 1.3-wave grid leaves two thirds of the part idle for the last third)
  const int64_t kc_cols = 8 * (sizeof(T) == 8 ? 2 : 4);   // columns per CTA of k_skinny_kc
  const int64_t ncg = g_c_contig ? (ncols + 256 * V - 1) / (256 * V) : (ncols + kc_cols - 1) / kc_cols;
  int64_t kquant = g_c_contig ? 64 : 32 * V * 2;
  const int64_t cap = a.capacity;
  // few output columns (300 x 4096 with 10 rows: one column group): K ranges of 64 would leave 64 CTAs for the whole
  // part; two batches of eight K steps per CTA still keep eight loads in flight
  if (g_c_contig && ncg * ((K + kquant - 1) / kquant) * 2 < cap) kquant = 16;
  const int64_t maxsplit = (K + kquant - 1) / kquant;
  int64_t nsplit = 1;
  if (ncg < cap) {
    nsplit = cap / ncg;   // one round
  } else {
    double best = 0.0;
    for (int64_t t = 1; t <= 8; ++t) {
      const int64_t items = ncg * t;
      const double eff = (double)items / (double)(((items + cap - 1) / cap) * cap);
      if (eff > best + 0.02) { best = eff; nsplit = t; }
    }
  }
  if (nsplit > maxsplit) nsplit = maxsplit;
  if (nsplit > 1024) nsplit = 1024;
  if (nsplit < 1) nsplit = 1;
  int64_t kper = ((K + nsplit - 1) / nsplit + kquant - 1) / kquant * kquant;
  nsplit = (K + kper - 1) / kper;
  const int RP = padded_mr(R);
  const int64_t Kp = (K + 32 * V - 1) / (32 * V) * (32 * V);
  void *ws = nullptr, *sp = nullptr;
  rc = nb_pool_alloc(ctx, (size_t)(nsplit * R * ncols) * sizeof(T), &ws);
  if (rc != NB_OK) return rc;
  rc = nb_pool_alloc(ctx, (size_t)RP * (size_t)Kp * sizeof(T), &sp);
  if (rc != NB_OK) { nb_pool_release(ctx, ws); return rc; }
  {
    const int64_t tot = (int64_t)RP * Kp;
    k_pack_s<T><<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(S, sr, sk, R, K, (T*)sp,
                                                                      g_c_contig ? 1 : Kp, g_c_contig ? RP : 1, RP, Kp);
    ctx->stats.kernel_launches++;
  }
  a.Sp = sp; a.Kp = Kp; a.ws = ws; a.kper = kper; a.nsplit = (int)nsplit; a.ncg = (int)ncg;
  rc = g_c_contig ? dispatch_mr<T, LaunchCC>(R, a) : dispatch_mr<T, LaunchKC>(R, a);
  if (rc == NB_OK) {
    const int ek = epi ? epi->kind : 0;
    k_skinny_finish<T><<<dim3((unsigned)((ncols + 255) / 256), (unsigned)R), 256, 0, ctx->stream>>>(
        (const T*)ws, (int)nsplit, R, ncols, (T)alpha, (T)beta, C, ldc, swapped ? 1 : 0, ek, epi ? epi->op : 0,
        epi ? (const T*)epi->bias : nullptr, epi ? (const T*)epi->y : nullptr, epi ? epi->ldy : 0);
    ctx->stats.kernel_launches++;
    if (cudaGetLastError() != cudaSuccess) rc = nb_fail(ctx, NB_ERR_CUDA, "k_skinny_finish launch failed");
  }
  nb_pool_release(ctx, sp);
  nb_pool_release(ctx, ws);
  return rc;
}

template <typename T, int KK>
int32_t launch_smallk(nb_ctx* ctx, const SmallK<T>& g) {
  constexpr int V = Vec<T>::N;
  dim3 grid((unsigned)((g.m + 256 * V - 1) / (256 * V)), (unsigned)((g.n + g.ncb - 1) / g.ncb));
  k_gemm_smallk<T, KK><<<grid, 256, 0, ctx->stream>>>(g);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

template <typename T>
int32_t run_smallk(nb_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, double alpha, const T* A,
                   int64_t lda, const T* B, int64_t ldb, double beta, T* C, int64_t ldc, const nb_gemm_epilogue* epi,
                   int32_t math_mode) {
  constexpr int V = Vec<T>::N;
  SmallK<T> g{};
  g.m = m; g.n = n; g.k = (int)k;
  g.A = A; g.a_si = tA ? lda : 1; g.a_sk = tA ? 1 : lda;
  g.B = B; g.b_sk = tB ? ldb : 1; g.b_sj = tB ? 1 : ldb;
  g.C = C; g.ldc = ldc;
  g.alpha = (T)alpha; g.beta = (T)beta;
  bool vec = al16(C) && (ldc % V == 0);
  void* rs_partial = nullptr;
  void *e_hi = nullptr, *e_lo = nullptr;
  float* e_sc = nullptr;
  int64_t e_ld = 0;
  // columns per CTA: 128, or fewer (down to 16) until the grid covers the part twice (500 x 4096 x 10 in Float64 is one
  // CTA row: 32 CTAs of 128 columns left four fifths of the SMs idle)
  g.ncb = 128;
  {
    const int64_t gx = (m + 256 * V - 1) / (256 * V);
    while (g.ncb > 16 && gx * ((n + g.ncb - 1) / g.ncb) < 2 * (int64_t)ctx->num_sms) g.ncb >>= 1;
  }
  const int nparts = (int)((n + g.ncb - 1) / g.ncb);
  int32_t rc = NB_OK;
  if (epi && epi->kind) {
    g.ekind = epi->kind; g.eop = epi->op;
    g.bias = (const T*)epi->bias;
    g.y = (const T*)epi->y; g.ldy = epi->ldy;
    if (g.y) vec = vec && al16(g.y) && (g.ldy % V == 0);
    if (epi->rowsum) {
      rc = nb_pool_alloc(ctx, (size_t)nparts * (size_t)m * sizeof(T), &rs_partial);
      if (rc != NB_OK) return rc;
      g.rowsum_partial = (T*)rs_partial;
    }
    if (sizeof(T) == 4 && epi->emit_split) {
      int32_t fp16 = 0;
      float scale = 1.f;
      if (nb_split_emit_begin(ctx, math_mode, epi->kind, epi->op, m, n, &e_hi, &e_lo, &e_ld, &fp16, &scale, &e_sc) ==
          NB_OK && e_hi) {
        g.s_hi = e_hi; g.s_lo = e_lo; g.ld_s = e_ld; g.s_fp16 = fp16; g.s_scale = scale;
        vec = vec && (e_ld % 4 == 0);
      }
    }
  }
  g.vec = vec ? 1 : 0;
  switch (k) {
    case 1: rc = launch_smallk<T, 1>(ctx, g); break;
    case 2: rc = launch_smallk<T, 2>(ctx, g); break;
    case 3: rc = launch_smallk<T, 3>(ctx, g); break;
    case 4: rc = launch_smallk<T, 4>(ctx, g); break;
    case 5: case 6: rc = launch_smallk<T, 6>(ctx, g); break;
    case 7: case 8: rc = launch_smallk<T, 8>(ctx, g); break;
    case 9: case 10: rc = launch_smallk<T, 10>(ctx, g); break;
    case 11: case 12: rc = launch_smallk<T, 12>(ctx, g); break;
    default: rc = launch_smallk<T, 16>(ctx, g); break;
  }
  if (rs_partial) {
    if (rc == NB_OK) {
      k_rowsum_finish_t<T><<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>((const T*)rs_partial, nparts, m,
                                                                                (T*)epi->rowsum, epi->rowsum_acc);
      ctx->stats.kernel_launches++;
      if (cudaGetLastError() != cudaSuccess) rc = nb_fail(ctx, NB_ERR_CUDA, "k_rowsum_finish_t launch failed");
    }
    nb_pool_release(ctx, rs_partial);
  }
  if (e_hi) nb_split_emit_end(ctx, rc == NB_OK, C, m, n, ldc, e_ld, e_hi, e_lo, e_sc);
  return rc;
}

template <typename T>
int32_t skinny_t(nb_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, double alpha, const T* A, int64_t lda,
                 const T* B, int64_t ldb, double beta, T* C, int64_t ldc, const nb_gemm_epilogue* epi,
                 int32_t math_mode) {
  const bool fused = epi && epi->kind != 0;
  if (k <= 16 && m * n >= 4096) {
    return run_smallk<T>(ctx, tA, tB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, epi, math_mode);
  }
  if (fused && epi->rowsum) return NB_ERR_UNSUPPORTED;   // (row sums of a skinny result: the general path has them)
  if (m <= 16 && m <= n && n * k >= 32768) {
    // C[r, c] = sum_k opA[r, k] * opB[k, c]: stream B.  opA[r, k] = A[r + k lda] (N) or A[k + r lda] (T)
    const int64_t sr = tA ? lda : 1, sk = tA ? 1 : lda;
    return run_stream<T>(ctx, /*g_c_contig=*/tB, B, ldb, n, k, A, sr, sk, (int)m, alpha, beta, C, ldc, false, epi);
  }
  if (n <= 16 && m * k >= 32768) {
    // C'[r = j, c = i] = sum_k opB[k, j] * opA[i, k]: stream A.  S[r, k] = B[k + r ldb] (N) or B[r + k ldb] (T)
    const int64_t sr = tB ? 1 : ldb, sk = tB ? ldb : 1;
    return run_stream<T>(ctx, /*g_c_contig=*/!tA, A, lda, m, k, B, sr, sk, (int)n, alpha, beta, C, ldc, true, epi);
  }
  return NB_ERR_UNSUPPORTED;
}

}  // namespace

// C = epilogue(alpha * op(A) * op(B) + beta * C) for the skinny shapes; NB_ERR_UNSUPPORTED (nothing launched) when the
// shape is not one of them.  `epi` may be NULL.
int32_t nb_gemm_skinny(nb_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, double alpha, const void* A,
                       int64_t lda, const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int32_t dtype,
                       const nb_gemm_epilogue* epi, int32_t math_mode) {
  static const bool on = !(getenv("NB_GEMM_SKINNY") && atoi(getenv("NB_GEMM_SKINNY")) == 0);
  if (!on || m <= 0 || n <= 0 || k <= 0) return NB_ERR_UNSUPPORTED;
  if (epi && epi->kind != 0 && beta != 0.0) return NB_ERR_UNSUPPORTED;
  if (dtype == NB_F32)
    return skinny_t<float>(ctx, tA, tB, m, n, k, alpha, (const float*)A, lda, (const float*)B, ldb, beta, (float*)C, ldc,
                           epi, math_mode);
  return skinny_t<double>(ctx, tA, tB, m, n, k, alpha, (const double*)A, lda, (const double*)B, ldb, beta, (double*)C,
                          ldc, epi, math_mode);
}
