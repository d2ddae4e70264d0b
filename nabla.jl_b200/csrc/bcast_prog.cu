// Composite scalar programs (user lambdas in broadcast / map / sum(f, .)) on the lean kernels.
//
// The reference differentiates an arbitrary scalar function per element with ForwardDiff duals
// (fmad, src/core.jl:285-292; broadcast rule functional.jl:59-95).  Here the function is a register program
// (nb_prog, traced once on the host) that the kernel evaluates and differentiates exactly by reverse
// accumulation.  The general kernels (bcast.cu: k_flat/k_tall/k_wcol with GenEv/GenEvS) pay ~700 instructions
// of launch-description bookkeeping per 128-bit vector around the interpreter (ncu: 180 thread-instructions per
// element for a two-instruction program, 10-15 % of the HBM peak).  For programs with one or two array operands
// this file runs the same interpreter inside the LEAN kernels (bcast_lean.cuh): compile-time operand count and
// wanted-sensitivity set, U vectors loaded before any is computed, slot values and adjoints in per-thread
// columns of dynamic shared memory (conflict-free 128-bit accesses), dense jump-table dispatch (nb_prog::hop),
// one decode per instruction for U vectors, and a reverse sweep that reads only what each partial needs and
// stores the first contribution to every adjoint (nb_prog::rflags).
#include "ops.cuh"
#include "bcast_types.cuh"

namespace {
using namespace nbb;
#include "bcast_lean.cuh"

__device__ __forceinline__ void unspecified(float& x) { asm("" : "=f"(x)); }
__device__ __forceinline__ void unspecified(double& x) { asm("" : "=d"(x)); }

template <typename T, int V, int NARGS, int MODE, int UU>
struct ProgPack {
  using Extra = nb_prog;
  using VT = typename VecOf<T>::type;
  static constexpr bool HAS_TAIL = false;  // programs keep the whole-vector restriction
  using Scalar = ProgPack;
  static constexpr int NOUT = MODE == LM_PULL_AB ? 2 : 1;
  static constexpr int U = UU;
  T x[NARGS][V];
  T yb[MODE != LM_FWD ? V : 1];
  T ys[MODE != LM_FWD ? V : 1];  // saved forward value, when it replaces the last (transcendental) instruction

  __device__ __forceinline__ void load(const MRP& P, const nb_prog&, int64_t r, int64_t c) {
#pragma unroll
    for (int k = 0; k < NARGS; ++k) lean_ld<T, V>(P.arg[k], r, c, x[k]);
    if constexpr (MODE != LM_FWD) {
      if (P.need_y) lean_ld<T, V>(P.ys, r, c, ys);
      if (P.has_ybar) lean_ld<T, V>(P.ybar, r, c, yb);
      else {
#pragma unroll
        for (int i = 0; i < V; ++i) yb[i] = T(1);
      }
    }
  }

  template <int N>
  static __device__ __forceinline__ void compute_many(const MRP& P, const nb_prog& G, const ProgPack (&x)[N],
                                                      T (&o)[N][2][V]) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // every lean kernel runs 256 threads per CTA: slot (s, u) of thread t sits at byte ((s*U + u)*256 + t)*16 of
    // the slot area — one 32-bit multiply-add per access, shared-window addressing
    constexpr uint32_t SLOT_STRIDE = 256 * 16;
    const uint32_t tid = threadIdx.y * blockDim.x + threadIdx.x;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw) + (uint32_t)P.slot_off + tid * 16u;
    const int n = G.n_instr, nslots = NARGS + n;
    auto ld = [&](int s, int u, T (&v)[V]) {
      const uint32_t addr = sbase + (uint32_t)(s * U + u) * SLOT_STRIDE;
      if constexpr (sizeof(T) == 4) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
      } else {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(addr));
      }
    };
    auto st = [&](int s, int u, const T (&v)[V]) {
      const uint32_t addr = sbase + (uint32_t)(s * U + u) * SLOT_STRIDE;
      if constexpr (sizeof(T) == 4) {
        asm volatile("st.shared.v4.f32 [%4], {%0, %1, %2, %3};" ::"f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "r"(addr) : "memory");
      } else {
        asm volatile("st.shared.v2.f64 [%2], {%0, %1};" ::"d"(v[0]), "d"(v[1]), "r"(addr) : "memory");
      }
    };
    auto get = [&](int idx, int u, T (&v)[V]) {
      if (idx >= 0) ld(idx, u, v);
      else {
        const T cst = T(G.consts[-idx - 1]);
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = cst;
      }
    };
#pragma unroll
    for (int k = 0; k < NARGS; ++k)
#pragma unroll
      for (int u = 0; u < N; ++u) st(k, u, x[u].x[k]);
    // ---- forward ----
    // In a pullback whose last instruction is a transcendental with f' = f'(y) (tanh, exp, logistic, ...) the
    // host passes the Branch's saved value (core.jl:76) and that instruction is not re-evaluated.
    const int nfwd = (MODE != LM_FWD && P.need_y) ? n - 1 : n;
    if constexpr (MODE != LM_FWD) {
      if (P.need_y) {
#pragma unroll
        for (int u = 0; u < N; ++u) st(NARGS + n - 1, u, x[u].ys);
      }
    }
    T last[N][V];
    const uint32_t skip = MODE != LM_FWD ? G.fwd_skip : 0u;   // values no computed partial reads (nb_prog_derive)
    for (int i = 0; i < nfwd; ++i) {
      if (skip >> i & 1u) continue;
      const int hop = G.hop[i], op = G.op[i], ia = G.a[i], ib = G.b[i];
      const bool bin = hop >= NB_HOP_ADD && hop != NB_HOP_COLD_UN;
#pragma unroll
      for (int u = 0; u < N; ++u) {
        T a[V], b[V];
        get(ia, u, a);
        if (bin) get(ib, u, b);
        else {
#pragma unroll
          for (int v = 0; v < V; ++v) unspecified(b[v]);
        }
        nb_eval_hop<T, V>(hop, op, a, b, last[u]);
        st(NARGS + i, u, last[u]);
      }
    }
    if constexpr (MODE == LM_FWD) {
#pragma unroll
      for (int u = 0; u < N; ++u)
#pragma unroll
        for (int v = 0; v < V; ++v) { o[u][0][v] = n ? last[u][v] : x[u].x[0][v]; o[u][1][v] = T(0); }
      return;
    } else {
      // ---- reverse: the adjoint of slot s lives in slot nslots + s ----
      for (int i = n - 1; i >= 0; --i) {
        const int fl = G.rflags[i];
        if (!(fl & NB_RF_LIVE)) continue;
        const int hop = G.hop[i], op = G.op[i], ia = G.a[i], ib = G.b[i];
        const bool same = (fl & NB_RF_SLOT_B) && ib == ia;  // f(x, x): both partials go to one adjoint
#pragma unroll
        for (int u = 0; u < N; ++u) {
          T g[V], a[V], b[V], y[V], da[V], db[V];
          if (i == n - 1) {
#pragma unroll
            for (int v = 0; v < V; ++v) g[v] = T(1);
          } else {
            ld(nslots + NARGS + i, u, g);
          }
          // inputs a partial does not read stay unspecified (no fill instructions): defined for the compiler only
#pragma unroll
          for (int v = 0; v < V; ++v) { unspecified(a[v]); unspecified(b[v]); unspecified(y[v]); }
          if (fl & NB_RF_NEED_A) get(ia, u, a);
          if (fl & NB_RF_NEED_B) get(ib, u, b);
          if (fl & NB_RF_NEED_Y) ld(NARGS + i, u, y);
          nb_partials_hop<T, V>(hop, op, a, b, y, da, db);
          if (fl & NB_RF_SLOT_A) {
            T t[V];
#pragma unroll
            for (int v = 0; v < V; ++v) t[v] = nb_gmul(g[v], same ? da[v] + db[v] : da[v], G.guard != 0);
            if (!(fl & NB_RF_FIRST_A)) {
              T old[V];
              ld(nslots + ia, u, old);
#pragma unroll
              for (int v = 0; v < V; ++v) t[v] += old[v];
            }
            st(nslots + ia, u, t);
          }
          if ((fl & NB_RF_SLOT_B) && !same) {
            T t[V];
#pragma unroll
            for (int v = 0; v < V; ++v) t[v] = nb_gmul(g[v], db[v], G.guard != 0);
            if (!(fl & NB_RF_FIRST_B)) {
              T old[V];
              ld(nslots + ib, u, old);
#pragma unroll
              for (int v = 0; v < V; ++v) t[v] += old[v];
            }
            st(nslots + ib, u, t);
          }
        }
      }
      constexpr int W0 = MODE == LM_PULL_B ? 1 : 0;  // operand of the first output
#pragma unroll
      for (int u = 0; u < N; ++u) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int w = k == 0 ? W0 : 1;
          if (k >= NOUT) {
#pragma unroll
            for (int v = 0; v < V; ++v) o[u][k][v] = T(0);
          } else if (n == 0) {
#pragma unroll
            for (int v = 0; v < V; ++v) o[u][k][v] = x[u].yb[v];
          } else if (G.arg_live & (1u << w)) {
            T adj[V];
            ld(nslots + w, u, adj);
#pragma unroll
            for (int v = 0; v < V; ++v) o[u][k][v] = x[u].yb[v] * adj[v];
          } else {
#pragma unroll
            for (int v = 0; v < V; ++v) o[u][k][v] = T(0);
          }
        }
      }
    }
  }
};

template <typename T, int NARGS, int MODE, int UU>
void prog_launch(int kind, MRP P, const nb_prog& G, const LeanLaunch& a, cudaStream_t st) {
  constexpr int V = VecOf<T>::V;
  using Pk = ProgPack<T, V, NARGS, MODE, UU>;
  const size_t own = kind == 1 ? a.smem : 0;
  const size_t off = (own + 15) & ~(size_t)15;
  P.slot_off = (int32_t)off;
  const size_t total = off + (size_t)2 * (G.nargs + G.n_instr) * UU * 16 * 256;
  if (kind == 0) {
    if (total > 32 * 1024) cudaFuncSetAttribute(k_flat_lean<T, Pk>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total);
    k_flat_lean<T, Pk><<<a.blocks, 256, total, st>>>(P, a.nvec, G);
  } else if (kind == 1) {
    if (total > 32 * 1024) cudaFuncSetAttribute(k_tall_lean<T, Pk>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total);
    k_tall_lean<T, Pk><<<a.grid, a.block, total, st>>>(P, a.cps, G);
  } else {
    if (total > 32 * 1024) cudaFuncSetAttribute(k_wcol_lean<T, Pk>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total);
    k_wcol_lean<T, Pk><<<a.blocks, 256, total, st>>>(P, G);
  }
}

template <typename T, int UU>
bool prog_dispatch(int kind, int mode, const MRP& P, const nb_prog& G, const LeanLaunch& a, cudaStream_t st) {
  if (G.nargs == 1) {
    if (mode == LM_FWD) { prog_launch<T, 1, LM_FWD, UU>(kind, P, G, a, st); return true; }
    if (mode == LM_PULL_A) { prog_launch<T, 1, LM_PULL_A, UU>(kind, P, G, a, st); return true; }
    return false;
  }
  if (G.nargs == 2) {
    if (mode == LM_FWD) { prog_launch<T, 2, LM_FWD, UU>(kind, P, G, a, st); return true; }
    if (mode == LM_PULL_A) { prog_launch<T, 2, LM_PULL_A, UU>(kind, P, G, a, st); return true; }
    if (mode == LM_PULL_B) { prog_launch<T, 2, LM_PULL_B, UU>(kind, P, G, a, st); return true; }
    if (mode == LM_PULL_AB) { prog_launch<T, 2, LM_PULL_AB, UU>(kind, P, G, a, st); return true; }
  }
  return false;
}

}  // namespace

// kind 0 = flat, 1 = tall, 2 = wcol; returns false when this program / mode has no lean kernel (caller falls back
// to the general kernels).  Slot storage: 2 * (nargs + n_instr) * U 16-byte columns per thread.
bool nb_lean_prog(int dtype, int kind, int mode, const nbb::MRP& P, const nb_prog& G, const nbb::LeanLaunch& a,
                  cudaStream_t st) {
  const int nslots = G.nargs + G.n_instr;
  const size_t per_u = (size_t)2 * nslots * 16 * 256;
  const size_t budget = 150 * 1024;
  if (per_u > budget) return false;
  const bool u2 = 2 * per_u <= 100 * 1024;  // two vectors per decode while two CTAs still fit an SM
  if (dtype == NB_F32) return u2 ? prog_dispatch<float, 2>(kind, mode, P, G, a, st) : prog_dispatch<float, 1>(kind, mode, P, G, a, st);
  return u2 ? prog_dispatch<double, 2>(kind, mode, P, G, a, st) : prog_dispatch<double, 1>(kind, mode, P, G, a, st);
}
