// Lean fast path of the broadcast family: single-primitive scalar functions from a hot set, with the
// opcode and the set of wanted partials fixed at COMPILE time.
//
// Why (ncu, profiles/r1_ncu_c3_f32_*_v0): the general kernels (runtime opcode switch, runtime operand
// masks, out-of-line catalogue) use 78-80 registers, reach 37 % occupancy, spend 43 % of the issue
// slots on bookkeeping and stop at 25-43 % of the HBM peak.  Here the compiler knows which operands
// the value / partials need (nb_partial_needs is constexpr for a fixed OP and MODE), so unused operand
// streams do not exist, and every thread issues the 128-bit loads of U = 4 independent vectors before
// it computes (memory-level parallelism instead of occupancy).
//
// Included by bcast.cu inside its anonymous namespace (uses MRP, Opnd, OutS, VecOf, warp_sum ...).
#pragma once

enum { LM_FWD = 0, LM_PULL_A = 1, LM_PULL_B = 2, LM_PULL_AB = 3 };

__host__ __device__ constexpr bool lean_hot_op(int op) {
  return op == NB_OP_IDENTITY || op == NB_OP_NEG || op == NB_OP_ABS2 || op == NB_OP_EXP || op == NB_OP_LOG ||
         op == NB_OP_TANH || op == NB_OP_LOGISTIC || op == NB_OP_SQRT || op == NB_OP_INV || op == NB_OP_ADD ||
         op == NB_OP_SUB || op == NB_OP_MUL || op == NB_OP_DIV;
}
__host__ __device__ constexpr bool lean_is_binary(int op) { return nb_op_is_binary(op); }

// static operand needs: bit0 a, bit1 b, bit2 y (saved forward value)
template <int OP, int MODE>
struct LeanNeeds {
  static constexpr int pull = (MODE & 1 ? nb_partial_needs(OP, 0) : 0) | (MODE & 2 ? nb_partial_needs(OP, 1) : 0);
  static constexpr int value = MODE == LM_FWD ? (lean_is_binary(OP) ? 3 : 1) : pull;
};

template <typename T, int V>
__device__ __forceinline__ void lean_ld(const Opnd& o, int64_t r, int64_t c, T (&x)[V]) {
  const T* p = (const T*)o.p + c * o.cs;
  if constexpr (V == 1) {  // scalar form: tail elements of a flat array, 0-dim operands
    x[0] = __ldg(o.rs ? p + r : p);
    return;
  }
  if (o.rs) {
    using VT = typename VecOf<T>::type;
    const VT v = __ldg(reinterpret_cast<const VT*>(p + r));
    const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int i = 0; i < V; ++i) x[i] = e[i];
  } else {
    const T s = __ldg(p);
#pragma unroll
    for (int i = 0; i < V; ++i) x[i] = s;
  }
}

struct LeanNoExtra {};

template <typename T, int V, int OP, int MODE>
struct LeanPack {
  using Extra = LeanNoExtra;
  using Scalar = LeanPack<T, 1, OP, MODE>;  // the same op on single elements (flat tails, 0-dim operands)
  static constexpr bool HAS_TAIL = true;
  static constexpr int NOUT = MODE == LM_PULL_AB ? 2 : 1;
  static constexpr int U = 4;  // independent vectors in flight per thread
  static constexpr int NEED = LeanNeeds<OP, MODE>::value;
  T a[(NEED & 1) ? V : 1], b[(NEED & 2) ? V : 1], y[(NEED & 4) ? V : 1], yb[MODE != LM_FWD ? V : 1];

  template <int N>
  static __device__ __forceinline__ void compute_many(const MRP&, const Extra&, const LeanPack (&x)[N], T (&o)[N][2][V]) {
#pragma unroll
    for (int u = 0; u < N; ++u) x[u].compute(o[u]);
  }

  // recompute flag: the partial needs y but no saved forward value was passed -> needs a and b
  __device__ __forceinline__ void load(const MRP& P, const Extra&, int64_t r, int64_t c) {
    if constexpr ((NEED & 1) != 0) {
      if (P.lean_mask & 1) lean_ld<T, V>(P.arg[0], r, c, a);
      else {
#pragma unroll
        for (int i = 0; i < V; ++i) a[i] = T(P.cval[0]);
      }
    }
    if constexpr ((NEED & 2) != 0) {
      if (P.lean_mask & 2) lean_ld<T, V>(P.arg[1], r, c, b);
      else {
#pragma unroll
        for (int i = 0; i < V; ++i) b[i] = T(P.cval[1]);
      }
    }
    if constexpr ((NEED & 4) != 0) lean_ld<T, V>(P.ys, r, c, y);
    if constexpr (MODE != LM_FWD) {
      if (P.has_ybar) lean_ld<T, V>(P.ybar, r, c, yb);
      else {
#pragma unroll
        for (int i = 0; i < V; ++i) yb[i] = T(1);
      }
    }
  }

  // o[0] = value (forward) or the first wanted sensitivity; o[1] = the second one (MODE == PULL_AB)
  __device__ __forceinline__ void compute(T (&o)[2][V]) const {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const T av = (NEED & 1) ? a[(NEED & 1) ? i : 0] : T(0);
      const T bv = (NEED & 2) ? b[(NEED & 2) ? i : 0] : T(0);
      const T yv = (NEED & 4) ? y[(NEED & 4) ? i : 0] : T(0);
      if constexpr (MODE == LM_FWD) {
        o[0][i] = nb_eval_full<T>(OP, av, bv);
        o[1][i] = T(0);
      } else if constexpr (!lean_is_binary(OP)) {
        o[0][i] = yb[i] * nb_dunary_full<T>(OP, av, yv);
        o[1][i] = T(0);
      } else if constexpr (MODE == LM_PULL_A) {
        o[0][i] = yb[i] * nb_dbinary<T>(OP, 0, av, bv, yv);
        o[1][i] = T(0);
      } else if constexpr (MODE == LM_PULL_B) {
        o[0][i] = yb[i] * nb_dbinary<T>(OP, 1, av, bv, yv);
        o[1][i] = T(0);
      } else {
        o[0][i] = yb[i] * nb_dbinary<T>(OP, 0, av, bv, yv);
        o[1][i] = yb[i] * nb_dbinary<T>(OP, 1, av, bv, yv);
      }
    }
  }
};

template <typename T, int V>
__device__ __forceinline__ void lean_store(const OutS& out, int64_t off, const T (&v)[V], T scale) {
  using VT = typename VecOf<T>::type;
  T* p = (T*)out.p + off;
  if constexpr (V == 1) {
    *p = out.acc ? *p + v[0] * scale : v[0] * scale;
    return;
  }
  VT w;
  T* e = reinterpret_cast<T*>(&w);
  if (out.acc) {
    w = *reinterpret_cast<const VT*>(p);
#pragma unroll
    for (int i = 0; i < V; ++i) e[i] += v[i] * scale;
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) e[i] = v[i] * scale;
  }
  *reinterpret_cast<VT*>(p) = w;
}

// ---- flat: cols == 1, N = P.rows elements, N % V == 0 handled by the caller (tail launch) ----------
template <typename T, typename Pk>
__global__ void __launch_bounds__(256) k_flat_lean(const MRP P, int64_t nvec, const typename Pk::Extra X) {
  constexpr int V = VecOf<T>::V;
  constexpr int U = Pk::U;
  constexpr int NOUT = Pk::NOUT;
  __shared__ T red[32];
  T accs[NOUT];
#pragma unroll
  for (int k = 0; k < NOUT; ++k) accs[k] = T(0);
  const int64_t S = (int64_t)gridDim.x * blockDim.x;
  const T scale = T(P.scale);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * S < nvec; i += U * S) {
    Pk x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) x[u].load(P, X, (i + u * S) * V, 0);
    T oo[U][2][V];
    Pk::template compute_many<U>(P, X, x, oo);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      T (&o)[2][V] = oo[u];
#pragma unroll
      for (int k = 0; k < NOUT; ++k) {
        if (P.out[k].mode == OM_ELEM) lean_store<T, V>(P.out[k], (i + u * S) * V, o[k], scale);
        else {
#pragma unroll
          for (int v = 0; v < V; ++v) accs[k] += o[k][v];
        }
      }
    }
  }
  for (; i < nvec; i += S) {
    Pk x1[1];
    x1[0].load(P, X, i * V, 0);
    T o1[1][2][V];
    Pk::template compute_many<1>(P, X, x1, o1);
    T (&o)[2][V] = o1[0];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      if (P.out[k].mode == OM_ELEM) lean_store<T, V>(P.out[k], i * V, o[k], scale);
      else {
#pragma unroll
        for (int v = 0; v < V; ++v) accs[k] += o[k][v];
      }
    }
  }
  if constexpr (Pk::HAS_TAIL) {
    // elements past the last whole vector (any length, incl. 0-dim operands: nvec == 0, one element)
    using P1 = typename Pk::Scalar;
    for (int64_t e = nvec * V + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < P.rows; e += S) {
      P1 x1[1];
      x1[0].load(P, X, e, 0);
      T o1[1][2][1];
      P1::template compute_many<1>(P, X, x1, o1);
#pragma unroll
      for (int k = 0; k < NOUT; ++k) {
        if (P.out[k].mode == OM_ELEM) lean_store<T, 1>(P.out[k], e, o1[0][k], scale);
        else accs[k] += o1[0][k][0];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NOUT; ++k) {
    if (P.out[k].mode != OM_ALL) continue;
    T s = block_sum<T>(accs[k], red);
    if (threadIdx.x == 0) {
      if (P.out[k].partial) ((T*)P.out[k].partial)[blockIdx.x] = s;
      else {  // single-CTA launch: no second phase, the result is final here
        T* dst = (T*)P.out[k].p;
        *dst = P.out[k].acc ? *dst + s * scale : s * scale;
      }
    }
  }
}

// ---- tall: rows % V == 0, threads along rows, CTAs tile rows x column slabs --------------------------
template <typename T, typename Pk>
__global__ void __launch_bounds__(256) k_tall_lean(const MRP P, int64_t cols_per_slab, const typename Pk::Extra X) {
  constexpr int V = VecOf<T>::V;
  constexpr int U = Pk::U;
  constexpr int NOUT = Pk::NOUT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
  const int64_t r0 = ((int64_t)blockIdx.x * TX + tx) * V;
  const bool row_ok = r0 < P.rows;
  const int64_t c0 = (int64_t)blockIdx.y * cols_per_slab;
  const int64_t c1 = min(P.cols, c0 + cols_per_slab);
  const T scale = T(P.scale);
  T acc[NOUT][V];
#pragma unroll
  for (int k = 0; k < NOUT; ++k)
#pragma unroll
    for (int v = 0; v < V; ++v) acc[k][v] = T(0);
  if (row_ok) {
    int64_t c = c0 + ty;
    for (; c + (int64_t)(U - 1) * TY < c1; c += (int64_t)U * TY) {
      Pk x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) x[u].load(P, X, r0, c + (int64_t)u * TY);
      T oo[U][2][V];
      Pk::template compute_many<U>(P, X, x, oo);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T (&o)[2][V] = oo[u];
        const int64_t cc = c + (int64_t)u * TY;
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
          if (P.out[k].mode == OM_ELEM) lean_store<T, V>(P.out[k], r0 * P.out[k].rs + cc * P.out[k].cs, o[k], scale);
          else {
#pragma unroll
            for (int v = 0; v < V; ++v) acc[k][v] += o[k][v];
          }
        }
      }
    }
    for (; c < c1; c += TY) {
      Pk x1[1];
      x1[0].load(P, X, r0, c);
      T o1[1][2][V];
      Pk::template compute_many<1>(P, X, x1, o1);
      T (&o)[2][V] = o1[0];
#pragma unroll
      for (int k = 0; k < NOUT; ++k) {
        if (P.out[k].mode == OM_ELEM) lean_store<T, V>(P.out[k], r0 * P.out[k].rs + c * P.out[k].cs, o[k], scale);
        else {
#pragma unroll
          for (int v = 0; v < V; ++v) acc[k][v] += o[k][v];
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NOUT; ++k) {
    const int mode = P.out[k].mode;
    if (mode == OM_KEEP_ROWS) {
      __syncthreads();
#pragma unroll
      for (int v = 0; v < V; ++v) smem[(ty * TX + tx) * V + v] = acc[k][v];
      __syncthreads();
      if (ty == 0 && row_ok) {
        T s[V];
#pragma unroll
        for (int v = 0; v < V; ++v) s[v] = T(0);
        for (int y = 0; y < TY; ++y)
#pragma unroll
          for (int v = 0; v < V; ++v) s[v] += smem[(y * TX + tx) * V + v];
        T* dst = (T*)P.out[k].partial + (int64_t)blockIdx.y * P.rows + r0;
#pragma unroll
        for (int v = 0; v < V; ++v) dst[v] = s[v];
      }
    } else if (mode == OM_ALL) {
      T s = T(0);
#pragma unroll
      for (int v = 0; v < V; ++v) s += acc[k][v];
      const int lane = (ty * TX + tx) & 31, w = (ty * TX + tx) >> 5;
      s = warp_sum(s);
      __syncthreads();
      if (lane == 0) smem[w] = s;
      __syncthreads();
      if (w == 0) {
        s = lane < ((TX * TY + 31) >> 5) ? smem[lane] : T(0);
        s = warp_sum(s);
        if (lane == 0) ((T*)P.out[k].partial)[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
      }
    }
  }
}

// ---- wcol: one warp per column, rows % V == 0; outputs ELEM / KEEP_COLS (reduce over the rows) ------
template <typename T, typename Pk>
__global__ void __launch_bounds__(256) k_wcol_lean(const MRP P, const typename Pk::Extra X) {
  constexpr int V = VecOf<T>::V;
  constexpr int U = Pk::U;
  constexpr int NOUT = Pk::NOUT;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const T scale = T(P.scale);
  constexpr int64_t STEP = 32 * V;
  for (int64_t c = warp; c < P.cols; c += nwarps) {
    T acc[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) acc[k] = T(0);
    int64_t r = (int64_t)lane * V;
    for (; r + (U - 1) * STEP < P.rows; r += U * STEP) {
      Pk x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) x[u].load(P, X, r + u * STEP, c);
      T oo[U][2][V];
      Pk::template compute_many<U>(P, X, x, oo);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T (&o)[2][V] = oo[u];
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
          if (P.out[k].mode == OM_ELEM)
            lean_store<T, V>(P.out[k], (r + u * STEP) * P.out[k].rs + c * P.out[k].cs, o[k], scale);
          else {
#pragma unroll
            for (int v = 0; v < V; ++v) acc[k] += o[k][v];
          }
        }
      }
    }
    for (; r < P.rows; r += STEP) {
      Pk x1[1];
      x1[0].load(P, X, r, c);
      T o1[1][2][V];
      Pk::template compute_many<1>(P, X, x1, o1);
      T (&o)[2][V] = o1[0];
#pragma unroll
      for (int k = 0; k < NOUT; ++k) {
        if (P.out[k].mode == OM_ELEM) lean_store<T, V>(P.out[k], r * P.out[k].rs + c * P.out[k].cs, o[k], scale);
        else {
#pragma unroll
          for (int v = 0; v < V; ++v) acc[k] += o[k][v];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      if (P.out[k].mode != OM_KEEP_COLS) continue;
      const T s = warp_sum(acc[k]);
      if (lane == 0) {
        T* dst = (T*)P.out[k].p + c * P.out[k].cs;
        *dst = P.out[k].acc ? *dst + s * scale : s * scale;
      }
    }
  }
}

template <typename T, int OP, int MODE>
void lean_launch_wcol(const MRP& P, unsigned blocks, cudaStream_t st) {
  k_wcol_lean<T, LeanPack<T, VecOf<T>::V, OP, MODE>><<<blocks, 256, 0, st>>>(P, LeanNoExtra{});
}

// ---- dispatch ------------------------------------------------------------------------------------
template <typename T, int OP, int MODE>
void lean_launch_flat(const MRP& P, int64_t nvec, unsigned blocks, cudaStream_t st) {
  k_flat_lean<T, LeanPack<T, VecOf<T>::V, OP, MODE>><<<blocks, 256, 0, st>>>(P, nvec, LeanNoExtra{});
}
template <typename T, int OP, int MODE>
void lean_launch_tall(const MRP& P, dim3 grid, dim3 block, size_t smem, int64_t cps, cudaStream_t st) {
  k_tall_lean<T, LeanPack<T, VecOf<T>::V, OP, MODE>><<<grid, block, smem, st>>>(P, cps, LeanNoExtra{});
}

// F is a functor template taking <OP, MODE>; returns false when (op, mode) is not a lean combination
template <typename T, typename F>
bool lean_dispatch(int op, int mode, F&& f) {
#define NB_LEAN_UN(OPC)                                              \
  case OPC:                                                          \
    if (mode == LM_FWD) { f.template operator()<OPC, LM_FWD>(); return true; }       \
    if (mode == LM_PULL_A) { f.template operator()<OPC, LM_PULL_A>(); return true; } \
    return false;
#define NB_LEAN_BIN(OPC)                                             \
  case OPC:                                                          \
    if (mode == LM_FWD) { f.template operator()<OPC, LM_FWD>(); return true; }         \
    if (mode == LM_PULL_A) { f.template operator()<OPC, LM_PULL_A>(); return true; }   \
    if (mode == LM_PULL_B) { f.template operator()<OPC, LM_PULL_B>(); return true; }   \
    if (mode == LM_PULL_AB) { f.template operator()<OPC, LM_PULL_AB>(); return true; } \
    return false;
  switch (op) {
    NB_LEAN_UN(NB_OP_IDENTITY) NB_LEAN_UN(NB_OP_NEG) NB_LEAN_UN(NB_OP_ABS2) NB_LEAN_UN(NB_OP_EXP)
    NB_LEAN_UN(NB_OP_LOG) NB_LEAN_UN(NB_OP_TANH) NB_LEAN_UN(NB_OP_LOGISTIC) NB_LEAN_UN(NB_OP_SQRT)
    NB_LEAN_UN(NB_OP_INV) NB_LEAN_BIN(NB_OP_ADD) NB_LEAN_BIN(NB_OP_SUB) NB_LEAN_BIN(NB_OP_MUL)
    NB_LEAN_BIN(NB_OP_DIV)
    default: return false;
  }
#undef NB_LEAN_UN
#undef NB_LEAN_BIN
}
