// Kernel families (1) and (2): broadcast forward, fused broadcast pullback + unbroadcast, and
// map-reductions (sum / sum(f,.) / mean / dot / mapreduce(f,+,.;dims)).
//
// Reference behaviour being replaced (paths relative to the reference checkout):
//   forward   Branch(broadcast, (f, args...))            functional.jl:48-51, core.jl:84-95
//   pullback  ∇(broadcast, Arg{N}, ...) -> broadcastsum   functional.jl:59-95: one full pass PER
//             operand, a full-size temporary + sum! when the operand was broadcast, then a separate
//             add!! pass when the slot is already assigned (core.jl:149-156, 203-205)
//   reductions reducedim.jl:4-72, reduce.jl:4-12, ChainRules sum/mean/dot
//
// Here ONE pass over the broadcast index space produces every wanted operand sensitivity, the
// singleton-dimension sum is fused into the same pass (register / warp-shuffle / shared-memory
// partials, then a deterministic second phase — no floating-point atomics), and `+=` is fused.
//
// The index space is collapsed to (rows, cols) — rows is the contiguous (column-major) group — and
// one of four thread mappings is chosen:
//   k_flat   cols == 1          grid-stride, 128-bit vectors, 4 vectors in flight per thread
//   k_tall   rows  > 32         threads along rows (vectorised), CTAs tile rows x column-slabs;
//                               handles outputs that keep rows (reduce over cols), e.g. bias grads
//   k_wcol   rows  > 32         one warp per column; handles outputs that keep cols (reduce over
//                               rows), e.g. `f ./ sum(f, dims=1)` with many rows
//   k_short  rows <= 32         lanes = (row, column) pairs (10 x B output layers)
//   k_nd     >2 collapsed dims  generic fallback, one thread per output element
#include "ops.cuh"
#include "bcast_types.cuh"

int32_t nb_gemm_skinny(nb_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, double alpha, const void* A,
                       int64_t lda, const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int32_t dtype,
                       const nb_gemm_epilogue* epi, int32_t math_mode);  // gemm_skinny.cu

namespace {
using namespace nbb;

template <typename T, int V>
__device__ __forceinline__ void load_vals(const void* base, int64_t off, bool contig, int nvalid,
                                          bool vec_ok, T (&x)[V]) {
  const T* p = (const T*)base + off;
  if (!contig) {
    T v = __ldg(p);
#pragma unroll
    for (int i = 0; i < V; ++i) x[i] = v;
  } else if (V == VecOf<T>::V && vec_ok && nvalid == V) {
    using VT = typename VecOf<T>::type;
    VT v = *reinterpret_cast<const VT*>(p);
    const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int i = 0; i < V; ++i) x[i] = e[i];
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) x[i] = i < nvalid ? p[i] : T(0);
  }
}

template <typename T, int V>
__device__ __forceinline__ void store_vals(void* base, int64_t off, int nvalid, bool vec_ok, bool acc,
                                           const T (&o)[V]) {
  T* p = (T*)base + off;
  if (V == VecOf<T>::V && vec_ok && nvalid == V) {
    using VT = typename VecOf<T>::type;
    VT v;
    T* e = reinterpret_cast<T*>(&v);
    if (acc) {
      v = *reinterpret_cast<const VT*>(p);
#pragma unroll
      for (int i = 0; i < V; ++i) e[i] += o[i];
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) e[i] = o[i];
    }
    *reinterpret_cast<VT*>(p) = v;
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i)
      if (i < nvalid) p[i] = acc ? p[i] + o[i] : o[i];
  }
}

// ---- evaluators: produce o[k][v] for V consecutive rows starting at (r, c) -------------------
template <typename T, int V>
struct FastEv {
  static constexpr int MAXOUT = 2;
  static __device__ __forceinline__ void run(const MRP& P, const nb_prog&, int64_t r, int64_t c,
                                             int nvalid, T (&o)[MAXOUT][V]) {
    T a[V], b[V], yb[V], ys[V];
    const bool vok = P.vec_ok;
#pragma unroll
    for (int i = 0; i < V; ++i) { a[i] = T(P.cval[0]); b[i] = T(P.cval[1]); yb[i] = T(1); ys[i] = T(0); }
    if (P.load_mask & 1) load_vals<T, V>(P.arg[0].p, r * P.arg[0].rs + c * P.arg[0].cs, P.arg[0].rs != 0, nvalid, vok, a);
    if (P.load_mask & 2) load_vals<T, V>(P.arg[1].p, r * P.arg[1].rs + c * P.arg[1].cs, P.arg[1].rs != 0, nvalid, vok, b);
    const int op = P.op;
    if (!P.pullback) {
#pragma unroll
      for (int i = 0; i < V; ++i) o[0][i] = nb_eval<T>(op, a[i], b[i]);
      return;
    }
    if (P.has_ybar) load_vals<T, V>(P.ybar.p, r * P.ybar.rs + c * P.ybar.cs, P.ybar.rs != 0, nvalid, vok, yb);
    if (P.need_y) {
      if (P.has_y) {
        load_vals<T, V>(P.ys.p, r * P.ys.rs + c * P.ys.cs, P.ys.rs != 0, nvalid, vok, ys);
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) ys[i] = nb_eval<T>(op, a[i], b[i]);
      }
    }
    const bool unary = nb_op_is_unary(op);
#pragma unroll
    for (int k = 0; k < MAXOUT; ++k) {
      if (k < P.nout) {
        const int w = P.out[k].wrt;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          T d = T(0);
          if (unary) {
            if (w & 1) d = nb_dunary<T>(op, a[i], ys[i]);
          } else {
            if (w & 1) d += nb_dbinary<T>(op, 0, a[i], b[i], ys[i]);
            if (w & 2) d += nb_dbinary<T>(op, 1, a[i], b[i], ys[i]);
          }
          o[k][i] = yb[i] * d;
        }
      }
    }
  }
};

template <typename T, int V>
struct GenEv {
  static constexpr int MAXOUT = NB_MAX_OUTS;
  static __device__ __forceinline__ void run(const MRP& P, const nb_prog& G, int64_t r, int64_t c,
                                             int nvalid, T (&o)[MAXOUT][V]) {
    T x[NB_MAX_ARGS][V], yb[V];
    const bool vok = P.vec_ok;
#pragma unroll
    for (int k = 0; k < NB_MAX_ARGS; ++k)
      if (k < P.nargs)
        load_vals<T, V>(P.arg[k].p, r * P.arg[k].rs + c * P.arg[k].cs, P.arg[k].rs != 0, nvalid, vok, x[k]);
#pragma unroll
    for (int i = 0; i < V; ++i) yb[i] = T(1);
    if (P.pullback && P.has_ybar)
      load_vals<T, V>(P.ybar.p, r * P.ybar.rs + c * P.ybar.cs, P.ybar.rs != 0, nvalid, vok, yb);
    // V-wide interpretation: one opcode decode per V elements (ops.cuh)
    T vals[NB_MAX_SLOTS][V];
#pragma unroll
    for (int k = 0; k < NB_MAX_ARGS; ++k)
      if (k < P.nargs) {
#pragma unroll
        for (int i = 0; i < V; ++i) vals[k][i] = x[k][i];
      }
    nb_prog_eval_vec<T, V>(G, vals);
    if (!P.pullback) {
      const int last = G.n_instr ? G.nargs + G.n_instr - 1 : 0;
#pragma unroll
      for (int i = 0; i < V; ++i) o[0][i] = vals[last][i];
    } else {
      T adj[NB_MAX_SLOTS][V];
      nb_prog_grad_vec<T, V>(G, vals, adj);
#pragma unroll
      for (int k = 0; k < MAXOUT; ++k)
        if (k < P.nout) {
          const int w = P.out[k].wrt;
#pragma unroll
          for (int i = 0; i < V; ++i) o[k][i] = yb[i] * adj[w][i];
        }
    }
  }
};

// Composite programs, shared-memory interpreter.  The previous evaluator (GenEv above, kept for programs whose
// slots do not fit) holds the slot values of a thread in local memory and decodes with compare chains: ncu
// (profiles/: k_tall<float, GenEv> on (a,b)->tanh(a+b)) counted 142 thread-instructions per ELEMENT for a
// two-instruction program, a quarter of them ISETP/BRA, at 10-12 % of the HBM peak.  Here
//   * every thread owns a 16-byte column per slot in dynamic shared memory (slot s of thread t at
//     base[(s*NT + t)*V]: conflict-free 128-bit accesses, no local memory),
//   * the opcode switch is a dense jump table (nb_prog::hop), decoded once per V elements,
//   * the reverse sweep reads only what each partial needs and stores — instead of zero-filling and
//     accumulating — the first contribution to every adjoint (nb_prog::rflags, computed once on the host).
template <typename T, int V>
struct GenEvS {
  static constexpr int MAXOUT = NB_MAX_OUTS;
  using VT = typename std::conditional<V == 1, T, typename VecOf<T>::type>::type;

  static __device__ __forceinline__ void run(const MRP& P, const nb_prog& G, int64_t r, int64_t c,
                                             int nvalid, T (&o)[MAXOUT][V]) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
    VT* base = reinterpret_cast<VT*>(smem_raw + P.slot_off) + tid;
    const int na = G.nargs, n = G.n_instr, nslots = na + n;
    auto ld = [&](int s, T (&x)[V]) {
      const VT v = base[(size_t)s * NT];
      const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int i = 0; i < V; ++i) x[i] = e[i];
    };
    auto st = [&](int s, const T (&x)[V]) {
      VT v;
      T* e = reinterpret_cast<T*>(&v);
#pragma unroll
      for (int i = 0; i < V; ++i) e[i] = x[i];
      base[(size_t)s * NT] = v;
    };
    auto get = [&](int idx, T (&x)[V]) {  // slot value or constant
      if (idx >= 0) ld(idx, x);
      else {
        const T cst = T(G.consts[-idx - 1]);
#pragma unroll
        for (int i = 0; i < V; ++i) x[i] = cst;
      }
    };
    const bool vok = P.vec_ok;
    T last[V];
    for (int k = 0; k < na; ++k) {
      load_vals<T, V>(P.arg[k].p, r * P.arg[k].rs + c * P.arg[k].cs, P.arg[k].rs != 0, nvalid, vok, last);
      st(k, last);
    }
    // forward (in a pullback only the values some computed partial reads: nb_prog::fwd_skip)
    const uint32_t skip = P.pullback ? G.fwd_skip : 0u;
    for (int i = 0; i < n; ++i) {
      if (skip >> i & 1u) continue;
      T a[V], b[V];
      get(G.a[i], a);
      if (G.hop[i] >= NB_HOP_ADD && G.hop[i] != NB_HOP_COLD_UN) get(G.b[i], b);
      else {
#pragma unroll
        for (int v = 0; v < V; ++v) b[v] = T(0);
      }
      nb_eval_hop<T, V>(G.hop[i], G.op[i], a, b, last);
      st(na + i, last);
    }
    if (!P.pullback) {
#pragma unroll
      for (int i = 0; i < V; ++i) o[0][i] = last[i];
      return;
    }
    // reverse: adjoint of slot s lives in slot nslots + s
    for (int i = n - 1; i >= 0; --i) {
      const int fl = G.rflags[i];
      if (!(fl & NB_RF_LIVE)) continue;
      const int ia = G.a[i], ib = G.b[i];
      T g[V], a[V], b[V], y[V], da[V], db[V];
      if (i == n - 1) {
#pragma unroll
        for (int v = 0; v < V; ++v) g[v] = T(1);
      } else {
        ld(nslots + na + i, g);
      }
#pragma unroll
      for (int v = 0; v < V; ++v) { a[v] = T(0); b[v] = T(0); y[v] = T(0); }
      if (fl & NB_RF_NEED_A) get(ia, a);
      if (fl & NB_RF_NEED_B) get(ib, b);
      if (fl & NB_RF_NEED_Y) ld(na + i, y);
      nb_partials_hop<T, V>(G.hop[i], G.op[i], a, b, y, da, db);
      if (fl & NB_RF_SLOT_A) {
        const bool same = (fl & NB_RF_SLOT_B) && ib == ia;  // f(x, x): both partials go to one adjoint
        T t[V];
#pragma unroll
        for (int v = 0; v < V; ++v) t[v] = nb_gmul(g[v], same ? da[v] + db[v] : da[v], G.guard != 0);
        if (!(fl & NB_RF_FIRST_A)) {
          T old[V];
          ld(nslots + ia, old);
#pragma unroll
          for (int v = 0; v < V; ++v) t[v] += old[v];
        }
        st(nslots + ia, t);
      }
      if ((fl & NB_RF_SLOT_B) && ib != ia) {
        T t[V];
#pragma unroll
        for (int v = 0; v < V; ++v) t[v] = nb_gmul(g[v], db[v], G.guard != 0);
        if (!(fl & NB_RF_FIRST_B)) {
          T old[V];
          ld(nslots + ib, old);
#pragma unroll
          for (int v = 0; v < V; ++v) t[v] += old[v];
        }
        st(nslots + ib, t);
      }
    }
    T yb[V];
#pragma unroll
    for (int i = 0; i < V; ++i) yb[i] = T(1);
    if (P.has_ybar) load_vals<T, V>(P.ybar.p, r * P.ybar.rs + c * P.ybar.cs, P.ybar.rs != 0, nvalid, vok, yb);
#pragma unroll
    for (int k = 0; k < MAXOUT; ++k)
      if (k < P.nout) {
        const int w = P.out[k].wrt;
        if (n == 0) {
#pragma unroll
          for (int i = 0; i < V; ++i) o[k][i] = yb[i];
        } else if (G.arg_live & (1u << w)) {
          T adj[V];
          ld(nslots + w, adj);
#pragma unroll
          for (int i = 0; i < V; ++i) o[k][i] = yb[i] * adj[i];
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) o[k][i] = T(0);
        }
      }
  }
};

// bytes of slot storage one thread needs (values + adjoints)
template <typename T, int V>
static inline size_t gen_slot_bytes_per_thread(const nb_prog& G) { return (size_t)2 * (G.nargs + G.n_instr) * V * sizeof(T); }
constexpr size_t GEN_SMEM_LIMIT = 160 * 1024;

// ---- k_flat: cols == 1 ---------------------------------------------------------------------
template <typename T, template <typename, int> class EvT>
__global__ void __launch_bounds__(256) k_flat(const MRP P, const nb_prog G) {
  constexpr int V = VecOf<T>::V;
  using Ev = EvT<T, V>;
  constexpr int U = Ev::MAXOUT > 2 ? 1 : 4;
  __shared__ T red[32];
  T accs[Ev::MAXOUT];
#pragma unroll
  for (int k = 0; k < Ev::MAXOUT; ++k) accs[k] = T(0);
  const int64_t N = P.rows;
  const int64_t nvec = (N + V - 1) / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const T scale = T(P.scale);
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * U) {
    T o[U][Ev::MAXOUT][V];
    int nv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      const int64_t r = i * V;
      nv[u] = i < nvec ? (int)min((int64_t)V, N - r) : 0;
      if (nv[u] > 0) Ev::run(P, G, r, 0, nv[u], o[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (nv[u] <= 0) continue;
      const int64_t r = (i0 + u * stride) * V;
#pragma unroll
      for (int k = 0; k < Ev::MAXOUT; ++k) {
        if (k >= P.nout || !P.out[k].active) continue;
        if (P.out[k].mode == OM_ELEM) {
          T t[V];
#pragma unroll
          for (int v = 0; v < V; ++v) t[v] = o[u][k][v] * scale;
          store_vals<T, V>(P.out[k].p, r, nv[u], P.vec_ok, P.out[k].acc, t);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (v < nv[u]) accs[k] += o[u][k][v];
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < Ev::MAXOUT; ++k) {
    if (k >= P.nout || !P.out[k].active || P.out[k].mode != OM_ALL) continue;
    T s = block_sum<T>(accs[k], red);
    if (threadIdx.x == 0) ((T*)P.out[k].partial)[blockIdx.x] = s;
  }
}

// ---- k_tall: rows > 32, outputs ELEM / KEEP_ROWS / ALL --------------------------------------
// blockDim = (TX, TY); grid = (row blocks, column slabs); cols_per_slab columns per CTA
template <typename T, template <typename, int> class EvT>
__global__ void __launch_bounds__(256) k_tall(const MRP P, const nb_prog G, int64_t cols_per_slab) {
  constexpr int V = VecOf<T>::V;
  using Ev = EvT<T, V>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
  const int64_t r0 = ((int64_t)blockIdx.x * TX + tx) * V;
  const int nvalid = r0 < P.rows ? (int)min((int64_t)V, P.rows - r0) : 0;
  const int64_t c0 = (int64_t)blockIdx.y * cols_per_slab;
  const int64_t c1 = min(P.cols, c0 + cols_per_slab);
  const T scale = T(P.scale);
  T acc[Ev::MAXOUT][V];
#pragma unroll
  for (int k = 0; k < Ev::MAXOUT; ++k)
#pragma unroll
    for (int v = 0; v < V; ++v) acc[k][v] = T(0);
  if (nvalid > 0) {
    for (int64_t c = c0 + ty; c < c1; c += TY) {
      T o[Ev::MAXOUT][V];
      Ev::run(P, G, r0, c, nvalid, o);
#pragma unroll
      for (int k = 0; k < Ev::MAXOUT; ++k) {
        if (k >= P.nout || !P.out[k].active) continue;
        if (P.out[k].mode == OM_ELEM) {
          T t[V];
#pragma unroll
          for (int v = 0; v < V; ++v) t[v] = o[k][v] * scale;
          store_vals<T, V>(P.out[k].p, r0 * P.out[k].rs + c * P.out[k].cs, nvalid, P.vec_ok,
                           P.out[k].acc, t);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) acc[k][v] += o[k][v];
        }
      }
    }
  }
  // reductions
#pragma unroll
  for (int k = 0; k < Ev::MAXOUT; ++k) {
    if (k >= P.nout || !P.out[k].active) continue;
    const int mode = P.out[k].mode;
    if (mode == OM_KEEP_ROWS) {
      // sum over ty through shared memory, then one partial row-vector per slab
      __syncthreads();
#pragma unroll
      for (int v = 0; v < V; ++v) smem[(ty * TX + tx) * V + v] = acc[k][v];
      __syncthreads();
      if (ty == 0 && nvalid > 0) {
        T s[V];
#pragma unroll
        for (int v = 0; v < V; ++v) s[v] = T(0);
        for (int y = 0; y < TY; ++y)
#pragma unroll
          for (int v = 0; v < V; ++v) s[v] += smem[(y * TX + tx) * V + v];
        T* dst = (T*)P.out[k].partial + (int64_t)blockIdx.y * P.rows + r0;
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (v < nvalid) dst[v] = s[v];
      }
    } else if (mode == OM_ALL) {
      T s = T(0);
#pragma unroll
      for (int v = 0; v < V; ++v) s += acc[k][v];
      // flatten (tx,ty) for block_sum
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

// ---- k_wcol: one warp per column; outputs ELEM / KEEP_COLS -----------------------------------
template <typename T, template <typename, int> class EvT>
__global__ void __launch_bounds__(256) k_wcol(const MRP P, const nb_prog G) {
  constexpr int V = VecOf<T>::V;
  using Ev = EvT<T, V>;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const T scale = T(P.scale);
  for (int64_t c = warp; c < P.cols; c += nwarps) {
    T acc[Ev::MAXOUT];
#pragma unroll
    for (int k = 0; k < Ev::MAXOUT; ++k) acc[k] = T(0);
    for (int64_t r0 = (int64_t)lane * V; r0 < P.rows; r0 += 32 * V) {
      const int nvalid = (int)min((int64_t)V, P.rows - r0);
      T o[Ev::MAXOUT][V];
      Ev::run(P, G, r0, c, nvalid, o);
#pragma unroll
      for (int k = 0; k < Ev::MAXOUT; ++k) {
        if (k >= P.nout || !P.out[k].active) continue;
        if (P.out[k].mode == OM_ELEM) {
          T t[V];
#pragma unroll
          for (int v = 0; v < V; ++v) t[v] = o[k][v] * scale;
          store_vals<T, V>(P.out[k].p, r0 * P.out[k].rs + c * P.out[k].cs, nvalid, P.vec_ok,
                           P.out[k].acc, t);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (v < nvalid) acc[k] += o[k][v];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < Ev::MAXOUT; ++k) {
      if (k >= P.nout || !P.out[k].active || P.out[k].mode != OM_KEEP_COLS) continue;
      T s = warp_sum(acc[k]);
      if (lane == 0) {
        T* dst = (T*)P.out[k].p + c * P.out[k].cs;
        *dst = P.out[k].acc ? *dst + s * scale : s * scale;
      }
    }
  }
}

// ---- k_short: rows <= 32; all output modes ----------------------------------------------------
// A warp covers 32 / RP columns at a time, RP = the power of two >= rows: lane -> (row = lane % RP, column = lane / RP).
// (One thread per column left a 10 x 4096 array — the output layer of examples/mlp.jl — with 4096 threads on the whole
// part, each walking its column serially: 20-38 us of pure latency per launch for 40 K elements.)  Column sums
// (KEEP_COLS) are xor-shuffle trees inside the RP lanes, row sums (KEEP_ROWS) per-thread registers combined through
// shared memory in a fixed order, OM_ALL a block sum: deterministic.
// dynamic smem: blockDim.x + 32 accumulators (+ the interpreter's slots)
template <typename T, template <typename, int> class EvT>
__global__ void __launch_bounds__(128) k_short(const MRP P, const nb_prog G, int rp) {
  using Ev = EvT<T, 1>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const int tid = threadIdx.x, BT = blockDim.x;
  const int rows = (int)P.rows;
  const T scale = T(P.scale);
  const int r = tid & (rp - 1), cl = tid / rp, cpb = BT / rp;
  const bool rok = r < rows;
  T* red = smem + BT;
  T acc[Ev::MAXOUT];   // KEEP_ROWS: this thread's row; OM_ALL: everything it saw
#pragma unroll
  for (int k = 0; k < Ev::MAXOUT; ++k) acc[k] = T(0);
  const int64_t stride = (int64_t)gridDim.x * cpb;
  for (int64_t c0 = (int64_t)blockIdx.x * cpb; c0 < P.cols; c0 += stride) {   // (uniform trip count per block)
    const int64_t c = c0 + cl;
    const bool ok = rok && c < P.cols;
    T o[Ev::MAXOUT][1];
#pragma unroll
    for (int k = 0; k < Ev::MAXOUT; ++k) o[k][0] = T(0);
    if (ok) Ev::run(P, G, r, c, 1, o);
#pragma unroll
    for (int k = 0; k < Ev::MAXOUT; ++k) {
      if (k >= P.nout || !P.out[k].active) continue;
      const int mode = P.out[k].mode;
      T val = ok ? o[k][0] : T(0);
      if (mode == OM_ELEM) {
        if (ok) {
          T* dst = (T*)P.out[k].p + r * P.out[k].rs + c * P.out[k].cs;
          *dst = P.out[k].acc ? *dst + val * scale : val * scale;
        }
      } else if (mode == OM_KEEP_COLS) {
        for (int sft = rp >> 1; sft > 0; sft >>= 1) val += __shfl_xor_sync(0xffffffffu, val, sft);
        if (r == 0 && c < P.cols) {
          T* dst = (T*)P.out[k].p + c * P.out[k].cs;
          *dst = P.out[k].acc ? *dst + val * scale : val * scale;
        }
      } else {
        acc[k] += val;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < Ev::MAXOUT; ++k) {
    if (k >= P.nout || !P.out[k].active) continue;
    if (P.out[k].mode == OM_KEEP_ROWS) {
      __syncthreads();
      smem[tid] = acc[k];
      __syncthreads();
      if (tid < rows) {
        T s = T(0);
        for (int j = 0; j < cpb; ++j) s += smem[j * rp + tid];
        ((T*)P.out[k].partial)[(int64_t)blockIdx.x * rows + tid] = s;
      }
    } else if (P.out[k].mode == OM_ALL) {
      T s = block_sum<T>(acc[k], red);
      if (tid == 0) ((T*)P.out[k].partial)[blockIdx.x] = s;
    }
  }
}

// ---- phase 2: out[i] (+)= scale * sum_s partial[s*count + i] ---------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_finish(const T* __restrict__ partial, int64_t nparts,
                                                int64_t count, T* out, T scale, int acc) {
  __shared__ T red[32];
  if (count == 1) {
    T s = T(0);
    for (int64_t i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[i];
    s = block_sum<T>(s, red);
    if (threadIdx.x == 0) out[0] = acc ? out[0] + s * scale : s * scale;
    return;
  }
  // 32 outputs per block (coalesced along i), the 8 warps split the partials, fixed combination order
  __shared__ T sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + tx;
  T s = T(0);
  if (i < count)
    for (int64_t p = ty; p < nparts; p += 8) s += partial[p * count + i];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i < count) {
    T t = T(0);
#pragma unroll
    for (int y = 0; y < 8; ++y) t += sm[y][tx];
    out[i] = acc ? out[i] + t * scale : t * scale;
  }
}

// ---- k_nd: generic fallback for > 2 collapsed groups (one output per launch) -----------------
struct NDP {
  int32_t ng;                 // number of collapsed groups (<= 4)
  int64_t ext[NB_MAX_DIMS];   // group extents
  int32_t red[NB_MAX_DIMS];   // 1 if this group is reduced for the output
  int64_t ost[NB_MAX_DIMS];   // output strides
  int64_t ast[NB_MAX_ARGS][NB_MAX_DIMS];
  int64_t ybst[NB_MAX_DIMS], ysst[NB_MAX_DIMS];
  int64_t n_out, n_red;
  int32_t k;                  // which output of P
};

template <typename T, template <typename, int> class EvT>
__global__ void __launch_bounds__(256) k_nd(MRP P, const nb_prog G, const NDP D) {
  using Ev = EvT<T, 1>;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= D.n_out) return;
  // decode kept coordinates
  int64_t coord[NB_MAX_DIMS] = {0, 0, 0, 0};
  int64_t rem = e;
  for (int g = 0; g < D.ng; ++g)
    if (!D.red[g]) { coord[g] = rem % D.ext[g]; rem /= D.ext[g]; }
  T acc = T(0);
  for (int64_t j = 0; j < D.n_red; ++j) {
    int64_t rj = j;
    for (int g = 0; g < D.ng; ++g)
      if (D.red[g]) { coord[g] = rj % D.ext[g]; rj /= D.ext[g]; }
    // point every operand at this element: rs = 0 so load_vals reads p[off] with off = 0
    MRP Q = P;
    for (int a = 0; a < P.nargs; ++a) {
      int64_t off = 0;
      for (int g = 0; g < D.ng; ++g) off += coord[g] * D.ast[a][g];
      Q.arg[a].p = (const T*)P.arg[a].p + off; Q.arg[a].rs = 0; Q.arg[a].cs = 0;
    }
    {
      int64_t off = 0, off2 = 0;
      for (int g = 0; g < D.ng; ++g) { off += coord[g] * D.ybst[g]; off2 += coord[g] * D.ysst[g]; }
      if (P.has_ybar) { Q.ybar.p = (const T*)P.ybar.p + off; Q.ybar.rs = Q.ybar.cs = 0; }
      if (P.has_y) { Q.ys.p = (const T*)P.ys.p + off2; Q.ys.rs = Q.ys.cs = 0; }
    }
    T o[Ev::MAXOUT][1];
    Ev::run(Q, G, 0, 0, 1, o);
    T val = T(0);
#pragma unroll
    for (int k = 0; k < Ev::MAXOUT; ++k)
      if (k == D.k) val = o[k][0];
    acc += val;
  }
  int64_t ooff = 0;
  for (int g = 0; g < D.ng; ++g)
    if (!D.red[g]) ooff += coord[g] * D.ost[g];
  T* dst = (T*)P.out[D.k].p + ooff;
  const T s = acc * T(P.scale);
  *dst = P.out[D.k].acc ? *dst + s : s;
}

#include "bcast_lean.cuh"

// =================================================================================================
// host-side dispatch
// =================================================================================================
struct Part {          // one participant of the call (operand, ybar, y_saved or output)
  const nb_tensor* t;
  bool present[NB_MAX_DIMS];  // per original dim: spans the full extent (and the extent is > 1)
};

struct Call {
  const nb_prog* prog;
  int nargs;
  const nb_tensor* args;
  const nb_tensor* ybar;     // pullback only (may be NULL => ones)
  const nb_tensor* ysaved;   // nullable
  bool pullback;
  int nout;
  const nb_tensor* outs[NB_MAX_OUTS];
  int out_wrt[NB_MAX_OUTS];  // operand index each output differentiates (pullback)
  bool out_acc[NB_MAX_OUTS];
  double scale;
};

static inline int64_t ext_of(const nb_tensor* t, int d) { return d < t->ndim ? t->shape[d] : 1; }

template <typename T>
struct LeanKindF {
  int kind; const MRP& P; const LeanLaunch& a; cudaStream_t st;
  template <int OP, int MODE> void operator()() {
    if (kind == 0) lean_launch_flat<T, OP, MODE>(P, a.nvec, a.blocks, st);
    else if (kind == 1) lean_launch_tall<T, OP, MODE>(P, a.grid, a.block, a.smem, a.cps, st);
    else lean_launch_wcol<T, OP, MODE>(P, a.blocks, st);
  }
};

// hot primitives: kernels of this translation unit; the rest of the catalogue: bcast_lean_cold_*.cu
template <typename T>
bool lean_any(int kind, int mode, const MRP& P, const LeanLaunch& a, cudaStream_t st) {
  if (nb_is_fop(P.op))
    return sizeof(T) == 4 ? nb_lean_fused_f32(kind, P.op, mode, P, a, st) : nb_lean_fused_f64(kind, P.op, mode, P, a, st);
  if (lean_hot_op(P.op)) return lean_dispatch<T>(P.op, mode, LeanKindF<T>{kind, P, a, st});
  if (nb_op_is_unary(P.op))
    return sizeof(T) == 4 ? nb_lean_cold_un_f32(kind, P.op, mode, P, a, st) : nb_lean_cold_un_f64(kind, P.op, mode, P, a, st);
  return sizeof(T) == 4 ? nb_lean_cold_bin_f32(kind, P.op, mode, P, a, st) : nb_lean_cold_bin_f64(kind, P.op, mode, P, a, st);
}

// lean_mode: LM_* when the call qualifies for the compile-time-specialised kernels, -1 otherwise
template <typename T, template <typename, int> class Ev>
int32_t launch_all(nb_ctx* ctx, MRP& P, const nb_prog& G, int ngroups, int lean_mode, bool prog_lean = false) {
  constexpr int V = VecOf<T>::V;
  cudaStream_t st = ctx->stream;
  const int sms = ctx->num_sms;
  // shared-memory interpreter: slot storage behind the kernel's own dynamic shared memory
  constexpr bool IS_S = std::is_same<Ev<T, 1>, GenEvS<T, 1>>::value;
  auto with_slots = [&](auto kern, size_t own, int nthreads, int veff) -> size_t {
    if (!IS_S) return own;
    const size_t off = (own + 15) & ~(size_t)15;
    P.slot_off = (int32_t)off;
    const size_t total = off + (size_t)2 * (G.nargs + G.n_instr) * veff * sizeof(T) * nthreads;
    if (total > 32 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total);
    return total;
  };
  // lean kernels: single primitives (this file and bcast_lean_cold_*.cu) or whole programs (bcast_prog.cu)
  auto lean_try = [&](int kind, const LeanLaunch& LL) -> bool {
    if (prog_lean) return nb_lean_prog(sizeof(T) == 4 ? NB_F32 : NB_F64, kind, lean_mode, P, G, LL, st);
    return lean_any<T>(kind, lean_mode, P, LL, st);
  };
  struct Pending { void* partial; int64_t nparts, count; int k; };
  std::vector<Pending> pend;
  auto finish = [&]() -> int32_t {
    for (auto& q : pend) {
      int64_t blocks = q.count == 1 ? 1 : (q.count + 31) / 32;
      k_finish<T><<<(unsigned)blocks, 256, 0, st>>>((const T*)q.partial, q.nparts, q.count,
                                                    (T*)P.out[q.k].p, T(P.scale), P.out[q.k].acc);
      NB_LAUNCH_CHECK(ctx);
      nb_pool_release(ctx, q.partial);
    }
    pend.clear();
    return NB_OK;
  };
  auto alloc_partial = [&](int k, int64_t nparts, int64_t count) -> int32_t {
    void* p = nullptr;
    int32_t rc = nb_pool_alloc(ctx, (size_t)(nparts * count) * sizeof(T), &p);
    if (rc != NB_OK) return rc;
    P.out[k].partial = p;
    pend.push_back({p, nparts, count, k});
    return NB_OK;
  };

  if (ngroups <= 1) {
    const int64_t nvec = (P.rows + V - 1) / V;
    int64_t blocks = (nvec + 256 * 4 - 1) / (256 * 4);
    const int64_t cap = (int64_t)sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    for (int k = 0; k < P.nout; ++k) P.out[k].active = 1;
    // single primitives: the lean flat kernel takes any length (scalar tail, 0-dim operands) and, when one CTA
    // covers the array, writes reduced results itself (no k_finish launch); programs need whole vectors
    const bool lean_ok = lean_mode >= 0 && (prog_lean ? P.rows % V == 0 : true);
    const bool direct = lean_ok && !prog_lean && blocks == 1;
    if (!direct)
      for (int k = 0; k < P.nout; ++k)
        if (P.out[k].mode == OM_ALL) {
          int32_t rc = alloc_partial(k, blocks, 1);
          if (rc != NB_OK) return rc;
        }
    if (lean_ok && lean_try(0, LeanLaunch{P.rows / V, (unsigned)blocks, dim3(), dim3(), 0, 0})) {
      NB_LAUNCH_CHECK(ctx);
      return finish();
    }
    if (direct)  // (not reached for catalogue ops; keeps the general kernel's contract if it ever is)
      for (int k = 0; k < P.nout; ++k)
        if (P.out[k].mode == OM_ALL) {
          int32_t rc = alloc_partial(k, blocks, 1);
          if (rc != NB_OK) return rc;
        }
    {
      const size_t sm = with_slots(k_flat<T, Ev>, 0, 256, V);
      k_flat<T, Ev><<<(unsigned)blocks, 256, sm, st>>>(P, G);
    }
    NB_LAUNCH_CHECK(ctx);
    return finish();
  }

  if (P.rows <= 32) {
    // lanes = (row, column) pairs, RP = power of two >= rows (k_short)
    const int BT = 128;
    int rp = 1;
    while (rp < P.rows) rp <<= 1;
    const int cpb = BT / rp;
    int64_t blocks = (P.cols + cpb - 1) / cpb;
    const int64_t cap = (int64_t)sms * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    for (int k = 0; k < P.nout; ++k) {
      P.out[k].active = 1;
      int32_t rc = NB_OK;
      if (P.out[k].mode == OM_KEEP_ROWS) rc = alloc_partial(k, blocks, P.rows);
      if (P.out[k].mode == OM_ALL) rc = alloc_partial(k, blocks, 1);
      if (rc != NB_OK) return rc;
    }
    const size_t smem = with_slots(k_short<T, Ev>, (BT + 32) * sizeof(T), BT, 1);
    k_short<T, Ev><<<(unsigned)blocks, BT, smem, st>>>(P, G, rp);
    NB_LAUNCH_CHECK(ctx);
    return finish();
  }

  // rows > 32: KEEP_COLS outputs go to the warp-per-column kernel, everything else to k_tall
  bool any_tall = false, any_wcol = false;
  for (int k = 0; k < P.nout; ++k) (P.out[k].mode == OM_KEEP_COLS ? any_wcol : any_tall) = true;
  if (any_wcol && lean_mode >= 0 && P.rows % V == 0) {
    // lean warp-per-column kernel: handles KEEP_COLS and ELEM outputs in ONE pass
    bool ok = true;
    for (int k = 0; k < P.nout; ++k) ok = ok && (P.out[k].mode == OM_KEEP_COLS || P.out[k].mode == OM_ELEM);
    if (ok) {
      for (int k = 0; k < P.nout; ++k) P.out[k].active = 1;
      int64_t blocks = (P.cols + 7) / 8;
      const int64_t cap = (int64_t)sms * 8;
      if (blocks > cap) blocks = cap;
      if (lean_try(2, LeanLaunch{0, (unsigned)blocks, dim3(), dim3(), 0, 0})) {
        NB_LAUNCH_CHECK(ctx);
        return NB_OK;
      }
    }
  }
  if (any_tall) {
    for (int k = 0; k < P.nout; ++k) P.out[k].active = P.out[k].mode != OM_KEEP_COLS;
    const int64_t rvec = (P.rows + V - 1) / V;
    int TX = 32;
    while (TX < 256 && TX < rvec) TX <<= 1;
    const int TY = 256 / TX;
    const int64_t rblocks = (rvec + TX - 1) / TX;
    // enough CTAs for ~8 per SM, each slab at least 4*TY columns
    int64_t slabs = ((int64_t)sms * 8 + rblocks - 1) / rblocks;
    const int64_t min_cols = (int64_t)TY * 4;
    if (slabs > (P.cols + min_cols - 1) / min_cols) slabs = (P.cols + min_cols - 1) / min_cols;
    if (slabs < 1) slabs = 1;
    if (slabs > 256) slabs = 256;  // bounds the second-phase work (k_finish sums `slabs` partials)
    const int64_t cps = (P.cols + slabs - 1) / slabs;
    slabs = (P.cols + cps - 1) / cps;
    for (int k = 0; k < P.nout; ++k) {
      if (!P.out[k].active) continue;
      int32_t rc = NB_OK;
      if (P.out[k].mode == OM_KEEP_ROWS) rc = alloc_partial(k, slabs, P.rows);
      if (P.out[k].mode == OM_ALL) rc = alloc_partial(k, slabs * rblocks, 1);
      if (rc != NB_OK) return rc;
    }
    dim3 grid((unsigned)rblocks, (unsigned)slabs), block(TX, TY);
    const size_t smem = (size_t)256 * V * sizeof(T);
    if (!(lean_mode >= 0 && !any_wcol && P.rows % V == 0 &&
          lean_try(1, LeanLaunch{0, 0, grid, block, smem, cps})))
    {
      const size_t sm = with_slots(k_tall<T, Ev>, smem, 256, V);
      k_tall<T, Ev><<<grid, block, sm, st>>>(P, G, cps);
    }
    NB_LAUNCH_CHECK(ctx);
    int32_t rc = finish();
    if (rc != NB_OK) return rc;
  }
  if (any_wcol) {
    for (int k = 0; k < P.nout; ++k) P.out[k].active = P.out[k].mode == OM_KEEP_COLS;
    int64_t blocks = (P.cols + 7) / 8;
    const int64_t cap = (int64_t)sms * 8;
    if (blocks > cap) blocks = cap;
    const size_t sm = with_slots(k_wcol<T, Ev>, 0, 256, V);
    k_wcol<T, Ev><<<(unsigned)blocks, 256, sm, st>>>(P, G);
    NB_LAUNCH_CHECK(ctx);
  }
  return NB_OK;
}

static int32_t run_call(nb_ctx* ctx, const Call& C) {
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  NB_REQUIRE(ctx, C.nargs >= 1 && C.nargs <= NB_MAX_ARGS, NB_ERR_INVALID, "bad operand count");
  NB_REQUIRE(ctx, C.nout >= 1 && C.nout <= NB_MAX_OUTS, NB_ERR_INVALID, "bad output count");
  const int32_t dt = C.outs[0]->dtype;
  NB_REQUIRE(ctx, dt == NB_F32 || dt == NB_F64, NB_ERR_INVALID, "bad dtype");
  std::vector<const nb_tensor*> parts;
  for (int a = 0; a < C.nargs; ++a) parts.push_back(&C.args[a]);
  if (C.ybar) parts.push_back(C.ybar);
  if (C.ysaved) parts.push_back(C.ysaved);
  for (int k = 0; k < C.nout; ++k) parts.push_back(C.outs[k]);
  for (auto* t : parts) {
    NB_REQUIRE(ctx, t->dtype == dt, NB_ERR_INVALID, "mixed dtypes in one call");
    NB_REQUIRE(ctx, t->ndim >= 0 && t->ndim <= NB_MAX_DIMS, NB_ERR_INVALID, "ndim out of range");
    NB_REQUIRE(ctx, t->data != nullptr, NB_ERR_INVALID, "NULL data pointer");
  }
  // full broadcast shape from the operands (and ybar / y_saved, which must be compatible)
  // (outputs take part too: an output larger than every operand is a broadcast fill)
  int64_t F[NB_MAX_DIMS];
  for (int d = 0; d < NB_MAX_DIMS; ++d) {
    int64_t e = 1;
    for (auto* t : parts) {
      int64_t x = ext_of(t, d);
      if (x != 1) {
        NB_REQUIRE(ctx, e == 1 || e == x, NB_ERR_SHAPE, "DimensionMismatch: operands do not broadcast");
        e = x;
      }
    }
    F[d] = e;
  }
  for (auto* t : parts)
    for (int d = 0; d < NB_MAX_DIMS; ++d) {
      int64_t x = ext_of(t, d);
      NB_REQUIRE(ctx, x == 1 || x == F[d], NB_ERR_SHAPE,
                 "DimensionMismatch: tensor extent is neither 1 nor the broadcast extent");
    }
  if (C.ysaved)
    for (int d = 0; d < NB_MAX_DIMS; ++d)
      NB_REQUIRE(ctx, ext_of(C.ysaved, d) == F[d], NB_ERR_SHAPE, "y_saved must have the full shape");
  for (int k = 0; k < C.nout; ++k) nb_note_write(ctx, C.outs[k]->data);
  int64_t total = 1;
  for (int d = 0; d < NB_MAX_DIMS; ++d) total *= F[d];
  if (total == 0) {
    // empty index space: element-wise outputs are empty too; a reduction over nothing is zero
    // (Julia: sum(Float64[]) == 0.0), so non-accumulating reduced outputs are zero-filled
    for (int k = 0; k < C.nout; ++k) {
      const int64_t cnt = nb_numel(C.outs[k]);
      if (cnt > 0 && !C.out_acc[k]) {
        int32_t rc = nb_fill(ctx, C.outs[k]->data, dt, (uint64_t)cnt, 0.0);
        if (rc != NB_OK) return rc;
      }
    }
    return NB_OK;
  }

  // collapse dims: merge neighbours whose presence pattern is identical for every participant
  const int np = (int)parts.size();
  std::vector<std::vector<char>> pres(np, std::vector<char>(NB_MAX_DIMS, 0));
  for (int i = 0; i < np; ++i)
    for (int d = 0; d < NB_MAX_DIMS; ++d) pres[i][d] = F[d] > 1 && ext_of(parts[i], d) == F[d];
  int ng = 0;
  int64_t gext[NB_MAX_DIMS] = {1, 1, 1, 1};
  std::vector<std::vector<char>> gpres(np, std::vector<char>(NB_MAX_DIMS, 0));
  int prev = -1;
  for (int d = 0; d < NB_MAX_DIMS; ++d) {
    if (F[d] == 1) continue;
    bool same = prev >= 0;
    if (same)
      for (int i = 0; i < np; ++i)
        if (pres[i][d] != pres[i][prev]) { same = false; break; }
    if (same) {
      gext[ng - 1] *= F[d];
    } else {
      gext[ng] = F[d];
      for (int i = 0; i < np; ++i) gpres[i][ng] = pres[i][d];
      ++ng;
    }
    prev = d;
  }

  // program classification
  static const nb_prog identity_prog = {1, 0, 0, {0}, {0}, {0}, {0}};
  const nb_prog& G = C.prog ? *C.prog : identity_prog;
  NB_REQUIRE(ctx, G.nargs == C.nargs, NB_ERR_INVALID, "program/operand count mismatch");
  // un(bin(x0, x1)) over at most two operands / constants (`tanh.(A .+ b)`, `log.(f .+ eps)`): one fused primitive of
  // the single-primitive kernels instead of a two-instruction program on the interpreter
  static const bool fop_on = !(getenv("NB_FUSED_OPS") && atoi(getenv("NB_FUSED_OPS")) == 0);
  const bool is_fop = fop_on && G.n_instr == 2 && G.nargs <= 2 && nb_op_is_binary(G.op[0]) && !nb_is_fop(G.op[0]) &&
                      nb_op_is_unary(G.op[1]) && G.a[1] == G.nargs && nb_fop_supported(G.op[1], G.op[0]);
  const bool fast = G.n_instr <= 1 || is_fop;

  MRP P;
  memset(&P, 0, sizeof(P));
  P.rows = ng >= 1 ? gext[0] : 1;
  P.cols = ng >= 2 ? gext[1] : 1;
  P.pullback = C.pullback;
  P.scale = C.scale;
  P.nout = C.nout;
  P.has_ybar = C.ybar != nullptr;
  P.has_y = C.ysaved != nullptr;
  auto strides = [&](int part, Opnd& o) {
    const bool pr = ng >= 1 && gpres[part][0], pc = ng >= 2 && gpres[part][1];
    o.p = parts[part]->data;
    o.rs = pr ? 1 : 0;
    o.cs = pc ? (pr ? P.rows : 1) : 0;
  };
  int pi = 0;
  const int arg_part0 = 0;
  pi += C.nargs;
  const int ybar_part = C.ybar ? pi++ : -1;
  const int ys_part = C.ysaved ? pi++ : -1;
  const int out_part0 = pi;
  if (ybar_part >= 0) strides(ybar_part, P.ybar);
  if (ys_part >= 0) strides(ys_part, P.ys);

  if (fast) {
    // map the single instruction's two operands onto arg[0], arg[1]
    int op = G.n_instr ? G.op[0] : NB_OP_IDENTITY;
    int src[2] = {G.n_instr ? G.a[0] : 0, (G.n_instr && nb_op_is_binary(op)) ? G.b[0] : -1000};
    if (is_fop)   // derivative of the outer function from the Branch's saved value when it is a function of y alone
      op = nb_fop(G.op[1], G.op[0], C.pullback && C.ysaved && nb_fop_un_y_only(G.op[1]));
    P.op = op;
    P.nargs = 2;
    for (int s = 0; s < 2; ++s) {
      if (src[s] == -1000) { P.is_const[s] = 1; P.cval[s] = 0; continue; }
      if (src[s] < 0) {
        NB_REQUIRE(ctx, -src[s] - 1 < G.n_consts, NB_ERR_INVALID, "bad constant index");
        P.is_const[s] = 1; P.cval[s] = G.consts[-src[s] - 1];
      } else {
        NB_REQUIRE(ctx, src[s] < C.nargs, NB_ERR_INVALID, "bad operand index");
        strides(arg_part0 + src[s], P.arg[s]);
      }
    }
    int needs = 0;  // bit0 a, bit1 b, bit2 y
    if (!C.pullback) {
      needs = 3;
    } else {
      for (int k = 0; k < C.nout; ++k) {
        int w = 0;
        if (src[0] == C.out_wrt[k]) w |= 1;
        if (src[1] == C.out_wrt[k]) w |= 2;
        P.out[k].wrt = w;
        if (w & 1) needs |= nb_partial_needs(op, 0);
        if (w & 2) needs |= nb_partial_needs(op, 1);
      }
    }
    P.need_y = (needs & 4) != 0;
    if (P.need_y && !P.has_y) needs |= 3;  // must recompute y from the operands
    if (!P.need_y) P.has_y = 0;
    P.load_mask = 0;
    for (int s = 0; s < 2; ++s)
      if (!P.is_const[s] && (needs & (1 << s))) P.load_mask |= 1 << s;
  } else {
    P.nargs = C.nargs;
    for (int a = 0; a < C.nargs; ++a) strides(arg_part0 + a, P.arg[a]);
    for (int k = 0; k < C.nout; ++k) P.out[k].wrt = C.pullback ? C.out_wrt[k] : 0;
    P.load_mask = (1 << C.nargs) - 1;
  }
  for (int k = 0; k < C.nout; ++k) {
    Opnd o;
    strides(out_part0 + k, o);
    const bool pr = ng >= 1 && gpres[out_part0 + k][0], pc = ng >= 2 && gpres[out_part0 + k][1];
    P.out[k].p = (void*)o.p;
    P.out[k].rs = o.rs;
    P.out[k].cs = o.cs;
    P.out[k].acc = C.out_acc[k];
    if (ng <= 1) P.out[k].mode = (ng == 0 || pr) ? OM_ELEM : OM_ALL;
    else P.out[k].mode = pr ? (pc ? OM_ELEM : OM_KEEP_ROWS) : (pc ? OM_KEEP_COLS : OM_ALL);
  }
  // vectorisation is legal when every contiguous pointer is 16B aligned and column strides keep it
  {
    const int V = dt == NB_F64 ? 2 : 4;
    bool ok = true;
    auto chk = [&](const void* p, int64_t rs, int64_t cs) {
      if (rs == 0) return;
      if (((uintptr_t)p) % 16) ok = false;
      if (cs % V) ok = false;
    };
    for (int a = 0; a < P.nargs; ++a)
      if (P.arg[a].p) chk(P.arg[a].p, P.arg[a].rs, P.arg[a].cs);
    if (P.has_ybar) chk(P.ybar.p, P.ybar.rs, P.ybar.cs);
    if (C.ysaved) chk(P.ys.p, P.ys.rs, P.ys.cs);
    for (int k = 0; k < P.nout; ++k) chk(P.out[k].p, P.out[k].rs, P.out[k].cs);
    P.vec_ok = ok;
  }

  if (ng <= 2) {
    // Lean eligibility: hot single primitive, vectorisable, saved y present when the partial needs it,
    // and the wanted sensitivities are plain d/da, d/db (or both, in operand order).
    int lean_mode = -1;
    if (fast && P.vec_ok && (nb_op_is_unary(P.op) || nb_op_is_binary(P.op))) {
      if (!C.pullback) lean_mode = LM_FWD;
      else if (P.nout == 1 && (P.out[0].wrt == 1 || P.out[0].wrt == 2)) lean_mode = P.out[0].wrt;
      else if (P.nout == 2 && P.out[0].wrt == 1 && P.out[1].wrt == 2) lean_mode = LM_PULL_AB;
      if (lean_mode > 0) {
        int needs = 0;
        if (lean_mode & 1) needs |= nb_partial_needs(P.op, 0);
        if (lean_mode & 2) needs |= nb_partial_needs(P.op, 1);
        if ((needs & 4) && !C.ysaved) lean_mode = -1;  // would have to recompute y: general kernel
      }
      if (lean_mode >= 0) {
        P.lean_mask = (P.is_const[0] ? 0 : 1) | (P.is_const[1] ? 0 : 2);
        P.has_y = C.ysaved != nullptr;
      }
    }
    // composite programs with one or two array operands: the interpreter inside the lean kernels (bcast_prog.cu)
    int pmode = -1;
    if (!fast && P.vec_ok && C.nargs <= 2 && !getenv("NB_GEN_NOLEAN")) {
      if (!C.pullback) pmode = LM_FWD;
      else if (C.nout == 1 && (C.out_wrt[0] == 0 || C.out_wrt[0] == 1)) pmode = C.out_wrt[0] == 0 ? LM_PULL_A : LM_PULL_B;
      else if (C.nout == 2 && C.out_wrt[0] == 0 && C.out_wrt[1] == 1) pmode = LM_PULL_AB;
      if (pmode == LM_PULL_B && C.nargs == 1) pmode = -1;
      // saved forward value instead of re-evaluating a transcendental last instruction
      P.need_y = 0;
      if (pmode > 0 && C.ysaved && G.n_instr > 0) {
        const int lop = G.op[G.n_instr - 1];
        const bool y_only = nb_op_is_unary(lop) && nb_partial_needs(lop, 0) == 4;
        const bool cheap = lop == NB_OP_SQRT || lop == NB_OP_INV;
        if (y_only && !cheap && (G.rflags[G.n_instr - 1] & NB_RF_LIVE)) P.need_y = 1;
      }
      P.has_y = P.need_y;
    }
    // a composite pullback runs on a copy of the program specialised for the operands this launch differentiates
    nb_prog Gs;
    const nb_prog* Gp = &G;
    if (!fast && C.pullback && G.n_instr > 1) {
      Gs = G;
      uint32_t wanted = 0;
      for (int k = 0; k < C.nout; ++k) wanted |= 1u << C.out_wrt[k];
      nb_prog_derive(Gs, wanted);
      Gp = &Gs;
    }
    // otherwise: shared-memory interpreter when the slots of a 256-thread CTA fit
    const bool use_s = (size_t)2 * (G.nargs + G.n_instr) * 16 * 256 <= GEN_SMEM_LIMIT && !getenv("NB_GEN_LOCAL");
    if (dt == NB_F32) return fast ? launch_all<float, FastEv>(ctx, P, G, ng, lean_mode)
                           : use_s ? launch_all<float, GenEvS>(ctx, P, *Gp, ng, pmode, true)
                                   : launch_all<float, GenEv>(ctx, P, G, ng, pmode, true);
    return fast ? launch_all<double, FastEv>(ctx, P, G, ng, lean_mode)
         : use_s ? launch_all<double, GenEvS>(ctx, P, *Gp, ng, pmode, true)
                 : launch_all<double, GenEv>(ctx, P, G, ng, pmode, true);
  }

  // ---- generic N-D fallback: one launch per output ------------------------------------------
  NDP D;
  memset(&D, 0, sizeof(D));
  D.ng = ng;
  for (int g = 0; g < ng; ++g) D.ext[g] = gext[g];
  auto nd_strides = [&](int part, int64_t* st) {
    int64_t run = 1;
    for (int g = 0; g < ng; ++g) {
      st[g] = gpres[part][g] ? run : 0;
      if (gpres[part][g]) run *= gext[g];
    }
  };
  if (fast) {
    int op = G.n_instr ? G.op[0] : NB_OP_IDENTITY;
    int src[2] = {G.n_instr ? G.a[0] : 0, (G.n_instr && nb_op_is_binary(op)) ? G.b[0] : -1000};
    for (int s = 0; s < 2; ++s)
      if (src[s] >= 0) nd_strides(arg_part0 + src[s], D.ast[s]);
  } else {
    for (int a = 0; a < C.nargs; ++a) nd_strides(arg_part0 + a, D.ast[a]);
  }
  if (ybar_part >= 0) nd_strides(ybar_part, D.ybst);
  if (ys_part >= 0) nd_strides(ys_part, D.ysst);
  for (int k = 0; k < C.nout; ++k) {
    D.k = k;
    D.n_out = 1; D.n_red = 1;
    nd_strides(out_part0 + k, D.ost);
    for (int g = 0; g < ng; ++g) {
      D.red[g] = !gpres[out_part0 + k][g];
      (D.red[g] ? D.n_red : D.n_out) *= gext[g];
    }
    const unsigned blocks = (unsigned)((D.n_out + 255) / 256);
    if (dt == NB_F32) {
      if (fast) k_nd<float, FastEv><<<blocks, 256, 0, ctx->stream>>>(P, G, D);
      else k_nd<float, GenEv><<<blocks, 256, 0, ctx->stream>>>(P, G, D);
    } else {
      if (fast) k_nd<double, FastEv><<<blocks, 256, 0, ctx->stream>>>(P, G, D);
      else k_nd<double, GenEv><<<blocks, 256, 0, ctx->stream>>>(P, G, D);
    }
    NB_LAUNCH_CHECK(ctx);
  }
  return NB_OK;
}

template <typename T>
__global__ void k_fill(T* p, int64_t n, T v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

}  // namespace

extern "C" {

int32_t nb_prog_create(int32_t nargs, int32_t n_instr, const int32_t* ops, const int32_t* a,
                       const int32_t* b, int32_t n_consts, const double* consts, nb_prog** out) {
  if (!out) return nb_fail(nullptr, NB_ERR_INVALID, "nb_prog_create: out is NULL");
  *out = nullptr;
  if (nargs < 1 || nargs > NB_MAX_ARGS) return nb_fail(nullptr, NB_ERR_INVALID, "nb_prog_create: nargs out of range");
  if (n_instr < 0 || n_instr > NB_MAX_INSTR) return nb_fail(nullptr, NB_ERR_UNSUPPORTED, "nb_prog_create: program too long");
  if (n_consts < 0 || n_consts > NB_MAX_CONSTS) return nb_fail(nullptr, NB_ERR_UNSUPPORTED, "nb_prog_create: too many constants");
  nb_prog* p = new nb_prog();
  memset(p, 0, sizeof(*p));
  p->nargs = nargs; p->n_instr = n_instr; p->n_consts = n_consts;
  for (int i = 0; i < n_consts; ++i) p->consts[i] = consts[i];
  for (int i = 0; i < n_instr; ++i) {
    const int op = ops[i];
    const bool un = nb_op_is_unary(op), bi = nb_op_is_binary(op);
    auto ok_idx = [&](int idx) { return idx >= 0 ? idx < nargs + i : (-idx - 1) < n_consts; };
    if (!(un || bi) || !ok_idx(a[i]) || (bi && !ok_idx(b[i]))) {
      delete p;
      return nb_fail(nullptr, NB_ERR_UNSUPPORTED, "nb_prog_create: unsupported opcode or bad operand index");
    }
    p->op[i] = op; p->a[i] = a[i]; p->b[i] = bi ? b[i] : 0;
  }
  nb_prog_derive(*p);
  *out = p;
  return NB_OK;
}

int32_t nb_prog_destroy(nb_prog* p) {
  delete p;
  return NB_OK;
}

int32_t nb_bcast_fwd(nb_ctx* ctx, const nb_prog* prog, int32_t nargs, const nb_tensor* args,
                     const nb_tensor* out) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && args && out, NB_ERR_INVALID, "NULL argument");
  Call C{};
  C.prog = prog; C.nargs = nargs; C.args = args; C.pullback = false; C.nout = 1;
  C.outs[0] = out; C.out_acc[0] = false; C.scale = 1.0;
  // forward results carry the full broadcast shape
  for (int d = 0; d < NB_MAX_DIMS; ++d) {
    int64_t e = 1;
    for (int a = 0; a < nargs && a < NB_MAX_ARGS; ++a)
      if (ext_of(&args[a], d) != 1) e = ext_of(&args[a], d);
    NB_REQUIRE(ctx, ext_of(out, d) == e, NB_ERR_SHAPE, "nb_bcast_fwd: out must have the broadcast shape");
  }
  return run_call(ctx, C);
}

int32_t nb_prog_pullback_plan(const nb_prog* prog, uint32_t wanted_mask, uint32_t* live_instr, uint32_t* fwd_skip,
                              uint32_t* arg_live) {
  if (!prog || !live_instr || !fwd_skip || !arg_live) return nb_fail(nullptr, NB_ERR_INVALID, "nb_prog_pullback_plan: NULL argument");
  nb_prog g = *prog;
  nb_prog_derive(g, wanted_mask);
  *live_instr = 0;
  for (int i = 0; i < g.n_instr; ++i)
    if (g.rflags[i] & NB_RF_LIVE) *live_instr |= 1u << i;
  *fwd_skip = g.fwd_skip;
  *arg_live = g.arg_live;
  return NB_OK;
}

int32_t nb_bcast_reduce(nb_ctx* ctx, const nb_prog* prog, int32_t nargs, const nb_tensor* args,
                        double scale, const nb_tensor* out, int32_t accumulate) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && args && out, NB_ERR_INVALID, "NULL argument");
  Call C{};
  C.prog = prog; C.nargs = nargs; C.args = args; C.pullback = false; C.nout = 1;
  C.outs[0] = out; C.out_acc[0] = accumulate != 0; C.scale = scale;
  return run_call(ctx, C);
}

int32_t nb_bcast_pullback(nb_ctx* ctx, const nb_prog* prog, const nb_tensor* ybar,
                          const nb_tensor* y_saved, int32_t nargs, const nb_tensor* args,
                          uint32_t want_mask, const nb_tensor* outs, uint32_t accumulate_mask) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && args && outs, NB_ERR_INVALID, "NULL argument");
  NB_REQUIRE(ctx, nargs >= 1 && nargs <= NB_MAX_ARGS, NB_ERR_INVALID, "bad operand count");
  Call C{};
  C.prog = prog; C.nargs = nargs; C.args = args; C.pullback = true; C.scale = 1.0;
  C.ybar = ybar; C.ysaved = y_saved;
  // fast path handles <= 2 outputs per launch; generic up to NB_MAX_OUTS
  const bool fast = !prog || prog->n_instr <= 1;
  const int per = fast ? 2 : NB_MAX_OUTS;
  int k = 0;
  while (k < nargs) {
    C.nout = 0;
    for (; k < nargs && C.nout < per; ++k) {
      if (!(want_mask >> k & 1)) continue;
      for (int d = 0; d < NB_MAX_DIMS; ++d)
        NB_REQUIRE(ctx, ext_of(&outs[k], d) == ext_of(&args[k], d), NB_ERR_SHAPE,
                   "nb_bcast_pullback: outs[k] must be shaped like args[k]");
      C.outs[C.nout] = &outs[k];
      C.out_wrt[C.nout] = k;
      C.out_acc[C.nout] = (accumulate_mask >> k & 1) != 0;
      C.nout++;
    }
    if (C.nout) {
      int32_t rc = run_call(ctx, C);
      if (rc != NB_OK) return rc;
    }
  }
  return NB_OK;
}

int32_t nb_fill(nb_ctx* ctx, void* dst, int32_t dtype, uint64_t count, double value) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && (dst || !count), NB_ERR_INVALID, "NULL argument");
  if (!count) return NB_OK;
  nb_note_write(ctx, dst);
  int64_t blocks = (count + 1023) / 1024;
  if (blocks > ctx->num_sms * 8) blocks = ctx->num_sms * 8;
  if (dtype == NB_F32) k_fill<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>((float*)dst, count, (float)value);
  else if (dtype == NB_F64) k_fill<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>((double*)dst, count, value);
  else return nb_fail(ctx, NB_ERR_INVALID, "bad dtype");
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

// y = alpha*op(A)*x + beta*y.  'N': out[i] = sum_j A[i,j]*x[j] is the KEEP_ROWS pattern (k_tall),
// 'T': out[j] = sum_i A[i,j]*x[i] is the KEEP_COLS pattern (k_wcol): both stream A once at HBM rate.
int32_t nb_gemv(nb_ctx* ctx, char tA, int64_t m, int64_t n, double alpha, const void* A,
                int64_t lda, const void* x, double beta, void* y, int32_t dtype) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && A && x && y, NB_ERR_INVALID, "NULL argument");
  NB_REQUIRE(ctx, lda >= m, NB_ERR_SHAPE, "nb_gemv: leading dimension too small (DimensionMismatch)");
  const bool tr = tA == 'T' || tA == 't' || tA == 'C' || tA == 'c';
  if (m > 0 && n > 0) {  // the streaming skinny-GEMM kernels (gemm_skinny.cu): A read once with 16-byte loads, any lda
    nb_note_write(ctx, y);
    int32_t rc = nb_gemm_skinny(ctx, tr, false, tr ? n : m, 1, tr ? m : n, alpha, A, lda, x, tr ? m : n, beta, y,
                                tr ? n : m, dtype, nullptr, 0);
    if (rc != NB_ERR_UNSUPPORTED) return rc;
  }
  NB_REQUIRE(ctx, lda == m, NB_ERR_UNSUPPORTED, "nb_gemv: small matrices must be dense (lda == m)");
  static const nb_prog mul = {2, 1, 0, {NB_OP_MUL}, {0}, {1}, {0}};
  const int64_t ylen = tr ? n : m;
  if (beta != 0.0 && beta != 1.0) {  // y *= beta first
    static const nb_prog scl = {1, 0, 0, {0}, {0}, {0}, {0}};
    nb_tensor yt{y, dtype, 1, {ylen, 1, 1, 1}};
    Call S{};
    S.prog = &scl; S.nargs = 1; S.args = &yt; S.nout = 1; S.outs[0] = &yt; S.scale = beta;
    int32_t rc = run_call(ctx, S);
    if (rc != NB_OK) return rc;
  }
  nb_tensor args[2];
  args[0] = nb_tensor{(void*)A, dtype, 2, {m, n, 1, 1}};
  args[1] = tr ? nb_tensor{(void*)x, dtype, 2, {m, 1, 1, 1}} : nb_tensor{(void*)x, dtype, 2, {1, n, 1, 1}};
  nb_tensor out = tr ? nb_tensor{y, dtype, 2, {1, n, 1, 1}} : nb_tensor{y, dtype, 2, {m, 1, 1, 1}};
  Call C{};
  C.prog = &mul; C.nargs = 2; C.args = args; C.nout = 1; C.outs[0] = &out;
  C.out_acc[0] = beta != 0.0; C.scale = alpha;
  return run_call(ctx, C);
}

// A = alpha*x*y' + beta*A (outer product): the elementwise pattern with two broadcast operands
int32_t nb_ger(nb_ctx* ctx, int64_t m, int64_t n, double alpha, const void* x, const void* y,
               double beta, void* A, int64_t lda, int32_t dtype) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && A && x && y, NB_ERR_INVALID, "NULL argument");
  NB_REQUIRE(ctx, lda >= m, NB_ERR_SHAPE, "nb_ger: leading dimension too small (DimensionMismatch)");
  if (m > 0 && n > 0) {  // k = 1 product: the small-K kernel of gemm_skinny.cu (any lda, any beta)
    nb_note_write(ctx, A);
    int32_t rc = nb_gemm_skinny(ctx, false, false, m, n, 1, alpha, x, m, y, 1, beta, A, lda, dtype, nullptr, 0);
    if (rc != NB_ERR_UNSUPPORTED) return rc;
  }
  NB_REQUIRE(ctx, lda == m, NB_ERR_UNSUPPORTED, "nb_ger: small matrices must be dense (lda == m)");
  NB_REQUIRE(ctx, beta == 0.0 || beta == 1.0, NB_ERR_UNSUPPORTED, "nb_ger: beta must be 0 or 1");
  static const nb_prog mul = {2, 1, 0, {NB_OP_MUL}, {0}, {1}, {0}};
  nb_tensor args[2] = {nb_tensor{(void*)x, dtype, 2, {m, 1, 1, 1}}, nb_tensor{(void*)y, dtype, 2, {1, n, 1, 1}}};
  nb_tensor out{A, dtype, 2, {m, n, 1, 1}};
  Call C{};
  C.prog = &mul; C.nargs = 2; C.args = args; C.nout = 1; C.outs[0] = &out;
  C.out_acc[0] = beta == 1.0; C.scale = alpha;
  return run_call(ctx, C);
}

}  // extern "C"
