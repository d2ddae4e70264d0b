// Types shared by the broadcast-family translation units (bcast.cu and the cold-catalogue lean kernels
// in bcast_lean_cold_*.cu): the collapsed (rows, cols) launch description and the reduction helpers.
#pragma once
#include "common.cuh"

namespace nbb {

enum { OM_ELEM = 0, OM_KEEP_ROWS = 1, OM_KEEP_COLS = 2, OM_ALL = 3 };

struct Opnd {
  const void* p;
  int64_t rs, cs;  // element strides along rows / cols of the collapsed space (rs in {0,1})
};

struct OutS {
  void* p;
  int64_t rs, cs;
  int32_t mode;
  int32_t acc;     // fused +=
  int32_t wrt;     // generic: operand index; fast: bit0 -> d/da, bit1 -> d/db
  int32_t active;  // handled by this launch
  void* partial;
};

// sensitivities one launch of the general kernels produces (a pullback that wants more runs several launches)
#define NB_MAX_OUTS 6
#define NB_FAST_NIN 2
struct MRP {
  int64_t rows, cols;
  int32_t nargs;     // operands to load (fast: 2 = A,B; generic: program operands)
  int32_t nout;
  int32_t pullback;
  int32_t op;        // fast path opcode
  int32_t load_mask; // which operands are actually read (general kernels: FastEv)
  int32_t lean_mask; // lean kernels: bit s = operand s is an array (not a baked-in constant); WHICH operands are read is
                     // fixed at compile time there (LeanNeeds).  Kept apart from load_mask: a launch that qualifies for
                     // the lean kernels can still end on a general one (rows <= 32), whose loads must stay limited to
                     // what the partials read — an operand nothing reads may be a shape-only stand-in (fusion.Shadow)
  int32_t has_ybar, has_y, need_y;
  int32_t vec_ok;    // all contiguous operands 16B-aligned with compatible strides
  int32_t slot_off;  // GenEvS: byte offset of the interpreter's slot storage in dynamic shared memory
  double cval[NB_FAST_NIN];  // fast: value of a constant operand
  int32_t is_const[NB_FAST_NIN];
  double scale;
  Opnd arg[NB_MAX_ARGS];
  Opnd ybar, ys;
  OutS out[NB_MAX_OUTS];
};

template <typename T> struct VecOf;
template <> struct VecOf<float> { static constexpr int V = 4; using type = float4; };
template <> struct VecOf<double> { static constexpr int V = 2; using type = double2; };

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// block-wide sum of one value per thread; result valid in thread 0.  smem: >= 32 T
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nw = (blockDim.x * blockDim.y + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < nw ? smem[lane] : T(0);
    v = warp_sum(v);
  }
  return v;
}

// launch geometry handed from bcast.cu's dispatcher to a lean kernel (whichever mapping is used)
struct LeanLaunch {
  int64_t nvec;      // flat: number of V-vectors
  unsigned blocks;   // flat / wcol grid
  dim3 grid, block;  // tall
  size_t smem;       // tall
  int64_t cps;       // tall: columns per slab
};

}  // namespace nbb

// Cold-catalogue lean kernels (bcast_lean_cold_*.cu): kind 0 = flat, 1 = tall, 2 = wcol.  Returns false
// when (op, mode, kind) has no compile-time-specialised kernel; the caller then uses the general kernels.
bool nb_lean_cold_un_f32(int kind, int op, int mode, const nbb::MRP& P, const nbb::LeanLaunch& a, cudaStream_t st);
bool nb_lean_cold_un_f64(int kind, int op, int mode, const nbb::MRP& P, const nbb::LeanLaunch& a, cudaStream_t st);
bool nb_lean_cold_bin_f32(int kind, int op, int mode, const nbb::MRP& P, const nbb::LeanLaunch& a, cudaStream_t st);
bool nb_lean_cold_bin_f64(int kind, int op, int mode, const nbb::MRP& P, const nbb::LeanLaunch& a, cudaStream_t st);
// fused un(bin(a, b)) primitives (bcast_lean_fused_*.cu; ops.cuh nb_fop)
bool nb_lean_fused_f32(int kind, int op, int mode, const nbb::MRP& P, const nbb::LeanLaunch& a, cudaStream_t st);
bool nb_lean_fused_f64(int kind, int op, int mode, const nbb::MRP& P, const nbb::LeanLaunch& a, cudaStream_t st);
// composite programs on the lean kernels (bcast_prog.cu)
struct nb_prog;
bool nb_lean_prog(int dtype, int kind, int mode, const nbb::MRP& P, const nb_prog& G, const nbb::LeanLaunch& a, cudaStream_t st);
