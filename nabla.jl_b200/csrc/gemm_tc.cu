// Family (3), tensor-core part: Float32 GEMM on the 5th-generation tensor cores.
//
//   C(m x n) = alpha * op(A) * op(B) + beta * C       (BLAS column-major, all four transposes)
//
// replaces OpenBLAS sgemm behind ChainRules' rrule(*) / BLAS.gemm adjoints (reached through
// src/sensitivities/chainrules.jl:179-191,286 and src/sensitivities/linalg/blas.jl:224-228):
// forward W*X, Abar = Ybar*B' ('N','T'), Bbar = A'*Ybar ('T','N'); beta = 1 is the in-place accumulate.
//
// Design (sm_100a):
//   * two persistent kernels.  k_gemm_tc: one CTA per SM, tile 128 x BN (BN = 256 or 128).  k_gemm_tc_pair
//     (large products, the default there): one CTA PAIR (2-CTA cluster, tcgen05 cta_group::2) per two SMs, tile
//     256 x 256, each CTA stages half of B and its own half of A, the leader's MMA thread issues for both and
//     releases the stages of both CTAs with multicast commits.  K step = one 128-byte swizzle row.
//     12 warps: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one elected thread) + TMEM owner, warps 2-3
//     idle (they give their registers away, setmaxnreg), warps 4..11 = epilogue (TMEM -> registers -> stores).
//   * operands are staged by TMA (cp.async.bulk.tensor.2d, 128-byte swizzles) straight from the
//     column-major arrays: a non-transposed A is "MN-major", a transposed A is "K-major" (the UMMA
//     descriptors carry the major-ness, so no transposed copy of any operand is ever made).
//   * tiles are handed out DYNAMICALLY: the producer thread takes the next tile index from a global
//     atomic counter and publishes it to the MMA and epilogue roles (of both CTAs of a pair) through a
//     shared-memory ring (mbarrier full/empty), so a CTA that starts late or shares its SM with an NCCL kernel
//     (the overlapped allreduce) simply processes fewer tiles instead of delaying the whole GEMM.  Tiles are
//     rasterised in groups of group_m tile rows sized to the L2, and consecutive tiles walk K in alternating
//     directions (serpentine) so the panel tail of one tile is the head of the next.
//   * accumulators live in TMEM (2 x BN columns, double-buffered so the epilogue of chunk i overlaps
//     the main loop of chunk i+1); smem ring of STAGES stages guarded by full/empty mbarriers;
//     tcgen05.commit releases smem stages and publishes finished accumulators.
//   * precision modes (nb_math_mode).  kind::tf32 keeps 10 mantissa bits of each operand: 7.7e-4
//     relative error, outside the 1e-4 bar.  The split modes write x = hi + lo and issue
//     lo*hi + hi*lo + hi*hi into one accumulator (three tensor passes per product):
//       NB_MATH_3XTF32  hi = tf32(x), lo = x - hi (both exact): 21 bits, 4e-7 error, 3 tf32 passes.
//       NB_MATH_BF16X3  (DEFAULT) hi = bf16(x), lo = bf16(x - hi): 16 bits (residual <= 2^-18 |x|),
//                       ~1e-5 error, 3 kind::f16 passes at twice the tf32 rate.  The tensor pipe is
//                       power-limited on this part (1 kW cap), so tensor work per useful FLOP is the
//                       cost that matters: half of 3xTF32's time and energy.  The split copies are this
//                       kernel's own temporaries (cached per operand buffer, or emitted by the epilogue of the
//                       product that wrote the operand), so their leading dimension is padded to a multiple of
//                       8 elements: ANY lda/ldb/alignment works (e.g. the 10-row output layer).
//   * the tensor core adds each product group into the FP32 accumulator with truncation, so the error
//     of one long accumulation grows linearly in K (measured 2.9e-5 at K = 4096 for tf32).  K is
//     therefore processed in chunks of KCHUNK: every chunk gets a fresh TMEM accumulator and the
//     epilogue warps fold it into a running sum held in REGISTERS with a correctly rounded FP32 add (C is
//     written once, after the last chunk), which keeps the error at the single-chunk level for any K
//     (the Abar GEMMs have K = batch = 32768).
//   * fused epilogues (nb_gemm_fused): + bias and activation on the way out (the forward Branches `.+ b`,
//     `tanh.`), `.* f'(y)` of a saved activation (their pullbacks), per-row sums (the bias sensitivity) and the
//     bf16 hi/lo split of the result for the next product — 16-byte stores after a 4 x 4 transpose inside each
//     quad of lanes.
//   * kind::i8 instances of the same kernels compute the exact int8 slice products of the Float64 Ozaki split
//     (ozaki.cu): int32 TMEM accumulators, Float64 epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstring>

#include "ops.cuh"

namespace {

constexpr int BM = 128;
// 12 warps = 3 warpgroups: WG0 = TMA producer (warp 0), MMA issuer (warp 1), two idle warps; WG1, WG2 = epilogue
// (each warp owns one TMEM lane quadrant = 32 rows of the tile and half of its columns).  The epilogue keeps the
// running sum of the K chunks in REGISTERS (BN/2 per thread), so WG0 gives registers away (setmaxnreg).
constexpr int NTHREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 256;
constexpr int EPI_WARPS = EPI_THREADS / 32;  // the accumulator is released by ONE arrive per warp (see the fold loops)
constexpr int REGS_WG0 = 80, REGS_EPI = 208;  // 128*80 + 256*208 <= 384*168

// Operand kinds.  Every staged row is 128 bytes (one swizzle row), so tile byte sizes are kind-independent.
// KIND_I8: signed 8-bit operands, exact int32 accumulation (kind::i8) — the slice products of the Float64 Ozaki
// split (ozaki.cu); K-major operands only, one pass, no K chunking (the int32 accumulator is exact).
// KIND_I8ALL: every level of the Ozaki split in ONE launch (chunk table in GemmArgs, Float64 running sum in the
// epilogue registers): mid-size Float64 products, where one launch and one read-modify-write of C per level dominate.
enum { KIND_TF32 = 0, KIND_3XTF32 = 1, KIND_BF16X3 = 2, KIND_I8 = 3, KIND_I8ALL = 4 };
__host__ __device__ constexpr bool kind_is_i8(int k) { return k == KIND_I8 || k == KIND_I8ALL; }
template <int KIND>
struct KindT {
  static constexpr int ESIZE = kind_is_i8(KIND) ? 1 : KIND == KIND_BF16X3 ? 2 : 4;  // bytes per staged element
  static constexpr int BK = 128 / ESIZE;                          // elements of K per stage (32 / 64)
  static constexpr int UMMA_K = 32 / ESIZE;                       // K of one tcgen05.mma (8 / 16)
  static constexpr int MN_CHUNK = 128 / ESIZE;                    // MN-elements per 128-byte row (32 / 64)
  static constexpr uint32_t CHUNK_BYTES = BK * 128;               // one MN-major chunk: MN_CHUNK x BK rows
  static constexpr int NSPLIT = (KIND == KIND_TF32 || kind_is_i8(KIND)) ? 1 : 2;  // staged copies per operand (hi, lo)
  static constexpr int NPASS = (KIND == KIND_TF32 || kind_is_i8(KIND)) ? 1 : 3;
  // instruction-descriptor a/b format: F32F16 family 1 = bf16, 2 = tf32; S8 family 1 = signed 8 bit
  static constexpr uint32_t FMT = (KIND == KIND_BF16X3 || kind_is_i8(KIND)) ? 1u : 2u;
  static constexpr uint32_t CFMT = kind_is_i8(KIND) ? 2u : 1u;     // accumulator: 1 = f32, 2 = s32
  // MN-major smem layout (see make_desc)
  static constexpr uint64_t MN_TYPE = KIND == KIND_BF16X3 ? 2 : 1;
  static constexpr uint64_t MN_SBO = KIND == KIND_BF16X3 ? 1024 : 512;
  static constexpr uint32_t MN_KSTEP = UMMA_K * 128;              // bytes to the next UMMA_K k-rows
  // K range accumulated in TMEM before the epilogue folds it into C.  Also bounds the L2 working set
  // of the tiles in flight (a 16 x ~9 block of tiles): ~35 MB per chunk for every kind.
  static constexpr int KCHUNK = KIND == KIND_3XTF32 ? 1024 : 2048;
};

// Fused epilogue (tape-level fusion of the Branches around a product):
//   EPI_FWD   C = f(alpha*A*B + bias[row])           `f.(W*X .+ b)`: the `*`, `.+` and `f.` Branches in one launch
//   EPI_PULL  C = (alpha*A*B) .* f'(y)               `Ybar .* f'(y)` on top of `Ybar = W'*Zbar`: the `*` pullback
//                                                    and the `f.` pullback (functional.jl:92-95) in one launch
// plus, for either, the row sums of C (the bias sensitivity: the `.+` pullback's unbroadcast, functional.jl:61-63)
// and the hi/lo operand split of C for the next product (k_split).  Applied by the epilogue of the LAST K chunk.
enum { EPI_NONE = 0, EPI_FWD = 1, EPI_PULL = 2 };
struct EpiArgs {
  int32_t kind, op;
  const float* bias;      // EPI_FWD, nullable
  const float* y;         // EPI_PULL: saved forward value (m x n, ldy)
  int64_t ldy;
  float* rowsum_partial;  // nullable: [2 * tiles_n][M] per-tile-half row sums, summed in order by k_rowsum_finish
  void* s_hi;             // nullable: split copies of C (bf16 for BF16x3, f32 for 3xTF32), leading dim ld_s
  void* s_lo;
  int64_t ld_s;
  int32_t s_fp16;         // the emitted 16-bit split is scaled binary16 (x * s_scale) instead of bfloat16
  float s_scale;
  void* s2_hi;            // nullable: a second, bfloat16 split of C (same ld_s) next to a binary16 one — forward values
  void* s2_lo;            // feed the next forward product (binary16) AND the adjoint products (bfloat16)
  int32_t last_launch;    // this launch covers the end of K
};

struct GemmArgs {
  int64_t M, N, K;
  int32_t a_mn, b_mn;  // operand is MN-major in memory (A: not transposed, B: transposed)
  // 16-bit split kinds: element format of each operand's hi/lo copies in the instruction descriptor (0 = binary16
  // scaled by a power of two per tensor, 1 = bfloat16) and the device-resident inverse scales (nullable = 1)
  uint32_t a_fmt, b_fmt;
  const float* a_sinv;
  const float* b_sinv;
  float alpha, beta;
  float* C;
  int64_t ldc;
  int32_t tiles_m, tiles_n;
  int32_t* tile_counter;  // zeroed before the launch; next tile index to hand out
  int32_t tile_begin, tile_end;  // this launch hands out tiles [tile_begin, tile_end) of the rasterised order
                                 // (tile_end = 0: all) — a product can be cut at any tile boundary at no extra traffic
  const int32_t* skip_if; // host side only, nullable: device word; when non-zero at launch time the counter starts beyond
                          // the last tile, so every CTA sees the end sentinel at once and the launch does nothing
  int32_t group_m;        // tile-rows per rasterisation group (see tile_coords)
  int64_t k_begin, k_end; // K range of this launch (multiples of the K step except at the very end)
  int32_t kchunk_kb;      // K blocks accumulated in TMEM before the epilogue folds them into C
  int32_t kchunk_first_kb;  // length of the FIRST TWO chunks of every tile (>= kchunk_kb): they are issued while the
                            // epilogue still stores the previous tile, so they are the MMA's runway (see chunk_len)
  uint64_t hint_a, hint_b;  // L2 eviction policies of the operand loads (0 = none)
  int32_t vec_store;        // epilogue may use 16-byte accesses along the rows (alignment checked on the host)
  int32_t serpentine;       // odd waves walk K backwards: the panel slices used last are still in the L2
  int32_t dbg_extra_loads;  // experiment: re-load the A-hi tile this many extra times per K block into scratch smem
  EpiArgs epi;
  // KIND_I8: C64[i,j] = beta64*C64[i,j] + scale64 * row_scale[i] * col_scale[j] * (int32 accumulator)
  double* C64;
  const double* row_scale;
  const double* col_scale;
  double scale64, beta64;
  // KIND_I8ALL: chunk c = one level of the split: tab_nkb[c] K blocks, A from K element tab_ka[c], B from tab_kb[c],
  // its int32 accumulator weighted by tab_scale[c] in the Float64 running sum
  int32_t nchunk_tab;
  int32_t tab_nkb[8], tab_ka[8], tab_kb[8];
  double tab_scale[8];
};

constexpr int SCHED_SLOTS = 4;

// Chunk schedule of one tile, identical in the MMA issuer and the epilogue: chunks 0 and 1 have kchunk_first_kb K
// blocks, the rest kchunk_kb.
__device__ __forceinline__ int chunk_len(const GemmArgs& g, int chunk_index) {
  return chunk_index < 2 ? g.kchunk_first_kb : g.kchunk_kb;
}

// Tile rasterisation: groups of `group_m` tile-rows are swept along n with m fastest inside the group, so the
// ~74 tiles in flight form a group_m x ~9 block that shares group_m A panels and ~9 B panels; long K is launched
// in ranges of 8192.  Measured (ncu, 8192x32768x8192 BF16x3, tools/gemm_knob_ab.sh): the operands are still read
// ~6x from DRAM (14-18 GB for 2.4 GB algorithmic, 88 GB of L2->SM traffic) for every group size, L2 eviction
// hint and K range tried — the panel set of a wave (17 panels x 8.4 MB) exceeds what the two-partition L2 keeps
// of data shared by all SMs — but DRAM runs at ~2 TB/s of 6.5 and the tensor pipe stays 99.6 % active; group_m = 8
// is 3 % faster than 5 under the power cap.
__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int group_m, int& tm, int& tn) {
  const int per_group = group_m * tiles_n;
  const int group = t / per_group;
  const int first_m = group * group_m;
  const int gsize = min(tiles_m - first_m, group_m);
  const int r = t - group * per_group;
  tm = first_m + r % gsize;
  tn = r / gsize;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 20000000000LL) __trap();  // ~10 s
  }
}

// L2 eviction policies for the operand loads (createpolicy.fractional encodings, fraction 1.0)
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            uint64_t hint) {
  if (hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
  }
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], M = 128, N and major-ness from idesc
template <int KIND>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  if (kind_is_i8(KIND)) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else if (KIND == KIND_BF16X3) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// UMMA shared-memory descriptors (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   K-major  (SWIZZLE_128B, layout type 2): rows of 128 B along K; 8-row swizzle atoms 1024 B apart
//            (SBO); LBO unused (1).
//   MN-major tf32 (SWIZZLE_128B_BASE32B, layout type 1 — the only MN-major layout kind::tf32 accepts,
//            cutlass sm100_common.inl "for mn-major tf32 operands, SW128_32B is the only available smem
//            layout"): rows of 128 B hold 32 MN-elements of one k; 4-row swizzle atoms (32-byte chunks
//            XOR row%4) 512 B apart (SBO); the next 32 MN-elements at LBO = CHUNK_BYTES.  TMA writes
//            exactly this with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
//   MN-major bf16: ordinary SWIZZLE_128B (type 2): rows hold 64 MN-elements, 8-row atoms, SBO = 1024,
//            the next 64 MN-elements at LBO = CHUNK_BYTES.
template <int KIND>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major) {
  using KT = KindT<KIND>;
  const uint64_t lbo = mn_major ? (KT::CHUNK_BYTES >> 4) : 1;
  const uint64_t sbo = mn_major ? (KT::MN_SBO >> 4) : (1024 >> 4);
  const uint64_t type = mn_major ? KT::MN_TYPE : 2;
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (type << 61);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 64 columns in one instruction: the fold of a short K chunk is a latency chain (barrier -> TMEM load -> add -> barrier)
// between two short MMA chunks, so fewer, wider loads matter more than register economy
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- operand splits -------------------------------------------------------------------------------
// tf32: hi = round-to-nearest tf32(x) (low 13 mantissa bits zero), lo = x - hi (exact), stored as f32
__device__ __forceinline__ void split1(float x, float& hi, float& lo) {
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));
  hi = __uint_as_float(t);
  lo = x - hi;
}
// bf16: hi = bf16(x), lo = bf16(x - hi); x - hi - lo <= 2^-18 |x|
__device__ __forceinline__ void split1(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// scaled binary16: hi = fp16(s*x), lo = fp16(s*x - hi): 22 significant bits for every element within 2^18 of the
// tensor's largest magnitude (s is the power of two that puts that magnitude just below 2^15), an absolute floor of
// 2^-40 of it below; the product's epilogue multiplies by the inverse scales
__device__ __forceinline__ void split1s(float x, float s, __half& hi, __half& lo) {
  const float xs = x * s;
  hi = __float2half_rn(xs);
  lo = __float2half_rn(xs - __half2float(hi));
}

__device__ __forceinline__ float eff_alpha(const float alpha, const float* a_sinv, const float* b_sinv) {
  float a = alpha;
  if (a_sinv) a *= __ldg(a_sinv);
  if (b_sinv) a *= __ldg(b_sinv);
  return a;
}

__device__ __forceinline__ float epi_act(int op, float x) {
  switch (op) {
    case NB_OP_IDENTITY: return x;
    case NB_OP_TANH: return tanhf(x);
    case NB_OP_LOGISTIC: return 1.f / (1.f + expf(-x));
    case NB_OP_EXP: return expf(x);
    default: return nb_eval_cold<float>(op, x, 0.f);
  }
}

__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_WG0)); }
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI)); }

// The epilogue's view of one output row (this thread's row gm) x NCOL columns.  K is accumulated in TMEM in
// chunks (truncating FP32 accumulate: the error of one long accumulation grows linearly in K); every finished
// chunk is folded into `r` with a correctly rounded FP32 add, the TMEM buffer is handed back to the MMA issuer
// right after the read, and C is written ONCE per tile (ncu on the previous version, which folded through C in
// global memory: 3.9 GB written and ~3 GB re-read per 8192x32768x8192 product for 1.07 GB of output, and the
// product ran 11 % faster with the folds switched off).
template <int KIND, int NCOL>
struct EpiAcc {
  float r[NCOL];

  __device__ __forceinline__ void fold(uint32_t taddr, bool first) {
    if constexpr (NCOL % 64 == 0 && !kind_is_i8(KIND)) {
#pragma unroll
      for (int c = 0; c < NCOL / 64; ++c) {
        uint32_t v[64];
        tmem_ld64(taddr + (uint32_t)(c * 64), v);
        if (first) {
#pragma unroll
          for (int j = 0; j < 64; ++j) r[c * 64 + j] = __uint_as_float(v[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 64; ++j) r[c * 64 + j] += __uint_as_float(v[j]);
        }
      }
      return;
    }
#pragma unroll
    for (int c = 0; c < NCOL / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(taddr + (uint32_t)(c * 32), v);
      if (first) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[c * 32 + j] = __uint_as_float(v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[c * 32 + j] += __uint_as_float(v[j]);
      }
    }
  }

  // 4 x 4 transpose inside a quad of lanes: on entry lane 4a+i holds v[c] = X[row 4a+i][col c]; on exit lane 4a+b
  // holds v[i] = X[row 4a+i][col b] — four consecutive rows of ONE column, i.e. a 16-byte vector of column-major C.
  static __device__ __forceinline__ void quad_transpose(float (&v)[4], int b) {
    const bool hi2 = (b & 2) != 0, hi1 = (b & 1) != 0;
    float s0 = hi2 ? v[0] : v[2], s1 = hi2 ? v[1] : v[3];
    s0 = __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 = __shfl_xor_sync(0xffffffffu, s1, 2);
    if (hi2) { v[0] = s0; v[1] = s1; } else { v[2] = s0; v[3] = s1; }
    s0 = hi1 ? v[0] : v[1];
    s1 = hi1 ? v[2] : v[3];
    s0 = __shfl_xor_sync(0xffffffffu, s0, 1);
    s1 = __shfl_xor_sync(0xffffffffu, s1, 1);
    if (hi1) { v[0] = s0; v[2] = s1; } else { v[1] = s0; v[3] = s1; }
  }

  // Vector form of `store` (M, the leading dimensions and the base addresses are multiples of 4 elements / 16 bytes;
  // chosen per launch on the host).  The row-per-lane layout that tcgen05.ld delivers makes every access a 4-byte
  // column-strided one: 512 memory instructions per thread and tile, and with 8 warps per SM the fused epilogues
  // (y loads, C, hi, lo stores) could not keep up with the MMA of the next tile.  After a quad transpose each lane
  // owns 4 consecutive rows of every 4th column: 16-byte loads/stores, a quarter of the instructions.
  // gw = first row of this warp's 32; rows 4a .. 4a+3, columns == b (mod 4).
  __device__ __forceinline__ void store_vec(const GemmArgs& g, int64_t gw, int64_t n0, int part, int lane) {
    using SplitT = typename std::conditional<KIND == KIND_BF16X3, __nv_bfloat16, float>::type;
    const int a = lane >> 2, b = lane & 3;
    const int64_t R0 = gw + 4 * a;
    const bool row_ok = R0 < g.M;
    const float alpha = eff_alpha(g.alpha, g.a_sinv, g.b_sinv), beta = g.beta;
    const int ek = g.epi.last_launch ? g.epi.kind : EPI_NONE;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ek == EPI_FWD && g.epi.bias && row_ok) bias4 = __ldg(reinterpret_cast<const float4*>(g.epi.bias + R0));
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr int SB = 16;
#pragma unroll
    for (int c = 0; c < NCOL / SB; ++c) {
      const int64_t nb = n0 + c * SB;  // warp-uniform
      if (nb >= g.N) break;
      float x[SB / 4][4];
#pragma unroll
      for (int gq = 0; gq < SB / 4; ++gq) {
#pragma unroll
        for (int i = 0; i < 4; ++i) x[gq][i] = alpha * r[c * SB + gq * 4 + i];
        quad_transpose(x[gq], b);
      }
      if (!row_ok) continue;
      if (beta != 0.f) {
        float4 old[SB / 4];
#pragma unroll
        for (int gq = 0; gq < SB / 4; ++gq) {
          const int64_t col = nb + gq * 4 + b;
          old[gq] = col < g.N ? __ldcg(reinterpret_cast<const float4*>(g.C + R0 + col * g.ldc)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int gq = 0; gq < SB / 4; ++gq) {
          x[gq][0] += beta * old[gq].x; x[gq][1] += beta * old[gq].y;
          x[gq][2] += beta * old[gq].z; x[gq][3] += beta * old[gq].w;
        }
      }
      if (ek == EPI_FWD) {
        const int op = g.epi.op;
#pragma unroll
        for (int gq = 0; gq < SB / 4; ++gq) {
          x[gq][0] = epi_act(op, x[gq][0] + bias4.x); x[gq][1] = epi_act(op, x[gq][1] + bias4.y);
          x[gq][2] = epi_act(op, x[gq][2] + bias4.z); x[gq][3] = epi_act(op, x[gq][3] + bias4.w);
        }
      } else if (ek == EPI_PULL) {
        const int op = g.epi.op;
        float4 ys[SB / 4];
#pragma unroll
        for (int gq = 0; gq < SB / 4; ++gq) {
          const int64_t col = nb + gq * 4 + b;
          ys[gq] = col < g.N ? __ldcs(reinterpret_cast<const float4*>(g.epi.y + R0 + col * g.epi.ldy)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int gq = 0; gq < SB / 4; ++gq) {
          x[gq][0] *= nb_dunary<float>(op, 0.f, ys[gq].x); x[gq][1] *= nb_dunary<float>(op, 0.f, ys[gq].y);
          x[gq][2] *= nb_dunary<float>(op, 0.f, ys[gq].z); x[gq][3] *= nb_dunary<float>(op, 0.f, ys[gq].w);
        }
      }
#pragma unroll
      for (int gq = 0; gq < SB / 4; ++gq) {
        const int64_t col = nb + gq * 4 + b;
        if (col < g.N) {
          *reinterpret_cast<float4*>(g.C + R0 + col * g.ldc) = make_float4(x[gq][0], x[gq][1], x[gq][2], x[gq][3]);
#pragma unroll
          for (int i = 0; i < 4; ++i) rs[i] += x[gq][i];
          if (ek != EPI_NONE && g.epi.s_hi) {
            __align__(16) SplitT h[4];
            __align__(16) SplitT l[4];
            if (KIND == KIND_BF16X3 && g.epi.s_fp16) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                split1s(x[gq][i], g.epi.s_scale, *reinterpret_cast<__half*>(&h[i]), *reinterpret_cast<__half*>(&l[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) split1(x[gq][i], h[i], l[i]);
            }
            SplitT* hp = (SplitT*)g.epi.s_hi + R0 + col * g.epi.ld_s;
            SplitT* lp = (SplitT*)g.epi.s_lo + R0 + col * g.epi.ld_s;
            if (sizeof(SplitT) == 2) {
              *reinterpret_cast<uint2*>(hp) = *reinterpret_cast<const uint2*>(h);
              *reinterpret_cast<uint2*>(lp) = *reinterpret_cast<const uint2*>(l);
            } else {
              *reinterpret_cast<uint4*>(hp) = *reinterpret_cast<const uint4*>(h);
              *reinterpret_cast<uint4*>(lp) = *reinterpret_cast<const uint4*>(l);
            }
            if (KIND == KIND_BF16X3 && g.epi.s2_hi) {
              __align__(8) __nv_bfloat16 h2[4];
              __align__(8) __nv_bfloat16 l2[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) split1(x[gq][i], h2[i], l2[i]);
              *reinterpret_cast<uint2*>((__nv_bfloat16*)g.epi.s2_hi + R0 + col * g.epi.ld_s) = *reinterpret_cast<const uint2*>(h2);
              *reinterpret_cast<uint2*>((__nv_bfloat16*)g.epi.s2_lo + R0 + col * g.epi.ld_s) = *reinterpret_cast<const uint2*>(l2);
            }
          }
        }
      }
    }
    if (ek != EPI_NONE && g.epi.rowsum_partial) {
      // this lane summed columns == b (mod 4); the quad's four lanes together cover the row
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 1);
        rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 2);
      }
      if (b == 0 && row_ok)
        *reinterpret_cast<float4*>(g.epi.rowsum_partial + (int64_t)part * g.M + R0) = make_float4(rs[0], rs[1], rs[2], rs[3]);
    }
  }

  // KIND_I8: r[] holds the exact int32 accumulator bits; scaled into the Float64 C (ozaki.cu)
  __device__ __forceinline__ void store_i8(const GemmArgs& g, int64_t gm, int64_t n0) {
    if (gm >= g.M) return;
    const double rs = g.scale64 * g.row_scale[gm];
    const double beta = g.beta64;
#pragma unroll
    for (int c = 0; c < NCOL / 8; ++c) {
      const int64_t nb = n0 + c * 8;
      if (nb >= g.N) break;
      double* cp = g.C64 + gm + nb * g.ldc;
      double old[8], cs[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = nb + j < g.N;
        cs[j] = ok ? g.col_scale[nb + j] : 0.0;
        old[j] = (ok && beta != 0.0) ? cp[(int64_t)j * g.ldc] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (nb + j < g.N) cp[(int64_t)j * g.ldc] = beta * old[j] + rs * cs[j] * (double)__float_as_int(r[c * 8 + j]);
    }
  }

  // C[gm, n0 .. n0+NCOL) = epilogue(alpha * r + beta * C); `part` indexes this thread's row-sum partial
  __device__ __forceinline__ void store(const GemmArgs& g, int64_t gm, int64_t n0, int part, int lane) {
    if constexpr (kind_is_i8(KIND)) { store_i8(g, gm, n0); return; }
    if (g.vec_store) { store_vec(g, gm - lane, n0, part, lane); return; }
    using SplitT = typename std::conditional<KIND == KIND_BF16X3, __nv_bfloat16, float>::type;
    if (gm >= g.M) return;
    const float alpha = eff_alpha(g.alpha, g.a_sinv, g.b_sinv), beta = g.beta;
    const int ek = g.epi.last_launch ? g.epi.kind : EPI_NONE;
    const float bias = (ek == EPI_FWD && g.epi.bias) ? __ldg(g.epi.bias + gm) : 0.f;
    float rsum = 0.f;
    constexpr int SB = 16;  // columns per store step (bounds the registers live next to r[])
#pragma unroll
    for (int c = 0; c < NCOL / SB; ++c) {
      const int64_t nb = n0 + c * SB;
      if (nb >= g.N) break;
      float* cp = g.C + gm + nb * g.ldc;
      const int jmax = (int)min((int64_t)SB, g.N - nb);
      float x[SB];
#pragma unroll
      for (int j = 0; j < SB; ++j) x[j] = alpha * r[c * SB + j];
      if (beta != 0.f) {
        // all SB loads first (independent, in flight together), then the arithmetic
        float old[SB];
#pragma unroll
        for (int j = 0; j < SB; ++j) old[j] = j < jmax ? __ldcg(cp + (int64_t)j * g.ldc) : 0.f;
#pragma unroll
        for (int j = 0; j < SB; ++j) x[j] += beta * old[j];
      }
      if (ek == EPI_FWD) {
        const int op = g.epi.op;
#pragma unroll
        for (int j = 0; j < SB; ++j) x[j] = epi_act(op, x[j] + bias);
      } else if (ek == EPI_PULL) {
        const float* yp = g.epi.y + gm + nb * g.epi.ldy;
        const int op = g.epi.op;
        float ys[SB];
#pragma unroll
        for (int j = 0; j < SB; ++j) ys[j] = j < jmax ? __ldcs(yp + (int64_t)j * g.epi.ldy) : 0.f;
#pragma unroll
        for (int j = 0; j < SB; ++j) x[j] *= nb_dunary<float>(op, 0.f, ys[j]);
      }
#pragma unroll
      for (int j = 0; j < SB; ++j)
        if (j < jmax) { cp[(int64_t)j * g.ldc] = x[j]; rsum += x[j]; }
      if (ek != EPI_NONE && g.epi.s_hi) {
        SplitT* hp = (SplitT*)g.epi.s_hi + gm + nb * g.epi.ld_s;
        SplitT* lp = (SplitT*)g.epi.s_lo + gm + nb * g.epi.ld_s;
#pragma unroll
        for (int j = 0; j < SB; ++j) {
          SplitT h, l;
          if (KIND == KIND_BF16X3 && g.epi.s_fp16)
            split1s(x[j], g.epi.s_scale, *reinterpret_cast<__half*>(&h), *reinterpret_cast<__half*>(&l));
          else
            split1(x[j], h, l);
          if (j < jmax) { hp[(int64_t)j * g.epi.ld_s] = h; lp[(int64_t)j * g.epi.ld_s] = l; }
          if (KIND == KIND_BF16X3 && g.epi.s2_hi && j < jmax) {
            __nv_bfloat16 h2, l2;
            split1(x[j], h2, l2);
            ((__nv_bfloat16*)g.epi.s2_hi)[gm + (nb + j) * g.epi.ld_s] = h2;
            ((__nv_bfloat16*)g.epi.s2_lo)[gm + (nb + j) * g.epi.ld_s] = l2;
          }
        }
      }
    }
    if (ek != EPI_NONE && g.epi.rowsum_partial) g.epi.rowsum_partial[(int64_t)part * g.M + gm] = rsum;
  }
};

// KIND_I8ALL: Float64 running sum over the levels of the Ozaki split for this thread's row x NCOL columns
template <int NCOL>
struct EpiAcc64 {
  double r[NCOL];
  __device__ __forceinline__ void fold(uint32_t taddr, double w, bool first) {
#pragma unroll
    for (int c = 0; c < NCOL / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(taddr + (uint32_t)(c * 32), v);
      if (first) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[c * 32 + j] = w * (double)(int)v[j];
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[c * 32 + j] += w * (double)(int)v[j];
      }
    }
  }
  __device__ __forceinline__ void store(const GemmArgs& g, int64_t gm, int64_t n0) {
    if (gm >= g.M) return;
    const double rs = g.scale64 * g.row_scale[gm];
    const double beta = g.beta64;
#pragma unroll
    for (int c = 0; c < NCOL / 8; ++c) {
      const int64_t nb = n0 + c * 8;
      if (nb >= g.N) break;
      double* cp = g.C64 + gm + nb * g.ldc;
      double old[8], cs[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = nb + j < g.N;
        cs[j] = ok ? g.col_scale[nb + j] : 0.0;
        old[j] = (ok && beta != 0.0) ? cp[(int64_t)j * g.ldc] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (nb + j < g.N) cp[(int64_t)j * g.ldc] = beta * old[j] + rs * cs[j] * r[c * 8 + j];
    }
  }
};

// deterministic second phase of the fused row sums: out[i] (+)= sum over tile columns, in order
__global__ void __launch_bounds__(256) k_rowsum_finish(const float* __restrict__ partial, int nparts, int64_t M,
                                                       float* __restrict__ out, int acc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(int64_t)p * M + i];
  out[i] = acc ? out[i] + s : s;
}

template <int BN, int KIND>
struct Cfg {
  static constexpr uint32_t A_BYTES = BM * 128;
  static constexpr uint32_t B_BYTES = BN * 128;
  static constexpr uint32_t NSPLIT = KindT<KIND>::NSPLIT;
  static constexpr uint32_t STAGE_BYTES = (A_BYTES + B_BYTES) * NSPLIT;
  static constexpr int STAGES = (int)(196608u / STAGE_BYTES);  // 192 KB of operand ring
  static constexpr uint32_t SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 512 /*barriers, tile ring*/;
  static constexpr uint32_t TMEM_COLS = 2 * BN;  // two accumulator stages
};

template <int BN, int KIND>
__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
          const __grid_constant__ CUtensorMap mB0, const __grid_constant__ CUtensorMap mB1, const GemmArgs g) {
  using C_ = Cfg<BN, KIND>;
  using KT = KindT<KIND>;
  constexpr int STAGES = C_::STAGES;
  constexpr int BK = KT::BK, NPASS = KT::NPASS, MNC = KT::MN_CHUNK;
  constexpr uint32_t CHUNK_BYTES = KT::CHUNK_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + STAGES * C_::STAGE_BYTES);
  uint64_t* full = bars;                     // [STAGES]
  uint64_t* empty = bars + STAGES;           // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;       // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]
  uint64_t* sfull = bars + 2 * STAGES + 4;                 // [SCHED_SLOTS] tile index published
  uint64_t* sempty = bars + 2 * STAGES + 4 + SCHED_SLOTS;  // [SCHED_SLOTS] tile index consumed by all readers
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4 + 2 * SCHED_SLOTS);
  volatile int32_t* sched_tile = (volatile int32_t*)(tmem_slot + 2);  // [SCHED_SLOTS]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = g.tile_end > 0 ? g.tile_end : g.tiles_m * g.tiles_n;
  const int nkb = (int)((g.k_end - g.k_begin + BK - 1) / BK);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], EPI_WARPS); }
    for (int q = 0; q < SCHED_SLOTS; ++q) { mbar_init(&sfull[q], 1); mbar_init(&sempty[q], 1 + EPI_THREADS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation is warp-collective; this warp also frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(C_::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < EPI_WARP0) {
  reg_dec();
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0, q = 0;
      uint32_t ph = 0, qph = 0;
      int t = atomicAdd(g.tile_counter, 1);
      for (;;) {
        // tile scheduler: publish this tile (a sentinel >= ntiles ends every role) and request the next
        // one right away, so the atomic's round trip is hidden behind this tile's loads
        mbar_wait(&sempty[q], qph ^ 1);
        sched_tile[q] = t;
        mbar_arrive(&sfull[q]);
        if (++q == SCHED_SLOTS) { q = 0; qph ^= 1; }
        if (t >= ntiles) break;
        const int t_cur = t;
        t = atomicAdd(g.tile_counter, 1);
        int tm, tn;
        tile_coords(t_cur, g.tiles_m, g.tiles_n, g.group_m, tm, tn);
        const int m0 = tm * BM, n0 = tn * BN;
        if constexpr (KIND == KIND_I8ALL) {
          for (int c = 0; c < g.nchunk_tab; ++c)
            for (int kb = 0; kb < g.tab_nkb[c]; ++kb) {
              mbar_wait(&empty[s], ph ^ 1);
              mbar_expect_tx(&full[s], C_::STAGE_BYTES);
              uint8_t* st = smem + s * C_::STAGE_BYTES;
              tma_load_2d(st, &mA0, &full[s], g.tab_ka[c] + kb * BK, m0, 0);
              tma_load_2d(st + C_::A_BYTES, &mB0, &full[s], g.tab_kb[c] + kb * BK, n0, 0);
              if (++s == STAGES) { s = 0; ph ^= 1; }
            }
          continue;
        }
        for (int kb = 0; kb < nkb; ++kb) {  // chunk boundaries do not concern the producer
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], C_::STAGE_BYTES);
          uint8_t* st = smem + s * C_::STAGE_BYTES;
          const int k0 = (int)g.k_begin + kb * BK;
#pragma unroll
          for (int h = 0; h < (int)C_::NSPLIT; ++h) {
            uint8_t* a_dst = st + h * C_::A_BYTES;
            uint8_t* b_dst = st + C_::NSPLIT * C_::A_BYTES + h * C_::B_BYTES;
            const CUtensorMap* ma = h ? &mA1 : &mA0;
            const CUtensorMap* mb = h ? &mB1 : &mB0;
            if (g.a_mn) {
#pragma unroll
              for (int c = 0; c < BM / MNC; ++c) tma_load_2d(a_dst + c * CHUNK_BYTES, ma, &full[s], m0 + MNC * c, k0, g.hint_a);
            } else {
              tma_load_2d(a_dst, ma, &full[s], k0, m0, g.hint_a);
            }
            if (g.b_mn) {
#pragma unroll
              for (int c = 0; c < BN / MNC; ++c) tma_load_2d(b_dst + c * CHUNK_BYTES, mb, &full[s], n0 + MNC * c, k0, g.hint_b);
            } else {
              tma_load_2d(b_dst, mb, &full[s], k0, n0, g.hint_b);
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (KT::CFMT << 4) | ((KIND == KIND_BF16X3 ? g.a_fmt : KT::FMT) << 7) |
                             ((KIND == KIND_BF16X3 ? g.b_fmt : KT::FMT) << 10) | ((uint32_t)g.a_mn << 15) |
                             ((uint32_t)g.b_mn << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_kstep = g.a_mn ? KT::MN_KSTEP : 32u;  // K-major: UMMA_K elements = 32 bytes
      const uint32_t b_kstep = g.b_mn ? KT::MN_KSTEP : 32u;
      int s = 0, a = 0, q = 0;
      uint32_t ph = 0, aph = 0, qph = 0;
      for (;;) {
        mbar_wait(&sfull[q], qph);
        const int t = sched_tile[q];
        mbar_arrive(&sempty[q]);
        if (++q == SCHED_SLOTS) { q = 0; qph ^= 1; }
        if (t >= ntiles) break;
        if constexpr (KIND == KIND_I8ALL) {
          for (int c = 0; c < g.nchunk_tab; ++c) {
            mbar_wait(&tempty[a], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
            for (int kb = 0; kb < g.tab_nkb[c]; ++kb) {
              mbar_wait(&full[s], ph);
              tc_fence_after();
              const uint32_t st = smem_u32(smem + s * C_::STAGE_BYTES);
#pragma unroll
              for (int kk = 0; kk < BK / KT::UMMA_K; ++kk)
                tc_mma<KIND>(d_tmem, make_desc<KIND>(st + kk * 32u, false),
                             make_desc<KIND>(st + C_::A_BYTES + kk * 32u, false), idesc, (uint32_t)((kb | kk) != 0));
              tc_commit(&empty[s]);
              if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            tc_commit(&tfull[a]);
            if (++a == 2) { a = 0; aph ^= 1; }
          }
          continue;
        }
        for (int kb0 = 0, ci = 0; kb0 < nkb; kb0 += chunk_len(g, ci), ++ci) {
          const int kb1 = min(nkb, kb0 + chunk_len(g, ci));
          mbar_wait(&tempty[a], aph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + s * C_::STAGE_BYTES);
            const uint32_t a_hi = st, a_lo = st + C_::A_BYTES;
            const uint32_t b_hi = st + C_::NSPLIT * C_::A_BYTES, b_lo = b_hi + C_::B_BYTES;
#pragma unroll
            for (int p = 0; p < NPASS; ++p) {
              // small terms first: lo*hi, hi*lo, then hi*hi
              const uint32_t ab = NPASS == 1 ? a_hi : (p == 0 ? a_lo : a_hi);
              const uint32_t bb = NPASS == 1 ? b_hi : (p == 1 ? b_lo : b_hi);
#pragma unroll
              for (int kk = 0; kk < BK / KT::UMMA_K; ++kk) {
                tc_mma<KIND>(d_tmem, make_desc<KIND>(ab + kk * a_kstep, g.a_mn),
                             make_desc<KIND>(bb + kk * b_kstep, g.b_mn), idesc,
                             (uint32_t)(((kb - kb0) | p | kk) != 0));
              }
            }
            tc_commit(&empty[s]);  // smem stage reusable once these MMAs have read it
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
          tc_commit(&tfull[a]);  // accumulator (one K chunk) complete
          if (++a == 2) { a = 0; aph ^= 1; }
        }
      }
    }
  }
  } else {
    reg_inc();
    // ===================== epilogue (warps 4..11) =====================
    const int q = warp & 3;                   // TMEM lane quadrant this warp may access
    const int half = (warp - EPI_WARP0) >> 2;  // which BN/2 columns
    int a = 0, sq = 0;
    uint32_t aph = 0, sqph = 0;
    EpiAcc<KIND, BN / 2> acc;
    for (;;) {
      mbar_wait(&sfull[sq], sqph);
      const int t = sched_tile[sq];
      mbar_arrive(&sempty[sq]);
      if (++sq == SCHED_SLOTS) { sq = 0; sqph ^= 1; }
      if (t >= ntiles) break;
      int tm, tn;
      tile_coords(t, g.tiles_m, g.tiles_n, g.group_m, tm, tn);
      if constexpr (KIND == KIND_I8ALL) {
        EpiAcc64<BN / 2> a64;
        for (int c = 0; c < g.nchunk_tab; ++c) {
          mbar_wait(&tfull[a], aph);
          tc_fence_after();
          a64.fold(tmem_base + (uint32_t)(a * BN + half * (BN / 2)) + ((uint32_t)(q * 32) << 16), g.tab_scale[c], c == 0);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[a]);
          if (++a == 2) { a = 0; aph ^= 1; }
        }
        a64.store(g, (int64_t)tm * BM + q * 32 + lane, (int64_t)tn * BN + half * (BN / 2));
        continue;
      }
      for (int kb0 = 0, ci = 0; kb0 < nkb; kb0 += chunk_len(g, ci), ++ci) {
        mbar_wait(&tfull[a], aph);
        tc_fence_after();
        acc.fold(tmem_base + (uint32_t)(a * BN + half * (BN / 2)) + ((uint32_t)(q * 32) << 16), kb0 == 0);
        tc_fence_before();
        // one arrive per WARP: with short K chunks the fold is a latency chain between two MMA chunks, and 256
        // (pair kernel: 512, half of them remote) serialised arrives on one mbarrier were a visible part of it
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[a]);  // the accumulator is free as soon as it is in registers
        if (++a == 2) { a = 0; aph ^= 1; }
      }
      acc.store(g, (int64_t)tm * BM + q * 32 + lane, (int64_t)tn * BN + half * (BN / 2), tn * 2 + half, lane);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C_::TMEM_COLS));
  }
}

// =================================================================================================
// CTA-pair variant (cta_group::2): two CTAs of one cluster — two SMs of one TPC — share a 256 x 256
// tile.  Each CTA stages its own 128 rows of A and its own 128 columns of B; one thread of the leader
// issues tcgen05.mma.cta_group::2 (M = 256), which reads both halves of B through the pair's shared
// memories and writes rows 0-127 / 128-255 of the accumulator into the two CTAs' TMEM.  Per CTA and K
// block the smem traffic drops from 240 KB (144 KB of operand reads + 96 KB of TMA writes, above the
// 128 B/clk port) to 160 KB, which is what capped the one-CTA kernel at ~75 % tensor-pipe activity.
//   full[s]   (leader)      : both producers' TMA bytes (cta_group::2 loads signal the leader's barrier)
//   empty[s]  (both CTAs)   : tcgen05.commit multicast to the pair
//   tfull[a]  (both CTAs)   : tcgen05.commit multicast — each CTA drains its own TMEM half
//   tempty[a] (leader)      : one arrive per epilogue warp of BOTH CTAs (the peer's arrive remotely)
// =================================================================================================
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  // non-.aligned forms: the role loops leave the lanes of warps 0 and 1 diverged
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0,
                                                 int c1, uint64_t hint) {
  if (hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(leader_bar) & PEER_BIT_MASK), "r"(c0), "r"(c1),
          "l"(hint)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(leader_bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
  }
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// Arrive on a barrier of CTA `cta` of the pair WITHOUT publishing memory: handing a TMEM accumulator back to the MMA
// issuer orders tcgen05 operations only (tcgen05.wait::ld + tcgen05.fence::before_thread_sync on this side,
// tcgen05.fence::after_thread_sync on the other).  The release.cluster form above compiles to MEMBAR.ALL.GPU + ERRBAR in
// front of the arrive — ncu on the 128-chunk forward product: 31 % of all warp-stall samples sat on that fence (it
// waits for the warp's outstanding global stores of the previous tile), in the middle of the fold's latency chain.
__device__ __forceinline__ void mbar_arrive_cta_nofence(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// wait on a LOCAL barrier whose arrivals (and the data they publish) come from the peer CTA: cluster-scope acquire
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 20000000000LL) __trap();
  }
}
__device__ __forceinline__ void st_shared_cta(const void* local_addr, uint32_t cta, int32_t v) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "st.shared::cluster.u32 [ra], %2;\n\t}"
      ::"r"(smem_u32(local_addr)), "r"(cta), "r"(v)
      : "memory");
}
__device__ __forceinline__ void tc_mma_i8_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int KIND>
struct Cfg2T {
  static constexpr int BN = 256;                                  // pair tile: 256 x 256
  static constexpr uint32_t A_BYTES = 128 * 128;                  // this CTA's 128 rows of A
  static constexpr uint32_t B_BYTES = 128 * 128;                  // this CTA's 128 columns of B
  static constexpr uint32_t NS = KindT<KIND>::NSPLIT;             // hi + lo (BF16x3) or one copy (I8)
  static constexpr uint32_t STAGE_BYTES = (A_BYTES + B_BYTES) * NS;
  static constexpr int STAGES = (int)(196608u / STAGE_BYTES);     // 3 (BF16x3) / 6 (I8)
  static constexpr uint32_t DBG_BYTES = 16384;  // scratch for the traffic-sensitivity experiment (NB_GEMM_DBG_EXTRA_LOAD)
  static constexpr uint32_t SMEM = STAGES * STAGE_BYTES + 1024 + 1024 + DBG_BYTES;
  static constexpr uint32_t TMEM_COLS = 512;
};
using Cfg2 = Cfg2T<KIND_BF16X3>;

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
k_gemm_tc_pair(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
               const __grid_constant__ CUtensorMap mB0, const __grid_constant__ CUtensorMap mB1, const GemmArgs g) {
  using KT = KindT<KIND>;
  using C_ = Cfg2T<KIND>;
  constexpr int NS = (int)C_::NS;
  constexpr int STAGES = C_::STAGES, BN = C_::BN;
  constexpr int BK = KT::BK, MNC = KT::MN_CHUNK;
  constexpr uint32_t CHUNK_BYTES = KT::CHUNK_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + STAGES * C_::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  // tile scheduler ring: the leader's producer takes pair-tile indices from a global counter and publishes them to
  // every role of BOTH CTAs (sched_tile / sfull in each CTA's shared memory, written remotely for the peer); all
  // consumers release a slot on the LEADER's sempty.  A pair that shares its SMs with an NCCL kernel (the
  // overlapped allreduce) then simply takes fewer tiles instead of delaying the whole product.
  uint64_t* sfull = bars + 2 * STAGES + 4;                 // [SCHED_SLOTS]
  uint64_t* sempty = bars + 2 * STAGES + 4 + SCHED_SLOTS;  // [SCHED_SLOTS] (leader's copy is the one used)
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4 + 2 * SCHED_SLOTS);
  volatile int32_t* sched_tile = (volatile int32_t*)(tmem_slot + 2);  // [SCHED_SLOTS]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int ntiles = g.tile_end > 0 ? g.tile_end : g.tiles_m * g.tiles_n;  // pair tiles (256 x 256)
  const int nkb = (int)((g.k_end - g.k_begin + BK - 1) / BK);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * EPI_WARPS); }
    // consumers of a slot: leader MMA thread + peer producer + the epilogue threads of both CTAs
    for (int q = 0; q < SCHED_SLOTS; ++q) { mbar_init(&sfull[q], 1); mbar_init(&sempty[q], 2 + 2 * EPI_THREADS); }
    mbar_init(sempty + SCHED_SLOTS + 4, 1);  // dbg_bar
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(C_::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < EPI_WARP0) {
  reg_dec();
  if (warp == 0) {
    // ===================== TMA producer (both CTAs, each its own halves) =====================
    if (lane == 0) {
      int s = 0, q = 0;
      uint32_t ph = 0, qph = 0, dph = 0;
      bool dbg_pending = false;
      uint64_t* dbg_bar = sempty + SCHED_SLOTS + 4;  // (a spare barrier word behind the scheduler ring)
      int t_next = leader ? atomicAdd(g.tile_counter, 1) : 0;
      for (;;) {
        int t;
        if (leader) {
          // publish this tile (a sentinel >= ntiles ends every role) and request the next one right away
          t = t_next;
          mbar_wait_cluster(&sempty[q], qph ^ 1);
          sched_tile[q] = t;
          st_shared_cta((const void*)&sched_tile[q], 1, t);
          mbar_arrive(&sfull[q]);
          mbar_arrive_cta(&sfull[q], 1);  // release.cluster: the peer sees sched_tile[q]
          if (t < ntiles) t_next = atomicAdd(g.tile_counter, 1);
        } else {
          mbar_wait_cluster(&sfull[q], qph);
          t = sched_tile[q];
          mbar_arrive_cta(&sempty[q], 0);
        }
        if (++q == SCHED_SLOTS) { q = 0; qph ^= 1; }
        if (t >= ntiles) break;
        int tm, tn;
        tile_coords(t, g.tiles_m, g.tiles_n, g.group_m, tm, tn);
        const int m0 = tm * 256 + (int)rank * 128, n0 = tn * BN + (int)rank * 128;
        const bool flip = g.serpentine && ((t / (int)(gridDim.x >> 1)) & 1);
        for (int kb = 0; kb < nkb; ++kb) {
          if (g.dbg_extra_loads) {
            // L2->SM traffic sensitivity experiment: the same A-hi tile again, into scratch nobody reads
            uint8_t* scratch = smem + STAGES * C_::STAGE_BYTES + 1024;
            if (dbg_pending) { mbar_wait(dbg_bar, dph); dph ^= 1; }
            const int kx = (int)g.k_begin + (flip ? nkb - 1 - kb : kb) * BK;
            mbar_expect_tx(dbg_bar, C_::A_BYTES * (uint32_t)g.dbg_extra_loads);
            for (int e = 0; e < g.dbg_extra_loads; ++e) {
              if (g.a_mn) {
                for (int c = 0; c < 128 / MNC; ++c) tma_load_2d(scratch + c * CHUNK_BYTES, &mA0, dbg_bar, m0 + MNC * c, kx, 0);
              } else {
                tma_load_2d(scratch, &mA0, dbg_bar, kx, m0, 0);
              }
            }
            dbg_pending = true;
          }
          mbar_wait(&empty[s], ph ^ 1);
          if (leader) mbar_expect_tx(&full[s], 2 * C_::STAGE_BYTES);  // both CTAs' bytes land on this barrier
          uint8_t* st = smem + s * C_::STAGE_BYTES;
          const int k0 = (int)g.k_begin + (flip ? nkb - 1 - kb : kb) * BK;
#pragma unroll
          for (int h = 0; h < NS; ++h) {
            uint8_t* a_dst = st + h * C_::A_BYTES;
            uint8_t* b_dst = st + NS * C_::A_BYTES + h * C_::B_BYTES;
            const CUtensorMap* ma = h ? &mA1 : &mA0;
            const CUtensorMap* mb = h ? &mB1 : &mB0;
            if (g.a_mn) {
#pragma unroll
              for (int c = 0; c < 128 / MNC; ++c)
                tma_load_2d_pair(a_dst + c * CHUNK_BYTES, ma, &full[s], m0 + MNC * c, k0, g.hint_a);
            } else {
              tma_load_2d_pair(a_dst, ma, &full[s], k0, m0, g.hint_a);
            }
            if (g.b_mn) {
#pragma unroll
              for (int c = 0; c < 128 / MNC; ++c)
                tma_load_2d_pair(b_dst + c * CHUNK_BYTES, mb, &full[s], n0 + MNC * c, k0, g.hint_b);
            } else {
              tma_load_2d_pair(b_dst, mb, &full[s], k0, n0, g.hint_b);
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
      if (dbg_pending) mbar_wait(dbg_bar, dph);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = (KT::CFMT << 4) | ((KIND == KIND_BF16X3 ? g.a_fmt : KT::FMT) << 7) |
                             ((KIND == KIND_BF16X3 ? g.b_fmt : KT::FMT) << 10) | ((uint32_t)g.a_mn << 15) |
                             ((uint32_t)g.b_mn << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t a_kstep = g.a_mn ? KT::MN_KSTEP : 32u;
      const uint32_t b_kstep = g.b_mn ? KT::MN_KSTEP : 32u;
      int s = 0, a = 0, q = 0;
      uint32_t ph = 0, aph = 0, qph = 0;
      for (;;) {
        mbar_wait(&sfull[q], qph);
        const int t = sched_tile[q];
        mbar_arrive(&sempty[q]);
        if (++q == SCHED_SLOTS) { q = 0; qph ^= 1; }
        if (t >= ntiles) break;
        for (int kb0 = 0, ci = 0; kb0 < nkb; kb0 += chunk_len(g, ci), ++ci) {
          const int kb1 = min(nkb, kb0 + chunk_len(g, ci));
          mbar_wait(&tempty[a], aph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t st = smem_u32(smem + s * C_::STAGE_BYTES);
            const uint32_t a_hi = st, a_lo = st + C_::A_BYTES;
            const uint32_t b_hi = st + NS * C_::A_BYTES, b_lo = b_hi + C_::B_BYTES;
#pragma unroll
            for (int p = 0; p < KT::NPASS; ++p) {
              const uint32_t ab = KT::NPASS == 1 ? a_hi : (p == 0 ? a_lo : a_hi);
              const uint32_t bb = KT::NPASS == 1 ? b_hi : (p == 1 ? b_lo : b_hi);
#pragma unroll
              for (int kk = 0; kk < BK / KT::UMMA_K; ++kk) {
                const uint32_t accf = (uint32_t)(((kb - kb0) | p | kk) != 0);
                if constexpr (KIND == KIND_I8)
                  tc_mma_i8_pair(d_tmem, make_desc<KIND>(ab + kk * a_kstep, g.a_mn),
                                 make_desc<KIND>(bb + kk * b_kstep, g.b_mn), idesc, accf);
                else
                  tc_mma_bf16_pair(d_tmem, make_desc<KIND>(ab + kk * a_kstep, g.a_mn),
                                   make_desc<KIND>(bb + kk * b_kstep, g.b_mn), idesc, accf);
              }
            }
            tc_commit_pair(&empty[s]);  // frees the stage in BOTH CTAs
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
          tc_commit_pair(&tfull[a]);    // accumulator halves ready in both CTAs
          if (++a == 2) { a = 0; aph ^= 1; }
        }
      }
    }
  }
  } else {
    reg_inc();
    // ===================== epilogue (warps 4..11 of both CTAs): this CTA's 128 rows =====================
    const int q = warp & 3;
    const int half = (warp - EPI_WARP0) >> 2;
    int a = 0, q2 = 0;
    uint32_t aph = 0, q2ph = 0;
    EpiAcc<KIND, BN / 2> acc;
    for (;;) {
      mbar_wait_cluster(&sfull[q2], q2ph);
      const int t = sched_tile[q2];
      mbar_arrive_cta(&sempty[q2], 0);
      if (++q2 == SCHED_SLOTS) { q2 = 0; q2ph ^= 1; }
      if (t >= ntiles) break;
      int tm, tn;
      tile_coords(t, g.tiles_m, g.tiles_n, g.group_m, tm, tn);
      for (int kb0 = 0, ci = 0; kb0 < nkb; kb0 += chunk_len(g, ci), ++ci) {
        mbar_wait(&tfull[a], aph);
        tc_fence_after();
        acc.fold(tmem_base + (uint32_t)(a * BN + half * (BN / 2)) + ((uint32_t)(q * 32) << 16), kb0 == 0);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cta_nofence(&tempty[a], 0);  // the leader's barrier collects both CTAs' epilogue warps
        if (++a == 2) { a = 0; aph ^= 1; }
      }
      acc.store(g, (int64_t)tm * 256 + (int64_t)rank * 128 + q * 32 + lane, (int64_t)tn * BN + half * (BN / 2),
                tn * 2 + half, lane);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while the pair still works
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C_::TMEM_COLS));
  }
}

// ---- operand splits: column-major (rows x cols, ld_in) -> hi, lo with leading dimension ld_out --------
template <typename O>
__global__ void __launch_bounds__(256) k_split(const float* __restrict__ x, int64_t rows, int64_t cols,
                                               int64_t ld_in, O* __restrict__ hi, O* __restrict__ lo,
                                               int64_t ld_out, int vec) {
  // grid.y strides over columns, grid.x * 256 threads over groups of 4 rows
  const int64_t r4 = (rows + 3) >> 2;
  for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) {
    const float* xc = x + c * ld_in;
    O* hc = hi + c * ld_out;
    O* lc = lo + c * ld_out;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < r4; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = i << 2;
      float v[4];
      if (vec && r + 3 < rows) {
        const float4 q = *reinterpret_cast<const float4*>(xc + r);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = r + e < rows ? xc[r + e] : 0.f;
      }
      __align__(16) O h[4];
      __align__(16) O l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split1(v[e], h[e], l[e]);
      if (r + 3 < rows) {  // ld_out is a multiple of 8 and r of 4: 8/16-byte aligned vector stores
        if (sizeof(O) == 2) {
          *reinterpret_cast<uint2*>(hc + r) = *reinterpret_cast<const uint2*>(h);
          *reinterpret_cast<uint2*>(lc + r) = *reinterpret_cast<const uint2*>(l);
        } else {
          *reinterpret_cast<uint4*>(hc + r) = *reinterpret_cast<const uint4*>(h);
          *reinterpret_cast<uint4*>(lc + r) = *reinterpret_cast<const uint4*>(l);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (r + e < rows) { hc[r + e] = h[e]; lc[r + e] = l[e]; }
      }
    }
  }
}

// Scaled binary16 split (NB_MATH_FP16X3): sc[0] = power-of-two scale, sc[1] = its inverse (device-resident, written
// by k_scale_from_absmax or fixed by the producer of the array).
__global__ void __launch_bounds__(256) k_split_h(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld_in,
                                                 __half* __restrict__ hi, __half* __restrict__ lo, int64_t ld_out,
                                                 int vec, const float* __restrict__ sc) {
  const float s = __ldg(sc);
  const int64_t r4 = (rows + 3) >> 2;
  for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) {
    const float* xc = x + c * ld_in;
    __half* hc = hi + c * ld_out;
    __half* lc = lo + c * ld_out;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < r4; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = i << 2;
      float v[4];
      if (vec && r + 3 < rows) {
        const float4 q = *reinterpret_cast<const float4*>(xc + r);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = r + e < rows ? xc[r + e] : 0.f;
      }
      __align__(8) __half h[4];
      __align__(8) __half l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split1s(v[e], s, h[e], l[e]);
      if (r + 3 < rows) {
        *reinterpret_cast<uint2*>(hc + r) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(lc + r) = *reinterpret_cast<const uint2*>(l);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (r + e < rows) { hc[r + e] = h[e]; lc[r + e] = l[e]; }
      }
    }
  }
}

// C = f.(C .+ bias) in place, plus the 16-bit operand splits of the result — the forward epilogue of nb_gemm_fused run
// as its own streaming pass (read 4 B, write 4 + 4 [+ 4] B per element) when the product itself must fold its TMEM
// accumulator every 128 K: the fused epilogue (tanh + two splits, ~45 us per 128 x 256 tile) would sit between the
// folds and stall the tensor pipe; a plain store does not.  s_hi/s_lo: first split (binary16 scaled by s_scale when
// s_fp16, else bfloat16); s2_hi/s2_lo: optional second, bfloat16 split.
__global__ void __launch_bounds__(256) k_bias_act_split(float* __restrict__ C, int64_t ldc, int64_t rows, int64_t cols,
                                                        const float* __restrict__ bias, int op, uint16_t* __restrict__ s_hi,
                                                        uint16_t* __restrict__ s_lo, int64_t ld_s, int s_fp16, float s_scale,
                                                        uint16_t* __restrict__ s2_hi, uint16_t* __restrict__ s2_lo, int vec) {
  const int64_t r4 = (rows + 3) >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < r4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i << 2;
    const bool full = vec && r + 3 < rows;
    float b[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) b[e] = (bias && r + e < rows) ? __ldg(bias + r + e) : 0.f;
    constexpr int UC = 4;   // columns in flight per thread
    for (int64_t c0 = (int64_t)blockIdx.y * UC; c0 < cols; c0 += (int64_t)gridDim.y * UC) {
      float v[UC][4];
#pragma unroll
      for (int u = 0; u < UC; ++u) {
        const int64_t c = c0 + u;
        if (c < cols && full) {
          const float4 q = __ldcs(reinterpret_cast<const float4*>(C + r + c * ldc));
          v[u][0] = q.x; v[u][1] = q.y; v[u][2] = q.z; v[u][3] = q.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[u][e] = (c < cols && r + e < rows) ? C[r + e + c * ldc] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < UC; ++u) {
        const int64_t c = c0 + u;
        if (c >= cols) continue;
        __align__(8) uint16_t h[4], l[4], h2[4], l2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x = epi_act(op, v[u][e] + b[e]);
          v[u][e] = x;
          if (s_fp16) {
            __half hh, ll;
            split1s(x, s_scale, hh, ll);
            h[e] = __half_as_ushort(hh); l[e] = __half_as_ushort(ll);
          } else {
            __nv_bfloat16 hh, ll;
            split1(x, hh, ll);
            h[e] = __bfloat16_as_ushort(hh); l[e] = __bfloat16_as_ushort(ll);
          }
          __nv_bfloat16 h3, l3;
          split1(x, h3, l3);
          h2[e] = __bfloat16_as_ushort(h3); l2[e] = __bfloat16_as_ushort(l3);
        }
        if (full) {
          *reinterpret_cast<float4*>(C + r + c * ldc) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
          if (s_hi) {
            *reinterpret_cast<uint2*>(s_hi + r + c * ld_s) = *reinterpret_cast<const uint2*>(h);
            *reinterpret_cast<uint2*>(s_lo + r + c * ld_s) = *reinterpret_cast<const uint2*>(l);
          }
          if (s2_hi) {
            *reinterpret_cast<uint2*>(s2_hi + r + c * ld_s) = *reinterpret_cast<const uint2*>(h2);
            *reinterpret_cast<uint2*>(s2_lo + r + c * ld_s) = *reinterpret_cast<const uint2*>(l2);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (r + e < rows) {
              C[r + e + c * ldc] = v[u][e];
              if (s_hi) { s_hi[r + e + c * ld_s] = h[e]; s_lo[r + e + c * ld_s] = l[e]; }
              if (s2_hi) { s2_hi[r + e + c * ld_s] = h2[e]; s2_lo[r + e + c * ld_s] = l2[e]; }
            }
        }
      }
    }
  }
}

// max |x| over a column-major (rows x cols, ld) array into *out (bit pattern of a non-negative float: unsigned
// integer order == float order; the word is zeroed before the launch).  max is order-independent: deterministic.
__global__ void __launch_bounds__(256) k_absmax_f32(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld,
                                                    int vec, unsigned int* __restrict__ out) {
  float m = 0.f;
  const int64_t r4 = (rows + 3) >> 2;
  for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) {
    const float* xc = x + c * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < r4; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = i << 2;
      if (vec && r + 3 < rows) {
        const float4 q = *reinterpret_cast<const float4*>(xc + r);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(q.x), fabsf(q.y))), fmaxf(fabsf(q.z), fabsf(q.w)));
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (r + e < rows) m = fmaxf(m, fabsf(xc[r + e]));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// sc[0] = 2^(15 - e) with max|x| = f * 2^e, f in [0.5, 1)  (so max|x| * sc[0] < 2^15 < 65504), sc[1] = 1 / sc[0]
__global__ void k_scale_from_absmax(const unsigned int* __restrict__ amax, float* __restrict__ sc) {
  const float mx = __uint_as_float(*amax);
  float s = 1.f;
  if (mx > 0.f && mx < INFINITY) {
    int e;
    frexpf(mx, &e);
    int p = 15 - e;
    p = p > 100 ? 100 : (p < -100 ? -100 : p);
    s = ldexpf(1.f, p);
  }
  sc[0] = s;
  sc[1] = 1.f / s;
}

__global__ void k_set_scale(float* __restrict__ sc, float s) {
  sc[0] = s;
  sc[1] = 1.f / s;
}

// ---- host ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

enum { FMT_FP16 = 0, FMT_BF16 = 1 };  // values of the instruction descriptor's a/b format field (kind::f16)

struct SplitEntry {
  const void* src;
  int64_t rows, cols, ld_in, ld_out;
  int kind;
  int fmt;      // KIND_BF16X3 only: FMT_BF16 or FMT_FP16 (scaled binary16)
  void* hi;
  void* lo;
  float* sc;    // FMT_FP16: {scale, 1/scale} on the device (owned by the entry), else nullptr
  uint64_t stamp;
};

struct GemmState {
  EncodeTiledFn encode = nullptr;
  int32_t* tile_counter = nullptr;  // device word shared by the (stream-ordered) GEMM launches of the ctx
  std::mutex cache_mu;              // nb_release may arrive from a finalizer thread
  std::vector<SplitEntry> cache;    // LRU by stamp
  uint64_t clock = 0;
  size_t max_entries = 24;
};

GemmState* state_of(nb_ctx* ctx) {
  if (!ctx->gemm_state) {
    GemmState* st = new GemmState();
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      st->encode = (EncodeTiledFn)fn;
    else
      cudaGetLastError();
    if (cudaMalloc(&st->tile_counter, 256) != cudaSuccess) {
      cudaGetLastError();
      st->encode = nullptr;
    }
    ctx->gemm_state = st;
  }
  return (GemmState*)ctx->gemm_state;
}

// 2-D map over a column-major (d0 x d1, ld) array; box = {box0, box1}, zero OOB fill
bool make_map(GemmState* st, CUtensorMap* map, const void* base, CUtensorMapDataType dt, int esize, int64_t d0,
              int64_t d1, int64_t ld, uint32_t box0, uint32_t box1, CUtensorMapSwizzle swz) {
  cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = st->encode(map, dt, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// tile counter of the next launch: 0, or past every tile when *skip_if is set (the launch then does nothing — how a
// product decided on the device, e.g. the Float64 split's range guard in ozaki.cu, is cancelled without a host sync)
__global__ void k_counter_init(int32_t* counter, const int32_t* skip_if, int32_t start) {
  *counter = (skip_if && *skip_if) ? 0x40000000 : start;
}

static int32_t reset_counter(nb_ctx* ctx, const GemmArgs& g) {
  if (g.skip_if || g.tile_begin) {
    k_counter_init<<<1, 1, 0, ctx->stream>>>(g.tile_counter, g.skip_if, g.tile_begin);
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
  }
  NB_CUDA_OK(ctx, cudaMemsetAsync(g.tile_counter, 0, sizeof(int32_t), ctx->stream));
  return NB_OK;
}

template <int BN, int KIND>
int32_t launch(nb_ctx* ctx, const CUtensorMap maps[4], const GemmArgs& g) {
  using C_ = Cfg<BN, KIND>;
  auto kern = k_gemm_tc<BN, KIND>;
  static bool attr_done[64] = {};  /* per device: cudaFuncSetAttribute applies to the current device only */
  if (!attr_done[ctx->device & 63]) {
    NB_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C_::SMEM));
    attr_done[ctx->device & 63] = true;
  }
  const int ntiles = (g.tile_end > 0 ? g.tile_end : g.tiles_m * g.tiles_n) - g.tile_begin;
  const int grid = ntiles < ctx->num_sms ? ntiles : ctx->num_sms;
  { const int32_t rc0 = reset_counter(ctx, g); if (rc0 != NB_OK) return rc0; }
  kern<<<grid, NTHREADS, C_::SMEM, ctx->stream>>>(maps[0], maps[1], maps[2], maps[3], g);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

template <int KIND>
int32_t launch_pair(nb_ctx* ctx, const CUtensorMap maps[4], const GemmArgs& g) {
  using C2 = Cfg2T<KIND>;
  static bool attr_done[64] = {};  /* per device: cudaFuncSetAttribute applies to the current device only */
  if (!attr_done[ctx->device & 63]) {
    NB_CUDA_OK(ctx, cudaFuncSetAttribute(k_gemm_tc_pair<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2::SMEM));
    attr_done[ctx->device & 63] = true;
  }
  const int ntiles = (g.tile_end > 0 ? g.tile_end : g.tiles_m * g.tiles_n) - g.tile_begin;
  int max_pairs = ctx->num_sms / 2;
  if (const char* e = getenv("NB_GEMM_PAIRS")) max_pairs = atoi(e) > 0 && atoi(e) < max_pairs ? atoi(e) : max_pairs;
  const int pairs = ntiles < max_pairs ? ntiles : max_pairs;
  { const int32_t rc0 = reset_counter(ctx, g); if (rc0 != NB_OK) return rc0; }
  k_gemm_tc_pair<KIND><<<2 * pairs, NTHREADS, C2::SMEM, ctx->stream>>>(maps[0], maps[1], maps[2], maps[3], g);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

template <typename O>
int32_t run_split(nb_ctx* ctx, const float* x, int64_t rows, int64_t cols, int64_t ld_in, O* hi, O* lo,
                  int64_t ld_out) {
  const int vec = ((uintptr_t)x % 16 == 0) && (ld_in % 4 == 0);
  int64_t gx = ((rows + 3) / 4 + 255) / 256;
  if (gx > 64) gx = 64;
  int64_t gy = cols;
  const int64_t want = (int64_t)ctx->num_sms * 16;
  if (gx * gy > want) gy = (want + gx - 1) / gx;
  if (gy > 65535) gy = 65535;
  if (gy < 1) gy = 1;
  k_split<O><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, ctx->stream>>>(x, rows, cols, ld_in, hi, lo, ld_out, vec);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

static void entry_release(nb_ctx* ctx, const SplitEntry& e) {
  nb_pool_release(ctx, e.hi);
  nb_pool_release(ctx, e.lo);
  if (e.sc) nb_pool_release(ctx, e.sc);
}

// the cache takes ownership of e.hi / e.lo / e.sc (LRU eviction releases them to the pool)
void cache_insert(nb_ctx* ctx, GemmState* st, SplitEntry e) {
  std::vector<SplitEntry> evict;
  {
    std::lock_guard<std::mutex> lk(st->cache_mu);
    if (st->cache.size() >= st->max_entries) {
      size_t v = 0;
      for (size_t i = 1; i < st->cache.size(); ++i)
        if (st->cache[i].stamp < st->cache[v].stamp) v = i;
      evict.push_back(st->cache[v]);
      st->cache.erase(st->cache.begin() + v);
    }
    e.stamp = ++st->clock;
    st->cache.push_back(e);
  }
  for (auto& v : evict) entry_release(ctx, v);
}

// Split of the column-major source (rows x cols, ld_in), cached until the source buffer is written or released.
// `fmt` is the 16-bit format (KIND_BF16X3): both operands of one tcgen05.mma must have the SAME format (a binary16 A
// with a bfloat16 B raises an illegal-instruction fault on sm_100a — measured), so an array that is multiplied in both
// modes has both splits cached.  The entry (and the cache) owns hi / lo / sc.
// Inside a CUDA-graph capture the cache works the same way — the entries then describe buffers of the capture and
// are dropped when it ends (nb_graph_begin / nb_graph_end / nb_graph_launch clear the cache).
int32_t get_split(nb_ctx* ctx, GemmState* st, int kind, int fmt, const float* src, int64_t rows,
                  int64_t cols, int64_t ld_in, int64_t ld_out, int esize, SplitEntry* out) {
  {
    std::lock_guard<std::mutex> lk(st->cache_mu);
    for (auto& e : st->cache)
      if (e.src == src && e.rows == rows && e.cols == cols && e.ld_in == ld_in && e.kind == kind &&
          (kind != KIND_BF16X3 || e.fmt == fmt)) {
        e.stamp = ++st->clock;
        *out = e;
        return NB_OK;
      }
  }
  SplitEntry e{src, rows, cols, ld_in, ld_out, kind, kind == KIND_BF16X3 ? fmt : FMT_BF16, nullptr, nullptr, nullptr, 0};
  const size_t bytes = (size_t)(ld_out * cols) * esize;
  int32_t rc = nb_pool_alloc(ctx, bytes, &e.hi);
  if (rc != NB_OK) return rc;
  rc = nb_pool_alloc(ctx, bytes, &e.lo);
  if (rc != NB_OK) { nb_pool_release(ctx, e.hi); return rc; }
  if (kind == KIND_BF16X3 && fmt == FMT_FP16) {
    void* scp = nullptr;
    rc = nb_pool_alloc(ctx, 16, &scp);
    if (rc != NB_OK) { nb_pool_release(ctx, e.hi); nb_pool_release(ctx, e.lo); return rc; }
    e.sc = (float*)scp;
    // power-of-two scale from max|x| (one extra read of the source), then the split itself
    unsigned int* amax = (unsigned int*)(e.sc + 2);
    cudaMemsetAsync(amax, 0, sizeof(unsigned int), ctx->stream);
    const int vec = ((uintptr_t)src % 16 == 0) && (ld_in % 4 == 0);
    int64_t gx = ((rows + 3) / 4 + 255) / 256;
    if (gx > 64) gx = 64;
    int64_t gy = cols;
    const int64_t want = (int64_t)ctx->num_sms * 16;
    if (gx * gy > want) gy = (want + gx - 1) / gx;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    k_absmax_f32<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, ctx->stream>>>(src, rows, cols, ld_in, vec, amax);
    ctx->stats.kernel_launches++;
    k_scale_from_absmax<<<1, 1, 0, ctx->stream>>>(amax, e.sc);
    ctx->stats.kernel_launches++;
    k_split_h<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, ctx->stream>>>(src, rows, cols, ld_in, (__half*)e.hi,
                                                                        (__half*)e.lo, ld_out, vec, e.sc);
    ctx->stats.kernel_launches++;
    if (cudaGetLastError() != cudaSuccess) rc = nb_fail(ctx, NB_ERR_CUDA, "operand split launch failed");
  } else if (kind == KIND_BF16X3) {
    rc = run_split<__nv_bfloat16>(ctx, src, rows, cols, ld_in, (__nv_bfloat16*)e.hi, (__nv_bfloat16*)e.lo, ld_out);
  } else {
    rc = run_split<float>(ctx, src, rows, cols, ld_in, (float*)e.hi, (float*)e.lo, ld_out);
  }
  if (rc != NB_OK) { entry_release(ctx, e); return rc; }
  cache_insert(ctx, st, e);
  *out = e;
  return NB_OK;
}

}  // namespace

// A device buffer is about to be written (or recycled): cached splits of it are stale.
void nb_note_write(nb_ctx* ctx, const void* dst) {
  GemmState* st = (GemmState*)ctx->gemm_state;
  if (!st || !dst) return;
  std::vector<SplitEntry> drop;
  {
    std::lock_guard<std::mutex> lk(st->cache_mu);
    if (st->cache.empty()) return;
    // by address RANGE: the ABI accepts interior pointers (sub-matrices, ld > rows), so a write anywhere inside a
    // cached source — or a source starting inside the written block's first element — invalidates the entry
    const char* d = (const char*)dst;
    for (size_t i = 0; i < st->cache.size();) {
      const SplitEntry& e = st->cache[i];
      const char* lo = (const char*)e.src;
      const char* hi = lo + (size_t)(e.ld_in * (e.cols > 0 ? e.cols - 1 : 0) + e.rows) * sizeof(float);
      if (d >= lo && d < hi) {
        drop.push_back(e);
        st->cache.erase(st->cache.begin() + i);
      } else {
        ++i;
      }
    }
  }
  for (auto& e : drop) entry_release(ctx, e);
}

// Same, for a write of `nbytes` starting at dst (every cached source overlapping the range is dropped).
void nb_note_write_range(nb_ctx* ctx, const void* dst, size_t nbytes) {
  GemmState* st = (GemmState*)ctx->gemm_state;
  if (!st || !dst) return;
  std::vector<SplitEntry> drop;
  {
    std::lock_guard<std::mutex> lk(st->cache_mu);
    if (st->cache.empty()) return;
    const char* d0 = (const char*)dst;
    const char* d1 = d0 + (nbytes ? nbytes : 1);
    for (size_t i = 0; i < st->cache.size();) {
      const SplitEntry& e = st->cache[i];
      const char* lo = (const char*)e.src;
      const char* hi = lo + (size_t)(e.ld_in * (e.cols > 0 ? e.cols - 1 : 0) + e.rows) * sizeof(float);
      if (d0 < hi && lo < d1) {
        drop.push_back(e);
        st->cache.erase(st->cache.begin() + i);
      } else {
        ++i;
      }
    }
  }
  for (auto& e : drop) entry_release(ctx, e);
}

void nb_split_cache_clear(nb_ctx* ctx) {
  GemmState* st = (GemmState*)ctx->gemm_state;
  if (!st) return;
  std::vector<SplitEntry> all;
  {
    std::lock_guard<std::mutex> lk(st->cache_mu);
    all.swap(st->cache);
  }
  for (auto& e : all) entry_release(ctx, e);
}

// Operand-split emission for kernels outside this file (gemm_skinny.cu): buffers for the 16-bit split of an m x n
// Float32 result that the NEXT product will consume, registered in the cache once the producing launch is enqueued.
// Sensitivities (and everything unbounded) are emitted as bfloat16; *hi stays NULL when nothing should be emitted.
int32_t nb_split_emit_begin(nb_ctx* ctx, int32_t math_mode, int32_t epi_kind, int32_t epi_op, int64_t m, int64_t n,
                            void** hi, void** lo, int64_t* ld, int32_t* fp16, float* scale, float** sc) {
  *hi = *lo = nullptr;
  *sc = nullptr;
  *fp16 = 0;
  *scale = 1.f;
  (void)epi_kind; (void)epi_op;
  if (math_mode == NB_MATH_TF32 || math_mode == NB_MATH_3XTF32 || math_mode == NB_MATH_FP32) return NB_OK;
  if (math_mode != NB_MATH_BF16X3) return NB_OK;   // binary16 consumers compute their own scale (absmax pass)
  GemmState* st = state_of(ctx);
  if (!st->encode) return NB_OK;
  const int64_t e_ld = (m + 7) / 8 * 8;
  const size_t bytes = (size_t)(e_ld * n) * 2;
  if (nb_pool_alloc(ctx, bytes, hi) != NB_OK) { *hi = nullptr; return NB_OK; }
  if (nb_pool_alloc(ctx, bytes, lo) != NB_OK) { nb_pool_release(ctx, *hi); *hi = *lo = nullptr; return NB_OK; }
  *ld = e_ld;
  return NB_OK;
}

void nb_split_emit_end(nb_ctx* ctx, bool ok, const void* C, int64_t m, int64_t n, int64_t ldc, int64_t ld, void* hi,
                       void* lo, float* sc) {
  if (!hi) return;
  SplitEntry e{C, m, n, ldc, ld, KIND_BF16X3, sc ? FMT_FP16 : FMT_BF16, hi, lo, sc, 0};
  if (ok) cache_insert(ctx, state_of(ctx), e);
  else entry_release(ctx, e);
}

void nb_gemm_state_free(nb_ctx* ctx) {
  nb_split_cache_clear(ctx);
  if (ctx->gemm_state && ((GemmState*)ctx->gemm_state)->tile_counter) cudaFree(((GemmState*)ctx->gemm_state)->tile_counter);
  delete (GemmState*)ctx->gemm_state;
  ctx->gemm_state = nullptr;
}

// Returns NB_ERR_UNSUPPORTED when the problem is too small to be worth a 128-row tensor tile or (single
// pass tf32 only, which reads the caller's arrays directly) when TMA cannot describe the operands.
int32_t nb_gemm_tcgen05(nb_ctx* ctx, bool tA, bool tB, int64_t m, int64_t n, int64_t k, float alpha,
                        const float* A, int64_t lda, const float* B, int64_t ldb, float beta, float* C,
                        int64_t ldc, int32_t math_mode, const nb_gemm_epilogue* epi) {
  GemmState* st = state_of(ctx);
  const bool fused = epi && epi->kind != 0;
  if (fused && beta != 0.f) return NB_ERR_UNSUPPORTED;
  if (!st->encode) return NB_ERR_UNSUPPORTED;
  if (m * n * k < (int64_t)1 << 18) return NB_ERR_UNSUPPORTED;  // tiny: one tile would idle 147 SMs
  if (m >= (1ll << 31) || n >= (1ll << 31) || k >= (1ll << 31)) return NB_ERR_UNSUPPORTED;
  const int kind = math_mode == NB_MATH_TF32 ? KIND_TF32 : math_mode == NB_MATH_3XTF32 ? KIND_3XTF32 : KIND_BF16X3;
  // 16-bit split kinds: NB_MATH_FP16X3 (and NB_MATH_DEFAULT) splits both operands as scaled binary16 (22 bits),
  // NB_MATH_BF16X3 as bfloat16 (16 bits, any dynamic range)
  const int want_fmt = math_mode == NB_MATH_BF16X3 ? FMT_BF16 : FMT_FP16;
  if (kind == KIND_TF32) {
    if (((uintptr_t)A | (uintptr_t)B) & 15) return NB_ERR_UNSUPPORTED;
    if ((lda & 3) || (ldb & 3)) return NB_ERR_UNSUPPORTED;
  }
  const bool a_mn = !tA, b_mn = tB;
  // stored extents: A is (a_r x a_c, lda), B is (b_r x b_c, ldb)
  const int64_t a_r = tA ? k : m, a_c = tA ? m : k;
  const int64_t b_r = tB ? n : k, b_c = tB ? k : n;
  int BN = 256;
  {
    const int64_t tm = (m + BM - 1) / BM;
    if (tm * ((n + 255) / 256) < ctx->num_sms && n > 128) BN = 128;  // more, smaller tiles to fill the SMs
    if (n <= 128) BN = 128;
    if (const char* e = getenv("NB_GEMM_BN")) BN = atoi(e) == 128 ? 128 : 256;
  }
  // CTA-pair kernel (256 x 256 tiles) for the 16-bit split kinds when there are enough pair tiles to fill the 74 pairs
  bool use_pair = false;
  if (kind == KIND_BF16X3) {
    const int64_t pt = ((m + 255) / 256) * ((n + 255) / 256);
    use_pair = pt >= ctx->num_sms / 2;
    if (const char* e = getenv("NB_GEMM_2CTA")) use_pair = atoi(e) != 0;
    if (use_pair) BN = 256;
  }
  const int esize = kind == KIND_BF16X3 ? 2 : 4;
  const void* src[4] = {A, A, B, B};  // A hi, A lo, B hi, B lo
  int64_t ld_a = lda, ld_b = ldb;
  SplitEntry ea{}, eb{};
  ea.fmt = eb.fmt = FMT_BF16;
  if (kind != KIND_TF32) {
    const int64_t pad = 16 / esize;  // leading dimensions of the split copies: 16-byte multiples
    ld_a = (a_r + pad - 1) / pad * pad;
    ld_b = (b_r + pad - 1) / pad * pad;
    int32_t rc = get_split(ctx, st, kind, want_fmt, A, a_r, a_c, lda, ld_a, esize, &ea);
    if (rc != NB_OK) return rc;
    rc = get_split(ctx, st, kind, want_fmt, B, b_r, b_c, ldb, ld_b, esize, &eb);
    if (rc != NB_OK) return rc;
    src[0] = ea.hi; src[1] = ea.lo; src[2] = eb.hi; src[3] = eb.lo;
    // the operands' buffers must outlive this launch even if the cache drops the entries before it returns
    // (stream-ordered pool: a released block is only re-used by LATER work on the same stream, so no retain needed)
  }
  const CUtensorMapDataType dt = kind == KIND_BF16X3 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  // (binary16 and bfloat16 tiles are moved as the same 16-bit type: TMA only copies and zero-fills)
  const uint32_t bk = 128 / esize, mnc = 128 / esize;
  // MN-major -> inner dim is the MN extent, box {mnc, bk}; K-major -> inner dim k, box {bk, tile rows}
  const CUtensorMapSwizzle swz_mn = kind == KIND_BF16X3 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  const uint32_t abox0 = a_mn ? mnc : bk, abox1 = a_mn ? bk : BM;
  const uint32_t bbox0 = b_mn ? mnc : bk, bbox1 = b_mn ? bk : (use_pair ? 128u : (uint32_t)BN);
  CUtensorMap maps[4];
  bool ok = true;
  for (int i = 0; i < 4 && ok; ++i) {
    const bool isA = i < 2;
    ok = make_map(st, &maps[i], src[i], dt, esize, isA ? a_r : b_r, isA ? a_c : b_c, isA ? ld_a : ld_b,
                  isA ? abox0 : bbox0, isA ? abox1 : bbox1,
                  (isA ? a_mn : b_mn) ? swz_mn : CU_TENSOR_MAP_SWIZZLE_128B);
  }
  int32_t rc = NB_OK;
  if (!ok) {
    rc = NB_ERR_UNSUPPORTED;
  } else {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.M = m; g.N = n; g.K = k;
    g.a_mn = a_mn; g.b_mn = b_mn;
    g.alpha = alpha; g.beta = beta;
    g.a_fmt = (uint32_t)ea.fmt; g.b_fmt = (uint32_t)eb.fmt;
    g.a_sinv = ea.sc ? ea.sc + 1 : nullptr;
    g.b_sinv = eb.sc ? eb.sc + 1 : nullptr;
    g.C = C; g.ldc = ldc;
    const int tile_rows = use_pair ? 256 : BM;
    g.tiles_m = (int32_t)((m + tile_rows - 1) / tile_rows);
    g.tiles_n = (int32_t)((n + BN - 1) / BN);
    g.tile_counter = st->tile_counter;
    memset(&g.epi, 0, sizeof(g.epi));
    // Forward epilogue as its own streaming pass (k_bias_act_split) when this product folds every 128 K
    static const bool unfuse_on = !(getenv("NB_GEMM_UNFUSE_FWD") && atoi(getenv("NB_GEMM_UNFUSE_FWD")) == 0);
    const bool unfuse = fused && unfuse_on && epi->kind == 1 && !epi->rowsum && math_mode != NB_MATH_BF16X3 &&
                        kind == KIND_BF16X3 && k > 256;
    void* rs_partial = nullptr;
    void *e_hi = nullptr, *e_lo = nullptr, *e2_hi = nullptr, *e2_lo = nullptr;
    float* e_sc = nullptr;
    int64_t e_ld = 0;
    if (fused) {
      g.epi.kind = unfuse ? EPI_NONE : epi->kind;
      g.epi.op = epi->op;
      g.epi.bias = (const float*)epi->bias;
      g.epi.y = (const float*)epi->y;
      g.epi.ldy = epi->ldy;
      if (epi->rowsum) {
        rc = nb_pool_alloc(ctx, (size_t)g.tiles_n * 2 * (size_t)m * sizeof(float), &rs_partial);
        g.epi.rowsum_partial = (float*)rs_partial;
      }
      if (rc == NB_OK && epi->emit_split && kind != KIND_TF32) {
        const int64_t pad = 16 / esize;
        e_ld = (m + pad - 1) / pad * pad;
        const size_t bytes = (size_t)(e_ld * n) * esize;
        rc = nb_pool_alloc(ctx, bytes, &e_hi);
        if (rc == NB_OK) rc = nb_pool_alloc(ctx, bytes, &e_lo);
        if (rc != NB_OK) {  // the split is an optimisation: go on without it
          if (e_hi) nb_pool_release(ctx, e_hi);
          e_hi = e_lo = nullptr;
          rc = NB_OK;
        }
        g.epi.s_hi = e_hi; g.epi.s_lo = e_lo; g.epi.ld_s = e_ld;
        // A forward value bounded by one in magnitude (tanh, logistic) is emitted as scaled binary16 with the fixed
        // scale 2^15 when this product itself runs in the binary16 mode: the next forward product then reads 22-bit
        // operands without an absmax pass.  Everything else (sensitivities, unbounded activations) stays bfloat16.
        if (e_hi && kind == KIND_BF16X3 && want_fmt == FMT_FP16 && epi->kind == 1 &&
            (epi->op == NB_OP_TANH || epi->op == NB_OP_LOGISTIC)) {
          void* scp = nullptr;
          if (nb_pool_alloc(ctx, 16, &scp) == NB_OK) {
            e_sc = (float*)scp;
            k_set_scale<<<1, 1, 0, ctx->stream>>>(e_sc, 32768.f);
            ctx->stats.kernel_launches++;
            g.epi.s_fp16 = 1;
            g.epi.s_scale = 32768.f;
            if (epi->emit_split & 2) {  // the adjoint products will want this forward value in bfloat16 too
              if (nb_pool_alloc(ctx, bytes, &e2_hi) == NB_OK && nb_pool_alloc(ctx, bytes, &e2_lo) == NB_OK) {
                g.epi.s2_hi = e2_hi; g.epi.s2_lo = e2_lo;
              } else {
                if (e2_hi) nb_pool_release(ctx, e2_hi);
                e2_hi = e2_lo = nullptr;
              }
            }
          }
        }
      }
    }
    {
      auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
      bool v = (m % 4 == 0) && (ldc % 4 == 0) && al16(C);
      if (fused) {
        if (g.epi.bias) v = v && al16(g.epi.bias);
        if (g.epi.y) v = v && al16(g.epi.y) && (g.epi.ldy % 4 == 0);
        if (g.epi.rowsum_partial) v = v && al16(g.epi.rowsum_partial);
        if (g.epi.s_hi) v = v && al16(g.epi.s_hi) && al16(g.epi.s_lo) && (g.epi.ld_s % 8 == 0);
        if (g.epi.s2_hi) v = v && al16(g.epi.s2_hi) && al16(g.epi.s2_lo);
      }
      if (const char* e = getenv("NB_GEMM_VEC_STORE")) v = v && atoi(e) != 0;
      g.vec_store = v ? 1 : 0;
      g.serpentine = 1;  // ncu, forward 8192x32768x8192: DRAM reads 8.95 -> 7.69 GB, L2 72.8 -> 70.1 GB
      if (const char* e = getenv("NB_GEMM_SERPENTINE")) g.serpentine = atoi(e) != 0;
      g.dbg_extra_loads = 0;
      if (const char* e = getenv("NB_GEMM_DBG_EXTRA_LOAD")) g.dbg_extra_loads = atoi(e) > 0 && atoi(e) <= 4 ? atoi(e) : 0;
    }
    // long K is launched in ranges so that the operand panel of a rasterisation group fits the L2
    int64_t KRANGE = 8192;
    if (const char* e = getenv("NB_GEMM_KRANGE")) KRANGE = atoll(e) >= 64 ? atoll(e) / 64 * 64 : KRANGE;
    // K accumulated in TMEM before the epilogue folds it into its register sum.  The tensor core adds every
    // tcgen05.mma into the FP32 accumulator with TRUNCATION, a bias that grows linearly with the number of adds:
    // measured per product (8192 x 512 x 8192, tools/sweep_kchunk.py) 7.1e-6 at 2048, 1.7e-6 at 512, 8.4e-7 at 256,
    // 4.1e-7 at 128 for the binary16 split (tf32: twice that, its MMA covers half the K).  The accurate modes
    // (FP16x3, 3xTF32: forward products, whose error saturating activations amplify ~80x at width 8192) fold every
    // 128; BF16x3 (adjoint products: operand-limited at 4.5e-6 anyway) keeps 2048.  Cost of 128: -15 % on a
    // 8192 x 32768 x 8192 product (the fold's TMEM reads sit between two short MMA chunks).
    int64_t kchunk = math_mode == NB_MATH_BF16X3 ? KindT<KIND_BF16X3>::KCHUNK : 128;
    if (const char* e = getenv("NB_GEMM_KCHUNK")) kchunk = atoll(e) >= 64 ? atoll(e) / 64 * 64 : kchunk;
    if (math_mode != NB_MATH_BF16X3)
      if (const char* e = getenv("NB_GEMM_KCHUNK_ACC")) kchunk = atoll(e) >= 64 ? atoll(e) / 64 * 64 : kchunk;
    g.kchunk_kb = (int32_t)(kchunk / (128 / esize));
    // The first two chunks of a tile are issued while the epilogue warps still STORE the previous tile (they cannot
    // fold meanwhile): with 128-chunks the MMA's runway would be 2 x 128 of K (~4 us) against a ~10 us plain store,
    // and every tile would stall.  Chunks of 384 there (13 us of runway) cost little accuracy: the error variance of a
    // tile is ~ sum c_i^3, so 2 x 384^3 + 58 x 128^3 at K = 8192 gives 5.8e-7 per product instead of 4.1e-7.
    int64_t kfirst = kchunk;
    if (math_mode != NB_MATH_BF16X3 && k >= 32 * kchunk) kfirst = 3 * kchunk;  // only where 2 long chunks weigh little
    if (const char* e = getenv("NB_GEMM_KCHUNK_FIRST")) kfirst = atoll(e) >= 64 ? atoll(e) / 64 * 64 : kfirst;
    if (kfirst < kchunk) kfirst = kchunk;
    g.kchunk_first_kb = (int32_t)(kfirst / (128 / esize));
    g.hint_a = g.hint_b = 0;
    if (const char* e = getenv("NB_GEMM_HINT")) {  // 1: A evict_last  2: B evict_last  +4: the other evict_first
      const int h = atoi(e);
      if (h & 1) g.hint_a = L2_EVICT_LAST;
      if (h & 2) g.hint_b = L2_EVICT_LAST;
      if (h & 4) { if (!(h & 1)) g.hint_a = L2_EVICT_FIRST; if (!(h & 2)) g.hint_b = L2_EVICT_FIRST; }
    }
    int64_t panel_budget = 70ll << 20;  // group of 8 pair-tile rows at K = 8192 (A/B on the forward shape: 3 % faster than 5)
    if (const char* e = getenv("NB_GEMM_PANEL_MB")) panel_budget = (int64_t)atoi(e) << 20;
    const int nsplit_bytes = esize * (kind == KIND_TF32 ? 1 : 2);
    for (int64_t kb = 0; kb < k && rc == NB_OK; kb += KRANGE) {
      g.k_begin = kb;
      g.k_end = kb + KRANGE < k ? kb + KRANGE : k;
      g.epi.last_launch = g.k_end == k;
      if (kb > 0) g.beta = 1.f;
      const int64_t panel = (int64_t)tile_rows * (g.k_end - g.k_begin) * nsplit_bytes;
      int64_t gm = panel_budget / (panel > 0 ? panel : 1);
      g.group_m = (int32_t)(gm < 1 ? 1 : gm > 32 ? 32 : gm);
      if (const char* e = getenv("NB_GEMM_GROUP_M")) g.group_m = atoi(e) > 0 ? atoi(e) : g.group_m;
      auto go = [&]() -> int32_t {
        if (use_pair) return launch_pair<KIND_BF16X3>(ctx, maps, g);
        if (kind == KIND_BF16X3) return BN == 256 ? launch<256, KIND_BF16X3>(ctx, maps, g) : launch<128, KIND_BF16X3>(ctx, maps, g);
        if (kind == KIND_3XTF32) return BN == 256 ? launch<256, KIND_3XTF32>(ctx, maps, g) : launch<128, KIND_3XTF32>(ctx, maps, g);
        return BN == 256 ? launch<256, KIND_TF32>(ctx, maps, g) : launch<128, KIND_TF32>(ctx, maps, g);
      };
      // A large collective was issued just before this product (comm.cu): its NCCL kernel can only start when CTAs
      // retire, and a persistent launch holds every SM until its last tile.  The first launch therefore ends after ONE
      // wave of tiles (they start together and take the same time, so nothing idles at that boundary), the collective's
      // CTAs take the SMs they need there (high-priority stream) and the second launch runs the remaining tiles.
      const int all_tiles = g.tiles_m * g.tiles_n;
      const int wave = use_pair ? ctx->num_sms / 2 : ctx->num_sms;
      if (ctx->cut_next_gemm && kb == 0) {
        ctx->cut_next_gemm = false;
        if (all_tiles >= 3 * wave) {
          g.tile_begin = 0; g.tile_end = wave;
          rc = go();
          g.tile_begin = wave; g.tile_end = all_tiles;
          if (rc == NB_OK) rc = go();
          g.tile_begin = 0; g.tile_end = 0;
          continue;
        }
      }
      rc = go();
    }
    if (unfuse && rc == NB_OK) {
      const int vec = (((uintptr_t)C & 15) == 0) && (ldc % 4 == 0) && (!g.epi.s_hi || g.epi.ld_s % 4 == 0) &&
                      (!epi->bias || ((uintptr_t)epi->bias & 15) == 0);
      int64_t gx = ((m + 3) / 4 + 255) / 256;
      if (gx > 64) gx = 64;
      int64_t gy = (n + 3) / 4;
      const int64_t want = (int64_t)ctx->num_sms * 8;
      if (gx * gy > want) gy = (want + gx - 1) / gx;
      if (gy < 1) gy = 1;
      if (gy > 65535) gy = 65535;
      k_bias_act_split<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, ctx->stream>>>(
          C, ldc, m, n, (const float*)epi->bias, epi->op, (uint16_t*)g.epi.s_hi, (uint16_t*)g.epi.s_lo, g.epi.ld_s,
          (kind == KIND_BF16X3) ? g.epi.s_fp16 : 0, g.epi.s_scale, (uint16_t*)g.epi.s2_hi, (uint16_t*)g.epi.s2_lo, vec);
      ctx->stats.kernel_launches++;
      if (cudaGetLastError() != cudaSuccess) rc = nb_fail(ctx, NB_ERR_CUDA, "k_bias_act_split launch failed");
    }
    if (rs_partial) {
      if (rc == NB_OK) {
        k_rowsum_finish<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>((const float*)rs_partial, 2 * g.tiles_n, m,
                                                                           (float*)epi->rowsum, epi->rowsum_acc);
        ctx->stats.kernel_launches++;
        if (cudaGetLastError() != cudaSuccess) rc = nb_fail(ctx, NB_ERR_CUDA, "k_rowsum_finish launch failed");
      }
      nb_pool_release(ctx, rs_partial);
    }
    if (e_hi) {
      SplitEntry ne{C, m, n, ldc, e_ld, kind, e_sc ? FMT_FP16 : FMT_BF16, e_hi, e_lo, e_sc, 0};
      if (rc == NB_OK) cache_insert(ctx, st, ne);
      else entry_release(ctx, ne);
    }
    if (e2_hi) {
      SplitEntry n2{C, m, n, ldc, e_ld, kind, FMT_BF16, e2_hi, e2_lo, nullptr, 0};
      if (rc == NB_OK) cache_insert(ctx, st, n2);
      else entry_release(ctx, n2);
    }
  }
  return rc;
}

// One level of the Float64 Ozaki split (ozaki.cu):  C64 = beta*C64 + scale * diag(row_scale) * (A8' * B8) * diag(col_scale)
// with A8 (kp x m) and B8 (kp x n) signed 8-bit, K-major (column-major with K contiguous, leading dimensions lda8 /
// ldb8 in elements = bytes, multiples of 16), kp a multiple of 128.  The int32 accumulation is exact as long as
// kp * 2^12 < 2^31 (|slice| <= 64): checked by the caller.
int32_t nb_gemm_i8(nb_ctx* ctx, int64_t m, int64_t n, int64_t kp, const int8_t* A8, int64_t lda8, const int8_t* B8,
                   int64_t ldb8, double scale, const double* row_scale, const double* col_scale, double beta,
                   double* C, int64_t ldc, const int32_t* skip_if) {
  GemmState* st = state_of(ctx);
  if (!st->encode) return NB_ERR_UNSUPPORTED;
  if (kp % 128 || (lda8 & 15) || (ldb8 & 15) || (((uintptr_t)A8 | (uintptr_t)B8) & 15)) return NB_ERR_UNSUPPORTED;
  int BN = 256;
  if (((m + BM - 1) / BM) * ((n + 255) / 256) < ctx->num_sms && n > 128) BN = 128;
  if (n <= 128) BN = 128;
  bool use_pair = ((m + 255) / 256) * ((n + 255) / 256) >= ctx->num_sms / 2;
  if (const char* e = getenv("NB_GEMM_2CTA")) use_pair = atoi(e) != 0;
  if (use_pair) BN = 256;
  CUtensorMap maps[4];
  bool ok = make_map(st, &maps[0], A8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, kp, m, lda8, 128, BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
            make_map(st, &maps[2], B8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, kp, n, ldb8, 128, use_pair ? 128u : (uint32_t)BN,
                     CU_TENSOR_MAP_SWIZZLE_128B);
  if (!ok) return NB_ERR_UNSUPPORTED;
  maps[1] = maps[0];
  maps[3] = maps[2];
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = m; g.N = n; g.K = kp;
  g.a_mn = 0; g.b_mn = 0;
  g.alpha = 1.f; g.beta = 0.f;
  g.C = nullptr; g.ldc = ldc;
  g.tiles_m = (int32_t)((m + (use_pair ? 256 : BM) - 1) / (use_pair ? 256 : BM));
  g.tiles_n = (int32_t)((n + BN - 1) / BN);
  g.tile_counter = st->tile_counter;
  g.k_begin = 0; g.k_end = kp;
  g.kchunk_kb = (int32_t)(kp / 128);  // one exact accumulation, no folds
  g.kchunk_first_kb = g.kchunk_kb;
  g.serpentine = 1;
  const int64_t panel = (int64_t)BM * kp;
  int64_t gm = (70ll << 20) / (panel > 0 ? panel : 1);
  g.group_m = (int32_t)(gm < 1 ? 1 : gm > 32 ? 32 : gm);
  g.C64 = C; g.row_scale = row_scale; g.col_scale = col_scale; g.scale64 = scale; g.beta64 = beta;
  g.skip_if = skip_if;
  if (use_pair) return launch_pair<KIND_I8>(ctx, maps, g);
  return BN == 256 ? launch<256, KIND_I8>(ctx, maps, g) : launch<128, KIND_I8>(ctx, maps, g);
}

// Every level of the Ozaki split in ONE launch (KIND_I8ALL): A8 (S*kp x m) and B8 (S*kp x n) as in nb_gemm_i8, levels
// 0 .. S-1 (A slices 0..l against B slices l..0) plus, with extra_level, level S (A slices 1..S-1 against B slices
// S-1..1: five more passes, error 2^-7 smaller).  C64 = beta*C64 + alpha * diag(row_scale) * (sum of levels) * diag(col_scale).
int32_t nb_gemm_i8all(nb_ctx* ctx, int64_t m, int64_t n, int64_t kp, int32_t S, int32_t extra_level, const int8_t* A8,
                      int64_t ld8, const int8_t* B8, double alpha, const double* row_scale, const double* col_scale,
                      double beta, double* C, int64_t ldc, const int32_t* skip_if) {
  GemmState* st = state_of(ctx);
  if (!st->encode) return NB_ERR_UNSUPPORTED;
  if (kp % 128 || (ld8 & 15) || (((uintptr_t)A8 | (uintptr_t)B8) & 15) || S < 1 || S + (extra_level ? 1 : 0) > 8)
    return NB_ERR_UNSUPPORTED;
  constexpr int BNL = 128;  // Float64 running sums: 64 columns per epilogue thread
  CUtensorMap maps[4];
  bool ok = make_map(st, &maps[0], A8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, (int64_t)S * kp, m, ld8, 128, BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
            make_map(st, &maps[2], B8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, (int64_t)S * kp, n, ld8, 128, BNL, CU_TENSOR_MAP_SWIZZLE_128B);
  if (!ok) return NB_ERR_UNSUPPORTED;
  maps[1] = maps[0];
  maps[3] = maps[2];
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = m; g.N = n; g.K = kp;
  g.alpha = 1.f;
  g.ldc = ldc;
  g.tiles_m = (int32_t)((m + BM - 1) / BM);
  g.tiles_n = (int32_t)((n + BNL - 1) / BNL);
  g.tile_counter = st->tile_counter;
  g.k_begin = 0; g.k_end = kp;
  g.kchunk_kb = 1;
  g.kchunk_first_kb = 1;
  g.group_m = 8;
  g.C64 = C; g.row_scale = row_scale; g.col_scale = col_scale; g.scale64 = alpha; g.beta64 = beta;
  const int32_t nkb0 = (int32_t)(kp / 128);
  int c = 0;
  for (int l = 0; l < S; ++l, ++c) {
    g.tab_nkb[c] = (l + 1) * nkb0;
    g.tab_ka[c] = 0;
    g.tab_kb[c] = (int32_t)((S - 1 - l) * kp);
    g.tab_scale[c] = ldexp(1.0, -7 * (l + 2));
  }
  if (extra_level && S > 1) {
    g.tab_nkb[c] = (S - 1) * nkb0;
    g.tab_ka[c] = (int32_t)kp;
    g.tab_kb[c] = 0;
    g.tab_scale[c] = ldexp(1.0, -7 * (S + 2));
    ++c;
  }
  g.nchunk_tab = c;
  g.skip_if = skip_if;
  return launch<BNL, KIND_I8ALL>(ctx, maps, g);
}
