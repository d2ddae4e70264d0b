// Array-structure rules: sub-matrix copies behind hcat / vcat / range getindex
// and their pullbacks, and gather / scatter-add behind getindex with an index vector.
//
// Reference: src/sensitivities/chainrules_ignored.jl:12-40 (hcat/vcat pullbacks are `copy(selectdim(ȳ,…))`),
// src/sensitivities/indexing.jl:3-11 + ChainRules rrule(getindex): x̄ = zero(x); x̄[inds...] .+= ȳ
// (duplicated indices accumulate: test/sensitivities/indexing.jl "Overlapping indices (#139)").
#include "common.cuh"

namespace {

template <typename T>
__global__ void __launch_bounds__(256) k_copy2d(int64_t rows, int64_t cols, const T* __restrict__ src, int64_t lds,
                                                T* __restrict__ dst, int64_t ldd, int acc) {
  // grid.y strides over columns, threads over rows (contiguous): coalesced on both sides
  for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) {
    const T* s = src + c * lds;
    T* d = dst + c * ldd;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
      d[r] = acc ? d[r] + s[r] : s[r];
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_gather(int64_t n, const T* __restrict__ src, const int64_t* __restrict__ idx,
                                                T* __restrict__ dst) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[idx[i]];
}

template <typename T>
__global__ void __launch_bounds__(256) k_scatter_add(int64_t n, const T* __restrict__ src,
                                                     const int64_t* __restrict__ idx, T* __restrict__ dst) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&dst[idx[i]], src[i]);
}

}  // namespace

extern "C" {

int32_t nb_copy2d(nb_ctx* ctx, int64_t rows, int64_t cols, const void* src, int64_t src_off, int64_t ld_src,
                  void* dst, int64_t dst_off, int64_t ld_dst, int32_t dtype, int32_t accumulate) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  NB_REQUIRE(ctx, rows >= 0 && cols >= 0 && src_off >= 0 && dst_off >= 0, NB_ERR_INVALID, "nb_copy2d: negative extent");
  NB_REQUIRE(ctx, dtype == NB_F32 || dtype == NB_F64, NB_ERR_INVALID, "bad dtype");
  if (rows == 0 || cols == 0) return NB_OK;
  NB_REQUIRE(ctx, src && dst, NB_ERR_INVALID, "nb_copy2d: NULL buffer");
  NB_REQUIRE(ctx, (cols == 1 || (ld_src >= rows && ld_dst >= rows)), NB_ERR_SHAPE, "nb_copy2d: leading dimension too small");
  nb_note_write(ctx, dst);
  int64_t gx = (rows + 255) / 256;
  if (gx > 64) gx = 64;
  int64_t gy = cols;
  const int64_t want = (int64_t)ctx->num_sms * 16;
  if (gx * gy > want) gy = (want + gx - 1) / gx;
  if (gy > 65535) gy = 65535;
  const dim3 grid((unsigned)gx, (unsigned)gy);
  if (dtype == NB_F32)
    k_copy2d<float><<<grid, 256, 0, ctx->stream>>>(rows, cols, (const float*)src + src_off, ld_src, (float*)dst + dst_off, ld_dst, accumulate);
  else
    k_copy2d<double><<<grid, 256, 0, ctx->stream>>>(rows, cols, (const double*)src + src_off, ld_src, (double*)dst + dst_off, ld_dst, accumulate);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int32_t nb_gather(nb_ctx* ctx, int64_t n, const void* src, const int64_t* idx_dev, void* dst, int32_t dtype) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && (n == 0 || (src && idx_dev && dst)), NB_ERR_INVALID, "NULL argument");
  NB_REQUIRE(ctx, dtype == NB_F32 || dtype == NB_F64, NB_ERR_INVALID, "bad dtype");
  if (n == 0) return NB_OK;
  nb_note_write(ctx, dst);
  int64_t blocks = (n + 255) / 256;
  if (blocks > ctx->num_sms * 8) blocks = ctx->num_sms * 8;
  if (dtype == NB_F32) k_gather<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, (const float*)src, idx_dev, (float*)dst);
  else k_gather<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, (const double*)src, idx_dev, (double*)dst);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int32_t nb_scatter_add(nb_ctx* ctx, int64_t n, const void* src, const int64_t* idx_dev, void* dst, int32_t dtype) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && (n == 0 || (src && idx_dev && dst)), NB_ERR_INVALID, "NULL argument");
  NB_REQUIRE(ctx, dtype == NB_F32 || dtype == NB_F64, NB_ERR_INVALID, "bad dtype");
  if (n == 0) return NB_OK;
  nb_note_write(ctx, dst);
  int64_t blocks = (n + 255) / 256;
  if (blocks > ctx->num_sms * 8) blocks = ctx->num_sms * 8;
  if (dtype == NB_F32) k_scatter_add<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, (const float*)src, idx_dev, (float*)dst);
  else k_scatter_add<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, (const double*)src, idx_dev, (double*)dst);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

}  // extern "C"
