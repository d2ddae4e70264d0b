// Shared internals of libnabla_cuda.so: context, pooled stream-ordered allocator, error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/nabla_cuda.h"

#define NB_NUM_SMS_DEFAULT 148

struct nb_block {
  size_t size;  // size class actually allocated
  int refs;
  uint32_t tag;  // 0 = general pool; otherwise the id of the captured graph that owns the block
};

struct nb_ctx {
  int device = 0;
  int num_sms = NB_NUM_SMS_DEFAULT;
  cudaStream_t stream = nullptr;
  std::mutex mu;  // guards the pool maps (nb_release may come from a GC finalizer thread)
  std::unordered_map<void*, nb_block> live;
  std::map<size_t, std::vector<void*>> free_bins;
  size_t free_bytes = 0;
  size_t max_cached = 0;  // 0 = unlimited
  nb_stats stats{};
  std::string err;
  bool capturing = false;
  // Blocks handed out while a sweep is being captured into a CUDA graph are baked into the graph
  // by address, so they must never serve anyone else while the graph lives: they are tagged with
  // the graph id, recycled only inside that capture (cap_bins) and parked until nb_graph_destroy.
  uint32_t cap_tag = 0, next_tag = 1;
  uint64_t cap_launch_base = 0;
  std::map<size_t, std::vector<void*>> cap_bins;
  std::unordered_map<uint32_t, std::vector<std::pair<void*, size_t>>> parked;
  // communicator (comm.cu)
  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  cudaStream_t comm_stream = nullptr;      // overlapped reductions (nb_allreduce_sum_async)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool comm_pending = false;
  bool cut_next_gemm = false;  // a large collective was just issued: the next tensor-core product ends its first launch
                               // after ONE wave of tiles, so the NCCL kernel finds free SMs a tile time later instead
                               // of after the whole persistent launch (gemm_tc.cu launch ranges)
  // tcgen05 GEMM staging (gemm_tf32.cu)
  void* gemm_state = nullptr;
};

struct nb_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  uint64_t launches_per_replay = 0;
  nb_ctx* ctx = nullptr;
  uint32_t tag = 0;
};

extern thread_local std::string nb_tls_error;

static inline int32_t nb_fail(nb_ctx* ctx, int32_t code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  nb_tls_error = msg;
  return code;
}

// Every entry point makes the context's device current for the calling thread first: kernel launches, cudaMalloc and
// cudaFuncSetAttribute act on the CURRENT device, one host thread may drive contexts on several GPUs, and other
// libraries in the process (the host's own CUDA code) may have switched devices in between.  cudaSetDevice to the
// device that is already current is a thread-local compare inside the runtime.
static inline void nb_enter(nb_ctx* ctx) {
  if (ctx) cudaSetDevice(ctx->device);
}

#define NB_CUDA_OK(ctx, expr)                                                            \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      return nb_fail((ctx), NB_ERR_CUDA,                                                 \
                     std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ \
                         ":" + std::to_string(__LINE__) + ")");                          \
    }                                                                                    \
  } while (0)

#define NB_REQUIRE(ctx, cond, code, msg)                 \
  do {                                                   \
    if (!(cond)) return nb_fail((ctx), (code), (msg));   \
  } while (0)

#define NB_LAUNCH_CHECK(ctx)                    \
  do {                                          \
    (ctx)->stats.kernel_launches++;             \
    NB_CUDA_OK((ctx), cudaGetLastError());      \
  } while (0)

static inline size_t nb_dtype_size(int32_t dt) { return dt == NB_F64 ? 8 : 4; }

// Operand-split cache of the tensor-core GEMM (gemm_tc.cu): every array of a sweep is multiplied twice
// (W in the forward product and in W'*Ybar, an activation in the forward product and in Ybar*h', Ybar in
// both adjoints), so the hi/lo copies are kept until the source buffer is written or released.  Every
// entry point that writes device memory reports the destination here.
void nb_note_write(nb_ctx* ctx, const void* dst);
void nb_note_write_range(nb_ctx* ctx, const void* dst, size_t nbytes);
void nb_split_cache_clear(nb_ctx* ctx);

// internal allocator entry points (ctx.cu)
int32_t nb_pool_alloc(nb_ctx* ctx, size_t nbytes, void** out);
int32_t nb_pool_release(nb_ctx* ctx, void* p);

static inline int64_t nb_numel(const nb_tensor* t) {
  int64_t n = 1;
  for (int d = 0; d < t->ndim; ++d) n *= t->shape[d];
  return n;
}
