// Batch-sharded gradient exchange: NCCL allreduce(sum) of parameter sensitivities on the sweep
// stream (the reference has no distributed layer at all).
//
// libnccl is dlopen'ed at nb_comm_init time — the torch-bundled libnccl.so.2 is already mapped
// when the host is Python, and a Julia host finds the system libnccl — so libnabla_cuda.so itself
// has no link-time NCCL dependency and loads on a CPU-only box.
#include <dlfcn.h>

#include "common.cuh"

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8, ncclSum = 0 };

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      a.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (a.h) break;
    }
    if (!a.h) return;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.h, "ncclAllReduce");
    a.GroupStart = (decltype(a.GroupStart))dlsym(a.h, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.h, "ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GroupStart &&
           a.GroupEnd && a.GetErrorString;
  });
  return a;
}

int32_t nccl_fail(nb_ctx* ctx, const char* what, ncclResult_t r) {
  return nb_fail(ctx, NB_ERR_COMM, std::string(what) + ": " + api().GetErrorString(r));
}
}  // namespace

extern "C" {

int32_t nb_comm_unique_id(uint8_t id_out[128]) {
  if (!api().ok) return nb_fail(nullptr, NB_ERR_COMM, "libnccl.so.2 could not be loaded");
  ncclUniqueId id;
  ncclResult_t r = api().GetUniqueId(&id);
  if (r != 0) return nccl_fail(nullptr, "ncclGetUniqueId", r);
  memcpy(id_out, id.internal, 128);
  return NB_OK;
}

int32_t nb_comm_init(nb_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t id_in[128]) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && id_in, NB_ERR_INVALID, "NULL argument");
  NB_REQUIRE(ctx, nranks >= 1 && rank >= 0 && rank < nranks, NB_ERR_INVALID, "bad rank/nranks");
  NB_REQUIRE(ctx, api().ok, NB_ERR_COMM, "libnccl.so.2 could not be loaded");
  NB_REQUIRE(ctx, !ctx->nccl_comm, NB_ERR_INVALID, "communicator already initialised");
  NB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(id.internal, id_in, 128);
  ncclComm_t comm = nullptr;
  ncclResult_t r = api().CommInitRank(&comm, nranks, id, rank);
  if (r != 0) return nccl_fail(ctx, "ncclCommInitRank", r);
  ctx->nccl_comm = comm;
  ctx->nranks = nranks;
  ctx->rank = rank;
  return NB_OK;
}

int32_t nb_allreduce_sum(nb_ctx* ctx, int32_t n_bufs, void* const* bufs, const int64_t* counts,
                         int32_t dtype) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && (n_bufs == 0 || (bufs && counts)), NB_ERR_INVALID, "NULL argument");
  NB_REQUIRE(ctx, dtype == NB_F32 || dtype == NB_F64, NB_ERR_INVALID, "bad dtype");
  if (ctx->nranks == 1 && !ctx->nccl_comm) return NB_OK;  // single rank: the sum is the identity
  NB_REQUIRE(ctx, ctx->nccl_comm, NB_ERR_COMM, "nb_allreduce_sum: nb_comm_init was not called");
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  for (int i = 0; i < n_bufs; ++i) nb_note_write(ctx, bufs[i]);
  ncclResult_t r = api().GroupStart();
  if (r != 0) return nccl_fail(ctx, "ncclGroupStart", r);
  for (int i = 0; i < n_bufs; ++i) {
    r = api().AllReduce(bufs[i], bufs[i], (size_t)counts[i], dtype == NB_F64 ? ncclFloat64 : ncclFloat32,
                        ncclSum, comm, ctx->stream);
    if (r != 0) {
      api().GroupEnd();
      return nccl_fail(ctx, "ncclAllReduce", r);
    }
  }
  r = api().GroupEnd();
  if (r != 0) return nccl_fail(ctx, "ncclGroupEnd", r);
  ctx->stats.kernel_launches++;
  return NB_OK;
}

static int32_t reduce_on(nb_ctx* ctx, cudaStream_t st, int32_t n_bufs, void* const* bufs, const int64_t* counts,
                         int32_t dtype) {
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  for (int i = 0; i < n_bufs; ++i) nb_note_write(ctx, bufs[i]);
  ncclResult_t r = api().GroupStart();
  if (r != 0) return nccl_fail(ctx, "ncclGroupStart", r);
  for (int i = 0; i < n_bufs; ++i) {
    r = api().AllReduce(bufs[i], bufs[i], (size_t)counts[i], dtype == NB_F64 ? ncclFloat64 : ncclFloat32,
                        ncclSum, comm, st);
    if (r != 0) {
      api().GroupEnd();
      return nccl_fail(ctx, "ncclAllReduce", r);
    }
  }
  r = api().GroupEnd();
  if (r != 0) return nccl_fail(ctx, "ncclGroupEnd", r);
  ctx->stats.kernel_launches++;
  return NB_OK;
}

int32_t nb_allreduce_sum_async(nb_ctx* ctx, int32_t n_bufs, void* const* bufs, const int64_t* counts,
                               int32_t dtype) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && (n_bufs == 0 || (bufs && counts)), NB_ERR_INVALID, "NULL argument");
  NB_REQUIRE(ctx, dtype == NB_F32 || dtype == NB_F64, NB_ERR_INVALID, "bad dtype");
  if (ctx->nranks == 1 && !ctx->nccl_comm) return NB_OK;
  NB_REQUIRE(ctx, ctx->nccl_comm, NB_ERR_COMM, "nb_allreduce_sum_async: nb_comm_init was not called");
  // (inside nb_graph_begin/end the fork below pulls the communicator stream into the capture; NCCL collectives are
  //  capturable, and nb_comm_join — required before nb_graph_end — joins it back)
  if (!ctx->comm_stream) {
    // highest priority: the persistent GEMM launches of the sweep take every SM, so a collective can only start when
    // CTAs retire — its CTAs must then win the freed SMs against the next product's
    int prio_lo = 0, prio_hi = 0;
    NB_CUDA_OK(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    NB_CUDA_OK(ctx, cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, prio_hi));
    NB_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    NB_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  }
  NB_CUDA_OK(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
  NB_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_fork, 0));
  ctx->comm_pending = true;
  {
    int64_t bytes = 0;
    for (int i = 0; i < n_bufs; ++i) bytes += counts[i] * (dtype == NB_F64 ? 8 : 4);
    // OFF by default (NB_COMM_CUT_BYTES=<bytes> turns it on): measured on 4 and 8 B200s the exchange then hides behind
    // the sweep (join wait 1.21 -> 0.61 ms at N = 8) but its CTAs — each one blocks a whole CTA PAIR of the GEMM —
    // slow the overlapped product by as much (profiles/README.md, round 2): NCCL's SM time is paid either way
    static const int64_t cut_bytes = getenv("NB_COMM_CUT_BYTES") ? atoll(getenv("NB_COMM_CUT_BYTES")) : 0;
    if (cut_bytes > 0 && bytes >= cut_bytes) ctx->cut_next_gemm = true;
  }
  return reduce_on(ctx, ctx->comm_stream, n_bufs, bufs, counts, dtype);
}

int32_t nb_comm_join(nb_ctx* ctx) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  if (!ctx->comm_pending) return NB_OK;
  NB_CUDA_OK(ctx, cudaEventRecord(ctx->ev_join, ctx->comm_stream));
  NB_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  ctx->comm_pending = false;
  return NB_OK;
}

int32_t nb_comm_destroy(nb_ctx* ctx) {
  if (ctx && ctx->comm_stream) {
    cudaStreamSynchronize(ctx->comm_stream);
    cudaStreamDestroy(ctx->comm_stream);
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
    ctx->comm_stream = nullptr;
    ctx->ev_fork = ctx->ev_join = nullptr;
    ctx->comm_pending = false;
  }
  if (ctx && ctx->nccl_comm) {
    api().CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
  }
  return NB_OK;
}

}  // extern "C"
