// Context, pooled stream-ordered allocator, copies, events and CUDA-graph capture.
//
// Replaces what Julia's GC does for the reference: every Branch value and every sensitivity is an
// `Array` allocated on demand (src/sensitivities/functional/functional.jl:62,77,85; core.jl:89).
// Here buffers are refcounted blocks from size-class bins; because all work of a sweep is ordered
// on ONE stream (the reference sweep is sequential, core.jl:160-166), a released block may be
// handed out again immediately: later kernels are stream-ordered after its last reader.
#include "common.cuh"

thread_local std::string nb_tls_error;

void nb_gemm_state_free(nb_ctx* ctx);  // gemm_tf32.cu

static size_t size_class(size_t n) {
  if (n < 512) return 512;
  if (n <= (1u << 20)) return (n + 511) & ~size_t(511);
  // keep 3 mantissa bits: waste <= 12.5 %
  int lg = 63 - __builtin_clzll((unsigned long long)n);
  size_t step = size_t(1) << (lg - 3);
  return (n + step - 1) & ~(step - 1);
}

static void trim_locked(nb_ctx* ctx) {
  for (auto& kv : ctx->free_bins)
    for (void* p : kv.second) cudaFree(p);
  ctx->free_bins.clear();
  ctx->stats.pool_bytes_reserved -= ctx->free_bytes;
  ctx->free_bytes = 0;
}

int32_t nb_pool_alloc(nb_ctx* ctx, size_t nbytes, void** out) {
  nb_enter(ctx);
  size_t cls = size_class(nbytes ? nbytes : 1);
  std::lock_guard<std::mutex> lk(ctx->mu);
  void* p = nullptr;
  if (ctx->cap_tag) {  // capture-private recycling first
    auto ct = ctx->cap_bins.find(cls);
    if (ct != ctx->cap_bins.end() && !ct->second.empty()) {
      p = ct->second.back();
      ct->second.pop_back();
    }
  }
  if (!p) {
    auto it = ctx->free_bins.find(cls);
    if (it != ctx->free_bins.end() && !it->second.empty()) {
      p = it->second.back();
      it->second.pop_back();
      ctx->free_bytes -= cls;
    } else {
      cudaError_t e = cudaMalloc(&p, cls);
      if (e != cudaSuccess && !ctx->cap_tag) {
        cudaGetLastError();
        trim_locked(ctx);
        e = cudaMalloc(&p, cls);
      }
      if (e != cudaSuccess) {
        cudaGetLastError();
        return nb_fail(ctx, NB_ERR_OOM, "nb_alloc: out of device memory requesting " +
                                            std::to_string(cls) + " bytes");
      }
      ctx->stats.cuda_mallocs++;
      ctx->stats.pool_bytes_reserved += cls;
    }
  }
  ctx->live[p] = nb_block{cls, 1, ctx->cap_tag};
  ctx->stats.pool_bytes_in_use += cls;
  if (ctx->stats.pool_bytes_in_use > ctx->stats.pool_high_water)
    ctx->stats.pool_high_water = ctx->stats.pool_bytes_in_use;
  *out = p;
  return NB_OK;
}

int32_t nb_pool_release(nb_ctx* ctx, void* p) {
  size_t block_bytes = 0;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto it = ctx->live.find(p);
    if (it == ctx->live.end()) return nb_fail(ctx, NB_ERR_INVALID, "nb_release: unknown buffer");
    if (it->second.refs > 1) {
      --it->second.refs;
      return NB_OK;
    }
    block_bytes = it->second.size;
  }
  // the block is about to be recycled: cached splits of any array inside it (interior pointers included) are stale
  nb_note_write_range(ctx, p, block_bytes);
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->live.find(p);
  if (it == ctx->live.end()) return nb_fail(ctx, NB_ERR_INVALID, "nb_release: unknown buffer");
  if (--it->second.refs > 0) return NB_OK;
  const size_t cls = it->second.size;
  const uint32_t tag = it->second.tag;
  ctx->live.erase(it);
  ctx->stats.pool_bytes_in_use -= cls;
  if (tag) {
    if (tag == ctx->cap_tag) ctx->cap_bins[cls].push_back(p);   // reusable inside this capture
    else if (ctx->parked.count(tag)) ctx->parked[tag].push_back({p, cls});  // graph still alive
    else { ctx->free_bins[cls].push_back(p); ctx->free_bytes += cls; }
    return NB_OK;
  }
  ctx->free_bins[cls].push_back(p);
  ctx->free_bytes += cls;
  if (ctx->max_cached && ctx->free_bytes > ctx->max_cached && !ctx->capturing) {
    cudaStreamSynchronize(ctx->stream);
    trim_locked(ctx);
  }
  return NB_OK;
}

extern "C" {

int32_t nb_abi_version(void) { return NB_ABI_VERSION; }

const char* nb_last_error(nb_ctx* ctx) { return ctx ? ctx->err.c_str() : nb_tls_error.c_str(); }

int32_t nb_device_count(int32_t* out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *out = 0;
    return nb_fail(nullptr, NB_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *out = n;
  return NB_OK;
}

int32_t nb_ctx_create(int32_t device, uint64_t pool_bytes, nb_ctx** out) {
  if (!out) return nb_fail(nullptr, NB_ERR_INVALID, "nb_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return nb_fail(nullptr, NB_ERR_CUDA,
                   "nb_ctx_create: no CUDA device available (libnabla_cuda has no CPU fallback)");
  }
  if (device < 0 || device >= n) return nb_fail(nullptr, NB_ERR_INVALID, "nb_ctx_create: bad device");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return nb_fail(nullptr, NB_ERR_CUDA, cudaGetErrorString(e));
  nb_ctx* ctx = new nb_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->num_sms = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete ctx;
    return nb_fail(nullptr, NB_ERR_CUDA, cudaGetErrorString(e));
  }
  ctx->max_cached = 0;
  if (pool_bytes) {  // pre-reserve one block so the first sweep does not pay cudaMalloc latency
    void* p = nullptr;
    if (nb_pool_alloc(ctx, pool_bytes, &p) == NB_OK) nb_pool_release(ctx, p);
  }
  *out = ctx;
  return NB_OK;
}

int32_t nb_comm_destroy(nb_ctx* ctx);

int32_t nb_ctx_destroy(nb_ctx* ctx) {
  if (!ctx) return NB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  nb_comm_destroy(ctx);
  nb_gemm_state_free(ctx);
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    trim_locked(ctx);
    for (auto& kv : ctx->live) cudaFree(kv.first);
    ctx->live.clear();
    for (auto& kv : ctx->parked)
      for (auto& pr : kv.second) cudaFree(pr.first);
    ctx->parked.clear();
    for (auto& kv : ctx->cap_bins)
      for (void* p : kv.second) cudaFree(p);
    ctx->cap_bins.clear();
  }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return NB_OK;
}

int32_t nb_sweep_begin(nb_ctx* ctx) {
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  NB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return NB_OK;
}

int32_t nb_sync(nb_ctx* ctx) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  // (a synchronisation inside nb_graph_begin / nb_graph_end would invalidate the capture with an opaque CUDA error)
  NB_REQUIRE(ctx, !ctx->capturing, NB_ERR_INVALID,
             "nb_sync: the context is capturing a CUDA graph — the captured function read a device value on the host or "
             "uploaded pageable host data (an index vector, a NumPy operand); create such operands before the capture");
  NB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return NB_OK;
}

int32_t nb_sweep_end(nb_ctx* ctx) { return nb_sync(ctx); }

int32_t nb_ctx_stream(nb_ctx* ctx, void** out) {
  NB_REQUIRE(ctx, ctx && out, NB_ERR_INVALID, "NULL argument");
  *out = (void*)ctx->stream;
  return NB_OK;
}

int32_t nb_ctx_stats(nb_ctx* ctx, nb_stats* out) {
  NB_REQUIRE(ctx, ctx && out, NB_ERR_INVALID, "NULL argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  *out = ctx->stats;
  return NB_OK;
}

int32_t nb_alloc(nb_ctx* ctx, uint64_t nbytes, void** out) {
  NB_REQUIRE(ctx, ctx && out, NB_ERR_INVALID, "NULL argument");
  return nb_pool_alloc(ctx, nbytes, out);
}

int32_t nb_retain(nb_ctx* ctx, void* buf) {
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->live.find(buf);
  if (it == ctx->live.end()) return nb_fail(ctx, NB_ERR_INVALID, "nb_retain: unknown buffer");
  it->second.refs++;
  return NB_OK;
}

int32_t nb_release(nb_ctx* ctx, void* buf) {
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  return nb_pool_release(ctx, buf);
}

int32_t nb_refcount(nb_ctx* ctx, void* buf, int32_t* out) {
  NB_REQUIRE(ctx, ctx && out, NB_ERR_INVALID, "NULL argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto it = ctx->live.find(buf);
  *out = it == ctx->live.end() ? 0 : it->second.refs;
  return NB_OK;
}

int32_t nb_h2d(nb_ctx* ctx, void* dst, const void* src, uint64_t nbytes) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  if (!nbytes) return NB_OK;
  if (ctx->capturing) {   // only pinned memory can be the source of a captured copy
    cudaPointerAttributes at{};
    const bool pinned = cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    NB_REQUIRE(ctx, pinned, NB_ERR_INVALID,
               "nb_h2d: pageable host memory cannot be copied inside a CUDA-graph capture (an index vector or a NumPy "
               "operand created by the captured function): upload it before the capture or use pinned memory");
  }
  nb_note_write(ctx, dst);
  NB_CUDA_OK(ctx, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->stats.h2d_bytes += nbytes;
  return NB_OK;
}

int32_t nb_d2h_async(nb_ctx* ctx, void* dst, const void* src, uint64_t nbytes) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  if (!nbytes) return NB_OK;
  NB_CUDA_OK(ctx, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->stats.d2h_bytes += nbytes;
  return NB_OK;
}

int32_t nb_d2h(nb_ctx* ctx, void* dst, const void* src, uint64_t nbytes) {
  NB_REQUIRE(ctx, !ctx || !ctx->capturing, NB_ERR_INVALID,
             "nb_d2h: the context is capturing a CUDA graph — a device value cannot be read on the host there");
  int32_t rc = nb_d2h_async(ctx, dst, src, nbytes);
  if (rc != NB_OK) return rc;
  NB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return NB_OK;
}

int32_t nb_d2d(nb_ctx* ctx, void* dst, const void* src, uint64_t nbytes) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  if (!nbytes) return NB_OK;
  nb_note_write(ctx, dst);
  NB_CUDA_OK(ctx, cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return NB_OK;
}

int32_t nb_buffer_written(nb_ctx* ctx, const void* ptr, uint64_t nbytes) {
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  if (ptr) nb_note_write_range(ctx, ptr, (size_t)nbytes);
  return NB_OK;
}

int32_t nb_host_alloc(uint64_t nbytes, void** out) {
  if (!out) return nb_fail(nullptr, NB_ERR_INVALID, "out is NULL");
  cudaError_t e = cudaMallocHost(out, nbytes ? nbytes : 1);
  if (e != cudaSuccess) return nb_fail(nullptr, NB_ERR_CUDA, cudaGetErrorString(e));
  return NB_OK;
}

int32_t nb_host_free(void* p) {
  if (p) cudaFreeHost(p);
  return NB_OK;
}

// ---- graphs ---------------------------------------------------------------------------------
int32_t nb_graph_begin(nb_ctx* ctx) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx, NB_ERR_INVALID, "ctx is NULL");
  NB_REQUIRE(ctx, !ctx->capturing, NB_ERR_INVALID, "nb_graph_begin: already capturing");
  NB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  nb_split_cache_clear(ctx);  // a replayed graph rewrites buffers without passing through the host
  NB_CUDA_OK(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->capturing = true;
  ctx->cap_tag = ctx->next_tag++;
  ctx->parked[ctx->cap_tag];
  ctx->cap_launch_base = ctx->stats.kernel_launches;
  return NB_OK;
}

int32_t nb_graph_end(nb_ctx* ctx, nb_graph** out) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && out, NB_ERR_INVALID, "NULL argument");
  NB_REQUIRE(ctx, ctx->capturing, NB_ERR_INVALID, "nb_graph_end: not capturing");
  nb_split_cache_clear(ctx);
  uint32_t tag;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    tag = ctx->cap_tag;
    ctx->capturing = false;
    ctx->cap_tag = 0;
    auto& park = ctx->parked[tag];  // blocks recycled inside the capture stay with the graph
    for (auto& kv : ctx->cap_bins)
      for (void* p : kv.second) park.push_back({p, kv.first});
    ctx->cap_bins.clear();
  }
  cudaGraph_t g = nullptr;
  NB_CUDA_OK(ctx, cudaStreamEndCapture(ctx->stream, &g));
  nb_graph* ng = new nb_graph();
  ng->graph = g;
  ng->ctx = ctx;
  ng->tag = tag;
  ng->launches_per_replay = ctx->stats.kernel_launches - ctx->cap_launch_base;
  cudaError_t e = cudaGraphInstantiate(&ng->exec, g, 0);
  if (e != cudaSuccess) {
    nb_graph_destroy(ng);
    return nb_fail(ctx, NB_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
  }
  *out = ng;
  return NB_OK;
}

int32_t nb_graph_launch(nb_ctx* ctx, nb_graph* g) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && g && g->exec, NB_ERR_INVALID, "NULL argument");
  nb_split_cache_clear(ctx);
  NB_CUDA_OK(ctx, cudaGraphLaunch(g->exec, ctx->stream));
  ctx->stats.kernel_launches += g->launches_per_replay;
  return NB_OK;
}

int32_t nb_graph_destroy(nb_graph* g) {
  if (!g) return NB_OK;
  if (g->ctx) {  // hand the graph's blocks back to the general pool (after the stream drains)
    nb_ctx* ctx = g->ctx;
    cudaStreamSynchronize(ctx->stream);
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto it = ctx->parked.find(g->tag);
    if (it != ctx->parked.end()) {
      for (auto& pr : it->second) {
        ctx->free_bins[pr.second].push_back(pr.first);
        ctx->free_bytes += pr.second;
      }
      ctx->parked.erase(it);
    }
    for (auto& kv : ctx->live)
      if (kv.second.tag == g->tag) kv.second.tag = 0;  // still referenced: becomes a normal block
  }
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
  return NB_OK;
}

// ---- events ---------------------------------------------------------------------------------
int32_t nb_event_create(void** out) {
  if (!out) return nb_fail(nullptr, NB_ERR_INVALID, "out is NULL");
  cudaEvent_t ev;
  cudaError_t e = cudaEventCreate(&ev);
  if (e != cudaSuccess) return nb_fail(nullptr, NB_ERR_CUDA, cudaGetErrorString(e));
  *out = (void*)ev;
  return NB_OK;
}

int32_t nb_event_record(nb_ctx* ctx, void* ev) {
  nb_enter(ctx);
  NB_REQUIRE(ctx, ctx && ev, NB_ERR_INVALID, "NULL argument");
  NB_CUDA_OK(ctx, cudaEventRecord((cudaEvent_t)ev, ctx->stream));
  return NB_OK;
}

int32_t nb_event_elapsed_ms(void* a, void* b, float* ms) {
  if (!a || !b || !ms) return nb_fail(nullptr, NB_ERR_INVALID, "NULL argument");
  cudaError_t e = cudaEventSynchronize((cudaEvent_t)b);
  if (e == cudaSuccess) e = cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b);
  if (e != cudaSuccess) return nb_fail(nullptr, NB_ERR_CUDA, cudaGetErrorString(e));
  return NB_OK;
}

int32_t nb_event_destroy(void* ev) {
  if (ev) cudaEventDestroy((cudaEvent_t)ev);
  return NB_OK;
}

}  // extern "C"
