// placeholder until the tcgen05 kernel lands (replaced below in this round)
#include "common.cuh"
int32_t nb_gemm_tcgen05(nb_ctx*, bool, bool, int64_t, int64_t, int64_t, float, const float*, int64_t,
                        const float*, int64_t, float, float*, int64_t, int32_t) {
  return NB_ERR_UNSUPPORTED;
}
void nb_gemm_state_free(nb_ctx*) {}
