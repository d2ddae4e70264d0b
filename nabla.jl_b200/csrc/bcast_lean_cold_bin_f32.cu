// see bcast_lean_cold.inc
#define NB_COLD_T float
#define NB_COLD_UNARY 0
#include "bcast_lean_cold.inc"

bool nb_lean_cold_bin_f32(int kind, int op, int mode, const nbb::MRP& P, const nbb::LeanLaunch& a, cudaStream_t st) {
  return cold_dispatch(kind, op, mode, P, a, st);
}
