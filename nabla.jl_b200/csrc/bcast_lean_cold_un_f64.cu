// see bcast_lean_cold.inc
#define NB_COLD_T double
#define NB_COLD_UNARY 1
#include "bcast_lean_cold.inc"

bool nb_lean_cold_un_f64(int kind, int op, int mode, const nbb::MRP& P, const nbb::LeanLaunch& a, cudaStream_t st) {
  return cold_dispatch(kind, op, mode, P, a, st);
}
