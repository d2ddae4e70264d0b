// fused un(bin(a, b)) lean kernels, float (see bcast_lean_fused.inc)
#define NB_FUSED_T float
#include "bcast_lean_fused.inc"

bool nb_lean_fused_f32(int kind, int op, int mode, const nbb::MRP& P, const nbb::LeanLaunch& a, cudaStream_t st) {
  return fused_dispatch(kind, op, mode, P, a, st);
}
