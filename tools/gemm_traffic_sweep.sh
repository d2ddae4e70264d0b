#!/bin/bash
# Sweep the GEMM L2 knobs: steady-state time (no profiler) and DRAM traffic (ncu) per configuration.
# usage: tools/gemm_traffic_sweep.sh   (writes gpurun_out/sweep_time.log, gpurun_out/sweep_traffic.csv)
mkdir -p gpurun_out
: > gpurun_out/sweep_time.log
: > gpurun_out/sweep_traffic.log
run() {  # name shapes env...
  name=$1; shapes=$2; shift 2
  echo "== $name [$shapes] $*" | tee -a gpurun_out/sweep_time.log
  env "$@" SHAPES=$shapes REPS=80 TAIL=40 python tools/bench_gemm.py 32768 bf16x3 >> gpurun_out/sweep_time.log 2>&1
  echo "== $name [$shapes] $*" >> gpurun_out/sweep_traffic.log
  env "$@" SHAPES=$shapes REPS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
     --clock-control none -k regex:k_gemm_tc -s 2 -c 8 --csv python tools/bench_gemm.py 32768 bf16x3 2>&1 \
     | grep -E "k_gemm" | awk -F'","' '{print $(NF-2)","$(NF)}' | tr -d '"' >> gpurun_out/sweep_traffic.log
}
run base        fwd  X=1
run kchunk8192  fwd  NB_GEMM_KCHUNK=8192
run hintA_gm7   fwd  NB_GEMM_HINT=1 NB_GEMM_PANEL_MB=64
run hintAf_gm7  fwd  NB_GEMM_HINT=5 NB_GEMM_PANEL_MB=64
run hintA_gm10  fwd  NB_GEMM_HINT=1 NB_GEMM_GROUP_M=10
run kr4096_gm8  fwd  NB_GEMM_KRANGE=4096 NB_GEMM_PANEL_MB=36
run kr4096_gm8h fwd  NB_GEMM_KRANGE=4096 NB_GEMM_PANEL_MB=36 NB_GEMM_HINT=1
run kr2048_gm32h fwd NB_GEMM_KRANGE=2048 NB_GEMM_PANEL_MB=70 NB_GEMM_HINT=1
run base        abar X=1
run hintA_gm7   abar NB_GEMM_HINT=1 NB_GEMM_PANEL_MB=64
run kr4096_gm8h abar NB_GEMM_KRANGE=4096 NB_GEMM_PANEL_MB=36 NB_GEMM_HINT=1
run base        bbar X=1
run hintA_gm7   bbar NB_GEMM_HINT=1 NB_GEMM_PANEL_MB=64
