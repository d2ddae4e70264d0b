#!/bin/bash
# A/B of GEMM launch knobs on the forward shape: steady-state time and DRAM traffic per configuration.
mkdir -p gpurun_out
: > gpurun_out/knob_time.log
: > gpurun_out/knob_traffic.log
run() {
  name=$1; shift
  echo "== $name $*" | tee -a gpurun_out/knob_time.log
  env "$@" SHAPES=fwd REPS=80 TAIL=40 python tools/bench_gemm.py 32768 bf16x3 >> gpurun_out/knob_time.log 2>&1
  echo "== $name $*" >> gpurun_out/knob_traffic.log
  env "$@" SHAPES=fwd REPS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
     --clock-control none -k regex:k_gemm_tc -s 2 -c 1 python tools/bench_gemm.py 32768 bf16x3 2>&1 \
     | grep -E "dram__|duration|lts__|tensor" >> gpurun_out/knob_traffic.log
}
for cfg in "$@"; do run $cfg; done
cat gpurun_out/knob_time.log | grep -E "==|avg"; cat gpurun_out/knob_traffic.log
