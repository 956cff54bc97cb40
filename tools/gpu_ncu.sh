#!/bin/bash
# ncu --set full captures of the kernels the roofline discussion is about (one launch each, after warm-up launches).
#   gpurun --timeout 1200 -- 'bash tools/gpu_ncu.sh'
mkdir -p gpurun_out
cap() {  # name, kernel regex, kernel_bench args...
  local name=$1 rx=$2; shift 2
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o gpurun_out/$name \
    python tools/kernel_bench.py "$@" --reps 2 > gpurun_out/$name.log 2>&1; echo "ncu $name rc=$?"
}
cap r02_ncu_fbwd_enc0 ss2d_bwd_kernel --dtype bf16 --fused-only --only enc0
cap r02_ncu_fbwd_dec96 ss2d_bwd_kernel --dtype bf16 --fused-only --only dec96
cap r02_ncu_fbwd_enc2 ss2d_bwd_kernel --dtype bf16 --fused-only --only enc2
cap r02_ncu_ffwd_enc0 ss2d_fwd_kernel --dtype bf16 --fused-only --only enc0
cap r02_ncu_sscan_bwd_bf16_enc0 sscan_bwd_regular --dtype bf16 --only enc0
cap r02_ncu_sscan_bwd_f32_enc0 sscan_bwd_regular --dtype f32 --only enc0
cap r02_ncu_sscan_fwd_bf16_enc0 sscan_fwd_regular --dtype bf16 --only enc0
ls -la gpurun_out
