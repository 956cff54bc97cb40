#!/bin/bash
# ncu --set full captures of the kernels the roofline discussion is about (one launch each, after warm-up launches);
# the reports are summarised ON THE BOX (raw counters, per-CUDA-line instruction / stall shares) and only the text
# comes back (gpurun_out is capped at 64 MiB).   gpurun --timeout 1500 -- 'bash tools/gpu_ncu.sh'
mkdir -p gpurun_out
cap() {  # name, kernel regex, kernel_bench args...
  local name=$1 rx=$2; shift 2
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -f -o /tmp/$name \
    python tools/kernel_bench.py "$@" --reps 2 > gpurun_out/$name.log 2>&1; echo "ncu $name rc=$?"
  python tools/ncu_raw_summary.py /tmp/$name.ncu-rep > gpurun_out/${name}_summary.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page source --print-source cuda,sass --csv > /tmp/$name.csv 2>/dev/null
  python tools/ncu_cuda_lines.py /tmp/$name.csv 70 > gpurun_out/${name}_cuda_lines.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page source --print-source sass --csv > /tmp/${name}_sass.csv 2>/dev/null
  python tools/ncu_src_summary.py /tmp/${name}_sass.csv 40 > gpurun_out/${name}_sass_mix.txt 2>&1
  python tools/ncu_stall_reasons.py 12 < /tmp/${name}_sass.csv > gpurun_out/${name}_stall_reasons.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum > gpurun_out/${name}_pipes.csv 2>/dev/null
  gzip -c /tmp/${name}_sass.csv > gpurun_out/${name}_sass.csv.gz   # per-instruction table (executed, stall columns) for offline reading
  rm -f /tmp/$name.ncu-rep /tmp/$name.csv /tmp/${name}_sass.csv
}
WHAT=${1:-all}
if [ "$WHAT" = all ] || [ "$WHAT" = fused ]; then
cap r02_ncu_fbwd_enc0 ss2d_bwd_kernel --dtype bf16 --fused-only --only enc0
cap r02_ncu_ffwd_enc0 ss2d_fwd_kernel --dtype bf16 --fused-only --only enc0
fi
if [ "$WHAT" = all ]; then
cap r02_ncu_fbwd_dec96 ss2d_bwd_kernel --dtype bf16 --fused-only --only dec96
cap r02_ncu_fbwd_enc2 ss2d_bwd_kernel --dtype bf16 --fused-only --only enc2
fi
if [ "$WHAT" = all ] || [ "$WHAT" = op ]; then
cap r02_ncu_sscan_bwd_bf16_enc0 sscan_bwd_regular --dtype bf16 --only enc0
cap r02_ncu_sscan_bwd_f32_enc0 sscan_bwd_regular --dtype f32 --only enc0
cap r02_ncu_sscan_fwd_bf16_enc0 sscan_fwd_regular --dtype bf16 --only enc0
fi
ls -la gpurun_out
