#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_model_parity.py -q -m gpu --timeout 300 > gpurun_out/test_model_parity.log 2>&1; echo "== model parity: $(tail -1 gpurun_out/test_model_parity.log)"
grep -E "^(FAILED|ERROR)|Error|max err" gpurun_out/test_model_parity.log | head
CMD="python tools/kernel_bench.py --dtype bf16 --fused-only --only enc0 --reps 3"
$CMD > gpurun_out/plain_kb.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ss2d_ -s 6 -c 2 -o gpurun_out/prof_fused -f $CMD > gpurun_out/ncu_fused.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_fused.log
