#!/bin/bash
# Runs every GPU test file in its own process (a CUDA fault in one cannot poison the rest), then the smoke test.
# The driver's own command is `python -m pytest tests -x -q -m gpu`; this script is the developer variant.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/gpu.txt
rc=0
for f in tests/test_*.py; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu -x --timeout 300 > gpurun_out/$n.log 2>&1
  code=$?
  [ $code -ne 0 ] && [ $code -ne 5 ] && rc=1   # 5 = no GPU tests in this file
  echo "== $n: $(tail -1 gpurun_out/$n.log)"
  grep -E "^(FAILED|ERROR)|Error|error:" gpurun_out/$n.log | head -8
done
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 || rc=1
echo "== smoke: $(tail -1 gpurun_out/smoke.log)"
exit $rc
