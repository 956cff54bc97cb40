#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_fused test_gpu_fused_sites; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu --timeout 300 > gpurun_out/$f.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/$f.log)"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/$f.log | head -12
done
timeout 300 python tools/fused_ab.py --dtype bf16 > gpurun_out/fused_ab_bf16.jsonl 2>&1; cat gpurun_out/fused_ab_bf16.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'])
print({k:(v['us_per_launch'],v['share_of_step']) for k,v in d['kernels'].items()})
PY
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-shadow > gpurun_out/bench_noshadow.json 2> gpurun_out/bench_noshadow.err; echo "bench noshadow rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_noshadow.json').read().strip().splitlines()[-1])
print("no-shadow:", d['value'], d['ms_per_step'])
PY
for c in 2 4 5; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_config$c.json 2> gpurun_out/bench_config$c.err; echo "config $c rc=$?"; head -c 700 gpurun_out/bench_config$c.json; echo; tail -2 gpurun_out/bench_config$c.err; done
