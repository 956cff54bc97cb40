#!/bin/bash
# Developer GPU pass: every GPU test file + smoke, the per-site fused microbench, the training-step bench with its
# kernel table, configs 2 / 4 / 5.   gpurun --timeout 1500 -- 'bash tools/gpu_suite.sh'
mkdir -p gpurun_out
bash run_gpu_tests.sh > gpurun_out/run_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; grep -E "^== |FAILED|^E  " gpurun_out/run_gpu_tests.log | head -60
timeout 300 python tools/kernel_bench.py --dtype bf16 --fused-only > gpurun_out/kernel_bench_fused_bf16.jsonl 2>&1; cat gpurun_out/kernel_bench_fused_bf16.jsonl | cut -c1-260
for dt in bf16 f32; do timeout 300 python tools/kernel_bench.py --dtype $dt > gpurun_out/kernel_bench_$dt.jsonl 2>&1; done; echo "kernel_bench all rc=$?"
timeout 200 python tools/side_bench.py > gpurun_out/side_bench.jsonl 2>&1; echo "side_bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget-s 60 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
timeout 600 python bench.py --steps 20 --warmup 5 --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:d['roofline'][k] for k in ('kernel','frac','us_per_launch','share_of_step','timed_in')}, d.get('cpu_baseline'))
print({k:(v['us_per_launch'],v['frac'],v['share_of_step']) for k,v in d['kernels'].items()})
PY
for c in 2 4 5; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_config$c.json 2> gpurun_out/bench_config$c.err; echo "config $c rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_config$c.json').read().strip().splitlines()[-1])
print(d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], {k:d['roofline'][k] for k in ('kernel','frac','us_per_launch','share_of_step')})
PY
done
