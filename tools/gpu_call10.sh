#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_fused test_gpu_fused_sites test_gpu_ss2d test_model_parity; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu --timeout 300 > gpurun_out/$f.log 2>&1
  echo "== $f: $(tail -1 gpurun_out/$f.log)"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/$f.log | head -12
done
timeout 300 python tools/fused_ab.py --dtype bf16 > gpurun_out/fused_ab_bf16.jsonl 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/fused_ab_bf16.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print({k:d[k] for k in d if k in ('site','v1_us','dx_us','fwd_v1_us','fwd_pair_us','fwd_pair_gbs_fused','y_rel_diff','us_per_image_step_weighted_by_calls')})
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel-table gpurun_out/kernel_table.txt > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:d['roofline'][k] for k in ('kernel','frac','us_per_launch','share_of_step','timed_in')})
print({k:(v['us_per_launch'],v['share_of_step']) for k,v in d['kernels'].items()})
PY
timeout 600 python bench.py --config 4 --steps 10 --warmup 3 > gpurun_out/bench_config4.json 2> gpurun_out/bench_config4.err; echo "config 4 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_config4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:d['roofline'][k] for k in ('kernel','frac','us_per_launch','share_of_step')})
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ss2d_fwd_pair -s 3 -c 1 -o gpurun_out/prof_fwd_pair -f python tools/fused_ab.py --dtype bf16 --only enc0 --reps 2 > gpurun_out/ncu_fwd.log 2>&1; echo "ncu rc=$?"
