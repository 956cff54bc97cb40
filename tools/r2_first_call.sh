#!/bin/bash
# First GPU call of a round: everything that was prepared without a GPU gets its verdict in one box lease.
#   gpurun --timeout 1500 -- 'bash tools/r2_first_call.sh'
# 1. the product's GPU test suite (file by file) + smoke         -> gpurun_out/test_*.log, smoke.log
# 2. candidate kernels vs product kernels, parity + time         -> gpurun_out/variant_bf16.jsonl, variant_f32.jsonl
# 3. the fused / model parity tests against the candidate build  -> gpurun_out/next_*.log
# 4. bench.py with the product and with the candidate library    -> gpurun_out/bench.json, bench_next.json
# 5. the per-site microbenches (config 2 block, side ops, configs 4/5)
mkdir -p gpurun_out
bash run_gpu_tests.sh > gpurun_out/run_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -25 gpurun_out/run_gpu_tests.log
for dt in bf16 f32; do
  timeout 300 python tools/variant_bench.py --dtype $dt --reps 20 > gpurun_out/variant_$dt.jsonl 2> gpurun_out/variant_$dt.err
  echo "variant_bench $dt rc=$?"; tail -2 gpurun_out/variant_$dt.jsonl
done
for f in test_gpu_fused test_gpu_ss2d test_model_parity; do
  TRAMBA_B200_LIB=next timeout 600 python -m pytest tests/$f.py -q -m gpu -x --timeout 300 > gpurun_out/next_$f.log 2>&1
  echo "== next/$f: $(tail -1 gpurun_out/next_$f.log)"
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
TRAMBA_B200_LIB=next timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_next.json 2> gpurun_out/bench_next.err
echo "bench(next) rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench.json", "gpurun_out/bench_next.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], "img/s", d["ms_per_step"], "ms;", {k: v["us_per_launch"] for k, v in d["kernels"].items() if k.startswith("ss2d")})
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout 300 python tools/ss2d_bench.py > gpurun_out/ss2d_bench.jsonl 2>&1; echo "ss2d_bench rc=$?"
timeout 300 python tools/kernel_bench.py --dtype bf16 --fused-only > gpurun_out/kernel_bench_fused_bf16.jsonl 2>&1
timeout 600 python tools/op_profile.py > gpurun_out/op_profile.txt 2> gpurun_out/op_profile.err; echo "op_profile rc=$?"; head -30 gpurun_out/op_profile.txt
# ncu of the candidate kernels on the enc0 site (only after the plain run above exited 0): same capture as the
# product's r01_ncu_f{bwd,fwd}_enc0_* so that instruction counts / stall reasons can be compared line by line
if grep -q '"all_parity_ok": true' gpurun_out/variant_bf16.jsonl; then
  for k in bwd fwd; do
    timeout 600 ncu --set full --import-source on --clock-control none -k regex:ss2d_${k}_cg_kernel -c 1 -f \
      -o gpurun_out/prof_f${k}_cg python tools/variant_bench.py --only enc0 --reps 1 > gpurun_out/ncu_f${k}_cg.log 2>&1
    echo "ncu ${k}_cg rc=$?"
  done
fi
# compute-sanitizer (race detection) on the small fused / scan / LayerNorm / dwconv cases: memcheck for
# out-of-bounds, racecheck for shared-memory hazards between the pipeline stages of the fused kernels
for tool in memcheck racecheck; do
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_fused.py -q -m gpu -x \
    -k "12 or 24" > gpurun_out/sanitizer_${tool}_fused.log 2>&1
  echo "compute-sanitizer $tool fused rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_fused.log | tail -1)"
done
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke \
  > gpurun_out/sanitizer_memcheck_smoke.log 2>&1
echo "compute-sanitizer memcheck smoke rc=$? : $(grep 'ERROR SUMMARY' gpurun_out/sanitizer_memcheck_smoke.log | tail -1)"
