#!/bin/bash
# N-GPU training-step bench (one process per GPU, NCCL): ZeRO-1 step (default) vs replicated step with a bf16 / fp32
# gradient all-reduce; the 1-GPU line of the same box first.
#   gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
run() {  # tag, extra args
  local tag=$1; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_${N}gpu_$tag.json 2> gpurun_out/bench_${N}gpu_$tag.err
  echo "bench N=$N $tag rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${N}gpu_$tag.json').read().strip().splitlines()[-1])
    print('$tag', d['n_gpus'], round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms', d.get('replica_check'), d['config'].get('collective'))
except Exception as e:
    print('unreadable', e); print(open('gpurun_out/bench_${N}gpu_$tag.err').read()[-1500:])
PY
}
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_1gpu_ref.json 2> gpurun_out/bench_1gpu_ref.err
python -c "import json; d=json.loads(open('gpurun_out/bench_1gpu_ref.json').read().strip().splitlines()[-1]); print('1 gpu', round(d['value'],1), round(d['ms_per_step'],3))"
run zero1
[ "${2:-}" = zero ] && exit 0
run bf16wire --no-zero
[ "${2:-}" = all ] && run f32wire --no-zero --allreduce-dtype f32
exit 0
