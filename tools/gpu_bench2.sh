#!/bin/bash
# 2-GPU call: DDP bench (ours) at N=2, reference arm under torchrun, then N=1 bench + short ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"; tail -1 gpurun_out/bench_n2.json | cut -c1-900; tail -3 gpurun_out/bench_n2.err
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "n1 rc=$?"; tail -1 gpurun_out/bench.json | cut -c1-2500; tail -2 gpurun_out/bench.err
