#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/op_profile.py > gpurun_out/op_profile.txt 2> gpurun_out/op_profile.err; echo "op_profile rc=$?"; sed -n 1,140p gpurun_out/op_profile.txt
