#!/bin/bash
# Runs the `-m gpu` test files WITHOUT a GPU, against the product's own C ABI and kernel sources executed on the CPU
# simulator (tools/cpu_sim; tests/conftest.py: TRAMBA_GPU_TESTS_ON_SIM).  A few minutes on 8 cores.
#   bash tools/run_gpu_tests_on_sim.sh
# Extra checks: TRAMBA_SIM_LAZY=1 (asynchronous copies land at their wait), TRAMBA_SIM_SCHEDULE=n (thread order).
mode=1
rc=0
for f in test_gpu_scan_merge test_gpu_proj_gather test_gpu_ss2d test_gpu_fused test_gpu_fused_sites test_dwconv test_layernorm2d \
         test_gpu_dct test_z_gpu_ablation; do
  out=$(TRAMBA_GPU_TESTS_ON_SIM=$mode timeout 1800 python -m pytest tests/$f.py -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -1)
  echo "== $f: $out"
  case "$out" in *failed*|*error*) rc=1;; esac
done
# the operator selective scan: everything but the two BASELINE-size cases (3.5 minutes on the simulator; add them with
# `TRAMBA_GPU_TESTS_ON_SIM=1 python -m pytest tests/test_gpu_sscan.py -m gpu -k "config2 or full"`)
out=$(TRAMBA_GPU_TESTS_ON_SIM=$mode timeout 1800 python -m pytest tests/test_gpu_sscan.py -m gpu -q --timeout 600 \
      -p no:cacheprovider -k "not config2 and not full and not 9216 and not 192" 2>&1 | tail -1)
echo "== test_gpu_sscan (small and medium cases): $out"
case "$out" in *failed*|*error*) rc=1;; esac
# whole-model parity on the simulated library
out=$(timeout 1800 python -m pytest tests/test_model_parity.py -q -k simulated -p no:cacheprovider 2>&1 | tail -1)
echo "== test_model_parity (simulated library): $out"
case "$out" in *failed*|*error*) rc=1;; esac
exit $rc
