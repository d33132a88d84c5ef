#!/bin/bash
# last check of the shipped library: per-layer check in both tensor modes, the gpu tests, the training-step bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for m in bf16x3 bf16_fp8; do
  WN_CHECK_MODE=$m timeout 200 python tools/umma_layer_check.py > gpurun_out/layer_check_$m.log 2>&1; echo "layer_check $m exit $?"
  grep -E "final|FAILED" gpurun_out/layer_check_$m.log | tr '\n' ' '; echo
done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_train.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench_train exit $?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_train.json') if l.startswith('{')][-1])
print([(r['batch'], r['height'], round(r['native_ms'],3)) for r in d['results']])"
