#!/bin/bash
# full gpu test-suite + training-step benchmark + forward bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/umma_layer_check.py > gpurun_out/layer_check.log 2>&1; echo "layer_check exit $?"; grep -E "final|FAILED" gpurun_out/layer_check.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_train.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench_train exit $?"; cat gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_bf16x3.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'clocks',d['clocks'])
print(d['kernel_ms_per_step'])
PY
