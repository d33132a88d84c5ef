#!/bin/bash
# quick validation run: per-layer check, gpu tests, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/umma_layer_check.py > gpurun_out/layer_check.log 2>&1; echo "layer_check exit $?" | tee -a gpurun_out/layer_check.log
grep -E "final|FAILED|exit" gpurun_out/layer_check.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_bf16x3.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'clocks',d['clocks'])
print(d['kernel_ms_per_step']); print(d['kernel_tflops']); print(d['roofline'])
PY
tail -3 gpurun_out/bench_bf16x3.err
