#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/umma_layer_check.py > gpurun_out/layer_check.log 2>&1; echo "layer_check exit $?"; grep -E "final|FAILED" gpurun_out/layer_check.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_bf16x3.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'clocks',d['clocks'])
print(d['kernel_ms_per_step']); print(d['kernel_tflops'])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 10 -c 10 -o gpurun_out/prof_r1_umma_v4 \
  python tools/profile_forward.py 1 1080 1920 bf16x3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_bf16x3_v4.csv \
  python bench.py --steps 1 --warmup 1 --batch 2 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1; echo "ncu list exit $?"
