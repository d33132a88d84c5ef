#!/bin/bash
# GPU run 2: tcgen05 conv path bring-up: rate probe, per-layer check, parity tests, bench, ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/probe2.log
for args in "128 200 1" "128 200 2" "128 200 4" "64 200 1" "64 200 2" "64 200 4" "64 200 8" "16 200 4" "16 200 8" "224 200 2" "96 200 4"; do
  timeout 60 tools/umma_probe rate2 $args >> gpurun_out/probe2.log 2>&1; echo "[rate2 $args exit $?]" >> gpurun_out/probe2.log
done
cat gpurun_out/probe2.log
timeout 300 python tools/umma_layer_check.py > gpurun_out/layer_check.log 2>&1; echo "layer_check exit $?" | tee -a gpurun_out/layer_check.log
cat gpurun_out/layer_check.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; echo "bench exit $?"
cat gpurun_out/bench_bf16x3.json; tail -5 gpurun_out/bench_bf16x3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_bf16x3.csv \
  python bench.py --steps 1 --warmup 1 --batch 2 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 10 -c 10 -o gpurun_out/prof_r1_umma \
  python tools/profile_forward.py 1 1080 1920 bf16x3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
tail -3 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"apply_kernel|stats_kernel" -s 2 -c 2 -o gpurun_out/prof_r1_pre \
  python tools/profile_forward.py 4 1080 1920 bf16x3 > gpurun_out/ncu_pre.log 2>&1; echo "ncu pre exit $?"
ls -la gpurun_out/
