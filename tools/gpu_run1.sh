#!/bin/bash
# GPU run 1: tcgen05/TMA probes, parity tests of the first CUDA path, first bench line, launch list.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt
: > gpurun_out/probe.log
for t in basic swapped conv conv_swapped n224 n64 n16 n256 tma; do
  timeout 60 tools/umma_probe $t >> gpurun_out/probe.log 2>&1; echo "[$t exit $?]" >> gpurun_out/probe.log
done
for n in 16 64 128 224 256; do
  timeout 60 tools/umma_probe rate $n 200 >> gpurun_out/probe.log 2>&1; echo "[rate $n exit $?]" >> gpurun_out/probe.log
  timeout 60 tools/umma_probe rate $n 200 1 >> gpurun_out/probe.log 2>&1; echo "[rate_conv $n exit $?]" >> gpurun_out/probe.log
done
cat gpurun_out/probe.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --mode fp32 --steps 2 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench exit $?"
cat gpurun_out/bench_fp32.json; tail -5 gpurun_out/bench_fp32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1_fp32.csv \
  python bench.py --mode fp32 --steps 1 --warmup 1 --batch 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?"
