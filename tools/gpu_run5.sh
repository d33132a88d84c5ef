#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 10 -c 10 -o gpurun_out/prof_r1_umma_v3 \
  python tools/profile_forward.py 1 1080 1920 bf16x3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_bf16x3.csv \
  python bench.py --steps 1 --warmup 1 --batch 2 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1; echo "ncu list exit $?"
timeout 600 python -m pytest tests -m gpu -q -k "gradients" 2>&1 | tail -3
