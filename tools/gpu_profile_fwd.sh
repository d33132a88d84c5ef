#!/bin/bash
# ncu evidence of the default forward mode only: full capture of the ten conv launches (one 1080p image) and the
# launch list of a short bench.  Condense with tools/summarize_ncu.py.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 10 -c 10 -o gpurun_out/prof_umma_default \
  python tools/profile_forward.py 1 1080 1920 default > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv \
  python bench.py --steps 1 --warmup 1 --batch 2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
