#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 11 -c 2 -o gpurun_out/prof_l1split \
  python tools/profile_forward.py 1 1080 1920 bf16x3 > gpurun_out/ncu_l1.log 2>&1; echo "ncu exit $?"
