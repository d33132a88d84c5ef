#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python tools/bench_train.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench_train exit $?"; cat gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wgrad_umma|gate_bwd|bias_grad" -s 20 -c 14 -o gpurun_out/prof_r1_bwd \
  python tools/bench_train.py > gpurun_out/ncu_bwd.log 2>&1; echo "ncu bwd exit $?"
timeout 1800 bash tools/cli_smoke.sh > gpurun_out/cli_smoke.log 2>&1; echo "cli exit $?"; tail -12 gpurun_out/cli_smoke.log
