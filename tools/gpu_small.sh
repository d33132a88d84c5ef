#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python tools/bench_small.py > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "exit $?"; cat gpurun_out/bench_small.json; tail -3 gpurun_out/bench_small.err
