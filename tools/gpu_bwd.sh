#!/bin/bash
# backward validation + training-step benchmark
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "backward or training or gradients" 2>&1 | tail -3
timeout 600 python tools/bench_train.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench_train exit $?"; cat gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
