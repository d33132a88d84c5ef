#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "graph or enhance or hub" 2>&1 | tail -4
timeout 600 python tools/bench_small.py > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "exit $?"; python -c "
import json
for r in json.load(open('gpurun_out/bench_small.json')): print(r)
"; tail -3 gpurun_out/bench_small.err
