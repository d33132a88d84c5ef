#!/bin/bash
# usage: gpu_bench_n.sh N   (run under gpurun --gpus N)
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
grep '^{' gpurun_out/bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value',d['value'],'n',d['n_gpus'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['clocks'])"
tail -3 gpurun_out/bench_n$N.err
