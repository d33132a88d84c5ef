#!/bin/bash
# N-GPU pass (gpurun --gpus N): the NCCL sharded == single-GPU test (2 ranks) and the bench line at N ranks.
# usage: gpu_r2_n.sh N
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N="${1:-2}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$N.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "nccl or peer_out" > gpurun_out/pytest_nccl.log 2>&1; echo "pytest nccl exit $?"; tail -3 gpurun_out/pytest_nccl.log
timeout ${T:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
for g in ${GATHERS:-peer nccl}; do
timeout ${T:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 10 --warmup 3 --gather $g > gpurun_out/bench_n${N}_$g.json 2> gpurun_out/bench_n${N}_$g.err; echo "bench n$N $g exit $?"
done
python - <<PY
import json
import os
for f in ("bench_n$N.json", "bench_n${N}_peer.json", "bench_n${N}_nccl.json"):
    if not os.path.exists("gpurun_out/" + f): continue
    d = json.loads([l for l in open("gpurun_out/" + f) if l.startswith("{")][-1])
    print(f, "value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], d["detail"]["collective"][:60]); print(d.get("multi_gpu"))
PY
tail -5 gpurun_out/bench_n$N.err
