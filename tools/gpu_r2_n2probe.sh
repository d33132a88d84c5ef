#!/bin/bash
# 2-GPU pass for the output exchange: the NCCL / peer-memory / fused tests, the cost probe, the bench line (fused and peer)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "nccl or peer_out" > gpurun_out/pytest_nccl.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_nccl.log
for g in fused; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --gather $g > gpurun_out/bench_n2_$g.json 2> gpurun_out/bench_n2_$g.err; echo "bench n2 $g exit $?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_n2_$g.json") if l.startswith("{")][-1])
print("$g value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"]); print(d.get("multi_gpu")); print(d["kernel_ms_per_step"])
PY
tail -3 gpurun_out/bench_n2_$g.err
done
