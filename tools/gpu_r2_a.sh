#!/bin/bash
# round 2, first GPU pass: the whole gpu test-suite, smoke(), both bench arms, then the 400-epoch training run
# whose checkpoint becomes the third parity weight set (tests/golden/trained_synthetic_400ep.npz)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log; grep -h "max rel err" gpurun_out/pytest_gpu.log | tail -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref exit $?"
python - <<'PY'
import json
for f in ("default", "reference"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_{f}.json") if l.startswith("{")][-1])
    except Exception as e:
        print(f, "no line", e); continue
    print(f, "value", d.get("value"), "e2e", d.get("e2e", {}).get("value"), "ms", d.get("ms_per_step"), d.get("clocks"))
    if f == "default": print(d.get("kernel_ms_per_step")); print(d.get("parity")); print(d["roofline"]["frac"], d["roofline"]["achieved"], d["gpu_launches"])
PY
rm -rf training
timeout 1500 python train.py --synthetic --epochs 400 --seed 0 > gpurun_out/train400.log 2>&1; echo "train exit $?"
tail -2 gpurun_out/train400.log
python - <<'PY'
import numpy as np, torch
sd = torch.load("training/0/last.pt", map_location="cpu")
np.savez_compressed("gpurun_out/trained_synthetic_400ep.npz", **{k: v.numpy() for k, v in sd.items()})
print("saved", len(sd), "tensors")
PY
cp training/0/metrics-val.csv gpurun_out/train400_metrics-val.csv; rm -rf training
