#!/bin/bash
# BASELINE configs[4]: train.py, 400 epochs, batch 16, 112x112, UIEB-shaped synthetic data (890 items).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -rf training
timeout 2400 python train.py --synthetic --epochs 400 --seed 0 > gpurun_out/train400.log 2>&1; echo "train exit $?"
grep -E "Train|Val" gpurun_out/train400.log | awk 'NR<=4 || NR%100<2' | head -40
tail -3 gpurun_out/train400.log
cp training/0/metrics-train.csv gpurun_out/train400_metrics-train.csv
cp training/0/metrics-val.csv gpurun_out/train400_metrics-val.csv
rm -rf training
