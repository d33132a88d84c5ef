#!/bin/bash
# End-to-end smoke of the reference-compatible CLIs on the GPU box (tiny synthetic data).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
set -e
python - <<'PY'
import numpy as np, cv2, os, torch, sys
sys.path.insert(0, '.')
from oracle.forward import synthetic_image, synthetic_state_dict
os.makedirs('/tmp/wn_src', exist_ok=True)
for i in range(3):
    cv2.imwrite(f'/tmp/wn_src/img{i}.png', synthetic_image(i, 120, 160, 'smooth')[..., ::-1])
torch.save(synthetic_state_dict(0, 3.0), '/tmp/wn_weights.pt')
PY
python inference.py --source /tmp/wn_src --weights /tmp/wn_weights.pt --name smoke
python inference.py --source /tmp/wn_src/img0.png --weights /tmp/wn_weights.pt --name smoke_split --show-split
ls -la output/smoke output/smoke_split
python - <<'PY'
import cv2, numpy as np, sys, torch
sys.path.insert(0, '.')
from oracle import forward as ofw, preprocess as opre
sd = torch.load('/tmp/wn_weights.pt')
rgb = cv2.imread('/tmp/wn_src/img1.png')[..., ::-1]
wb, gc, he = opre.transform(np.ascontiguousarray(rgb))
ins = [torch.from_numpy(opre.arr2ten(a).copy()) for a in (rgb, wb, he, gc)]
ref = opre.ten2arr(ofw.waternet_forward(sd, *ins).numpy())[0]
got = cv2.imread('output/smoke/img1.png')[..., ::-1]
d = np.abs(got.astype(int) - ref.astype(int))
print('inference.py vs oracle: max diff', d.max(), 'frac differing', (d != 0).mean())
assert d.max() <= 1
PY
timeout 900 python train.py --synthetic --epochs 1 --batch-size 16 --height 64 --width 64 --seed 0 2>&1 | tail -8
ls training/*/
timeout 600 python score.py --synthetic --weights training/0/last.pt --height 64 --width 64 2>&1 | tail -3
rm -rf output training
