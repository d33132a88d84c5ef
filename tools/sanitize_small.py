"""Tiny end-to-end pass (preprocess, all forward modes incl. the fused-tail default mode on a multi-tile-per-CTA
shape, sub-modules, resize, training forward + backward) for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from waternet_b200 import _lib
from waternet_b200.engine import get_engine
from waternet_b200.net import WaterNet

torch.manual_seed(0)
eng = get_engine("cuda:0")
m = WaterNet().cuda()
eng.pack_weights(m._ordered_params())
rgb = torch.randint(0, 256, (2, 37, 53, 3), dtype=torch.uint8, device="cuda")
for mode in (_lib.MODE_DEFAULT, _lib.MODE_BF16X3, _lib.MODE_FP32_SIMT):
    out = eng.enhance(rgb, mode=mode)
big = torch.randint(0, 256, (1, 130, 200, 3), dtype=torch.uint8, device="cuda")   # several tiles per CTA pair? (13 x 9 tiles)
eng.set_chunk_pixels(130 * 200)
out = eng.enhance(torch.cat([big, big]), mode=_lib.MODE_DEFAULT)                    # two passes, fused tails
eng.set_chunk_pixels(0)
eng.set_debug_flags(1792 + 2048)
out = eng.enhance(big, mode=_lib.MODE_DEFAULT)                                      # the unfused / plain-first-layer forms
eng.set_debug_flags(0)
res = eng.resize_batch([big[0], rgb[0]], 48, 64, swap_rb=True)
gray = eng.white_balance_gray(rgb[..., 0].contiguous())
pre = eng.preprocess(rgb)
ins = [pre[k] for k in ("x", "wb", "he", "gc")]
m.train()
loss = m(*ins).square().mean()
loss.backward()
torch.cuda.synchronize()
print("ok", float(loss), float(m.cmg.conv2.weight.grad.abs().sum()))
