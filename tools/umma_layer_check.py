"""Per-layer comparison of the tensor-core path against the fp32 CUDA-core path (bring-up aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from waternet_b200 import _lib
from waternet_b200.engine import get_engine
from waternet_b200.net import WaterNet

# WN_CHECK_MODE=bf16_fp8 checks the fp8-correction scheme instead of plain bf16x3
TENSOR_MODE = {"bf16x3": _lib.MODE_BF16X3, "bf16_fp8": _lib.MODE_BF16_FP8}[os.environ.get("WN_CHECK_MODE", "bf16x3")]
torch.manual_seed(0)
eng = get_engine("cuda:0")
m = WaterNet().cuda().eval()
with torch.no_grad():
    for p in m.parameters():
        p.mul_(3.0)
eng.pack_weights(m._ordered_params())
torch.cuda.synchronize()
names = ["cmg.conv1", "cmg.conv2", "cmg.conv3", "cmg.conv4", "cmg.conv5", "cmg.conv6", "cmg.conv7", "cm(sigmoid)",
         "refiner conv1 x3", "refiner conv2 x3"]
for shape in [(1, 32, 48), (2, 37, 53), (1, 40, 40), (1, 300, 500)]:  # odd tile count; several tiles per CTA  # the last one: several tiles per CTA
    n, h, w = shape
    ins = [torch.rand(n, 3, h, w, device="cuda") for _ in range(4)]
    print("shape", shape)
    for layer in [0, 8, 1, 2, 3, 4, 5, 6, 7, 9]:
        try:
            a = eng.debug_layer(*ins, layer=layer, mode=_lib.MODE_FP32_SIMT)
            b = eng.debug_layer(*ins, layer=layer, mode=TENSOR_MODE)
            torch.cuda.synchronize()
            err = (a - b).abs().max().item() / max(a.abs().max().item(), 1e-30)
            bad = ((a - b).abs() > 1e-3 * a.abs().max()).float().mean().item()
            print(f"  layer {layer:2d} {names[layer]:18s} max|ref|={a.abs().max().item():.4g} rel err={err:.3e} frac bad={bad:.4f}")
            if err > 1e-3:
                d = (a - b).abs()
                idx = torch.nonzero(d > 1e-3 * a.abs().max())[:6]
                for i in idx.tolist():
                    print("     ", i, float(a[tuple(i)]), float(b[tuple(i)]))
        except Exception as e:
            print(f"  layer {layer} FAILED: {e}")
            sys.exit(1)
    out_a = eng.forward(*ins, mode=_lib.MODE_FP32_SIMT)
    out_b = eng.forward(*ins, mode=TENSOR_MODE)
    torch.cuda.synchronize()
    print("  final rel err", ((out_a - out_b).abs().max() / out_a.abs().max()).item())
