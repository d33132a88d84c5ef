"""One warm-up + one measured forward at a given size, for ncu captures of the conv kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from waternet_b200 import _lib
from waternet_b200.engine import get_engine
from waternet_b200.net import WaterNet

n, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (1, 1080, 1920)))
mode = {"fp32": _lib.MODE_FP32_SIMT, "bf16x3": _lib.MODE_BF16X3, "bf16_fp8": _lib.MODE_BF16_FP8,
        "default": _lib.MODE_DEFAULT}[sys.argv[4] if len(sys.argv) > 4 else "default"]
torch.manual_seed(0)
eng = get_engine("cuda:0")  # honours WATERNET_B200_DEBUG_FLAGS
m = WaterNet().cuda().eval()
eng.pack_weights(m._ordered_params())
rgb = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
out = torch.empty_like(rgb)
for _ in range(2):
    eng.enhance(rgb, mode=mode, out_u8=out)
torch.cuda.synchronize()
print("done", eng.launch_count)
