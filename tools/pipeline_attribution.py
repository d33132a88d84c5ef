"""Attribute per-layer time of the tensor-core forward to pipeline pieces by switching them off
(wn_debug_set_flags).  Prints per-layer ms for one 1080p image under each flag combination."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from waternet_b200 import _lib
from waternet_b200.engine import get_engine
from waternet_b200.net import WaterNet

torch.manual_seed(0)
eng = get_engine("cuda:0")
m = WaterNet().cuda().eval()
eng.pack_weights(m._ordered_params())
n, h, w = 2, 1080, 1920
rgb = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
pre = eng.preprocess(rgb)
ins = [pre[k] for k in ("x", "wb", "he", "gc")]
out = torch.empty((n, 3, h, w), device="cuda")
names = ["L1", "c2", "c3", "c4", "c5", "c6", "c7", "c8", "-", "r2", "r3"]
rows = {}
for flags, label in [(0, "normal"), (1, "no epilogue stores"), (2, "no weight refetch"), (3, "neither"),
                     (4, "hi*hi pass only"), (7, "hi*hi only, no stores, no refetch")]:
    eng.lib.wn_debug_set_flags(eng.handle, flags)
    for _ in range(2):
        eng.forward(*ins, mode=_lib.MODE_BF16X3, out=out)
    eng.enable_timing(True)
    for _ in range(3):
        eng.forward(*ins, mode=_lib.MODE_BF16X3, out=out)
    ms, cnt = eng.read_timings()
    eng.enable_timing(False)
    rows[label] = {names[i]: round(ms[i] / 3 / n, 3) for i in range(11) if cnt[i]}
eng.lib.wn_debug_set_flags(eng.handle, 0)
print(json.dumps(rows))
for label, r in rows.items():
    print(f"{label:36s}", " ".join(f"{k}={v:6.3f}" for k, v in r.items()), " total=%.2f" % sum(r.values()))
