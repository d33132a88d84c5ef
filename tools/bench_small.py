"""Latency/throughput of the forward and of the end-to-end enhance at small sizes (BASELINE configs[1])."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from waternet_b200 import _lib
from waternet_b200.engine import get_engine
from waternet_b200.net import WaterNet


def timed(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (time.perf_counter() - t0) * 1e3 / iters


torch.manual_seed(0)
eng = get_engine("cuda:0")
m = WaterNet().cuda().eval()
eng.pack_weights(m._ordered_params())
res = []
for (n, h, w) in [(1, 112, 112), (16, 112, 112), (1, 480, 720), (1, 1080, 1920)]:
    rgb = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty_like(rgb)
    pre = eng.preprocess(rgb)
    ins = [pre[k] for k in ("x", "wb", "he", "gc")]
    fo = torch.empty((n, 3, h, w), device="cuda")
    for mode, name in ((_lib.MODE_BF16X3, "bf16x3"), (_lib.MODE_FP32_SIMT, "fp32")):
        dev_ms, wall_ms = timed(lambda: eng.forward(*ins, mode=mode, out=fo), 50 if h < 500 else 10)
        e2e_ms, e2e_wall = timed(lambda: eng.enhance(rgb, mode=mode, out_u8=out), 50 if h < 500 else 10)
        res.append({"batch": n, "h": h, "w": w, "mode": name, "forward_ms": round(dev_ms, 4), "forward_wall_ms": round(wall_ms, 4),
                    "enhance_ms": round(e2e_ms, 4), "images_per_s_enhance": round(n / (e2e_ms * 1e-3), 1)})
# host-buffer call with and without CUDA-graph replay (small frames are launch-bound)
import numpy as np
from waternet_b200.api import Enhancer
for (n, h, w) in [(1, 112, 112), (1, 240, 320), (1, 480, 720)]:
    frame = np.random.default_rng(0).integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    row = {"batch": n, "h": h, "w": w, "mode": "bf16x3 host-buffer call"}
    for flag in (False, True):
        enh = Enhancer(m, cuda_graph=flag)
        pin_in = torch.from_numpy(frame).pin_memory()
        pin_out = torch.empty_like(pin_in).pin_memory()
        for _ in range(5):
            enh.enhance_pinned(pin_in, pin_out)
        t0 = time.perf_counter()
        for _ in range(100):
            enh.enhance_pinned(pin_in, pin_out)
        row["graph_ms" if flag else "launch_ms"] = round((time.perf_counter() - t0) * 10, 4)
    res.append(row)
print(json.dumps(res))
