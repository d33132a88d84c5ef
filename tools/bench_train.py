"""Training-step micro-benchmark: native forward_train + backward vs torch autograd (cuDNN) on the GPU.

BASELINE configs[4] shape: batch 16, 112x112 (plus a larger case).  Times forward + backward of the
WaterNet parameters only (no VGG loss, no optimizer), CUDA events, after warm-up.
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from waternet_b200.net import WaterNet

# MACs per pixel: forward 1,089,824; backward = dgrad (all but the first layer of each stack) + wgrad
FWD_MACS = 1089824
DGRAD_MACS = FWD_MACS - 75264 - 3 * 9408
WGRAD_MACS = FWD_MACS


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    torch.manual_seed(0)
    res = []
    for (n, h, w, iters) in [(16, 112, 112, 20), (4, 512, 512, 5)]:
        m = WaterNet().cuda().train()
        ins = [torch.rand(n, 3, h, w, device="cuda") for _ in range(4)]
        tgt = torch.rand(n, 3, h, w, device="cuda")

        def native():
            m.zero_grad(set_to_none=True)
            torch.nn.functional.mse_loss(m(*ins), tgt).backward()

        def graph():
            m.zero_grad(set_to_none=True)
            torch.nn.functional.mse_loss(m._graph(*ins), tgt).backward()

        t_native = timed(native, iters)
        torch.backends.cudnn.allow_tf32 = True
        t_tf32 = timed(graph, iters)
        torch.backends.cudnn.allow_tf32 = False
        t_fp32 = timed(graph, iters)
        torch.backends.cudnn.allow_tf32 = True
        flops = 2.0 * (FWD_MACS + DGRAD_MACS + WGRAD_MACS) * n * h * w
        res.append({"batch": n, "height": h, "width": w, "native_ms": t_native, "torch_cudnn_tf32_ms": t_tf32,
                    "torch_cudnn_fp32_ms": t_fp32, "native_tflops_algorithmic": flops / (t_native * 1e-3) / 1e12,
                    "speedup_vs_tf32": t_tf32 / t_native, "speedup_vs_fp32": t_fp32 / t_native})
    print(json.dumps({"metric": "training step (forward + backward of the 34 WaterNet parameters), ms",
                      "note": "native = wn_forward_train + wn_backward (bf16x3 tensor cores); torch = autograd over "
                              "F.conv2d (cuDNN)", "results": res}))


if __name__ == "__main__":
    main()
