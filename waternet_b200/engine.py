"""Per-device engine: owns the C-ABI handle, packed weights and workspaces.

torch is plumbing here (device memory, streams); all arithmetic happens in
libwaternet_b200.so.
"""
from __future__ import annotations

import ctypes
import threading
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib

_engines: Dict[int, "Engine"] = {}
_engines_lock = threading.Lock()


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.WaterNetLibraryError(
            "waternet_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback")
    dev = torch.device("cuda" if device is None else device)
    if dev.type != "cuda":
        raise _lib.WaterNetLibraryError(f"waternet_b200 runs on CUDA devices only, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def get_engine(device=None) -> "Engine":
    dev = _require_cuda(device)
    with _engines_lock:
        eng = _engines.get(dev.index)
        if eng is None:
            eng = Engine(dev)
            _engines[dev.index] = eng
        return eng


def _stream_ptr(device: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    def __init__(self, device: torch.device):
        self.lib = _lib.load()
        self.device = device
        handle = ctypes.c_void_p()
        _lib.check(self.lib.wn_create(device.index, ctypes.byref(handle)), "wn_create")
        self.handle = handle
        self._ws: Dict[str, torch.Tensor] = {}
        self._weights_key = None
        self._weights_keepalive = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.wn_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- plumbing ----------------------------------------------------------
    def _workspace(self, tag: str, nbytes: int) -> torch.Tensor:
        buf = self._ws.get(tag)
        if buf is None or buf.numel() < nbytes:
            self._ws[tag] = None
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._ws[tag] = buf
        return buf

    def release_workspaces(self) -> None:
        self._ws.clear()

    @property
    def launch_count(self) -> int:
        return int(self.lib.wn_launch_count(self.handle))

    def enable_timing(self, on: bool = True) -> None:
        _lib.check(self.lib.wn_enable_timing(self.handle, 1 if on else 0), "wn_enable_timing")

    def read_timings(self):
        """(ms[slot], count[slot]) accumulated since the last read; synchronises the device first."""
        torch.cuda.synchronize(self.device)
        ms = (ctypes.c_float * _lib.NUM_TIMING_SLOTS)()
        cnt = (ctypes.c_int * _lib.NUM_TIMING_SLOTS)()
        _lib.check(self.lib.wn_read_timings(self.handle, ms, cnt), "wn_read_timings")
        return list(ms), list(cnt)

    # ---- weights -------------------------------------------------------------
    def pack_weights(self, params: Sequence[torch.Tensor], key=None) -> None:
        """params: the 34 tensors in state-dict order (net.py:12-42,62-70,94-97)."""
        if len(params) != _lib.NUM_PARAMS:
            raise ValueError(f"expected {_lib.NUM_PARAMS} parameter tensors, got {len(params)}")
        if key is not None and key == self._weights_key:
            return
        staged = [p.detach().to(device=self.device, dtype=torch.float32).contiguous() for p in params]
        arr = (ctypes.c_void_p * _lib.NUM_PARAMS)(*[t.data_ptr() for t in staged])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.wn_pack_weights(self.handle, arr, _stream_ptr(self.device)), "wn_pack_weights")
        self._weights_keepalive = staged  # until the async pack kernels have consumed them
        self._weights_key = key

    # ---- forward -------------------------------------------------------------
    def forward(self, x, wb, he, gc, mode: int = _lib.MODE_DEFAULT, out: Optional[torch.Tensor] = None):
        """WaterNet.forward (net.py:99-108) on (N,3,H,W) fp32 CUDA tensors of any strides."""
        ins = []
        for t in (x, wb, he, gc):
            if t.device != self.device:
                raise ValueError(f"input on {t.device}, engine on {self.device}")
            if t.dim() != 4 or t.shape[1] != 3:
                raise ValueError(f"expected (N,3,H,W) inputs, got {tuple(t.shape)}")
            ins.append(t.detach() if t.dtype == torch.float32 else t.detach().float())
        n, _, h, w = ins[0].shape
        for t in ins[1:]:
            if t.shape != ins[0].shape:
                raise ValueError("the four inputs must have the same shape")
        if out is None:
            out = torch.empty((n, 3, h, w), dtype=torch.float32, device=self.device)
        if n == 0 or h == 0 or w == 0:  # empty batch: nothing to launch (torch's convs return empty too)
            return out
        strides = (ctypes.c_int64 * 16)(*[s for t in ins for s in t.stride()])
        nbytes = self.lib.wn_forward_workspace_bytes(n, h, w, mode)
        ws = self._workspace("forward", nbytes)
        with torch.cuda.device(self.device):
            rc = self.lib.wn_forward(self.handle, ins[0].data_ptr(), ins[1].data_ptr(), ins[2].data_ptr(),
                                     ins[3].data_ptr(), strides, out.data_ptr(), n, h, w, mode,
                                     ws.data_ptr(), ws.numel(), _stream_ptr(self.device))
        _lib.check(rc, "wn_forward")
        return out

    LAYER_CHANNELS = (128, 128, 128, 64, 64, 64, 64, 3, 96, 96)

    def debug_layer(self, x, wb, he, gc, layer: int, mode: int) -> torch.Tensor:
        """Test aid (wn_debug_forward_layer): an intermediate activation as fp32 (N,C,H,W)."""
        ins = [t.detach().float() for t in (x, wb, he, gc)]
        n, _, h, w = ins[0].shape
        dst = torch.empty((n, self.LAYER_CHANNELS[layer], h, w), dtype=torch.float32, device=self.device)
        strides = (ctypes.c_int64 * 16)(*[s for t in ins for s in t.stride()])
        ws = self._workspace("forward", self.lib.wn_forward_workspace_bytes(n, h, w, mode))
        with torch.cuda.device(self.device):
            rc = self.lib.wn_debug_forward_layer(self.handle, ins[0].data_ptr(), ins[1].data_ptr(), ins[2].data_ptr(),
                                                 ins[3].data_ptr(), strides, n, h, w, mode, layer, dst.data_ptr(),
                                                 ws.data_ptr(), ws.numel(), _stream_ptr(self.device))
        _lib.check(rc, "wn_debug_forward_layer")
        return dst

    # ---- training step (wn_forward_train / wn_backward) ---------------------------------------
    def forward_train(self, x, wb, he, gc):
        """Tensor-core forward that keeps every activation.  Returns (out, saved-workspace tensor)."""
        ins = [t.detach() if t.dtype == torch.float32 else t.detach().float() for t in (x, wb, he, gc)]
        n, _, h, w = ins[0].shape
        out = torch.empty((n, 3, h, w), dtype=torch.float32, device=self.device)
        strides = (ctypes.c_int64 * 16)(*[s for t in ins for s in t.stride()])
        ws = torch.empty(self.lib.wn_train_workspace_bytes(n, h, w), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.wn_forward_train(self.handle, ins[0].data_ptr(), ins[1].data_ptr(), ins[2].data_ptr(),
                                           ins[3].data_ptr(), strides, out.data_ptr(), n, h, w, ws.data_ptr(),
                                           ws.numel(), _stream_ptr(self.device))
        _lib.check(rc, "wn_forward_train")
        return out, ws

    def backward(self, grad_out: torch.Tensor, saved_ws: torch.Tensor, shapes, want_input_grads: bool = False):
        """d(loss)/d(out) + the workspace of forward_train -> the 34 parameter gradients (state-dict order)
        and, on request, the gradients of the four input images."""
        g = grad_out.detach().to(self.device, torch.float32).contiguous()
        n, _, h, w = g.shape
        grads = [torch.empty(tuple(s), dtype=torch.float32, device=self.device) for s in shapes]
        arr = (ctypes.c_void_p * _lib.NUM_PARAMS)(*[t.data_ptr() for t in grads])
        gin, gin_arr = None, None
        if want_input_grads:
            gin = [torch.empty((n, 3, h, w), dtype=torch.float32, device=self.device) for _ in range(4)]
            gin_arr = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in gin])
        with torch.cuda.device(self.device):
            rc = self.lib.wn_backward(self.handle, g.data_ptr(), arr, gin_arr, n, h, w, saved_ws.data_ptr(),
                                      saved_ws.numel(), _stream_ptr(self.device))
        _lib.check(rc, "wn_backward")
        return (grads, gin) if want_input_grads else grads

    # ---- preprocess / postprocess ----------------------------------------------
    def preprocess(self, rgb_u8: torch.Tensor, tensors: bool = True, images: bool = False):
        """rgb_u8: uint8 (N,H,W,3) CUDA tensor.  Returns dict with the requested outputs.

        tensors -> 'x','wb','he','gc' fp32 (N,3,H,W); images -> 'wb_u8','he_u8','gc_u8' uint8 NHWC.
        """
        if rgb_u8.dtype != torch.uint8 or rgb_u8.dim() != 4 or rgb_u8.shape[3] != 3:
            raise ValueError(f"expected uint8 (N,H,W,3), got {rgb_u8.dtype} {tuple(rgb_u8.shape)}")
        rgb_u8 = rgb_u8.to(self.device).contiguous()
        n, h, w, _ = rgb_u8.shape
        if n == 0 or h == 0 or w == 0:
            res = {}
            if tensors:
                res.update({k: torch.empty((n, 3, h, w), dtype=torch.float32, device=self.device) for k in ("x", "wb", "he", "gc")})
            if images:
                res.update({k: torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device) for k in ("wb_u8", "he_u8", "gc_u8")})
            return res
        res = {}
        ptr = {k: None for k in ("x", "wb", "he", "gc", "wb_u8", "he_u8", "gc_u8")}
        if tensors:
            for k in ("x", "wb", "he", "gc"):
                res[k] = torch.empty((n, 3, h, w), dtype=torch.float32, device=self.device)
                ptr[k] = res[k].data_ptr()
        if images:
            for k in ("wb_u8", "he_u8", "gc_u8"):
                res[k] = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device)
                ptr[k] = res[k].data_ptr()
        ws = self._workspace("preprocess", self.lib.wn_preprocess_workspace_bytes(n, h, w))
        with torch.cuda.device(self.device):
            rc = self.lib.wn_preprocess_u8(self.handle, rgb_u8.data_ptr(), n, h, w, ptr["x"], ptr["wb"], ptr["he"],
                                           ptr["gc"], ptr["wb_u8"], ptr["he_u8"], ptr["gc_u8"],
                                           ws.data_ptr(), ws.numel(), _stream_ptr(self.device))
        _lib.check(rc, "wn_preprocess_u8")
        return res

    def postprocess(self, out: torch.Tensor) -> torch.Tensor:
        """ten2arr on the device: fp32 (N,3,H,W) -> uint8 (N,H,W,3) CUDA tensor."""
        out = out.detach().to(self.device, torch.float32).contiguous()
        n, c, h, w = out.shape
        if c != 3:
            raise ValueError("expected (N,3,H,W)")
        res = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device)
        if res.numel() == 0:
            return res
        with torch.cuda.device(self.device):
            rc = self.lib.wn_postprocess_u8(self.handle, out.data_ptr(), res.data_ptr(), n, h, w,
                                            _stream_ptr(self.device))
        _lib.check(rc, "wn_postprocess_u8")
        return res

    def enhance(self, rgb_u8: torch.Tensor, mode: int = _lib.MODE_DEFAULT, out_u8: Optional[torch.Tensor] = None,
                out_f32: Optional[torch.Tensor] = None) -> torch.Tensor:
        """preprocess -> forward -> postprocess on uint8 (N,H,W,3) CUDA input; returns uint8 NHWC."""
        if rgb_u8.dtype != torch.uint8 or rgb_u8.dim() != 4 or rgb_u8.shape[3] != 3:
            raise ValueError(f"expected uint8 (N,H,W,3), got {rgb_u8.dtype} {tuple(rgb_u8.shape)}")
        rgb_u8 = rgb_u8.to(self.device).contiguous()
        n, h, w, _ = rgb_u8.shape
        if out_u8 is None:
            out_u8 = torch.empty((n, h, w, 3), dtype=torch.uint8, device=self.device)
        if out_u8.numel() == 0:
            return out_u8
        ws = self._workspace("enhance", self.lib.wn_enhance_workspace_bytes(n, h, w, mode))
        with torch.cuda.device(self.device):
            rc = self.lib.wn_enhance_u8(self.handle, rgb_u8.data_ptr(), out_u8.data_ptr(),
                                        None if out_f32 is None else out_f32.data_ptr(), n, h, w, mode,
                                        ws.data_ptr(), ws.numel(), _stream_ptr(self.device))
        _lib.check(rc, "wn_enhance_u8")
        return out_u8
