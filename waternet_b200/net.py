"""WaterNet as an ``nn.Module`` whose forward runs on the B200 kernels.

Mirror of the reference's ``waternet/net.py`` public surface (class names,
submodule and parameter names, forward signature and argument order) so that the
reference's checkpoints load with strict ``load_state_dict`` and callers
(``hubconf.py:75``, ``inference.py:88,191``, ``train.py:241,108``) do not change.

* ``WaterNet.forward(x, wb, ce, gc)`` (reference ``net.py:99-108``) dispatches to
  ``libwaternet_b200.so`` (``wn_forward``).  There is no CPU path: CPU tensors
  raise.
* When autograd needs a graph (training, ``train.py:100-133``), forward values and
  the 34 parameter gradients both come from the CUDA library (``wn_forward_train`` /
  ``wn_backward``: tensor-core data-gradient and weight-gradient kernels), including
  gradients of the input images when they require grad.  Only the fp32 CUDA-core mode
  re-evaluates the network with torch ops for its backward pass.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .engine import get_engine

# (in, out, kernel) of the confidence-map stack (reference net.py:12-42) and of a refiner (net.py:62-70)
CMG_SPEC = [(12, 128, 7), (128, 128, 5), (128, 128, 3), (128, 64, 1), (64, 64, 7), (64, 64, 5), (64, 64, 3), (64, 3, 3)]
REFINER_SPEC = [(6, 32, 7), (32, 32, 5), (32, 3, 3)]


def _same_conv(cin: int, cout: int, k: int) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, kernel_size=k, stride=1, dilation=1, padding=k // 2)


class _ConvStack(nn.Module):
    """conv1..convK attributes (the names the reference's state dict uses)."""

    spec: List[tuple] = []

    def __init__(self):
        super().__init__()
        for i, (cin, cout, k) in enumerate(self.spec, start=1):
            setattr(self, f"conv{i}", _same_conv(cin, cout, k))

    def convs(self):
        return [getattr(self, f"conv{i}") for i in range(1, len(self.spec) + 1)]

    def forward(self, *args, **kwargs):
        raise NotImplementedError(
            f"{type(self).__name__} is evaluated inside WaterNet.forward by the fused CUDA path; "
            "call the WaterNet module (reference net.py:99) instead of its submodules")


class ConfidenceMapGenerator(_ConvStack):
    """Eight convs, ReLU after the first seven, sigmoid after the last (net.py:7-56)."""

    spec = CMG_SPEC

    def _graph(self, x, wb, ce, gc):
        out = torch.cat([x, wb, ce, gc], dim=1)
        layers = self.convs()
        for conv in layers[:-1]:
            out = F.relu(conv(out))
        return torch.sigmoid(layers[-1](out))


class Refiner(_ConvStack):
    """Three conv+ReLU on cat[x, x_bar] (net.py:59-80); the last ReLU is part of it."""

    spec = REFINER_SPEC

    def _graph(self, x, xbar):
        out = torch.cat([x, xbar], dim=1)
        for conv in self.convs():
            out = F.relu(conv(out))
        return out


class _KernelForward(torch.autograd.Function):
    """Forward values and all gradients (34 parameters, and the four input images when they require
    grad) from the CUDA library (wn_forward_train / wn_backward).  Only the fp32 CUDA-core mode obtains
    its gradients by re-evaluating the torch graph.
    """

    @staticmethod
    def forward(ctx, model, mode, x, wb, ce, gc, *params):
        ctx.model = model
        ctx.native = mode != _lib.MODE_FP32_SIMT
        ctx.input_needs_grad = [t.requires_grad for t in (x, wb, ce, gc)]
        if ctx.native:
            eng = model._engine_with_weights(x)
            out, ws = eng.forward_train(x, wb, ce, gc)
            ctx.engine, ctx.saved_ws = eng, ws
            ctx.weights_key = eng._weights_key
            return out
        ctx.save_for_backward(x, wb, ce, gc)
        return model._kernel_forward(x, wb, ce, gc, mode)

    @staticmethod
    def backward(ctx, grad_out):
        model = ctx.model
        params = list(model.parameters())
        if ctx.native:
            eng = ctx.engine
            if eng._weights_key != ctx.weights_key:  # parameters changed between forward and backward
                raise RuntimeError("model parameters were modified between forward and backward")
            want_in = any(ctx.input_needs_grad)
            res = eng.backward(grad_out, ctx.saved_ws, [p.shape for p in params], want_input_grads=want_in)
            grads, gin = res if want_in else (res, [None] * 4)
            ctx.saved_ws = None
            gpar = [g if p.requires_grad else None for g, p in zip(grads, params)]
            gin = [g if need else None for g, need in zip(gin, ctx.input_needs_grad)]
            return (None, None, *gin, *gpar)
        x, wb, ce, gc = ctx.saved_tensors
        with torch.enable_grad():
            ins = [t.detach().requires_grad_(t.requires_grad) for t in (x, wb, ce, gc)]
            out = model._graph(*ins)
            wanted = [t for t in ins if t.requires_grad] + [p for p in params if p.requires_grad]
            grads = torch.autograd.grad(out, wanted, grad_out, allow_unused=True)
        it = iter(grads)
        gin = [next(it) if t.requires_grad else None for t in ins]
        gpar = [next(it) if p.requires_grad else None for p in params]
        return (None, None, *gin, *gpar)


class WaterNet(nn.Module):
    """
    Gated fusion network (reference ``net.py:83-108``)::

        model = WaterNet().cuda()
        out = model(x, wb, he, gc)      # four (N,3,H,W) tensors -> (N,3,H,W)

    ``precision``: ``"default"`` = ``"bf16_fp8"`` (tensor cores: bf16 products of the operands' high parts,
    the two correction terms of the heavy layers as one fp8 MMA; ~4e-4 of the fp32 result, inside the 1e-3
    parity bar), ``"bf16x3"`` (all three terms in bf16, ~3e-5; what training always uses),
    ``"fp32"`` (CUDA-core fp32 FMA).
    """

    def __init__(self, precision: str = "default"):
        super().__init__()
        self.cmg = ConfidenceMapGenerator()
        self.wb_refiner = Refiner()
        self.ce_refiner = Refiner()
        self.gc_refiner = Refiner()
        self.precision = precision

    # -- plumbing -----------------------------------------------------------------
    def _mode(self) -> int:
        table = {"default": _lib.MODE_DEFAULT, "fp32": _lib.MODE_FP32_SIMT, "bf16x3": _lib.MODE_BF16X3,
                 "bf16_fp8": _lib.MODE_BF16_FP8}
        if self.precision not in table:
            raise ValueError(f"unknown precision {self.precision!r}; choose from {sorted(table)}")
        return table[self.precision]

    def _ordered_params(self):
        """The 34 tensors in state-dict order (what wn_pack_weights expects)."""
        out = []
        for stack in (self.cmg, self.wb_refiner, self.ce_refiner, self.gc_refiner):
            for conv in stack.convs():
                out += [conv.weight, conv.bias]
        return out

    def _engine_with_weights(self, x):
        if not x.is_cuda:
            raise _lib.WaterNetLibraryError(
                "WaterNet.forward got CPU tensors: waternet_b200 has no CPU path; move the model and inputs to a "
                "CUDA device (B200)")
        eng = get_engine(x.device)
        params = self._ordered_params()
        if params[0].device != x.device:
            raise RuntimeError(f"model parameters on {params[0].device}, inputs on {x.device}")
        key = tuple((p.data_ptr(), p._version) for p in params)
        eng.pack_weights(params, key=key)
        return eng

    def _kernel_forward(self, x, wb, ce, gc, mode):
        return self._engine_with_weights(x).forward(x, wb, ce, gc, mode)

    def _graph(self, x, wb, ce, gc):
        """Differentiable torch-op evaluation, used only to obtain gradients."""
        cm = self.cmg._graph(x, wb, ce, gc)
        r_wb = self.wb_refiner._graph(x, wb)
        r_ce = self.ce_refiner._graph(x, ce)
        r_gc = self.gc_refiner._graph(x, gc)
        return r_wb * cm[:, 0:1] + r_ce * cm[:, 1:2] + r_gc * cm[:, 2:3]

    # -- reference signature: forward(x, wb, ce, gc), ce == histogram-equalised image ---
    def forward(self, x, wb, ce, gc):
        mode = self._mode()
        needs_graph = torch.is_grad_enabled() and (
            any(t.requires_grad for t in (x, wb, ce, gc)) or any(p.requires_grad for p in self.parameters()))
        if needs_graph:
            return _KernelForward.apply(self, mode, x, wb, ce, gc, *self.parameters())
        return self._kernel_forward(x, wb, ce, gc, mode)
