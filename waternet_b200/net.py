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

import threading
import weakref
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .engine import new_engine

MODES = {"default": _lib.MODE_DEFAULT, "fp32": _lib.MODE_FP32_SIMT, "bf16x3": _lib.MODE_BF16X3,
         "bf16_fp8": _lib.MODE_BF16_FP8}

# model -> {device index: Engine}.  Every module that can be called on its own (WaterNet and, like in the
# reference, its sub-modules) has a private engine per device = its own packed-weight slot in the library.
# Kept outside the module so that copy.deepcopy / pickling of a model never touches a C handle.
_model_engines = weakref.WeakKeyDictionary()
_model_engines_lock = threading.Lock()


def _param_version(p) -> int:
    try:
        return p._version
    except RuntimeError:  # tensors created under torch.inference_mode() do not track versions
        return -1


class _PackedWeightsMixin:
    """Packed-weight cache of a callable module.

    The kernels consume re-packed copies of the 34 tensors (``wn_pack_weights``).  The cache key is
    ``(data_ptr, _version)`` of every parameter plus an epoch that ``load_state_dict``, ``.to()/.cuda()/.float()``
    (``_apply``) and :meth:`invalidate_packed_weights` advance.  In-place updates through autograd-visible ops
    (optimizer steps, ``p.copy_()``, ``p.mul_()``) bump ``_version`` and are picked up automatically; writes through
    ``p.data`` (``p.data.copy_(ema)``) are invisible to ``_version`` -- call ``invalidate_packed_weights()`` after them.
    """

    def invalidate_packed_weights(self) -> None:
        object.__setattr__(self, "_pack_epoch", getattr(self, "_pack_epoch", 0) + 1)
        for child in self.children():
            if isinstance(child, _PackedWeightsMixin):
                child.invalidate_packed_weights()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_packed_weights()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_packed_weights()
        return out

    def _engine_for(self, x, params):
        """This module's private engine on x's device with ``params`` (34 tensors, state-dict order) packed."""
        if not x.is_cuda:
            raise _lib.WaterNetLibraryError(
                f"{type(self).__name__}.forward got CPU tensors: waternet_b200 has no CPU path; move the model and "
                "inputs to a CUDA device (B200)")
        if params[0].device != x.device:
            raise RuntimeError(f"model parameters on {params[0].device}, inputs on {x.device}")
        with _model_engines_lock:
            per_dev = _model_engines.setdefault(self, {})
            eng = per_dev.get(x.device.index)
            if eng is None:
                eng = per_dev[x.device.index] = new_engine(x.device)
        key = (getattr(self, "_pack_epoch", 0),) + tuple((p.data_ptr(), _param_version(p)) for p in params)
        eng.pack_weights(params, key=key)
        return eng

# (in, out, kernel) of the confidence-map stack (reference net.py:12-42) and of a refiner (net.py:62-70)
CMG_SPEC = [(12, 128, 7), (128, 128, 5), (128, 128, 3), (128, 64, 1), (64, 64, 7), (64, 64, 5), (64, 64, 3), (64, 3, 3)]
REFINER_SPEC = [(6, 32, 7), (32, 32, 5), (32, 3, 3)]


def _same_conv(cin: int, cout: int, k: int) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, kernel_size=k, stride=1, dilation=1, padding=k // 2)


class _ConvStack(_PackedWeightsMixin, nn.Module):
    """conv1..convK attributes (the names the reference's state dict uses).

    Like the reference's sub-modules (``net.py:45-56``, ``:75-80``) a stack can be called on its own.  Inside a
    ``WaterNet`` it runs with the parent's packed state dict (``wn_confidence_maps`` / ``wn_refine``); a
    free-standing instance packs its own tensors into the state-dict slots of its kind, zeros elsewhere.
    Sub-module calls are inference entry points: with autograd recording they evaluate the torch graph instead
    (the fused training path is ``WaterNet.forward``).
    """

    spec: List[tuple] = []
    precision = "default"

    def __init__(self):
        super().__init__()
        for i, (cin, cout, k) in enumerate(self.spec, start=1):
            setattr(self, f"conv{i}", _same_conv(cin, cout, k))
        object.__setattr__(self, "_parent_ref", None)   # weakref to the owning WaterNet (not a registered child)
        object.__setattr__(self, "_slot", 0)

    def convs(self):
        return [getattr(self, f"conv{i}") for i in range(1, len(self.spec) + 1)]

    def _own_params(self):
        return [t for conv in self.convs() for t in (conv.weight, conv.bias)]

    def _bind(self, parent, slot: int) -> None:
        object.__setattr__(self, "_parent_ref", weakref.ref(parent))
        object.__setattr__(self, "_slot", slot)

    def __getstate__(self):  # pickling (torch.save(model)): the weakref to the parent is re-created by WaterNet
        state = self.__dict__.copy()
        state["_parent_ref"] = None
        return state

    def __deepcopy__(self, memo):  # a copy is free-standing until a WaterNet re-binds it (weakrefs do not deep-copy)
        ref, self.__dict__["_parent_ref"] = self.__dict__.get("_parent_ref"), None
        try:
            cls = type(self)
            new = cls.__new__(cls)
            memo[id(self)] = new
            import copy as _copy
            new.__dict__.update({k: _copy.deepcopy(v, memo) for k, v in self.__dict__.items()})
            return new
        finally:
            self.__dict__["_parent_ref"] = ref

    def _mode_and_engine(self, x, zero_layout):
        """(mode, engine with the right state dict packed, slot).  zero_layout(own) -> the 34-tensor list of a
        free-standing stack."""
        parent = self._parent_ref() if self._parent_ref is not None else None
        if parent is not None:
            return parent._mode(), parent._engine_with_weights(x), self._slot
        if self.precision not in MODES:
            raise ValueError(f"unknown precision {self.precision!r}; choose from {sorted(MODES)}")
        own = self._own_params()
        return MODES[self.precision], self._engine_for(x, zero_layout(own)), 0

    @staticmethod
    def _needs_graph(tensors, params):
        return torch.is_grad_enabled() and (any(t.requires_grad for t in tensors) or any(p.requires_grad for p in params))


def _zeros_like_spec(spec, ref):
    out = []
    for cin, cout, k in spec:
        out += [torch.zeros((cout, cin, k, k), dtype=torch.float32, device=ref.device),
                torch.zeros((cout,), dtype=torch.float32, device=ref.device)]
    return out


class ConfidenceMapGenerator(_ConvStack):
    """Eight convs, ReLU after the first seven, sigmoid after the last (net.py:7-56)."""

    spec = CMG_SPEC

    def _graph(self, x, wb, ce, gc):
        out = torch.cat([x, wb, ce, gc], dim=1)
        layers = self.convs()
        for conv in layers[:-1]:
            out = F.relu(conv(out))
        return torch.sigmoid(layers[-1](out))

    def forward(self, x, wb, ce, gc):
        """Returns the three (N,1,H,W) maps ``out1, out2, out3`` like ``net.py:55-56``."""
        if self._needs_graph((x, wb, ce, gc), self._own_params()):
            maps = self._graph(x, wb, ce, gc)
        else:
            mode, eng, _ = self._mode_and_engine(
                x, lambda own: own + 3 * _zeros_like_spec(REFINER_SPEC, own[0]))
            maps = eng.confidence_maps(x, wb, ce, gc, mode)
        return torch.split(maps, [1, 1, 1], dim=1)


class Refiner(_ConvStack):
    """Three conv+ReLU on cat[x, x_bar] (net.py:59-80); the last ReLU is part of it."""

    spec = REFINER_SPEC

    def _graph(self, x, xbar):
        out = torch.cat([x, xbar], dim=1)
        for conv in self.convs():
            out = F.relu(conv(out))
        return out

    def forward(self, x, xbar):
        if self._needs_graph((x, xbar), self._own_params()):
            return self._graph(x, xbar)
        mode, eng, slot = self._mode_and_engine(
            x, lambda own: _zeros_like_spec(CMG_SPEC, own[0]) + own + 2 * _zeros_like_spec(REFINER_SPEC, own[0]))
        return eng.refine(slot, x, xbar, mode)


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


class WaterNet(_PackedWeightsMixin, nn.Module):
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
        self._bind_children()

    def _bind_children(self) -> None:
        self.cmg._bind(self, 0)
        for slot, ref in enumerate((self.wb_refiner, self.ce_refiner, self.gc_refiner)):
            ref._bind(self, slot)

    def __deepcopy__(self, memo):
        cls = type(self)
        new = cls.__new__(cls)
        memo[id(self)] = new
        import copy as _copy
        new.__dict__.update({k: _copy.deepcopy(v, memo) for k, v in self.__dict__.items()})
        new._bind_children()
        return new

    # -- plumbing -----------------------------------------------------------------
    def _mode(self) -> int:
        if self.precision not in MODES:
            raise ValueError(f"unknown precision {self.precision!r}; choose from {sorted(MODES)}")
        return MODES[self.precision]

    def _ordered_params(self):
        """The 34 tensors in state-dict order (what wn_pack_weights expects)."""
        out = []
        for stack in (self.cmg, self.wb_refiner, self.ce_refiner, self.gc_refiner):
            for conv in stack.convs():
                out += [conv.weight, conv.bias]
        return out

    def _engine_with_weights(self, x):
        """This model's private engine on x's device, its current parameters packed (no-op when unchanged)."""
        return self._engine_for(x, self._ordered_params())

    def engine(self):
        """The private engine on the device the parameters live on, current parameters packed."""
        class _On:  # what _engine_for looks at
            pass
        on = _On()
        on.device = self.cmg.conv1.weight.device
        on.is_cuda = on.device.type == "cuda"
        if on.is_cuda and on.device.index is None:
            on.device = torch.device("cuda", torch.cuda.current_device())
        return self._engine_for(on, self._ordered_params())

    def __setstate__(self, state):
        super().__setstate__(state)
        self._bind_children()

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
        if x.numel() == 0 and x.is_cuda:  # empty batch: nothing to launch (torch's convs return empty too)
            return self._engine_with_weights(x).forward(x, wb, ce, gc, mode)
        needs_graph = torch.is_grad_enabled() and (
            any(t.requires_grad for t in (x, wb, ce, gc)) or any(p.requires_grad for p in self.parameters()))
        if needs_graph:
            return _KernelForward.apply(self, mode, x, wb, ce, gc, *self.parameters())
        return self._kernel_forward(x, wb, ce, gc, mode)
