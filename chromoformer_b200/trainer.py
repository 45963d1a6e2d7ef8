"""Fused training step (train.py:182-196) and its data-parallel form.

``TrainStep`` chains, on one stream and without autograd bookkeeping,
    forward (activations kept) -> loss + dlogits kernel -> backward into the flat gradient
    buffer -> [NCCL all-reduce of the gradient-receiving prefix] -> one fused AdamW launch.
The only collective of the whole path is that all-reduce: one flat FP32 bucket
of ``model.n_active`` values, summed, with the 1/world mean folded into the AdamW kernel.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .model import _BatchIO
from .parallel import allreduce_gradients
from .synthetic import FORWARD_KEYS


class TrainStep:
    def __init__(self, model, lr=3e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, regression=False,
                 process_group=None, distributed=None):
        self.model = model
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.regression = regression
        self.group = process_group
        self.world = 1
        if distributed is None:
            distributed = dist.is_available() and dist.is_initialized()
        if distributed:
            self.world = dist.get_world_size(process_group)
        flat = model.flat_params
        self.grad = torch.zeros_like(flat)
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.loss = torch.zeros(1, device=flat.device)
        self.step_count = 0
        self.flags = _lib.F_TRAINING | model._precision_flag()
        self._ws = None

    def __call__(self, batch, target):
        """batch: dict with the six forward arguments on the model's device; target: labels.
        Returns the (device) loss tensor of this rank's micro-batch; no host sync."""
        lib = _lib.load()
        model = self.model
        io = _BatchIO(model, *[[batch[k][b] for b in model.binsizes] if isinstance(batch[k], dict) else batch[k]
                               for k in FORWARD_KEYS])
        flat = model.flat_params
        dev = flat.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        need = _lib.check(lib.chromo_workspace_floats(ctypes.byref(io.cfg), io.batch, self.flags), "workspace")
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.float32, device=dev)
        ws = self._ws
        n_out = int(io.cfg.n_out)
        logits = torch.empty(io.batch, n_out, dtype=torch.float32, device=dev)
        dlogits = torch.empty_like(logits)
        cfg = ctypes.byref(io.cfg)
        _lib.check(lib.chromo_forward(cfg, flat.data_ptr(), ctypes.byref(io.struct), logits.data_ptr(),
                                      ws.data_ptr(), ws.numel(), self.flags, stream), "chromo_forward")
        if self.regression:
            t = target.to(torch.float32).contiguous()
            _lib.check(lib.chromo_mse_loss(logits.data_ptr(), t.data_ptr(), io.batch * n_out, 1.0,
                                           self.loss.data_ptr(), dlogits.data_ptr(), stream), "chromo_mse_loss")
        else:
            t = target.to(torch.int64).contiguous()
            _lib.check(lib.chromo_ce_loss(logits.data_ptr(), t.data_ptr(), io.batch, n_out, 1.0,
                                          self.loss.data_ptr(), dlogits.data_ptr(), stream), "chromo_ce_loss")
        self.grad[:model.n_active].zero_()
        _lib.check(lib.chromo_backward(cfg, flat.data_ptr(), ctypes.byref(io.struct), dlogits.data_ptr(),
                                       self.grad.data_ptr(), ws.data_ptr(), ws.numel(), self.flags, stream),
                   "chromo_backward")
        scale = allreduce_gradients(self.grad, model.n_active, self.group) if self.world > 1 else 1.0
        self.step_count += 1
        _lib.check(lib.chromo_adamw(flat.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                    model.n_active, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
                                    self.step_count, scale, stream), "chromo_adamw")
        model.mark_parameters_changed()
        self.logits = logits
        return self.loss

    def set_lr(self, lr):
        """StepLR hook (train.py:158,344)."""
        self.lr = float(lr)
