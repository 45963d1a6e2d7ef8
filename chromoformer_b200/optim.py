"""Fused AdamW over the model's flat parameter buffer (one kernel launch per step).

Drop-in for ``torch.optim.AdamW(model.parameters(), lr=...)`` as used at train.py:157,196:
same hyper-parameter defaults, same decoupled update, tensors without a gradient are
skipped (no state, no decay — SURVEY A.4), and ``state_dict()`` has the stock layout
(``state[i] = {step, exp_avg, exp_avg_sq}`` indexed by position in ``model.parameters()``)
because the per-parameter moments are views into two flat buffers.
"""
import ctypes

import torch

from . import _lib


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0):
        if not hasattr(model, "flat_params"):
            raise TypeError("FusedAdamW needs a chromoformer_b200 model (flat parameter buffer)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(list(model.parameters()), defaults)
        self._model = model
        self._m = None
        self._v = None
        self._step = 0
        self.grad_scale = grad_scale

    def _ensure_state(self, keep_state=False):
        flat = self._model.flat_params
        if self._m is None or self._m.device != flat.device or self._m.numel() != flat.numel():
            self._m = torch.zeros_like(flat)
            self._v = torch.zeros_like(flat)
            if not keep_state:
                self.state.clear()

    def load_state_dict(self, state_dict):
        """Resume (the `optimizer` entry of train.py:322-343 checkpoints, written by this class or by stock
        torch.optim.AdamW): the loaded per-parameter moments are copied into the flat buffers the kernel updates and
        the state entries are re-pointed at views of those buffers; the step count is restored."""
        super().load_state_dict(state_dict)
        self._m = self._v = None
        self._ensure_state(keep_state=True)
        steps = []
        for (_, p, off, numel) in self._model._slots:
            st = self.state.get(p)
            if not st or "exp_avg" not in st:
                continue
            if off >= self._model.n_active:
                raise _lib.ChromoLibError("FusedAdamW.load_state_dict: optimiser state for a tensor that never "
                                          "receives a gradient")
            m_view = self._m[off:off + numel].view(p.shape)
            v_view = self._v[off:off + numel].view(p.shape)
            m_view.copy_(st["exp_avg"])
            v_view.copy_(st["exp_avg_sq"])
            st["exp_avg"], st["exp_avg_sq"] = m_view, v_view
            steps.append(int(float(st["step"])))
        if steps and min(steps) != max(steps):
            raise _lib.ChromoLibError("FusedAdamW.load_state_dict: per-parameter step counts differ "
                                      f"({min(steps)}..{max(steps)}); the fused kernel keeps one step for all tensors")
        self._step = steps[0] if steps else 0

    def _publish_state(self):
        """Expose the flat moments through the stock per-parameter state layout."""
        n_active = self._model.n_active
        for (_, p, off, numel) in self._model._slots:
            if off >= n_active or p.grad is None:
                continue
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = self._m[off:off + numel].view(p.shape)
                st["exp_avg_sq"] = self._v[off:off + numel].view(p.shape)
            st["step"] = torch.tensor(float(self._step))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        model = self._model
        grads = model.flat_grads
        active = model.active_parameters()
        if grads is None or all(p.grad is None for p in active):
            return loss            # nothing has a gradient yet (train.py:160-161 calls step() first)
        base = grads.data_ptr()
        for (_, p, off, _) in model._slots:
            if off < model.n_active and (p.grad is None or p.grad.data_ptr() != base + 4 * off):
                raise _lib.ChromoLibError("FusedAdamW: gradients are not the model's flat gradient views")
        self._ensure_state()
        group = self.param_groups[0]
        self._step += 1
        lib = _lib.load()
        flat = model.flat_params
        stream = torch.cuda.current_stream(flat.device).cuda_stream
        _lib.check(lib.chromo_adamw(flat.data_ptr(), grads.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                                    model.n_active, float(group["lr"]), float(group["betas"][0]),
                                    float(group["betas"][1]), float(group["eps"]), float(group["weight_decay"]),
                                    self._step, float(self.grad_scale), stream), "chromo_adamw")
        model.mark_parameters_changed()
        self._publish_state()
        return loss
