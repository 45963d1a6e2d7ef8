"""Batched inference over many genes (BASELINE configs[1] and [4]).

Genes are independent, so a sweep is just contiguous chunks through the same kernels; the
host path double-buffers pinned-host -> device copies on a side stream under the compute
of the previous chunk.  No collective is involved at any GPU count: each rank
takes a contiguous gene range (see ``shard_range``)."""
import torch

from . import synthetic
from .parallel import shard_range  # noqa: F401  (re-exported)

_KEYS = synthetic.FORWARD_KEYS


def _slice(batch, lo, hi):
    return {k: ({b: t[lo:hi] for b, t in batch[k].items()} if isinstance(batch[k], dict) else batch[k][lo:hi])
            for k in _KEYS}


def _args(batch):
    return [batch[k] for k in _KEYS]


def pin_batch(batch):
    """Pinned-host copy of the forward arguments (setup cost, outside any timed region)."""
    pin = lambda t: t.contiguous().pin_memory()
    return {k: ({b: pin(t) for b, t in batch[k].items()} if isinstance(batch[k], dict) else pin(batch[k]))
            for k in _KEYS}


def batch_nbytes(batch):
    tot = 0
    for k in _KEYS:
        v = batch[k]
        for t in (v.values() if isinstance(v, dict) else (v,)):
            tot += t.numel() * t.element_size()
    return tot


class InferenceEngine:
    def __init__(self, model, chunk=4096):
        self.model = model
        self.chunk = int(chunk)
        self.device = model.flat_params.device
        self._copy_stream = None
        self._stage = None

    def to_device(self, batch):
        mv = lambda t: t.to(self.device).contiguous()
        return {k: ({b: mv(t) for b, t in batch[k].items()} if isinstance(batch[k], dict) else mv(batch[k]))
                for k in _KEYS}

    @torch.no_grad()
    def predict_device(self, batch, out=None):
        """Inputs already resident in HBM.  Returns logits [N, n_out] on the device."""
        n = batch["interaction_freq"].size(0)
        if out is None:
            out = torch.empty(n, int(self.model._cfg.n_out), dtype=torch.float32, device=self.device)
        for lo in range(0, n, self.chunk):
            hi = min(n, lo + self.chunk)
            out[lo:hi] = self.model.forward_batch(_slice(batch, lo, hi))
        return out

    def _staging(self, batch):
        """Two sets of device staging buffers shaped like one chunk of `batch`."""
        sig = tuple((k, b, tuple(t.shape[1:]), t.dtype) for k in _KEYS
                    for b, t in (batch[k].items() if isinstance(batch[k], dict) else ((0, batch[k]),)))
        if self._stage is not None and self._stage[0] == sig:
            return self._stage[1]
        mk = lambda t: torch.empty((self.chunk,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
        sets = [{k: ({b: mk(t) for b, t in batch[k].items()} if isinstance(batch[k], dict) else mk(batch[k]))
                 for k in _KEYS} for _ in range(2)]
        self._stage = (sig, sets)
        return sets

    @torch.no_grad()
    def predict_host(self, batch):
        """Host (ideally pinned) inputs -> host logits.  H2D of chunk i+1 overlaps compute of chunk i."""
        n = batch["interaction_freq"].size(0)
        sets = self._staging(batch)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cs = self._copy_stream
        main = torch.cuda.current_stream(self.device)
        out = torch.empty(n, int(self.model._cfg.n_out), dtype=torch.float32, device=self.device)
        bounds = [(lo, min(n, lo + self.chunk)) for lo in range(0, n, self.chunk)]
        copied = [torch.cuda.Event() for _ in bounds]
        freed = [torch.cuda.Event() for _ in bounds]

        def upload(i):
            lo, hi = bounds[i]
            dst = sets[i % 2]
            if i >= 2:
                cs.wait_event(freed[i - 2])
            else:
                cs.wait_stream(main)
            with torch.cuda.stream(cs):
                for k in _KEYS:
                    if isinstance(batch[k], dict):
                        for b, t in batch[k].items():
                            dst[k][b][:hi - lo].copy_(t[lo:hi], non_blocking=True)
                    else:
                        dst[k][:hi - lo].copy_(batch[k][lo:hi], non_blocking=True)
                copied[i].record(cs)

        upload(0)
        for i, (lo, hi) in enumerate(bounds):
            if i + 1 < len(bounds):
                upload(i + 1)
            main.wait_event(copied[i])
            out[lo:hi] = self.model.forward_batch(_slice(sets[i % 2], 0, hi - lo))
            freed[i].record(main)
        host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        host.copy_(out, non_blocking=True)
        main.synchronize()
        return host
