"""Validation / training-log metrics computed where the logits are.

The reference moves every batch's output to the host (``out.cpu()``, train.py:198-251, 286-320) and calls
sklearn / scipy on the concatenation.  `DeviceMetrics` keeps logits and labels in device buffers — `update()` is an
asynchronous device-to-device copy, no synchronisation — and `compute()` runs the metric kernels of
``csrc/metrics.cu`` (exact pair-count AUROC / average precision, FP64 r² / Pearson r) and reads back ONE small
vector.  Values are fractions; the reference's logging multiplies by 100.
"""
import torch

from . import _lib

__all__ = ["DeviceMetrics", "classification_metrics", "regression_metrics"]


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def classification_metrics(logits, labels):
    """logits [n, 2] (FP32, CUDA), labels [n] -> ({"acc", "auc", "ap"} as 0-dim FP64 device tensors, score [n]).
    Same definitions as sklearn.metrics.accuracy_score / roc_auc_score / average_precision_score on
    ``softmax(logits)[:, 1]`` (train.py:214-222)."""
    if logits.device.type != "cuda":
        raise _lib.ChromoLibError("metrics run on the device the logits are on; got %s (no CPU fallback)" % logits.device)
    lib = _lib.load()
    logits = logits.detach().to(torch.float32).contiguous()
    labels = labels.detach().to(torch.int64).contiguous().view(-1)
    n, c = logits.shape
    if labels.numel() != n:
        raise ValueError(f"{n} logits rows, {labels.numel()} labels")
    score = torch.empty(n, dtype=torch.float32, device=logits.device)
    out = torch.empty(4, dtype=torch.float64, device=logits.device)
    scratch = torch.empty(4, dtype=torch.float64, device=logits.device)
    _lib.check(lib.chromo_clf_metrics(logits.data_ptr(), labels.data_ptr(), n, c, score.data_ptr(), out.data_ptr(),
                                      scratch.data_ptr(), _stream(logits.device)), "chromo_clf_metrics")
    return {"acc": out[0], "auc": out[1], "ap": out[2]}, score


def regression_metrics(pred, labels):
    """pred, labels [n] -> {"r2", "r", "mse"} as 0-dim FP64 device tensors: sklearn.metrics.r2_score(labels, pred) and
    scipy.stats.pearsonr(labels, pred)[0] (train.py:203-206)."""
    if pred.device.type != "cuda":
        raise _lib.ChromoLibError("metrics run on the device the predictions are on; got %s (no CPU fallback)" % pred.device)
    lib = _lib.load()
    pred = pred.detach().to(torch.float32).contiguous().view(-1)
    labels = labels.detach().to(torch.float32).contiguous().view(-1)
    if labels.numel() != pred.numel():
        raise ValueError(f"{pred.numel()} predictions, {labels.numel()} labels")
    out = torch.empty(3, dtype=torch.float64, device=pred.device)
    _lib.check(lib.chromo_reg_metrics(pred.data_ptr(), labels.data_ptr(), pred.numel(), out.data_ptr(),
                                      _stream(pred.device)), "chromo_reg_metrics")
    return {"r2": out[0], "r": out[1], "mse": out[2]}


class DeviceMetrics:
    """Accumulates (logits, labels) of successive batches on the device and evaluates the reference's metrics there.

    >>> m = DeviceMetrics(regression=False, capacity=len(val_set))
    >>> for d in loader: m.update(model(...), d["label"])        # no host synchronisation
    >>> m.compute()   # {"acc": .., "auc": .., "ap": ..} in percent, as train.py prints them (one 32-byte D2H copy)
    """

    def __init__(self, regression, capacity=4096, n_out=None, device="cuda"):
        self.regression = bool(regression)
        self.n_out = (1 if regression else 2) if n_out is None else int(n_out)
        self.device = torch.device(device)
        self._cap = 0
        self._logits = self._labels = None
        self._n = 0
        self.score = None
        self._reserve(int(capacity))

    def _reserve(self, cap):
        if cap <= self._cap:
            return
        logits = torch.empty(cap, self.n_out, dtype=torch.float32, device=self.device)
        labels = torch.empty(cap, dtype=torch.float32 if self.regression else torch.int64, device=self.device)
        if self._n:
            logits[:self._n].copy_(self._logits[:self._n])
            labels[:self._n].copy_(self._labels[:self._n])
        self._logits, self._labels, self._cap = logits, labels, cap

    def reset(self):
        self._n = 0

    def __len__(self):
        return self._n

    def update(self, logits, labels):
        b = int(logits.shape[0])
        if self._n + b > self._cap:
            self._reserve(max(2 * self._cap, self._n + b))
        self._logits[self._n:self._n + b].copy_(logits.detach().view(b, self.n_out), non_blocking=True)
        self._labels[self._n:self._n + b].copy_(labels.detach().view(b), non_blocking=True)
        self._n += b

    @property
    def logits(self):
        return self._logits[:self._n]

    @property
    def labels(self):
        return self._labels[:self._n]

    def compute_device(self):
        """Metrics as 0-dim device tensors (fractions): nothing leaves the GPU."""
        if self._n == 0:
            raise ValueError("DeviceMetrics.compute() before any update()")
        if self.regression:
            self.score = self.logits.view(-1)
            return regression_metrics(self.score, self.labels)
        out, self.score = classification_metrics(self.logits, self.labels)
        return out

    def compute(self):
        """The reference's log line values: percent, Python floats (the one host read of the epoch)."""
        dev = self.compute_device()
        keys = list(dev)
        vals = torch.stack([dev[k] for k in keys]).cpu().tolist()
        return {k: (v if k == "mse" else v * 100.0) for k, v in zip(keys, vals)}
