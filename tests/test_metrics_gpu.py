"""Device-side validation metrics (csrc/metrics.cu) against the libraries the reference calls
(train.py:198-251, 286-320: sklearn.metrics.accuracy_score / roc_auc_score / average_precision_score / r2_score and
scipy.stats.pearsonr), to 1e-6, ties included."""
import numpy as np
import pytest
import torch
from scipy import stats
from sklearn import metrics as skm

from chromoformer_b200.metrics import DeviceMetrics, classification_metrics, regression_metrics

pytestmark = pytest.mark.gpu


def _clf_case(n, seed, ties):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, 2, generator=g) * 2.0
    if ties:                                    # heavy ties: a handful of distinct logit pairs, and saturated softmax values
        logits = (logits * 1.5).round() / 1.5
        logits[: n // 8, 1] += 40.0
    labels = (torch.rand(n, generator=g) < torch.sigmoid(0.7 * (logits[:, 1] - logits[:, 0]))).long()
    labels[0], labels[1] = 0, 1
    return logits, labels


@pytest.mark.parametrize("n,ties", [(2, False), (100, False), (100, True), (4739, False), (4739, True), (18955, True)])
def test_classification_metrics_match_sklearn(n, ties):
    logits, labels = _clf_case(n, n + int(ties), ties)
    dev, score = classification_metrics(logits.cuda(), labels.cuda())
    score = score.cpu().numpy()
    # the score itself: softmax in FP32 (what the reference hands to sklearn)
    assert np.abs(score - logits.softmax(dim=1)[:, 1].numpy()).max() < 2e-7
    lab = labels.numpy()
    want = {"acc": skm.accuracy_score(lab, logits.argmax(dim=1).numpy()), "auc": skm.roc_auc_score(lab, score),
            "ap": skm.average_precision_score(lab, score)}
    for k, v in want.items():
        assert abs(dev[k].item() - v) < 1e-6, (k, dev[k].item(), v)


@pytest.mark.parametrize("n", [2, 64, 4739, 18955])
def test_regression_metrics_match_sklearn_scipy(n):
    g = torch.Generator().manual_seed(n)
    labels = torch.randn(n, generator=g) * 3.0 + 5.0            # log2(TPM + 1)-like: a large mean against the spread
    pred = 0.8 * labels + torch.randn(n, generator=g) + 1.0
    dev = regression_metrics(pred.cuda(), labels.cuda())
    p, l = pred.numpy().astype(np.float64), labels.numpy().astype(np.float64)
    assert abs(dev["r2"].item() - skm.r2_score(l, p)) < 1e-6
    if n > 2:
        assert abs(dev["r"].item() - stats.pearsonr(l, p)[0]) < 1e-6
    assert abs(dev["mse"].item() - np.mean((l - p) ** 2)) < 1e-6 * max(1.0, np.mean((l - p) ** 2))


def test_device_metrics_accumulates_batches_without_host_reads():
    logits, labels = _clf_case(1000, 7, True)
    m = DeviceMetrics(regression=False, capacity=128)           # grows past its initial capacity
    for i in range(0, 1000, 64):
        m.update(logits[i:i + 64].cuda(), labels[i:i + 64].cuda())
    assert len(m) == 1000
    got = m.compute()
    score = m.score.cpu().numpy()
    assert abs(got["auc"] - 100 * skm.roc_auc_score(labels.numpy(), score)) < 1e-4
    assert abs(got["ap"] - 100 * skm.average_precision_score(labels.numpy(), score)) < 1e-4
    assert abs(got["acc"] - 100 * skm.accuracy_score(labels.numpy(), logits.argmax(dim=1).numpy())) < 1e-4
    r = DeviceMetrics(regression=True)
    y = torch.randn(300)
    r.update((y * 0.5 + 0.1).view(-1, 1).cuda(), y.cuda())
    out = r.compute()
    assert abs(out["r"] - 100.0) < 1e-4 and abs(out["r2"] - 100 * skm.r2_score(y.numpy(), (y * 0.5 + 0.1).numpy())) < 1e-4


def test_metrics_refuse_cpu_tensors():
    from chromoformer_b200 import _lib
    with pytest.raises(_lib.ChromoLibError):
        classification_metrics(torch.zeros(4, 2), torch.zeros(4, dtype=torch.long))
