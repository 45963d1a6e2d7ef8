"""Direct oracle parity of exactly the configurations bench.py times ("What's weak" #1):

  (a) the headline: BF16 tensor path, DENSE batch (k = 8 pCREs, 400 valid bins), 18,955 genes through InferenceEngine as
      ONE launch chain (bench.py's `value`, --chunk 18955) and in chunks of 4096 (its e2e arm) — first 64 + last 64 genes
      against the CPU oracle (as-written FP32 algorithm), < 1e-2 absolute (BASELINE.json north_star tolerance for BF16
      compute / FP32 accumulate); the two chunkings agree with each other;
  (b) the same for the dense i_max = 16 ablation (17-token Regulation attention);
  (c) stress-scaled weights: the measured BF16 error is asserted against the bound committed in
      profiles/r02_precision.md, and label agreement >= 99.9 % on the genes whose FP32 margin exceeds the tolerance.
"""
import pytest
import torch

from _util import KWS
from chromoformer_b200 import ChromoformerClassifier, synthetic
from chromoformer_b200.engine import InferenceEngine
from oracle import chromoformer_oracle as oracle

pytestmark = pytest.mark.gpu
BF16_TOL = 1e-2


def _mk(seed=123):
    return ChromoformerClassifier(7, 128, 128, dict(KWS[0]), dict(KWS[1]), dict(KWS[2]), seed=seed)


def _oracle_logits(sd, batch, lo, hi):
    part = synthetic.expand_full_masks(synthetic.slice_batch(batch, lo, hi))
    with torch.no_grad():
        return oracle.chromoformer_forward(sd, *synthetic.forward_args(part))


def test_headline_bf16_dense_sweep_vs_oracle():
    """bench.py's `value`: make_batch(18955, ragged=False, seed=0) through InferenceEngine(chunk=18955), BF16; its e2e arm:
    the same genes in chunks of 4096."""
    n = 18955
    model = _mk()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batch = synthetic.make_batch(n, ragged=False, seed=0)
    model.cuda().eval()
    model.precision = "bf16"
    eng = InferenceEngine(model, chunk=4096, device_chunk=4096)
    dev = eng.to_device(batch)
    got = eng.predict_device(dev).cpu()
    whole = InferenceEngine(model, chunk=n).predict_device(dev).cpu()        # the whole sweep as one launch chain
    assert got.shape == whole.shape == (n, 2) and torch.isfinite(got).all() and torch.isfinite(whole).all()
    for lo, hi in ((0, 64), (n - 64, n)):                 # first chunk's head, the short last chunk's tail
        want = _oracle_logits(sd, batch, lo, hi)
        for res in (got, whole):
            err = (res[lo:hi] - want).abs().max().item()
            assert err < BF16_TOL, (lo, hi, err)
    # how the sweep is cut moves a gene to another row of its 128-row tiles, which decides between equivalent BF16
    # roundings of a few intermediates (tools/debug_chunk_invariance.py): measured <= 1.5e-4 between the two chunkings,
    # most genes bit-identical - two orders of magnitude inside the tolerance
    assert (got - whole).abs().max().item() < 1e-3
    del dev
    # the e2e arm of the bench (pinned host -> device -> host) returns the same numbers (same chunk boundaries)
    part = synthetic.slice_batch(batch, 3 * 4096, n)
    host = eng.predict_host(part)
    assert torch.equal(host, got[3 * 4096:])


def test_bf16_dense_imax16_vs_oracle():
    """configs[3] ablation, dense: i_max = 16 (17 tokens per gene), two chunks with a short tail."""
    n = 300
    model = _mk(seed=7)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batch = synthetic.make_batch(n, i_max=16, ragged=False, seed=1)
    model.cuda().eval()
    model.precision = "bf16"
    eng = InferenceEngine(model, chunk=256)
    got = eng.predict_device(eng.to_device(batch)).cpu()
    for lo, hi in ((0, 24), (n - 24, n)):
        want = _oracle_logits(sd, batch, lo, hi)
        assert (got[lo:hi] - want).abs().max().item() < BF16_TOL


def _stress(model):
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 2:
                p.mul_(8.0 if name == "fc_head.2.weight" else 2.0)


def test_stress_weights_bf16_vs_oracle_and_label_agreement():
    """Stress-scaled weights (2-D weights x2, last layer x8: logits span about +-13, a CHAOTIC amplifier rather than a
    trained model).  Measured (profiles/r02_precision.md): BF16 path 1.6e-1 absolute = 1.2 % of the logit range, 99.74 %
    label agreement on 10,240 genes (100 % where the FP32 margin exceeds twice the error); the reference's own
    torch.autocast(bfloat16) is off by 8.8e-2 at HALF this logit range.  The north_star bound of 1e-2
    ABSOLUTE holds at the untrained / briefly-trained logit ranges (test_headline..., profiles/r02_precision.md) and is
    asserted here RELATIVE to the logit range.  Oracle comparison on 96 genes; label agreement on 10,240 genes against
    the strict-FP32 CUDA path (itself within 5e-4 of the oracle at this scale, test_forward_gpu.test_stress_scaled_weights)."""
    model = _mk(seed=123)
    _stress(model)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    n = 10240
    batch = synthetic.make_batch(n, ragged=True, seed=17, stress=True)
    model.cuda().eval()
    eng = InferenceEngine(model, chunk=4096)
    dev = eng.to_device(batch)
    model.precision = "fp32"
    want = eng.predict_device(dev).cpu()
    model.precision = "bf16"
    got = eng.predict_device(dev).cpu()
    ora = _oracle_logits(sd, batch, 0, 96)
    scale = want.abs().max().item()
    assert scale > 2.0, scale
    err_o = (got[:96] - ora).abs().max().item()
    err = (got - want).abs().max().item()
    assert err_o < 1.5e-2 * scale, (err_o, scale)
    assert err < 1.5e-2 * scale, (err, scale)
    margin = (want[:, 1] - want[:, 0]).abs()
    same = (got.argmax(1) == want.argmax(1))
    assert same[margin > 2 * err].all()
    assert same.float().mean().item() >= 0.997, same.float().mean().item()
