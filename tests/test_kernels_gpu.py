"""GPU parity of the stand-alone C-ABI kernels: fused AdamW, losses, input path."""
import ctypes

import numpy as np
import pytest
import torch

from _util import BINS, demo_batch, golden
from chromoformer_b200 import _lib
from oracle import chromoformer_oracle as oracle

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


def test_adamw_matches_torch_and_oracle():
    lib = _lib.load()
    n = 4 * 100003
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(n, generator=g); grads = [torch.randn(n, generator=g) * 0.1 for _ in range(4)]
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=3e-5)
    po, mo, vo = p0.clone(), torch.zeros(n), torch.zeros(n)
    p, m, v = p0.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step, gr in enumerate(grads, 1):
        ref.grad = gr.clone(); opt.step()
        po, mo, vo = oracle.adamw_step(po, gr, mo, vo, step)
        gd = gr.cuda()
        _lib.check(lib.chromo_adamw(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, 3e-5, 0.9, 0.999,
                                    1e-8, 0.01, step, 1.0, _stream()))
    torch.cuda.synchronize()
    # <= 1 ULP per step (the kernel contracts x - s*(m/d) into one FMA, torch's CPU path does not)
    assert ((p.cpu() - ref.detach()).abs() / ref.detach().abs().clamp(min=1.0)).max().item() < 6e-7
    assert ((p.cpu() - po).abs() / po.abs().clamp(min=1.0)).max().item() < 6e-7
    assert (m.cpu() - opt.state[ref]["exp_avg"]).abs().max().item() < 1e-7
    assert (v.cpu() - opt.state[ref]["exp_avg_sq"]).abs().max().item() < 1e-8


def test_losses_match_torch():
    lib = _lib.load()
    g = torch.Generator().manual_seed(1)
    y = torch.randn(64, 1, generator=g); t = torch.randn(64, 1, generator=g)
    yd, td = y.cuda(), t.cuda()
    loss = torch.zeros(1, device="cuda"); dy = torch.zeros_like(yd)
    _lib.check(lib.chromo_mse_loss(yd.data_ptr(), td.data_ptr(), 64, 1.0, loss.data_ptr(), dy.data_ptr(), _stream()))
    yr = y.clone().requires_grad_(True)
    lr = torch.nn.functional.mse_loss(yr, t); lr.backward()
    assert abs(loss.item() - lr.item()) < 1e-6 and (dy.cpu() - yr.grad).abs().max().item() < 1e-7
    y = torch.randn(64, 2, generator=g) * 3; lab = (torch.rand(64, generator=g) < 0.5).long()
    yd, ld = y.cuda(), lab.cuda()
    dy = torch.zeros_like(yd)
    _lib.check(lib.chromo_ce_loss(yd.data_ptr(), ld.data_ptr(), 64, 2, 1.0, loss.data_ptr(), dy.data_ptr(), _stream()))
    yr = y.clone().requires_grad_(True)
    lr = torch.nn.functional.cross_entropy(yr, lab); lr.backward()
    assert abs(loss.item() - lr.item()) < 1e-6 and (dy.cpu() - yr.grad).abs().max().item() < 1e-7


def test_input_path_matches_reference_items():
    """data.py:68-113 on the device: raw FP16 regions of 4 demo genes -> the reference's items."""
    from chromoformer_b200.data import bin_regions_device
    raw = golden("raw_regions.npz")
    for gi, idx in enumerate(raw["index"]):
        regions = [(raw[f"g{gi}_promoter"], 0, 40000, str(raw["strands"][gi]) == "-")]
        ci = 0
        while f"g{gi}_pcre{ci}" in raw.files:
            a = raw[f"g{gi}_pcre{ci}"]
            regions.append((a, 0, a.shape[1], False))
            ci += 1
        feats, spans = bin_regions_device(regions, BINS, 40000, device="cuda")
        ref = demo_batch(int(idx), int(idx) + 1)
        for r, b in enumerate(BINS):
            f = feats[r].cpu()
            assert (f[0] - ref["promoter_feats"][b][0, 0]).abs().max().item() < 2e-6
            for c in range(ci):
                assert (f[1 + c] - ref["pcre_feats"][b][0, c]).abs().max().item() < 2e-6
            want = golden("demo_items.npz")
            sp = spans[r].cpu().numpy()
            assert tuple(sp[0]) == tuple(want[f"span_p_{b}"][idx])
            for c in range(ci):
                assert tuple(sp[1 + c]) == tuple(want[f"span_c_{b}"][idx][c])


def test_input_path_ragged_lengths_vs_oracle():
    """Odd lengths (partial last bin, 1 bp, > chunk), unaligned starts, strand flip."""
    from chromoformer_b200.data import bin_regions_device
    rng = np.random.default_rng(0)
    regions = []
    for L, start, width, flip in ((40000, 0, 40000, True), (1801, 0, 1801, False), (1, 0, 1, False),
                                  (39999, 3, 39990, False), (17003, 1, 17001, True), (100, 0, 100, False),
                                  (2000, 0, 2000, False), (33333, 7, 16001, False)):
        a = (rng.random((7, L)) < 0.4) * rng.exponential(0.9, (7, L))
        regions.append((a.astype(np.float16), start, width, flip))
    feats, spans = bin_regions_device(regions, BINS, 40000, device="cuda")
    for i, (a, start, width, flip) in enumerate(regions):
        x = a[:, start:start + width].astype(np.float32)
        for r, b in enumerate(BINS):
            n = 40000 // b
            want, lp, nb, rp = oracle.bin_and_pad(x, b, n)
            if flip:
                want, lp = want[:, ::-1], rp
            got = feats[r][i].cpu().numpy().T
            assert np.abs(got - want).max() < 3e-6, (i, b)
            assert tuple(spans[r][i].cpu().numpy()) == (lp, nb)


def test_gene_batcher_matches_reference_items(tmp_path):
    """§8(f1): GeneBatcher (raw .npy -> device batch with centre-row masks, binned on the GPU) reproduces
    the reference DataLoader items of 4 demo genes and the model output on them."""
    import pandas as pd
    from chromoformer_b200 import ChromoformerClassifier, synthetic
    from chromoformer_b200.data import ChromoformerDataset, GeneBatcher
    raw = golden("raw_regions.npz")
    rows = []
    for gi, gene in enumerate(raw["genes"]):
        tss = 100000 * (gi + 1)
        np.save(tmp_path / f"chrT:{tss - 20000}-{tss + 20000}.npy", raw[f"g{gi}_promoter"])
        names = []
        ci = 0
        while f"g{gi}_pcre{ci}" in raw.files:
            a = raw[f"g{gi}_pcre{ci}"]
            s0 = 10_000_000 * (gi + 1) + 100_000 * ci
            names.append(f"chrT:{s0}-{s0 + a.shape[1]}")
            np.save(tmp_path / f"{names[-1]}.npy", a)
            ci += 1
        rows.append(dict(gene_id=str(gene), expression=1.0, eid="E003", label=gi % 2, chrom="chrT", start=tss,
                         end=tss + 1, strand=str(raw["strands"][gi]), split=1, neighbors=";".join(names),
                         scores=";".join(str(float(s)) for s in raw[f"g{gi}_scores"])))
    meta = tmp_path / "meta.csv"
    pd.DataFrame(rows).to_csv(meta, index=False)
    ds = ChromoformerDataset(str(meta), str(tmp_path), [r["gene_id"] for r in rows])
    batch = GeneBatcher(ds, device="cuda").batch(list(range(len(rows))))
    ref = [demo_batch(int(i), int(i) + 1) for i in raw["index"]]
    for b in BINS:
        n = 40000 // b
        for gi in range(len(rows)):
            assert (batch["promoter_feats"][b][gi].cpu() - ref[gi]["promoter_feats"][b][0]).abs().max().item() < 3e-6
            assert (batch["pcre_feats"][b][gi].cpu() - ref[gi]["pcre_feats"][b][0]).abs().max().item() < 3e-6
            assert torch.equal(batch["pcre_pad_masks"][b][gi].cpu(), ref[gi]["pcre_pad_masks"][b][0, :, 0, n // 2])
            assert torch.equal(batch["promoter_pad_masks"][b][gi].cpu(), ref[gi]["promoter_pad_masks"][b][0, :, 0, n // 2])
            assert torch.equal(batch["interaction_masks"][b][gi].cpu(), ref[gi]["interaction_masks"][b][0])
    # host items of the same dataset and the device batch give the same logits
    model = ChromoformerClassifier(seed=123).cuda().eval()
    with torch.no_grad():
        got = model(*[batch[k] for k in synthetic.FORWARD_KEYS]).cpu()
        items = [ds[i] for i in range(len(rows))]
        coll = {k: {b: torch.stack([it[k][b] for it in items]).cuda() for b in BINS} for k in synthetic.FORWARD_KEYS[:5]}
        want = model(*[coll[k] for k in synthetic.FORWARD_KEYS[:5]],
                     torch.stack([it["interaction_freq"] for it in items]).cuda()).cpu()
    assert (got - want).abs().max().item() < 2e-5
    # DeviceRegionStore: the raw dataset resident in HBM, batches by index (any order, repeats allowed)
    from chromoformer_b200.data import DeviceRegionStore
    store = DeviceRegionStore(ds, device="cuda")
    order = [2, 0, 3, 1, 2]
    sb = store.batch(order)
    for key in synthetic.FORWARD_KEYS:
        v = sb[key]
        if isinstance(v, dict):
            for b in BINS:
                assert torch.equal(v[b], batch[key][b][order]), (key, b)
        else:
            assert torch.equal(v, batch[key][order])
    assert torch.equal(sb["label"], batch["label"][order])
    with torch.no_grad():
        got2 = model(*[sb[k] for k in synthetic.FORWARD_KEYS]).cpu()
    assert torch.equal(got2, got[order])
