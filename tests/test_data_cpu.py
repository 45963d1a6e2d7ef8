"""Host side of the input path (data.py:23-215) on CPU: the re-built ChromoformerDataset against the
reference items committed in tests/golden/, and the invariants of the synthetic batch generator."""
import numpy as np
import pandas as pd
import torch

from _util import BINS, demo_batch, golden
from chromoformer_b200 import synthetic
from chromoformer_b200.data import ChromoformerDataset


def _write_fixture(tmp_path):
    raw = golden("raw_regions.npz")
    rows = []
    for gi, gene in enumerate(raw["genes"]):
        tss = 100000 * (gi + 1)
        np.save(tmp_path / f"chrT:{tss - 20000}-{tss + 20000}.npy", raw[f"g{gi}_promoter"])
        names, ci = [], 0
        while f"g{gi}_pcre{ci}" in raw.files:
            a = raw[f"g{gi}_pcre{ci}"]
            s0 = 10_000_000 * (gi + 1) + 100_000 * ci
            names.append(f"chrT:{s0}-{s0 + a.shape[1]}")
            np.save(tmp_path / f"{names[-1]}.npy", a)
            ci += 1
        rows.append(dict(gene_id=str(gene), expression=2.5 + gi, eid="E003", label=gi % 2, chrom="chrT", start=tss,
                         end=tss + 1, strand=str(raw["strands"][gi]), split=1, neighbors=";".join(names),
                         scores=";".join(str(float(s)) for s in raw[f"g{gi}_scores"])))
    meta = tmp_path / "meta.csv"
    pd.DataFrame(rows).to_csv(meta, index=False)
    return meta, rows, raw


def test_dataset_items_match_reference(tmp_path):
    meta, rows, raw = _write_fixture(tmp_path)
    ds = ChromoformerDataset(str(meta), str(tmp_path), [r["gene_id"] for r in rows])
    assert len(ds) == len(rows)
    for gi, idx in enumerate(raw["index"]):
        it = ds[gi]
        ref = demo_batch(int(idx), int(idx) + 1)
        assert it["label"].dtype == torch.int64 and int(it["label"]) == gi % 2
        for b in BINS:
            n = 40000 // b
            assert it["promoter_feats"][b].shape == (1, n, 7) and it["pcre_feats"][b].shape == (8, n, 7)
            assert (it["promoter_feats"][b] - ref["promoter_feats"][b][0]).abs().max().item() < 2e-6
            assert (it["pcre_feats"][b] - ref["pcre_feats"][b][0]).abs().max().item() < 2e-6
            assert torch.equal(it["promoter_pad_masks"][b], ref["promoter_pad_masks"][b][0])
            assert torch.equal(it["pcre_pad_masks"][b], ref["pcre_pad_masks"][b][0])
            assert torch.equal(it["interaction_masks"][b], ref["interaction_masks"][b][0])
        assert torch.equal(it["interaction_freq"], ref["interaction_freq"][0])
    reg = ChromoformerDataset(str(meta), str(tmp_path), [rows[0]["gene_id"]], regression=True)
    assert reg[0]["label"].dtype == torch.float32
    assert abs(float(reg[0]["label"]) - np.log2(2.5 + 1)) < 1e-6           # data.py:45-48
    coll = next(iter(torch.utils.data.DataLoader(ds, batch_size=4)))         # dicts keyed by int bin size
    assert set(coll["promoter_feats"]) == set(BINS) and coll["pcre_pad_masks"][100].shape == (4, 8, 1, 400, 400)


def test_synthetic_batch_follows_dataset_rules():
    b = synthetic.make_batch(64, ragged=True, full_masks=True, seed=3)
    k = b["n_partners"]
    assert k.min() >= 0 and k.max() <= 8 and len(set(k.tolist())) > 3
    for size in BINS:
        n = 40000 // size
        mp, mc, im = b["promoter_pad_masks"][size], b["pcre_pad_masks"][size], b["interaction_masks"][size]
        assert mp.shape == (64, 1, 1, n, n) and not mp.any()                  # w_prom == w_max: nothing padded
        assert mc.shape == (64, 8, 1, n, n) and im.shape == (64, 1, 9, 9)
        xc = b["pcre_feats"][size]
        for g in range(64):
            kk = int(k[g])
            assert mc[g, kk:].all()                                           # dummy slots fully masked (data.py:196-198)
            assert (xc[g, kk:] == 0).all()                                    # and zero features (data.py:192-194)
            assert not im[g, 0, :kk + 1, :kk + 1].any() and im[g, 0, kk + 1:].all() and im[g, 0, :, kk + 1:].all()
            for s in range(kk):
                row = ~mc[g, s, 0, n // 2]
                idx = torch.nonzero(row).flatten()
                nb = idx.numel()
                assert nb >= 1 and idx[-1] - idx[0] + 1 == nb                 # one contiguous valid span
                assert int(idx[0]) == (n - nb + 1) // 2                       # centred: left pad = ceil((n - nb)/2)
                assert (xc[g, s][~row] == 0).all()
        assert (b["interaction_freq"][:, 1:] == 0).all()                      # only row 0 carries scores (data.py:190)
    compact = synthetic.make_batch(64, ragged=True, full_masks=False, seed=3)
    full = synthetic.expand_full_masks(compact)
    for size in BINS:
        assert torch.equal(full["pcre_pad_masks"][size], b["pcre_pad_masks"][size])
        assert torch.equal(compact["pcre_feats"][size], b["pcre_feats"][size])
