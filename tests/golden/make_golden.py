"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run here (the build container), never on the GPU box:

    python tests/golden/make_golden.py

It imports ``chromoformer`` from /root/reference ($CHROMOFORMER_REF) with ``.cuda()``
patched to identity (the reference hard-codes it, net.py:52,129), and writes

  smoke_main.npz      known answers of net.py:431-568 (-3.1917 / -3.1917 / -0.1900) + logits
  demo_items.npz      ChromoformerDataset items of the 100 demo genes (features, valid spans,
                      interaction_freq, labels) — masks are stored as spans and verified to
                      reconstruct the reference masks exactly
  demo_logits.npz     reference logits of ChromoformerClassifier(seed=123) on those items and
                      the `prediction` column of demo/random_prediction.out
  raw_regions.npz     raw FP16 [7,L] depth of 4 demo genes (0, 1, 5 and 8 pCREs; both strands)
                      for the input-path parity tests
  train_golden.npz    reference regressor/classifier: loss, logits, per-tensor gradient
                      checksums, a few full gradients and 3 AdamW steps on a seeded synthetic batch
"""
import os
import sys

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CHROMOFORMER_REF", "/root/reference")
sys.path.insert(0, ROOT)


def import_reference():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    import chromoformer.net as net      # noqa
    import chromoformer.data as data    # noqa
    assert os.path.realpath(net.__file__).startswith(os.path.realpath(REF)), net.__file__
    return net, data


def smoke_inputs():
    """The RNG draw sequence of net.py:437-466 (must follow the three constructors)."""
    bsz, i_max = 8, 8
    ns = (20, 80, 400)
    x_p = [torch.randn([bsz, 1, n, 7]) for n in ns]
    x_c = [torch.randn([bsz, i_max, n, 7]) for n in ns]
    m_p = [torch.randn([bsz, 1, 1, n, n]).bool() for n in ns]
    m_c = [torch.randn([bsz, i_max, 1, n, n]).bool() for n in ns]
    i_m = [torch.randn([bsz, 1, 1 + i_max, 1 + i_max]).bool() for _ in ns]
    freq = torch.randn([bsz, 1 + i_max, 1 + i_max])
    return x_p, m_p, x_c, m_c, i_m, freq


def make_smoke(net):
    m0, m1, m2 = net.Chromoformer(), net.ChromoformerClassifier(), net.ChromoformerRegressor()
    x_p, m_p, x_c, m_c, i_m, freq = smoke_inputs()
    flat = []
    for r in range(3):
        flat += [x_p[r], m_p[r], x_c[r], m_c[r], i_m[r]]
    bins = (2000, 500, 100)
    d = lambda lst: {b: t for b, t in zip(bins, lst)}
    with torch.no_grad():
        o0 = m0(*flat, freq)
        o1 = m1(d(x_p), d(m_p), d(x_c), d(m_c), d(i_m), freq)
        o2 = m2(d(x_p), d(m_p), d(x_c), d(m_c), d(i_m), freq)
    print("smoke sums", o0.sum().item(), o1.sum().item(), o2.sum().item())
    assert abs(o0.sum().item() + 3.1917) < 1e-4 and abs(o1.sum().item() + 3.1917) < 1e-4
    assert abs(o2.sum().item() + 0.1900) < 1e-4
    np.savez_compressed(os.path.join(HERE, "smoke_main.npz"), legacy=o0.numpy(), classifier=o1.numpy(),
                        regressor=o2.numpy(),
                        input_checksum=np.array([x_p[2].double().sum().item(), x_c[2].double().sum().item(),
                                                 freq.double().sum().item(),
                                                 float(sum(int(m.sum()) for m in m_c))]))


def spans_from_mask(mask2d):
    """(first, count) of the un-masked span in the centre row; (0, 0) if everything is masked."""
    n = mask2d.shape[-1]
    row = ~mask2d[n // 2]
    idx = torch.nonzero(row).flatten()
    if idx.numel() == 0:
        return 0, 0
    return int(idx[0]), int(idx.numel())


def make_demo(net, data):
    meta = os.path.join(REF, "demo", "demo_meta.csv")
    genes = pd.read_csv(meta).gene_id.tolist()
    ds = data.ChromoformerDataset(meta, os.path.join(REF, "demo", "demo_data"), genes, n_feats=7, i_max=8,
                                  w_prom=40000, w_max=40000)
    bins = (2000, 500, 100)
    xp = {b: [] for b in bins}; xc = {b: [] for b in bins}
    sp_p = {b: [] for b in bins}; sp_c = {b: [] for b in bins}
    freq, labels, kk = [], [], []
    items = []
    for i in range(len(ds)):
        it = ds[i]
        items.append(it)
        k = int((~it["interaction_masks"][2000][0, 0]).sum()) - 1
        kk.append(k)
        for b in bins:
            n = 40000 // b
            xp[b].append(it["promoter_feats"][b].numpy()); xc[b].append(it["pcre_feats"][b].numpy())
            f, c = spans_from_mask(it["promoter_pad_masks"][b][0, 0])
            sp_p[b].append((f, c))
            rebuilt = torch.ones(n, n, dtype=torch.bool); rebuilt[f:f + c, f:f + c] = False
            assert torch.equal(rebuilt, it["promoter_pad_masks"][b][0, 0])
            row = []
            for s in range(8):
                fc, cc = spans_from_mask(it["pcre_pad_masks"][b][s, 0])
                rebuilt = torch.ones(n, n, dtype=torch.bool)
                if cc:
                    rebuilt[f:f + c, fc:fc + cc] = False
                assert torch.equal(rebuilt, it["pcre_pad_masks"][b][s, 0]), (i, b, s)
                row.append((fc, cc))
            sp_c[b].append(row)
            im = torch.ones(9, 9, dtype=torch.bool); im[:k + 1, :k + 1] = False
            assert torch.equal(im, it["interaction_masks"][b][0])
        freq.append(it["interaction_freq"].numpy()); labels.append(int(it["label"]))
    out = {"freq": np.stack(freq), "labels": np.array(labels), "n_partners": np.array(kk)}
    for b in bins:
        out[f"xp_{b}"] = np.stack(xp[b]); out[f"xc_{b}"] = np.stack(xc[b])
        out[f"span_p_{b}"] = np.array(sp_p[b], dtype=np.int32); out[f"span_c_{b}"] = np.array(sp_c[b], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "demo_items.npz"), **out)

    model = net.ChromoformerClassifier(7, 128, 128,
                                       {"n_layers": 1, "n_heads": 2, "d_model": 128, "d_ff": 128},
                                       {"n_layers": 2, "n_heads": 2, "d_model": 128, "d_ff": 256},
                                       {"n_layers": 6, "n_heads": 8, "d_model": 256, "d_ff": 256}, seed=123)
    model.eval()
    logits = []
    with torch.no_grad():
        for lo in range(0, len(items), 20):
            chunk = items[lo:lo + 20]
            coll = lambda key: {b: torch.stack([it[key][b] for it in chunk]) for b in bins}
            logits.append(model(coll("promoter_feats"), coll("promoter_pad_masks"), coll("pcre_feats"),
                                coll("pcre_pad_masks"), coll("interaction_masks"),
                                torch.stack([it["interaction_freq"] for it in chunk])))
    logits = torch.cat(logits).numpy()
    pred_file = pd.read_csv(os.path.join(REF, "demo", "random_prediction.out")).prediction.to_numpy()
    pred = 1.0 / (1.0 + np.exp(-logits[:, 1].astype(np.float64)))
    print("demo: max |sigmoid(logit1) - random_prediction.out| =", np.abs(pred - pred_file).max())
    assert np.abs(pred - pred_file).max() < 1e-6
    np.savez_compressed(os.path.join(HERE, "demo_logits.npz"), logits=logits, random_prediction=pred_file)
    return genes


def make_raw(genes):
    meta = pd.read_csv(os.path.join(REF, "demo", "demo_meta.csv"))
    meta["k"] = meta.neighbors.fillna("").apply(lambda s: len([x for x in s.split(";") if x]))
    picks = []
    for k, strand in ((0, None), (1, "-"), (5, None), (8, "+")):
        sub = meta[(meta.k == k) & ((meta.strand == strand) if strand else True)]
        picks.append(sub.iloc[0])
    out = {"genes": np.array([p.gene_id for p in picks]), "strands": np.array([p.strand for p in picks]),
           "index": np.array([genes.index(p.gene_id) for p in picks])}
    for gi, p in enumerate(picks):
        raw = np.load(os.path.join(REF, "demo", "demo_data", f"{p.chrom}:{p.start - 20000}-{p.start + 20000}.npy"))
        out[f"g{gi}_promoter"] = raw
        names = [x for x in (p.neighbors if isinstance(p.neighbors, str) else "").split(";") if x]
        scores = [float(s) for s in (p.scores if isinstance(p.scores, str) else "").split(";") if s]
        out[f"g{gi}_scores"] = np.array(scores, dtype=np.float64)
        for ci, name in enumerate(names):
            out[f"g{gi}_pcre{ci}"] = np.load(os.path.join(REF, "demo", "demo_data", f"{name}.npy"))
    np.savez_compressed(os.path.join(HERE, "raw_regions.npz"), **out)


def make_train(net):
    from chromoformer_b200 import synthetic
    kws = ({"n_layers": 1, "n_heads": 2, "d_model": 128, "d_ff": 128},
           {"n_layers": 2, "n_heads": 2, "d_model": 128, "d_ff": 256},
           {"n_layers": 6, "n_heads": 8, "d_model": 256, "d_ff": 256})
    batch = synthetic.make_batch(6, ragged=True, full_masks=True, seed=7)
    args = synthetic.forward_args(batch)
    out = {"input_checksum": np.array([batch["pcre_feats"][100].double().sum().item(),
                                       batch["interaction_freq"].double().sum().item(),
                                       float(batch["n_partners"].sum())])}
    for tag, cls, target in (("reg", net.ChromoformerRegressor, batch["labels_reg"].view(-1, 1)),
                             ("clf", net.ChromoformerClassifier, batch["labels_clf"])):
        model = cls(7, 128, 128, dict(kws[0]), dict(kws[1]), dict(kws[2]), seed=123)
        crit = torch.nn.MSELoss() if tag == "reg" else torch.nn.CrossEntropyLoss()
        opt = torch.optim.AdamW(model.parameters(), lr=3e-5)
        names = [n for n, _ in model.named_parameters()]
        losses = []
        for step in range(3):
            opt.zero_grad()
            logits = model(*args)
            loss = crit(logits, target)
            loss.backward()
            if step == 0:
                out[f"{tag}_logits"] = logits.detach().numpy().copy()
                gn, gs, has = [], [], []
                for n, p in model.named_parameters():
                    has.append(p.grad is not None)
                    gn.append(0.0 if p.grad is None else p.grad.double().norm().item())
                    gs.append(0.0 if p.grad is None else p.grad.double().sum().item())
                out[f"{tag}_grad_norm"] = np.array(gn); out[f"{tag}_grad_sum"] = np.array(gs)
                out[f"{tag}_has_grad"] = np.array(has)
                sd = dict(model.named_parameters())
                for key in ("fc_head.2.weight", "embed.100.lin_proj.weight", "pairwise_interaction.100.lin_proj_pcre.weight",
                            "regulation.500.transformer.layers.3.self_att.gamma_f",
                            "pairwise_interaction.2000.transformer.layers.1.self_att.ff.bias",
                            "embed.500.transformer.layers.0.ff.ln.weight"):
                    out[f"{tag}_grad::{key}"] = sd[key].grad.numpy().copy()
            opt.step()
            losses.append(loss.item())
        out[f"{tag}_losses"] = np.array(losses)
        out[f"{tag}_param_sum_after3"] = np.array([p.detach().double().sum().item() for p in model.parameters()])
        out[f"{tag}_param_after3::fc_head.0.bias"] = dict(model.named_parameters())["fc_head.0.bias"].detach().numpy().copy()
        out[f"{tag}_param_after3::embed.100.lin_proj.weight"] = \
            dict(model.named_parameters())["embed.100.lin_proj.weight"].detach().numpy().copy()
        print(tag, "losses", losses, "grad-less tensors", int((~out[f"{tag}_has_grad"]).sum()))
    out["param_names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "train_golden.npz"), **out)


if __name__ == "__main__":
    net, data = import_reference()
    make_smoke(net)
    genes = make_demo(net, data)
    make_raw(genes)
    make_train(net)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
