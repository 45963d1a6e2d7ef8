"""Synthetic gene batches of the default-config shape.

Feature law  x = Bernoulli(0.63) * log1p(Exponential(mean 0.85))  imitates the demo's binned
depth (mean ~0.34, 37 % exact zeros); pCRE counts follow the demo histogram and pCRE
lengths a clipped log-normal (median 5.9 kb, 1.8-40 kb) in the *ragged* variant, or are
full (k = i_max, 40 kb) in the *dense* variant.  Masks follow data.py:156-203 exactly.

Pad masks are emitted as the centre query row only ([B,regions,n]) unless
``full_masks=True`` ([B,regions,1,n,n] as the reference's DataLoader collates them) — the
n x n form is 1.5 MB/gene and only its row n//2 can influence the logits.
"""
import math

import torch

DEMO_K_HIST = (0.11, 0.08, 0.03, 0.03, 0.04, 0.05, 0.04, 0.04, 0.58)   # P(k pCREs), k = 0..8


def _features(shape, gen):
    u = torch.rand(shape, generator=gen)
    e = -0.85 * torch.log1p(-torch.rand(shape, generator=gen))         # Exponential(mean 0.85)
    return torch.where(u < 0.63, torch.log1p(e), torch.zeros(()))


def make_batch(n_genes, i_max=8, binsizes=(2000, 500, 100), w_max=40000, n_feats=7, ragged=False,
               full_masks=False, seed=0, stress=False):
    """Returns a dict with the six forward arguments (dicts keyed by int bin size, CPU tensors),
    plus ``n_partners`` [B] and ``labels_clf`` / ``labels_reg``."""
    gen = torch.Generator().manual_seed(seed)
    B, I, S = n_genes, i_max, i_max + 1
    if ragged:
        hist = torch.tensor(DEMO_K_HIST[:min(len(DEMO_K_HIST), I + 1)], dtype=torch.float64)
        if I + 1 > len(DEMO_K_HIST):
            hist = torch.cat([hist, torch.zeros(I + 1 - len(DEMO_K_HIST), dtype=torch.float64)])
            hist[-1], hist[len(DEMO_K_HIST) - 1] = hist[len(DEMO_K_HIST) - 1].clone(), 0.0
        k = torch.multinomial(hist / hist.sum(), B, replacement=True, generator=gen)
        length = torch.exp(math.log(5900.0) + 0.6 * torch.randn(B, I, generator=gen)).clamp(1800, w_max)
        length = length.round().long()
    else:
        k = torch.full((B,), I, dtype=torch.long)
        length = torch.full((B, I), w_max, dtype=torch.long)
    slot = torch.arange(I).unsqueeze(0)
    live = slot < k.unsqueeze(1)                                          # [B,I] real pCRE slots

    out = {"promoter_feats": {}, "promoter_pad_masks": {}, "pcre_feats": {}, "pcre_pad_masks": {},
           "interaction_masks": {}, "n_partners": k}
    for b in binsizes:
        n = w_max // b
        nb = torch.div(length + b - 1, b, rounding_mode="floor").clamp(max=n)    # ceil(L / bin), data.py:75
        lp = torch.div(n - nb + 1, 2, rounding_mode="floor")                       # ceil((n - nb)/2), data.py:87
        pos = torch.arange(n).view(1, 1, n)
        valid = (pos >= lp.unsqueeze(2)) & (pos < (lp + nb).unsqueeze(2)) & live.unsqueeze(2)   # [B,I,n]
        xp = _features((B, 1, n, n_feats), gen)
        xc = _features((B, I, n, n_feats), gen) * valid.unsqueeze(3)
        if stress:   # heavier tail: 1 % of the entries in U(2, 4.7)
            for t, v in ((xp, None), (xc, valid)):
                hit = torch.rand(t.shape, generator=gen) < 0.01
                if v is not None:
                    hit &= v.unsqueeze(3)
                t[hit] = 2.0 + 2.7 * torch.rand(int(hit.sum()), generator=gen)
        out["promoter_feats"][b] = xp
        out["pcre_feats"][b] = xc
        row_p = torch.zeros(B, 1, n, dtype=torch.bool)                   # w_prom == w_max: nothing padded
        row_c = ~valid                                                    # data.py:180-198
        if full_masks:
            out["promoter_pad_masks"][b] = row_p.view(B, 1, 1, 1, n).expand(B, 1, 1, n, n).contiguous()
            out["pcre_pad_masks"][b] = row_c.view(B, I, 1, 1, n).expand(B, I, 1, n, n).contiguous()
        else:
            out["promoter_pad_masks"][b] = row_p
            out["pcre_pad_masks"][b] = row_c
        idx = torch.arange(S)
        inside = (idx.view(1, S, 1) <= k.view(B, 1, 1)) & (idx.view(1, 1, S) <= k.view(B, 1, 1))
        out["interaction_masks"][b] = (~inside).unsqueeze(1)             # [B,1,S,S], data.py:200-203
    freq = torch.zeros(B, S, S)
    score = 1.5 + 1.5 * torch.rand(B, I, generator=gen)
    freq[:, 0, 1:] = score * live
    out["interaction_freq"] = freq
    out["labels_clf"] = (torch.rand(B, generator=gen) < 0.5).long()
    out["labels_reg"] = torch.randn(B, generator=gen)
    return out


FORWARD_KEYS = ("promoter_feats", "promoter_pad_masks", "pcre_feats", "pcre_pad_masks", "interaction_masks",
                "interaction_freq")


def forward_args(batch, device=None, non_blocking=False):
    """The six positional arguments of ``ChromoformerBase.forward`` from a make_batch() dict."""
    def mv(t):
        return t.to(device, non_blocking=non_blocking) if device is not None else t
    args = []
    for key in FORWARD_KEYS:
        v = batch[key]
        args.append({b: mv(t) for b, t in v.items()} if isinstance(v, dict) else mv(v))
    return args


def slice_batch(batch, lo, hi):
    out = {}
    for key, v in batch.items():
        out[key] = {b: t[lo:hi] for b, t in v.items()} if isinstance(v, dict) else v[lo:hi]
    return out


def expand_full_masks(batch):
    """Centre-row masks -> the reference's [B,regions,1,n,n] layout (every query row identical,
    which is what data.py:156-198 produces when the promoter is unpadded)."""
    out = dict(batch)
    for key in ("promoter_pad_masks", "pcre_pad_masks"):
        out[key] = {}
        for b, m in batch[key].items():
            if m.dim() == 3:
                B, R, n = m.shape
                m = m.view(B, R, 1, 1, n).expand(B, R, 1, n, n).contiguous()
            out[key][b] = m
    return out
