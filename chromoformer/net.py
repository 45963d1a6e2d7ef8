"""``chromoformer.net`` names (reference net.py) served by the sm_100a implementation.

``python -m chromoformer.net`` replays the reference's ``__main__`` smoke (net.py:431-568): the three
model classes on the same seeded random inputs; expected sums -3.1917, -3.1917 (legacy and dict API are
identical) and -0.1900."""
from chromoformer_b200.model import (Chromoformer, ChromoformerBase, ChromoformerClassifier,  # noqa: F401
                                     ChromoformerRegressor, EmbeddingTransformer,
                                     PairwiseInteractionTransformer, RegulationTransformer)


def _smoke():
    import torch
    models = [Chromoformer().cuda(), ChromoformerClassifier().cuda(), ChromoformerRegressor().cuda()]
    bsz, i_max, bins = 8, 8, (2000, 500, 100)
    n_of = {2000: 20, 500: 80, 100: 400}
    # same draw order as the reference script: promoter feats, pCRE feats, three mask families, frequencies
    x_p = {b: torch.randn(bsz, 1, n_of[b], 7) for b in bins}
    x_c = {b: torch.randn(bsz, i_max, n_of[b], 7) for b in bins}
    m_p = {b: torch.randn(bsz, 1, 1, n_of[b], n_of[b]).bool() for b in bins}
    m_c = {b: torch.randn(bsz, i_max, 1, n_of[b], n_of[b]).bool() for b in bins}
    i_m = {b: torch.randn(bsz, 1, 1 + i_max, 1 + i_max).bool() for b in bins}
    freq = torch.randn(bsz, 1 + i_max, 1 + i_max).cuda()
    for d in (x_p, x_c, m_p, m_c, i_m):
        for b in bins:
            d[b] = d[b].cuda()
    flat = [t for b in bins for t in (x_p[b], m_p[b], x_c[b], m_c[b], i_m[b])]
    with torch.no_grad():
        outs = [models[0](*flat, freq), models[1](x_p, m_p, x_c, m_c, i_m, freq), models[2](x_p, m_p, x_c, m_c, i_m, freq)]
    for out in outs:
        print(out.sum())
        print(out.shape)


if __name__ == "__main__":
    _smoke()
