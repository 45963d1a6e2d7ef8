"""How far does the UNMODIFIED reference move when IT computes in BF16?  (yardstick for the BF16 parity bar)

Runs the reference ChromoformerClassifier(seed=123) on the 100 demo genes twice — FP32 and under
``torch.autocast("cpu", dtype=torch.bfloat16)`` (the reference's own mixed-precision mode, train.py uses
torch.cuda.amp) — and stores the autocast logits in ``demo_autocast.npz``.  Run in the build container only:

    python tests/golden/make_autocast_yardstick.py
"""
import os
import sys

import numpy as np
import pandas as pd
import torch
from sklearn import metrics

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, import_reference  # noqa: E402


def main():
    net, data = import_reference()
    meta = os.path.join(REF, "demo", "demo_meta.csv")
    genes = pd.read_csv(meta).gene_id.tolist()
    ds = data.ChromoformerDataset(meta, os.path.join(REF, "demo", "demo_data"), genes, n_feats=7, i_max=8,
                                  w_prom=40000, w_max=40000)
    bins = (2000, 500, 100)
    items = [ds[i] for i in range(len(ds))]
    model = net.ChromoformerClassifier(7, 128, 128,
                                       {"n_layers": 1, "n_heads": 2, "d_model": 128, "d_ff": 128},
                                       {"n_layers": 2, "n_heads": 2, "d_model": 128, "d_ff": 256},
                                       {"n_layers": 6, "n_heads": 8, "d_model": 256, "d_ff": 256}, seed=123).eval()

    def run(autocast):
        out = []
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            for lo in range(0, len(items), 20):
                chunk = items[lo:lo + 20]
                coll = lambda key: {b: torch.stack([it[key][b] for it in chunk]) for b in bins}
                out.append(model(coll("promoter_feats"), coll("promoter_pad_masks"), coll("pcre_feats"),
                                 coll("pcre_pad_masks"), coll("interaction_masks"),
                                 torch.stack([it["interaction_freq"] for it in chunk])).float())
        return torch.cat(out).numpy()

    fp32, amp = run(False), run(True)
    labels = np.array([int(it["label"]) for it in items])
    sig = lambda z: 1.0 / (1.0 + np.exp(-z[:, 1].astype(np.float64)))
    a32, a16 = metrics.roc_auc_score(labels, sig(fp32)), metrics.roc_auc_score(labels, sig(amp))
    p32, p16 = metrics.average_precision_score(labels, sig(fp32)), metrics.average_precision_score(labels, sig(amp))
    print("max |logit_autocast - logit_fp32| =", np.abs(amp - fp32).max())
    print("AUROC fp32 %.6f autocast %.6f (diff %.2e)" % (a32, a16, abs(a32 - a16)))
    print("AP    fp32 %.6f autocast %.6f (diff %.2e)" % (p32, p16, abs(p32 - p16)))
    np.savez_compressed(os.path.join(HERE, "demo_autocast.npz"), logits_autocast=amp, logits_fp32=fp32)


if __name__ == "__main__":
    main()
