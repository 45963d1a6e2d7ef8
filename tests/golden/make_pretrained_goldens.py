"""Fixture generator: the reference's published predictions of its two PRETRAINED demo checkpoints
(demo/prediction.out = classifier E003 conf1 fold1, demo/prediction.reg.out = regressor) as one small .npz, so that
tests/test_pretrained_demo_gpu.py can switch on wherever the `.pt` files are supplied (they are absent from
/root/reference: .MISSING_LARGE_BLOBS).  Run in the build container:  python tests/golden/make_pretrained_goldens.py"""
import os

import numpy as np
import pandas as pd

REF = os.environ.get("CHROMOFORMER_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

clf = pd.read_csv(os.path.join(REF, "demo", "prediction.out"))
reg = pd.read_csv(os.path.join(REF, "demo", "prediction.reg.out"))
meta = pd.read_csv(os.path.join(REF, "demo", "demo_meta.csv"))
assert list(clf.gene_id) == list(meta.gene_id) == list(reg.gene_id)
np.savez_compressed(os.path.join(HERE, "demo_pretrained_predictions.npz"),
                    gene_id=np.array(meta.gene_id.tolist()), label=meta.label.to_numpy(np.int64),
                    expression=meta.expression.to_numpy(np.float64),
                    prediction_clf=clf.prediction.to_numpy(np.float64), prediction_reg=reg.prediction.to_numpy(np.float64))
print("wrote demo_pretrained_predictions.npz:", len(meta), "genes")
