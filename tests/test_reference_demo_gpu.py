"""config[0]: the reference's own `demo/run_demo.py`, UNMODIFIED, executed against this package
(PYTHONPATH = repo root so that `import chromoformer` resolves to the sm_100a implementation).
Needs the staged copy under baseline/_ref/demo (python tools/stage_reference_demo.py); skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

from _util import ROOT

pytestmark = pytest.mark.gpu
DEMO = os.path.join(ROOT, "baseline", "_ref", "demo")


@pytest.mark.skipif(not os.path.exists(os.path.join(DEMO, "run_demo.py")), reason="reference demo not staged")
def test_unmodified_run_demo_script(tmp_path):
    out = tmp_path / "pred.csv"
    env = dict(os.environ, PYTHONPATH=ROOT)
    proc = subprocess.run([sys.executable, os.path.join(DEMO, "run_demo.py"), "-m", os.path.join(DEMO, "demo_meta_head.csv"),
                           "-d", os.path.join(DEMO, "demo_data"), "-o", str(out)], env=env, capture_output=True, text=True,
                          timeout=600, cwd=str(tmp_path))
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert "ROC-AUC" in proc.stdout and "Accuracy" in proc.stdout
    got = pd.read_csv(out)
    want = pd.read_csv(os.path.join(DEMO, "random_prediction.out")).iloc[:len(got)]
    assert list(got.gene_id) == list(want.gene_id)
    # untrained seed-123 weights, FP32 path: the reference's published predictions for these genes
    assert np.abs(got.prediction.to_numpy() - want.prediction.to_numpy()).max() < 2e-5
