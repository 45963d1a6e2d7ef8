"""The reference's own callers, UNMODIFIED, executed against this package (SURVEY §8b "callers that must keep
working unchanged"): `demo/run_demo.py`, `demo/run_demo_regression.py` and the text of `chromoformer/train.py`.
PYTHONPATH puts this repo's `chromoformer` shim first, so `import chromoformer` resolves to the sm_100a
implementation.  Needs the staged copies under baseline/_ref (python tools/stage_reference_demo.py and the
`pip install --target baseline/_ref` of DESIGN.md §2); skipped otherwise."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest
import torch
import yaml

from _util import ROOT

pytestmark = pytest.mark.gpu
DEMO = os.path.join(ROOT, "baseline", "_ref", "demo")
REF_PKG = os.path.join(ROOT, "baseline", "_ref", "chromoformer")


def _run(script, tmp_path, precision, extra=()):
    out = tmp_path / f"pred_{precision}.csv"
    env = dict(os.environ, PYTHONPATH=ROOT, CHROMO_PRECISION=precision)
    proc = subprocess.run([sys.executable, os.path.join(DEMO, script), "-m", os.path.join(DEMO, "demo_meta_head.csv"),
                           "-d", os.path.join(DEMO, "demo_data"), "-o", str(out), *extra], env=env, capture_output=True,
                          text=True, timeout=600, cwd=str(tmp_path))
    assert proc.returncode == 0, proc.stderr[-2000:]
    return proc.stdout, pd.read_csv(out)


@pytest.mark.skipif(not os.path.exists(os.path.join(DEMO, "run_demo.py")), reason="reference demo not staged")
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unmodified_run_demo_script(tmp_path, precision):
    """config[0].  CHROMO_PRECISION=bf16 sends the unmodified script down the tcgen05 path (no code change)."""
    stdout, got = _run("run_demo.py", tmp_path, precision)
    assert "ROC-AUC" in stdout and "Accuracy" in stdout
    want = pd.read_csv(os.path.join(DEMO, "random_prediction.out")).iloc[:len(got)]
    assert list(got.gene_id) == list(want.gene_id)
    err = np.abs(got.prediction.to_numpy() - want.prediction.to_numpy()).max()
    # untrained seed-123 weights: the reference's published predictions for these genes.  FP32: exact to print
    # precision; BF16: sigmoid'(x) <= 1/4, so 1e-2 on the logits is 2.5e-3 on the prediction
    assert err < (2e-5 if precision == "fp32" else 2.5e-3), err
    if precision == "bf16":
        assert err > 0.0, "CHROMO_PRECISION=bf16 had no effect"


@pytest.mark.skipif(not os.path.exists(os.path.join(DEMO, "run_demo_regression.py")), reason="reference demo not staged")
def test_unmodified_run_demo_regression_script(tmp_path):
    """run_demo_regression.py as shipped.  It passes i_max in the dataset's n_feats slot (run_demo_regression.py:79-81),
    which makes the reference's own dataset fail on its dummy-slot zeros; this package's dataset sizes the pCRE array
    by the data's mark count instead, so the script runs.  No regression golden exists for untrained weights, so
    the check is: the script runs, prints its two metrics, and its predictions equal this package's own
    ChromoformerRegressor(seed=123) on the same items."""
    stdout, got = _run("run_demo_regression.py", tmp_path, "fp32")
    assert "R2" in stdout and "Pearson" in stdout
    from chromoformer import ChromoformerDataset, ChromoformerRegressor
    meta = os.path.join(DEMO, "demo_meta_head.csv")
    genes = pd.read_csv(meta).gene_id.tolist()
    ds = ChromoformerDataset(meta, os.path.join(DEMO, "demo_data"), genes, 8, w_prom=40000, w_max=40000)
    d = next(iter(torch.utils.data.DataLoader(ds, batch_size=len(genes))))
    model = ChromoformerRegressor(seed=123).cuda().eval()
    mv = lambda v: {b: t.cuda() for b, t in v.items()}
    with torch.no_grad():
        want = model(mv(d["promoter_feats"]), mv(d["promoter_pad_masks"]), mv(d["pcre_feats"]), mv(d["pcre_pad_masks"]),
                     mv(d["interaction_masks"]), d["interaction_freq"].cuda()).cpu().numpy().ravel()
    assert np.abs(got.prediction.to_numpy() - want).max() < 1e-5


def _tiny_dataset(root, n_genes=24, seed=0):
    rng = np.random.default_rng(seed)
    rows = []
    for g in range(n_genes):
        tss = 1_000_000 + 100_000 * g
        depth = (rng.random((7, 40000)) < 0.3) * rng.exponential(1.0, (7, 40000))
        np.save(root / f"chr1:{tss - 20000}-{tss + 20000}.npy", depth.astype(np.float16))
        names, scores = [], []
        for c in range(int(rng.integers(0, 4))):
            length = int(rng.integers(1800, 9000))
            s0 = 50_000_000 + 1_000_000 * g + 20_000 * c
            names.append(f"chr1:{s0}-{s0 + length}")
            scores.append(f"{1.5 + rng.random():.4f}")
            np.save(root / f"{names[-1]}.npy",
                    ((rng.random((7, length)) < 0.3) * rng.exponential(1.0, (7, length))).astype(np.float16))
        rows.append(dict(gene_id=f"ENSG{g:011d}", expression=float(rng.exponential(3.0)), eid="E003",
                         label=(g // 4) % 2, chrom="chr1", start=tss, end=tss + 1, strand="+-"[g % 2],
                         split=1 + g % 4, neighbors=";".join(names), scores=";".join(scores)))
    meta = root / "train.csv"
    pd.DataFrame(rows).to_csv(meta, index=False)
    return meta


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_PKG, "train.py")), reason="reference package not staged")
@pytest.mark.parametrize("regression", [False, True])
def test_reference_train_script_text_runs_unchanged(tmp_path, regression):
    """`python -m chromoformer.train` with the REFERENCE's train.py text (copied at test time from baseline/_ref into
    a scratch package whose other modules are this repo's shims): stock torch.optim.AdamW, anomaly mode, tqdm, wandb
    (disabled), sklearn metrics - all as written - on top of the sm_100a modules.  Checks the checkpoint layout of
    train.py:322-343 and that it reloads."""
    pytest.importorskip("wandb")
    pkg = tmp_path / "site" / "chromoformer"
    pkg.mkdir(parents=True)
    for f in ("__init__.py", "net.py", "data.py", "util.py"):
        shutil.copy(os.path.join(ROOT, "chromoformer", f), pkg / f)
    shutil.copy(os.path.join(REF_PKG, "train.py"), pkg / "train.py")            # the reference's text, untouched
    data = tmp_path / "data"
    data.mkdir()
    meta = _tiny_dataset(data)
    cfg = yaml.safe_load(open(os.path.join(ROOT, "chromoformer", "configs", "default.yaml")))
    cfg.update(num_epoch=3, bsz=4)
    cfg_path = tmp_path / "config.yaml"
    yaml.safe_dump(cfg, open(cfg_path, "w"))
    out = tmp_path / "ckpt.pt"
    env = dict(os.environ, PYTHONPATH=f"{tmp_path / 'site'}{os.pathsep}{ROOT}", WANDB_MODE="disabled", WANDB_SILENT="true")
    cmd = [sys.executable, "-m", "chromoformer.train", "-o", str(out), "-c", str(cfg_path), "--exp-id", "t", "-m", str(meta),
           "-d", str(data), "--fold", "1"] + (["--regression"] if regression else [])
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert proc.returncode == 0, proc.stderr[-3000:]
    assert "Validation loss=" in proc.stdout
    ckpt = torch.load(out, map_location="cpu", weights_only=False)
    want = {"net", "optimizer", "epoch", "last_val_loss", "val_score", "val_label", "last_val_r2" if regression else "last_val_auc"}
    assert set(ckpt) == want and ckpt["epoch"] == 2
    assert len(ckpt["optimizer"]["state"]) == 334
    from chromoformer import ChromoformerClassifier, ChromoformerRegressor
    model = (ChromoformerRegressor if regression else ChromoformerClassifier)(seed=1)
    model.load_state_dict(ckpt["net"])
