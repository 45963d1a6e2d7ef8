"""``python -m chromoformer.train`` — same command line, config.yaml schema, fold logic,
optimiser/scheduler settings and checkpoint layout as the reference script (train.py:26-37,
52-68, 79-103, 156-158, 322-343), running on the sm_100a kernels.

Differences that do not change results: the fused AdamW (one launch) replaces
torch.optim.AdamW, anomaly detection is off, wandb is imported only with --use-wandb, and the
metrics of train.py:198-251,286-320 (sklearn / scipy on out.cpu()) are evaluated on the device
(`chromoformer_b200.metrics.DeviceMetrics`: logits never leave the GPU; one small host read per
log line instead of one per step).
"""
import argparse
import os

import pandas as pd
import torch
import yaml

from chromoformer_b200.data import ChromoformerDataset
from chromoformer_b200.metrics import DeviceMetrics
from chromoformer_b200.model import ChromoformerClassifier, ChromoformerRegressor
from chromoformer_b200.optim import FusedAdamW

from .util import seed_everything


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("-o", "--output", required=True)
    ap.add_argument("-c", "--config", required=True)
    ap.add_argument("--exp-id", required=True)
    ap.add_argument("-m", "--meta", required=True)
    ap.add_argument("-d", "--npy-dir", required=True)
    ap.add_argument("--fold", type=int, required=True)
    ap.add_argument("--binsizes", nargs="+", default=[2000, 500, 100])
    ap.add_argument("--regression", action="store_true", default=False)
    ap.add_argument("--use-wandb", action="store_true", default=False)
    ap.add_argument("--num-workers", type=int, default=8)
    ap.add_argument("--stock-adamw", action="store_true", help="use torch.optim.AdamW instead of the fused kernel")
    return ap.parse_args(argv)


def to_device(batch, device):
    return {k: ({b: t.to(device, non_blocking=True) for b, t in v.items()} if isinstance(v, dict)
                else v.to(device, non_blocking=True)) for k, v in batch.items()}


def call_model(model, d):
    return model(d["promoter_feats"], d["promoter_pad_masks"], d["pcre_feats"], d["pcre_pad_masks"],
                 d["interaction_masks"], d["interaction_freq"])


def main(argv=None):
    args = parse_args(argv)
    if not args.use_wandb:
        os.environ["WANDB_MODE"] = "disabled"
    with open(args.config) as f:
        cfg = yaml.safe_load(f)
    print(cfg)
    cfg["exp_id"] = args.exp_id
    binsizes = [int(b) for b in args.binsizes]
    seed_everything(cfg["seed"])
    wandb = None
    if args.use_wandb:
        import wandb
        wandb.init(project="chromoformer-refactoring", group=args.exp_id)
        wandb.config.update(cfg)

    meta = pd.read_csv(args.meta).sample(frac=1, random_state=cfg["seed"]).reset_index(drop=True)
    if args.regression and "expression" not in meta.columns:
        raise ValueError("`expression` column is required for training ChromoformerRegression model.")
    print("Target genes:", meta.gene_id.nunique())
    quarters = [meta[meta.split == s].gene_id.tolist() for s in (1, 2, 3, 4)]
    train_genes = [g for off in range(3) for g in quarters[(args.fold + off) % 4]]
    val_genes = quarters[(args.fold + 3) % 4]
    print(len(train_genes), len(val_genes))

    def dataset(genes):
        return ChromoformerDataset(args.meta, args.npy_dir, genes, cfg["n_feats"], cfg["i_max"], binsizes,
                                   cfg["w_prom"], cfg["w_max"], regression=args.regression)
    loader_kw = dict(batch_size=cfg["bsz"], num_workers=args.num_workers, pin_memory=True)
    train_loader = torch.utils.data.DataLoader(dataset(train_genes), shuffle=True, drop_last=True, **loader_kw)
    val_loader = torch.utils.data.DataLoader(dataset(val_genes), **loader_kw)

    cls = ChromoformerRegressor if args.regression else ChromoformerClassifier
    model = cls(cfg["n_feats"], cfg["embed"]["d_model"], cfg["d_head"], cfg["embed"], cfg["pairwise_interaction"],
                cfg["regulation"], binsizes=binsizes, seed=42, i_max=cfg["i_max"], w_max=cfg["w_max"])
    model.cuda()
    device = model.flat_params.device
    criterion = torch.nn.MSELoss() if args.regression else torch.nn.CrossEntropyLoss()
    lr = float(cfg["lr"])
    optimizer = torch.optim.AdamW(model.parameters(), lr=lr) if args.stock_adamw else FusedAdamW(model, lr=lr)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=1, gamma=cfg["gamma"])
    optimizer.zero_grad()
    optimizer.step()

    val_loss = val_key = None
    for epoch in range(1, cfg["num_epoch"]):
        model.train()
        running = torch.zeros((), device=device)
        log = DeviceMetrics(args.regression, capacity=10 * cfg["bsz"], device=device)
        for i, d in enumerate(train_loader, 1):
            d = to_device(d, device)
            if args.regression:
                d["label"] = d["label"].view(-1, 1)
            optimizer.zero_grad()
            out = call_model(model, d)
            loss = criterion(out, d["label"])
            loss.backward()
            optimizer.step()
            running += loss.detach()
            log.update(out, d["label"])
            if i % 10 == 0:
                m = log.compute()
                m.pop("mse", None)
                text = ", ".join(f"{k}={v:.4f}" for k, v in m.items())
                mean_loss = running.item() / 10.0
                print(f"E{epoch} [{i}/{len(train_loader)}] {mean_loss:.4f}, lr={optimizer.param_groups[0]['lr']}, {text}")
                if wandb:
                    wandb.log({"train/loss": mean_loss, **{f"train/{k}": v for k, v in m.items()}})
                running.zero_()
                log.reset()

        model.eval()
        val = DeviceMetrics(args.regression, capacity=len(val_genes), device=device)
        with torch.no_grad():
            for d in val_loader:
                d = to_device(d, device)
                val.update(call_model(model, d), d["label"])
        val_out, val_label = val.logits, val.labels
        val_loss = criterion(val_out, val_label.view(-1, 1) if args.regression else val_label).cpu()
        m = val.compute()
        m.pop("mse", None)
        print(f"Validation loss={val_loss:.4f}, " + ", ".join(f"{k}={v:.4f}" for k, v in m.items()))
        if wandb:
            wandb.log({"val/loss": val_loss, "val/epoch": epoch, **{f"val/{k}": v for k, v in m.items()}})
        val_key = ("last_val_r2", m["r2"]) if args.regression else ("last_val_auc", m["auc"])
        torch.save({"net": model.state_dict(), "optimizer": optimizer.state_dict(), "epoch": epoch,
                    "last_val_loss": val_loss, val_key[0]: val_key[1], "val_score": val.score.cpu().numpy(),
                    "val_label": val_label.cpu().numpy()}, args.output)
        scheduler.step()
    if wandb and val_key:
        wandb.summary.update({"last_val_loss": val_loss, val_key[0]: val_key[1]})


if __name__ == "__main__":
    main()
