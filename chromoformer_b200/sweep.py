"""Ensemble sweep: every checkpoint x every gene (BASELINE configs[4]: 11 cell types x 4 folds =
44 checkpoints over 18,955 genes).

All checkpoints live on the GPU as flat FP32 parameter buffers (44 x 21.4 MB = 0.94 GB); switching
checkpoint is one device-to-device copy into the model's flat buffer.  Work is cut into
(checkpoint, gene-chunk) units and the units are dealt to the ranks as contiguous ranges
(`parallel.shard_range`) — no collective on the data path; results are gathered on the host.
Checkpoint files follow the reference layout (`torch.save({'net': state_dict, ...})`,
train.py:322-343; legacy key names are accepted through the `Chromoformer` class)."""
import torch

from . import parallel
from .engine import InferenceEngine, _KEYS


def _slice(batch, lo, hi):
    return {k: ({b: t[lo:hi] for b, t in batch[k].items()} if isinstance(batch[k], dict) else batch[k][lo:hi])
            for k in _KEYS}


class EnsembleSweep:
    def __init__(self, model, chunk=4096):
        self.model = model
        self.engine = InferenceEngine(model, chunk=chunk)
        self.flats = []

    def add_state_dict(self, state_dict):
        """Register one checkpoint (a reference ``ckpt['net']``); returns its index."""
        self.model.load_state_dict(state_dict)
        self.flats.append(self.model.flat_params.detach().clone())
        return len(self.flats) - 1

    def add_checkpoint(self, path):
        ckpt = torch.load(path, map_location="cpu")
        return self.add_state_dict(ckpt["net"] if "net" in ckpt else ckpt)

    def _activate(self, index):
        with torch.no_grad():
            self.model.flat_params.copy_(self.flats[index])
        self.model.mark_parameters_changed()

    @torch.no_grad()
    def run(self, batch, rank=0, world=1):
        """batch: device-resident forward arguments of ALL genes.  Returns (units, logits) for this rank:
        units = [(checkpoint, gene_lo, gene_hi)], logits = list of [gene_hi - gene_lo, n_out] device tensors."""
        n = batch["interaction_freq"].size(0)
        units = parallel.sweep_units(len(self.flats), n, self.engine.chunk)
        lo, hi = parallel.shard_range(len(units), rank, world)
        mine, out, active = units[lo:hi], [], None
        for (ck, g0, g1) in mine:
            if ck != active:
                self._activate(ck)
                active = ck
            out.append(self.model.forward_batch(_slice(batch, g0, g1), dense=bool(batch.get("dense", False))))
        return mine, out

    @staticmethod
    def assemble(n_checkpoints, n_genes, n_out, per_rank):
        """Host-side merge of the (units, logits) pairs of all ranks into [checkpoints, genes, n_out]."""
        full = torch.empty(n_checkpoints, n_genes, n_out)
        for units, logits in per_rank:
            for (ck, g0, g1), t in zip(units, logits):
                full[ck, g0:g1] = t.cpu()
        return full
