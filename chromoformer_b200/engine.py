"""Batched inference over many genes (BASELINE configs[1] and [4]).

Genes are independent, so a sweep is just contiguous chunks through the same kernels; the
host path double-buffers pinned-host -> device copies on a side stream under the compute
of the previous chunk.  No collective is involved at any GPU count: each rank
takes a contiguous gene range (see ``shard_range``)."""
import ctypes

import torch

from . import _lib, synthetic
from .parallel import shard_range  # noqa: F401  (re-exported)

_KEYS = synthetic.FORWARD_KEYS


def _slice(batch, lo, hi):
    return {k: ({b: t[lo:hi] for b, t in batch[k].items()} if isinstance(batch[k], dict) else batch[k][lo:hi])
            for k in _KEYS}


def _args(batch):
    return [batch[k] for k in _KEYS]


def pin_batch(batch):
    """Pinned-host copy of the forward arguments (setup cost, outside any timed region)."""
    pin = lambda t: t.contiguous().pin_memory()
    return {k: ({b: pin(t) for b, t in batch[k].items()} if isinstance(batch[k], dict) else pin(batch[k]))
            for k in _KEYS}


def batch_nbytes(batch):
    tot = 0
    for k in _KEYS:
        v = batch[k]
        for t in (v.values() if isinstance(v, dict) else (v,)):
            tot += t.numel() * t.element_size()
    return tot


# ---------------------------------------------------------------------------------------------- wire format
# What actually has to cross PCIe for a gene (run_demo.py:100-105 ships 1.63 MB: FP32 features + n x n masks):
#   features as FP16 [B,regions,n,F]  (the raw depth is FP16 on disk; ln(mean+1) <= 11 rounds at 2^-11 relative),
#   pad masks as (first valid bin, count) spans [B,regions,2] int32 (data.py:156-198 only ever produces such masks;
#   anything else travels as centre-row bytes), interaction masks as bytes, interaction_freq as FP32.
# 64 kB per gene at the default configuration; chromo_unpack_wire widens it on the device in one launch.
_WIRE_BINS = "bins"


def _centre_rows(m, n):
    """[B,R,1,n,n] (reference collation) or [B,R,n] -> [B,R,n] bool centre query rows."""
    if m.dim() == 5:
        return m[:, :, 0, n // 2, :]
    return m.reshape(m.size(0), -1, n)


def _spans_of(rows):
    """Centre-row masks [B,R,n] (True = padded) -> ([B,R,2] int32 spans, exact?)."""
    valid = ~rows
    n = rows.size(-1)
    cnt = valid.sum(-1)
    lo = torch.where(cnt > 0, valid.to(torch.uint8).argmax(-1), torch.zeros_like(cnt))
    pos = torch.arange(n).view(1, 1, n)
    exact = bool(torch.equal(valid, (pos >= lo.unsqueeze(-1)) & (pos < (lo + cnt).unsqueeze(-1))))
    return torch.stack([lo, cnt], -1).to(torch.int32).contiguous(), exact


def _sparse_pack(t16, fin):
    """FP16 tensor -> (occupancy bitmap as int32 words, its non-zero values in order, non-zeros in front of every block of
    1024 values [blocks + 1]) for chromo_unpack_sparse."""
    import numpy as np
    flat = t16.reshape(-1)
    nz = (flat != 0)
    total = flat.numel()
    n_blk = (total + 1023) // 1024
    pad = torch.zeros(n_blk * 1024, dtype=torch.bool)
    pad[:total] = nz
    bits = torch.from_numpy(np.packbits(pad.numpy(), bitorder="little").view(np.int32).copy())
    off = torch.zeros(n_blk + 1, dtype=torch.int64)
    off[1:] = pad.view(n_blk, 1024).sum(1).cumsum(0)
    return fin(bits), fin(flat[nz]), fin(off.to(torch.int32))


def pack_wire(batch, pin=True, compact=False, sparse=False):
    """Forward arguments (host tensors, reference layout) -> the FP16 / span wire format in (pinned) host memory.
    This is producer-side work (a DataLoader collate_fn or `GeneBatcher` would emit it directly); it is NOT part of
    the transport being timed.

    `compact`: pCRE features travel as their valid bins only (`xc_flat` [bins, F] FP16 in (gene, slot, bin) order + `off_c`
    [B*I + 1] first bin of every region): data.py:86-97 pads every pCRE to w_max / bin_size bins with zeros and
    data.py:175-177 fills dummy slots with zeros, so nothing else carries information - 14 kB instead of 64 kB per gene of
    the demo set.  Falls back to the full tensors for a resolution whose masks are not spans or whose padded bins are not
    all zero.

    `sparse`: zero-suppressed features (`*_bits` occupancy bitmap, `*_vals` non-zero FP16 values, `*_off` running counts per
    1024 values; chromo_unpack_sparse): ln(mean + 1) is exactly 0 wherever no read fell (data.py:75-80), 37 % of the demo's
    bins.  Lossless."""
    fin = (lambda t: t.contiguous().pin_memory()) if pin else (lambda t: t.contiguous())
    bins = list(batch["promoter_feats"].keys())
    wire = {_WIRE_BINS: bins, "xp": {}, "xc": {}, "span_p": {}, "span_c": {}, "rows_p": {}, "rows_c": {}, "imask": {},
            "xc_flat": {}, "off_c": {}}
    for k in ("xp", "xc"):
        for suffix in ("_bits", "_vals", "_off"):
            wire[k + suffix] = {}
    for b in bins:
        xp, xc = batch["promoter_feats"][b], batch["pcre_feats"][b]
        n = xp.size(-2)
        if sparse:
            wire["xp_bits"][b], wire["xp_vals"][b], wire["xp_off"][b] = _sparse_pack(xp.to(torch.float16), fin)
        else:
            wire["xp"][b] = fin(xp.to(torch.float16))
        done = False
        for key, src in (("p", batch["promoter_pad_masks"][b]), ("c", batch["pcre_pad_masks"][b])):
            rows = _centre_rows(src.bool(), n)
            spans, exact = _spans_of(rows)
            if exact:
                wire["span_" + key][b] = fin(spans)
            else:                                   # arbitrary mask: ship the centre rows as bytes
                wire["rows_" + key][b] = fin(rows.contiguous())
            if key == "c" and compact and exact and not bool((xc * rows.unsqueeze(-1)).any()):
                valid = ~rows                                            # [B, I, n]
                wire["xc_flat"][b] = fin(xc[valid].to(torch.float16))     # (gene, slot, bin) order
                off = torch.zeros(valid.size(0) * valid.size(1) + 1, dtype=torch.int64)
                off[1:] = valid.sum(-1).reshape(-1).cumsum(0)
                wire["off_c"][b] = fin(off.to(torch.int32))
                done = True
        if not done and sparse:
            wire["xc_bits"][b], wire["xc_vals"][b], wire["xc_off"][b] = _sparse_pack(xc.to(torch.float16), fin)
        elif not done:
            wire["xc"][b] = fin(xc.to(torch.float16))
        wire["imask"][b] = fin(batch["interaction_masks"][b].bool())
    wire["freq"] = fin(batch["interaction_freq"].to(torch.float32))
    wire["_shape_c"] = [(b, tuple(batch["pcre_feats"][b].shape[1:])) for b in bins]     # [I, n, F] of every resolution
    wire["_shape_p"] = [(b, tuple(batch["promoter_feats"][b].shape[1:])) for b in bins]
    return wire


def wire_nbytes(wire):
    tot = 0
    for k, v in wire.items():
        if k == _WIRE_BINS or k.startswith("_"):
            continue
        for t in (v.values() if isinstance(v, dict) else (v,)):
            tot += t.numel() * t.element_size()
    return tot


class InferenceEngine:
    """`chunk`: genes per host->device copy / forward step of the host paths (the copy of chunk i+1 runs under the forward
    of chunk i).  `device_chunk`: genes per launch chain when the inputs are already in HBM; default 8 x chunk - the
    persistent kernels amortise their set-up over a launch (one B200, 18,955 genes: 2.98 M genes/s in chunks of 4096,
    3.25 M at 9478, 3.38 M as one chain; 0.44 MB of workspace per gene)."""

    def __init__(self, model, chunk=4096, device_chunk=None):
        self.model = model
        self.chunk = int(chunk)
        self.device_chunk = int(device_chunk) if device_chunk else 8 * self.chunk
        self.device = model.flat_params.device
        self._copy_stream = None
        self._stage = None

    def to_device(self, batch):
        """The forward arguments in HBM.  Also notes (on the host, once) whether the batch carries any padding at all: a
        batch without dummy pCRE slots and padded bins is run with CHROMO_F_DENSE (no ragged plan; same results)."""
        mv = lambda t: t.to(self.device).contiguous()
        dev = {k: ({b: mv(t) for b, t in batch[k].items()} if isinstance(batch[k], dict) else mv(batch[k]))
               for k in _KEYS}
        dev["dense"] = not any(bool(t.any()) for k in ("pcre_pad_masks", "interaction_masks") for t in batch[k].values())
        return dev

    @torch.no_grad()
    def predict_device(self, batch, out=None):
        """Inputs already resident in HBM.  Returns logits [N, n_out] on the device."""
        n = batch["interaction_freq"].size(0)
        if out is None:
            out = torch.empty(n, int(self.model._cfg.n_out), dtype=torch.float32, device=self.device)
        for lo in range(0, n, self.device_chunk):
            hi = min(n, lo + self.device_chunk)
            out[lo:hi] = self.model.forward_batch(_slice(batch, lo, hi), dense=bool(batch.get("dense", False)))
        return out

    def _staging(self, batch):
        """Two sets of device staging buffers shaped like one chunk of `batch`."""
        sig = tuple((k, b, tuple(t.shape[1:]), t.dtype) for k in _KEYS
                    for b, t in (batch[k].items() if isinstance(batch[k], dict) else ((0, batch[k]),)))
        if self._stage is not None and self._stage[0] == sig:
            return self._stage[1]
        mk = lambda t: torch.empty((self.chunk,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
        sets = [{k: ({b: mk(t) for b, t in batch[k].items()} if isinstance(batch[k], dict) else mk(batch[k]))
                 for k in _KEYS} for _ in range(2)]
        self._stage = (sig, sets)
        return sets

    @torch.no_grad()
    def predict_host(self, batch):
        """Host (ideally pinned) inputs -> host logits.  H2D of chunk i+1 overlaps compute of chunk i."""
        n = batch["interaction_freq"].size(0)
        sets = self._staging(batch)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cs = self._copy_stream
        main = torch.cuda.current_stream(self.device)
        out = torch.empty(n, int(self.model._cfg.n_out), dtype=torch.float32, device=self.device)
        bounds = [(lo, min(n, lo + self.chunk)) for lo in range(0, n, self.chunk)]
        copied = [torch.cuda.Event() for _ in bounds]
        freed = [torch.cuda.Event() for _ in bounds]

        def upload(i):
            lo, hi = bounds[i]
            dst = sets[i % 2]
            if i >= 2:
                cs.wait_event(freed[i - 2])
            else:
                cs.wait_stream(main)
            with torch.cuda.stream(cs):
                for k in _KEYS:
                    if isinstance(batch[k], dict):
                        for b, t in batch[k].items():
                            dst[k][b][:hi - lo].copy_(t[lo:hi], non_blocking=True)
                    else:
                        dst[k][:hi - lo].copy_(batch[k][lo:hi], non_blocking=True)
                copied[i].record(cs)

        upload(0)
        for i, (lo, hi) in enumerate(bounds):
            if i + 1 < len(bounds):
                upload(i + 1)
            main.wait_event(copied[i])
            out[lo:hi] = self.model.forward_batch(_slice(sets[i % 2], 0, hi - lo))
            freed[i].record(main)
        host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        host.copy_(out, non_blocking=True)
        main.synchronize()
        return host

    # ---- FP16 / span wire ------------------------------------------------------------------------------------
    def _sparse_slices(self, wire, n):
        """Per zero-suppressed tensor (key, bin): values per gene and, for every chunk, (first block, last block + 1, first
        value, last value + 1) of its slice of the stream."""
        import math
        out = {}
        shapes = {"xp": dict(wire["_shape_p"]), "xc": dict(wire["_shape_c"])}
        for k in ("xp", "xc"):
            for b, off in wire[k + "_off"].items():
                V = math.prod(shapes[k][b])
                if (self.chunk * V) % 1024 and n > self.chunk:
                    raise ValueError(f"zero-suppressed wire: chunk * {V} values per gene must be a multiple of 1024")
                rows = []
                for lo in range(0, n, self.chunk):
                    hi = min(n, lo + self.chunk)
                    b0, b1 = lo * V // 1024, (hi * V + 1023) // 1024
                    rows.append((b0, b1))
                idx = torch.tensor([r[0] for r in rows] + [rows[-1][1]])
                cnt = off[idx].tolist()
                out[(k, b)] = (V, [(b0, b1, cnt[i], off[b1].item() if i + 1 == len(rows) else cnt[i + 1])
                                   for i, (b0, b1) in enumerate(rows)])
        return out

    def _wire_staging(self, wire):
        shape_c = dict(wire["_shape_c"])
        shape_p = dict(wire["_shape_p"])
        n_genes = wire["freq"].size(0)
        sparse = self._sparse_slices(wire, n_genes)
        sp_caps = {kb: max(max(v1 - v0 for (_, _, v0, v1) in rows), 1) for kb, (V, rows) in sparse.items()}
        # compact pCRE streams: a staging set holds the longest chunk of THIS wire
        caps = {}
        for b, off in wire["off_c"].items():
            I = shape_c[b][0]
            n = (off.numel() - 1) // I
            edges = off[torch.arange(0, n + self.chunk, self.chunk).clamp(max=n) * I].long()
            caps[b] = int((edges[1:] - edges[:-1]).max().item()) if n else 0
        sig = tuple((k, b, tuple(t.shape[1:]), t.dtype) for k, v in wire.items() if isinstance(v, dict)
                    for b, t in v.items()) + (tuple(wire["freq"].shape[1:]), tuple(sorted(caps.items())),
                                              tuple(sorted(sp_caps.items())))
        if getattr(self, "_wstage", None) is not None and self._wstage[0] == sig:
            return self._wstage[1], self._wstage[2]
        mk = lambda t, dt=None: torch.empty((self.chunk,) + tuple(t.shape[1:]), dtype=dt or t.dtype, device=self.device)
        sets = []
        for _ in range(2):
            st = {k: {b: mk(t) for b, t in v.items()} for k, v in wire.items()
                  if isinstance(v, dict) and k not in ("xc_flat", "off_c") and not k.endswith(("_bits", "_vals", "_off"))}
            for k in ("xp", "xc"):
                st[k + "_bits"], st[k + "_vals"], st[k + "_off"] = {}, {}, {}
            for (k, b), (V, rows) in sparse.items():
                nb = (self.chunk * V + 1023) // 1024 + 1
                st[k + "_bits"][b] = torch.empty(nb * 32, dtype=torch.int32, device=self.device)
                st[k + "_off"][b] = torch.empty(nb + 1, dtype=torch.int32, device=self.device)
                st[k + "_vals"][b] = torch.empty(sp_caps[(k, b)], dtype=torch.float16, device=self.device)
            st["xc_flat"] = {b: torch.empty((max(caps[b], 1),) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
                             for b, t in wire["xc_flat"].items()}
            st["off_c"] = {b: torch.empty(self.chunk * shape_c[b][0] + 1, dtype=torch.int32, device=self.device)
                           for b in wire["off_c"]}
            st["freq"] = mk(wire["freq"])
            sets.append(st)
        wide = {"promoter_feats": {b: torch.empty((self.chunk,) + shape_p[b], dtype=torch.float32, device=self.device)
                                   for b in wire[_WIRE_BINS]},
                "pcre_feats": {b: torch.empty((self.chunk,) + shape_c[b], dtype=torch.float32, device=self.device)
                               for b in wire[_WIRE_BINS]},
                "promoter_pad_masks": {}, "pcre_pad_masks": {}}
        for b in wire[_WIRE_BINS]:
            n = shape_p[b][-2]
            for key, name, reg in (("p", "promoter_pad_masks", 1), ("c", "pcre_pad_masks", shape_c[b][0])):
                if b in wire["span_" + key]:
                    wide[name][b] = torch.empty(self.chunk, reg, n, dtype=torch.bool, device=self.device)
        self._wstage = (sig, sets, wide)
        return sets, wide

    def _unpack(self, st, wide, m, bins, base=None, sparse_base=None):
        """One chromo_unpack_wire launch: FP16 -> FP32 features, spans -> centre-row masks, for the first m genes; compact
        pCRE streams (valid bins only, `base[b]` = stream position the staged chunk starts at) by chromo_unpack_compact."""
        lib = _lib.load()
        src, dst, cnt, spans, masks, rows, nb = [], [], [], [], [], [], []
        c_src, c_sp, c_off, c_base, c_dst, c_rows, c_nb, feats = [], [], [], [], [], [], [], 0
        for b in bins:
            for k16, k32 in (("xp", "promoter_feats"), ("xc", "pcre_feats")):
                if b not in st[k16]:
                    continue
                t = st[k16][b]
                src.append(t.data_ptr()); dst.append(wide[k32][b].data_ptr()); cnt.append(m * t[0].numel())
            if b in st["xc_flat"]:
                w = wide["pcre_feats"][b]
                c_src.append(st["xc_flat"][b].data_ptr()); c_sp.append(st["span_c"][b].data_ptr())
                c_off.append(st["off_c"][b].data_ptr()); c_base.append(int(base[b])); c_dst.append(w.data_ptr())
                c_rows.append(m * w.size(1)); c_nb.append(w.size(2)); feats = w.size(3)
            for key, name in (("p", "promoter_pad_masks"), ("c", "pcre_pad_masks")):
                if b in st["span_" + key]:
                    sp = st["span_" + key][b]
                    spans.append(sp.data_ptr()); masks.append(wide[name][b].data_ptr())
                    rows.append(m * sp.size(1)); nb.append(wide[name][b].size(-1))
        arr = lambda ty, xs: (ty * max(1, len(xs)))(*xs)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(lib.chromo_unpack_wire(len(src), arr(ctypes.c_void_p, src), arr(ctypes.c_void_p, dst),
                                          arr(ctypes.c_int64, cnt), len(spans), arr(ctypes.c_void_p, spans),
                                          arr(ctypes.c_void_p, masks), arr(ctypes.c_int32, rows), arr(ctypes.c_int32, nb),
                                          stream), "chromo_unpack_wire")
        s_bits, s_vals, s_off, s_base, s_dst, s_cnt = [], [], [], [], [], []
        for k16, k32 in (("xp", "promoter_feats"), ("xc", "pcre_feats")):
            for b in st[k16 + "_bits"]:
                w = wide[k32][b]
                s_bits.append(st[k16 + "_bits"][b].data_ptr()); s_vals.append(st[k16 + "_vals"][b].data_ptr())
                s_off.append(st[k16 + "_off"][b].data_ptr()); s_base.append(int(sparse_base[(k16, b)]))
                s_dst.append(w.data_ptr()); s_cnt.append(m * w[0].numel())
        for i in range(0, len(s_bits), 8):                      # (at most 8 segments per launch)
            j = slice(i, i + 8)
            _lib.check(lib.chromo_unpack_sparse(len(s_bits[j]), arr(ctypes.c_void_p, s_bits[j]), arr(ctypes.c_void_p, s_vals[j]),
                                                arr(ctypes.c_void_p, s_off[j]), arr(ctypes.c_int32, s_base[j]),
                                                arr(ctypes.c_void_p, s_dst[j]), arr(ctypes.c_int64, s_cnt[j]), stream),
                       "chromo_unpack_sparse")
        if c_src:
            _lib.check(lib.chromo_unpack_compact(len(c_src), arr(ctypes.c_void_p, c_src), arr(ctypes.c_void_p, c_sp),
                                                 arr(ctypes.c_void_p, c_off), arr(ctypes.c_int32, c_base),
                                                 arr(ctypes.c_void_p, c_dst), arr(ctypes.c_int32, c_rows),
                                                 arr(ctypes.c_int32, c_nb), int(feats), stream), "chromo_unpack_compact")

    @torch.no_grad()
    def predict_wire(self, wire):
        """`pack_wire` output (pinned host) -> host logits: H2D of chunk i+1 (copy stream) overlaps the unpack + forward
        of chunk i."""
        bins = wire[_WIRE_BINS]
        n = wire["freq"].size(0)
        sets, wide = self._wire_staging(wire)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cs = self._copy_stream
        main = torch.cuda.current_stream(self.device)
        out = torch.empty(n, int(self.model._cfg.n_out), dtype=torch.float32, device=self.device)
        bounds = [(lo, min(n, lo + self.chunk)) for lo in range(0, n, self.chunk)]
        copied = [torch.cuda.Event() for _ in bounds]
        freed = [torch.cuda.Event() for _ in bounds]
        shape_c = dict(wire["_shape_c"])
        sparse = self._sparse_slices(wire, n)
        # stream positions of the chunk boundaries of every compact pCRE stream (host integers, once per call)
        edges = {b: off[torch.tensor([lo for lo, _ in bounds] + [n]) * shape_c[b][0]].tolist() for b, off in wire["off_c"].items()}

        def upload(i):
            lo, hi = bounds[i]
            dst = sets[i % 2]
            if i >= 2:
                cs.wait_event(freed[i - 2])
            else:
                cs.wait_stream(main)
            with torch.cuda.stream(cs):
                for k, v in wire.items():
                    if isinstance(v, dict) and k not in ("xc_flat", "off_c") and not k.endswith(("_bits", "_vals", "_off")):
                        for b, t in v.items():
                            dst[k][b][:hi - lo].copy_(t[lo:hi], non_blocking=True)
                for (k, b), (V, rows) in sparse.items():             # zero-suppressed features: the chunk's blocks
                    b0, b1, v0, v1 = rows[i]
                    dst[k + "_bits"][b][:(b1 - b0) * 32].copy_(wire[k + "_bits"][b][b0 * 32:b1 * 32], non_blocking=True)
                    dst[k + "_off"][b][:b1 - b0 + 1].copy_(wire[k + "_off"][b][b0:b1 + 1], non_blocking=True)
                    if v1 > v0:
                        dst[k + "_vals"][b][:v1 - v0].copy_(wire[k + "_vals"][b][v0:v1], non_blocking=True)
                for b, off in wire["off_c"].items():                 # compact pCRE stream: the chunk's slice of it
                    I = shape_c[b][0]
                    r0, r1 = edges[b][i], edges[b][i + 1]
                    dst["off_c"][b][:(hi - lo) * I + 1].copy_(off[lo * I:hi * I + 1], non_blocking=True)
                    if r1 > r0:
                        dst["xc_flat"][b][:r1 - r0].copy_(wire["xc_flat"][b][r0:r1], non_blocking=True)
                dst["freq"][:hi - lo].copy_(wire["freq"][lo:hi], non_blocking=True)
                copied[i].record(cs)

        upload(0)
        for i, (lo, hi) in enumerate(bounds):
            if i + 1 < len(bounds):
                upload(i + 1)
            main.wait_event(copied[i])
            st, m = sets[i % 2], hi - lo
            self._unpack(st, wide, m, bins, base={b: edges[b][i] for b in edges},
                         sparse_base={kb: rows[i][2] for kb, (V, rows) in sparse.items()})
            batch = {"promoter_feats": {b: wide["promoter_feats"][b][:m] for b in bins},
                     "pcre_feats": {b: wide["pcre_feats"][b][:m] for b in bins},
                     "promoter_pad_masks": {b: (wide["promoter_pad_masks"][b][:m] if b in st["span_p"] else st["rows_p"][b][:m])
                                            for b in bins},
                     "pcre_pad_masks": {b: (wide["pcre_pad_masks"][b][:m] if b in st["span_c"] else st["rows_c"][b][:m])
                                        for b in bins},
                     "interaction_masks": {b: st["imask"][b][:m] for b in bins},
                     "interaction_freq": st["freq"][:m]}
            out[lo:hi] = self.model.forward_batch(batch)
            freed[i].record(main)
        host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        host.copy_(out, non_blocking=True)
        main.synchronize()
        return host
