"""Host side of the wire formats (engine.pack_wire), no GPU: what the producer lays out is decoded here with plain numpy
exactly as chromo_unpack_wire / chromo_unpack_sparse / chromo_unpack_compact define it (include/chromoformer_b200.h), and
must give back the FP16-rounded features and the centre-row masks of the batch."""
import numpy as np
import torch

from chromoformer_b200 import synthetic
from chromoformer_b200.engine import pack_wire, wire_nbytes, batch_nbytes

BINS = (2000, 500, 100)


def _decode_sparse(bits, vals, off, count):
    words = bits.numpy().view(np.uint32)
    on = np.unpackbits(words.view(np.uint8), bitorder="little")[:count].astype(bool)
    out = np.zeros(count, dtype=np.float32)
    out[on] = vals.numpy().astype(np.float32)
    # the running counts every 1024 values are what lets a chunk start anywhere block-aligned
    want = np.concatenate([[0], np.cumsum(np.pad(on, (0, (-count) % 1024)).reshape(-1, 1024).sum(1))])
    assert np.array_equal(off.numpy(), want)
    return out


def _decode_compact(flat, off, spans, shape):
    I, n, F = shape
    rows = spans.shape[0] * spans.shape[1]
    sp = spans.reshape(rows, 2).numpy()
    out = np.zeros((rows, n, F), dtype=np.float32)
    o = off.numpy()
    for r in range(rows):
        lo, cnt = sp[r]
        assert o[r + 1] - o[r] == cnt
        out[r, lo:lo + cnt] = flat.numpy()[o[r]:o[r + 1]].astype(np.float32)
    return out.reshape(-1, I, n, F)


def _masks_from_spans(spans, n):
    sp = spans.numpy()
    pos = np.arange(n)
    return ~((pos >= sp[..., :1]) & (pos < sp[..., :1] + sp[..., 1:2]))


def test_sparse_wire_round_trip():
    batch = synthetic.make_batch(37, ragged=True, seed=3, stress=True)
    wire = pack_wire(batch, pin=False, sparse=True)
    for b in BINS:
        for key, name in (("xp", "promoter_feats"), ("xc", "pcre_feats")):
            want = batch[name][b].half().float().numpy().reshape(-1)
            got = _decode_sparse(wire[key + "_bits"][b], wire[key + "_vals"][b], wire[key + "_off"][b], want.size)
            assert np.array_equal(got, want)
        assert np.array_equal(_masks_from_spans(wire["span_c"][b], 40000 // b), batch["pcre_pad_masks"][b].numpy())
    assert wire_nbytes(wire) < wire_nbytes(pack_wire(batch, pin=False))


def test_compact_wire_round_trip_and_fallback():
    batch = synthetic.make_batch(29, ragged=True, seed=4)
    wire = pack_wire(batch, pin=False, compact=True)
    shapes = dict(wire["_shape_c"])
    for b in BINS:
        want = batch["pcre_feats"][b].half().float().numpy()
        got = _decode_compact(wire["xc_flat"][b], wire["off_c"][b], wire["span_c"][b], shapes[b])
        assert np.array_equal(got, want)
    assert wire_nbytes(wire) < 0.3 * wire_nbytes(pack_wire(batch, pin=False)) < 0.15 * batch_nbytes(batch)
    # data under a pad mask, or a mask that is not a span: that resolution keeps its full tensor / its mask bytes
    batch["pcre_feats"][500][0, 7, 0, 0] = 1.0
    batch["pcre_pad_masks"][500][0, 7, 0] = True
    batch["pcre_pad_masks"][2000][1, 0, ::2] = True
    batch["pcre_pad_masks"][2000][1, 0, 1::2] = False
    wire = pack_wire(batch, pin=False, compact=True)
    assert set(wire["xc_flat"]) == {100} and set(wire["xc"]) == {2000, 500}
    assert set(wire["rows_c"]) == {2000} and set(wire["span_c"]) == {500, 100}
