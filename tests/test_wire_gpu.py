"""The FP16 / span wire format (engine.pack_wire -> chromo_unpack_wire -> forward; transport of run_demo.py:100-105):
bit-identical to the ordinary path on FP16-representable features, and within tolerance of the CPU oracle fed with the
ORIGINAL FP32 features (so the FP16 rounding of ln(mean+1) is inside the bound, not outside it)."""
import pytest
import torch

from _util import BINS, KWS
from chromoformer_b200 import ChromoformerClassifier, synthetic
from chromoformer_b200.engine import InferenceEngine, pack_wire, wire_nbytes, batch_nbytes
from oracle import chromoformer_oracle as oracle

pytestmark = pytest.mark.gpu


def _mk(seed=123):
    return ChromoformerClassifier(7, 128, 128, dict(KWS[0]), dict(KWS[1]), dict(KWS[2]), seed=seed)


@pytest.mark.parametrize("precision,ragged", [("fp32", True), ("bf16", True), ("bf16", False)])
def test_wire_path_equals_device_path_on_fp16_features(precision, ragged):
    model = _mk().cuda().eval()
    model.precision = precision
    n = 700
    batch = synthetic.make_batch(n, ragged=ragged, seed=21, stress=True)
    rounded = dict(batch)
    for key in ("promoter_feats", "pcre_feats"):
        rounded[key] = {b: t.half().float() for b, t in batch[key].items()}
    eng = InferenceEngine(model, chunk=256, device_chunk=256)    # 256, 256, 188: two staging sets + a short tail; same chunks on both paths
    want = eng.predict_device(eng.to_device(rounded)).cpu()
    wire = pack_wire(batch)
    assert wire_nbytes(wire) < 0.5 * batch_nbytes(batch)
    got = eng.predict_wire(wire)
    assert torch.equal(got, want)
    got2 = eng.predict_wire(pack_wire(synthetic.expand_full_masks(batch)))   # reference-collated n x n masks pack the same
    assert torch.equal(got2, want)


def test_wire_path_vs_oracle_with_fp32_features():
    model = _mk(seed=5)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batch = synthetic.make_batch(40, ragged=True, seed=22, stress=True)
    want = oracle.chromoformer_forward(sd, *synthetic.forward_args(synthetic.expand_full_masks(batch)))
    model.cuda().eval()
    eng = InferenceEngine(model, chunk=64)
    wire = pack_wire(batch)
    model.precision = "fp32"
    err32 = (eng.predict_wire(wire) - want).abs().max().item()
    model.precision = "bf16"
    err16 = (eng.predict_wire(wire) - want).abs().max().item()
    assert err32 < 1e-3, err32            # FP16 rounding of the features alone (2^-11 relative)
    assert err16 < 1e-2, err16            # north_star tolerance, wire + BF16 tensor path together


def test_wire_falls_back_to_mask_bytes_for_arbitrary_masks():
    model = _mk(seed=8).cuda().eval()
    batch = synthetic.make_batch(33, ragged=False, seed=23)
    gen = torch.Generator().manual_seed(0)
    for b in BINS:
        batch["pcre_pad_masks"][b] = torch.rand(batch["pcre_pad_masks"][b].shape, generator=gen) < 0.6
    batch["pcre_pad_masks"][100][1] = True                       # a fully masked gene
    wire = pack_wire(batch)
    assert set(wire["rows_c"]) == set(BINS) and not wire["span_c"] and set(wire["span_p"]) == set(BINS)
    rounded = dict(batch)
    for key in ("promoter_feats", "pcre_feats"):
        rounded[key] = {b: t.half().float() for b, t in batch[key].items()}
    eng = InferenceEngine(model, chunk=16)
    assert torch.equal(eng.predict_wire(wire), eng.predict_device(eng.to_device(rounded)).cpu())


def test_compact_wire_ships_valid_bins_only():
    """engine.pack_wire(compact=True): pCRE features as their valid bins (chromo_unpack_compact puts them back between
    zeros) - bit-identical to the ordinary path, a quarter of the bytes; a resolution whose padded bins are NOT all zero
    keeps the full tensor."""
    model = _mk(seed=4).cuda().eval()
    model.precision = "bf16"
    n = 700
    batch = synthetic.make_batch(n, ragged=True, seed=31, stress=True)
    rounded = dict(batch)
    for key in ("promoter_feats", "pcre_feats"):
        rounded[key] = {b: t.half().float() for b, t in batch[key].items()}
    eng = InferenceEngine(model, chunk=256, device_chunk=256)
    want = eng.predict_device(eng.to_device(rounded)).cpu()
    wire = pack_wire(batch, compact=True)
    assert set(wire["xc_flat"]) == set(BINS) and not wire["xc"]
    assert wire_nbytes(wire) < 0.3 * wire_nbytes(pack_wire(batch))
    got = eng.predict_wire(wire)
    assert torch.equal(got, want)
    # a second wire with other chunk lengths through the same engine (staging is re-sized, offsets are per chunk)
    part = synthetic.slice_batch(batch, 100, 700)
    part_r = synthetic.slice_batch(rounded, 100, 700)
    assert torch.equal(eng.predict_wire(pack_wire(part, compact=True)), eng.predict_device(eng.to_device(part_r)).cpu())
    # data in a padded bin: that resolution travels as the full tensor
    batch["pcre_feats"][500][3, 2, 0, 1] = 0.5
    assert bool(batch["pcre_pad_masks"][500][3, 2, 0])
    wire2 = pack_wire(batch, compact=True)
    assert set(wire2["xc_flat"]) == {2000, 100} and set(wire2["xc"]) == {500}
    rounded["pcre_feats"][500] = batch["pcre_feats"][500].half().float()
    assert torch.equal(eng.predict_wire(wire2), eng.predict_device(eng.to_device(rounded)).cpu())


@pytest.mark.parametrize("ragged", [False, True])
def test_zero_suppressed_wire_is_lossless(ragged):
    """engine.pack_wire(sparse=True): occupancy bitmap + non-zero FP16 values (chromo_unpack_sparse) - bit-identical to the
    ordinary path, fewer bytes in proportion to the exact zeros of ln(mean + 1)."""
    model = _mk(seed=6).cuda().eval()
    model.precision = "bf16"
    n = 700
    batch = synthetic.make_batch(n, ragged=ragged, seed=41, stress=True)
    rounded = dict(batch)
    for key in ("promoter_feats", "pcre_feats"):
        rounded[key] = {b: t.half().float() for b, t in batch[key].items()}
    eng = InferenceEngine(model, chunk=256, device_chunk=256)
    want = eng.predict_device(eng.to_device(rounded)).cpu()
    wire = pack_wire(batch, sparse=True)
    assert not wire["xp"] and not wire["xc"] and set(wire["xc_bits"]) == set(BINS)
    assert wire_nbytes(wire) < 0.75 * wire_nbytes(pack_wire(batch))
    assert torch.equal(eng.predict_wire(wire), want)
    # together with the compact pCRE stream of ragged genes (promoters zero-suppressed, pCREs as valid bins)
    both = pack_wire(batch, sparse=True, compact=True)
    assert torch.equal(eng.predict_wire(both), want)
    # a chunk whose value count is not a multiple of 1024 is refused, not mis-sliced
    with pytest.raises(ValueError):
        InferenceEngine(model, chunk=100).predict_wire(wire)
