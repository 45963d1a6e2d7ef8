"""One Regulation-transformer layer through chromo_regulation_layer (FP32 kernels and the fused
tcgen05 kernel) against the oracle's AttentionBlock restatement (modules.py:104-111, gate=True)."""
import ctypes

import pytest
import torch

from _util import KWS
from chromoformer_b200 import ChromoformerClassifier, _lib
from oracle import chromoformer_oracle as oracle

pytestmark = pytest.mark.gpu


def _layer(model, layer, x, imasks, freq, flags):
    lib = _lib.load()
    cfg = _lib.Config.from_buffer_copy(model._cfg)
    S = freq.shape[1]
    cfg.i_max = S - 1
    B = freq.shape[0]
    y = torch.empty_like(x)
    n = _lib.check(lib.chromo_workspace_floats(ctypes.byref(cfg), B, flags), "ws")
    ws = torch.empty(n, device="cuda")
    ptrs = (ctypes.c_void_p * 3)(*[m.data_ptr() for m in imasks])
    _lib.check(lib.chromo_regulation_layer(ctypes.byref(cfg), model.flat_params.data_ptr(), layer, x.data_ptr(),
                                           y.data_ptr(), x.shape[1] * x.shape[2], ptrs, freq.data_ptr(), B,
                                           ws.data_ptr(), n, flags, torch.cuda.current_stream().cuda_stream),
               "chromo_regulation_layer")
    torch.cuda.synchronize()
    return y


@pytest.mark.parametrize("B,S,layer", [(5, 9, 0), (300, 9, 3), (40, 17, 5)])
def test_regulation_layer_vs_oracle(B, S, layer):
    model = ChromoformerClassifier(7, 128, 128, dict(KWS[0]), dict(KWS[1]), dict(KWS[2]), seed=77)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(B)
    x = torch.randn(3, B * S, 128, generator=g)
    freq = torch.zeros(B, S, S); freq[:, 0, 1:] = 1.5 + 1.5 * torch.rand(B, S - 1, generator=g)
    k = torch.randint(0, S, (B,), generator=g)
    idx = torch.arange(S)
    inside = (idx.view(1, S, 1) <= k.view(B, 1, 1)) & (idx.view(1, 1, S) <= k.view(B, 1, 1))
    masks = [(~inside).clone() for _ in range(3)]
    masks[1][0] = True                                   # one fully masked gene
    want = []
    for r, b in enumerate((2000, 500, 100)):
        pre = f"regulation.{b}.transformer.layers.{layer}."
        h = oracle.self_attention_block(x[r].view(B, S, 128), masks[r].unsqueeze(1), freq, sd, pre + "self_att.", 8, True)
        want.append(oracle.feed_forward_block(h, sd, pre + "ff.").reshape(B * S, 128))
    want = torch.stack(want)
    model.cuda()
    xd, fd, md = x.cuda(), freq.cuda(), [m.cuda().contiguous() for m in masks]
    got32 = _layer(model, layer, xd, md, fd, 0).cpu()
    assert (got32 - want).abs().max().item() < 2e-5
    got16 = _layer(model, layer, xd, md, fd, _lib.F_BF16).cpu()
    assert not torch.equal(got16, got32)
    # BF16 operands on unit-variance activations: a few 1e-2 absolute after two LayerNorms
    assert (got16 - want).abs().max().item() < 6e-2
    assert (got16 - want).abs().mean().item() < 5e-3


@pytest.mark.parametrize("B,S", [(33, 9), (150, 9), (20, 17)])
def test_all_layers_in_one_launch(B, S):
    """layer = -1: the six layers in ONE launch of the fused kernel (operand tile resident in shared memory, residual
    rows parked in scratch between layers) against the FP32 kernels layer by layer and against six single-layer
    launches of the same fused kernel."""
    model = ChromoformerClassifier(7, 128, 128, dict(KWS[0]), dict(KWS[1]), dict(KWS[2]), seed=78).cuda()
    g = torch.Generator().manual_seed(B + S)
    x = torch.randn(3, B * S, 128, generator=g).cuda()
    freq = torch.zeros(B, S, S); freq[:, 0, 1:] = 1.5 + 1.5 * torch.rand(B, S - 1, generator=g)
    k = torch.randint(0, S, (B,), generator=g)
    idx = torch.arange(S)
    inside = (idx.view(1, S, 1) <= k.view(B, 1, 1)) & (idx.view(1, 1, S) <= k.view(B, 1, 1))
    md = [(~inside).clone().cuda().contiguous() for _ in range(3)]
    fd = freq.cuda()
    want, step = x, x
    for layer in range(6):
        want = _layer(model, layer, want, md, fd, 0)
        step = _layer(model, layer, step, md, fd, _lib.F_BF16)
    got = _layer(model, -1, x, md, fd, _lib.F_BF16)
    assert torch.isfinite(got).all()
    assert (got - step).abs().max().item() < 1e-5          # same arithmetic, layer by layer or in one launch
    assert (got - want).abs().max().item() < 1.5e-1 and (got - want).abs().mean().item() < 1e-2
    # FP32 kernels have no all-layer form
    lib = _lib.load()
    cfg = _lib.Config.from_buffer_copy(model._cfg); cfg.i_max = S - 1
    ptrs = (ctypes.c_void_p * 3)(*[m.data_ptr() for m in md])
    ws = torch.empty(1 << 20, device="cuda")
    rc = lib.chromo_regulation_layer(ctypes.byref(cfg), model.flat_params.data_ptr(), -1, x.data_ptr(), x.data_ptr(),
                                     x.shape[1] * 128, ptrs, fd.data_ptr(), B, ws.data_ptr(), 1 << 40, 0, None)
    assert rc != 0
