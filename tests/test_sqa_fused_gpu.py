"""The fused single-query attention kernel (sqa_fused.cu) alone, through chromo_single_query_attention.

Reference: float64 torch with the operands the tensor pipe sees in BF16 (qk, the position table and W_in in the score
term) rounded the same way; what remains is the BF16 rounding of the probabilities (2^-9 relative) and FP32
accumulation order."""
import math

import pytest
import torch

from chromoformer_b200 import _lib
from chromoformer_b200.model import sinusoid_table

pytestmark = pytest.mark.gpu


def _reference(qk, x, mask, w_in, pe, scale):
    """modules.py:16-30 for one query: keys = values = W_in x_j + PE_j."""
    regions, n, _ = x.shape
    qk_b, pe_b = qk.bfloat16().double(), pe.bfloat16().double()
    u = qk_b @ w_in.bfloat16().double()                                     # [rows, 7]   (a tensor-core GEMM in the kernel)
    s = qk_b @ pe_b.t() + torch.einsum("rf,rjf->rj", u, x.double().repeat_interleave(2, 0))
    s = s * scale
    s = s.masked_fill(mask.bool().repeat_interleave(2, 0), -1e9)
    p = torch.softmax(s, dim=1)
    xbar = torch.einsum("rj,rjf->rf", p, x.double().repeat_interleave(2, 0))
    return p @ pe_b + xbar @ w_in.double().t()


def _run(qk, x, mask, w_in, pe, scale, tau=None, monkeypatch=None):
    lib = _lib.load()
    regions, n, _ = x.shape
    ns = 32 if n <= 32 else (n + 15) // 16 * 16
    ws = torch.empty(64 * ns + 1024 + 8192 * ((regions + 63) // 64), device="cuda")
    out = torch.full((regions * 2, 128), float("nan"), device="cuda")
    dev = [t.cuda().contiguous() for t in (qk, x, mask.to(torch.uint8), w_in, pe)]
    if tau is not None:
        monkeypatch.setenv("CHROMO_SQA_TAU", str(tau))
    _lib.check(lib.chromo_single_query_attention(regions, n, *[t.data_ptr() for t in dev], scale, out.data_ptr(),
                                                 ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream),
               "chromo_single_query_attention")
    torch.cuda.synchronize()
    return out.cpu().double()


def _inputs(regions, n, seed, qscale=1.0):
    g = torch.Generator().manual_seed(seed)
    qk = torch.randn(regions * 2, 128, generator=g) * qscale
    x = torch.rand(regions, n, 7, generator=g) * (torch.rand(regions, n, 7, generator=g) < 0.6)
    mask = torch.rand(regions, n, generator=g) < 0.3
    w_in = torch.randn(128, 7, generator=g) * 0.3
    pe = sinusoid_table(n, 128)
    return qk, x, mask, w_in, pe


@pytest.mark.parametrize("regions,n", [(1, 20), (64, 20), (65, 80), (200, 400), (37, 400), (300, 80), (9, 48), (5, 16),
                                       (3, 4), (70, 384)])
def test_random_rows(regions, n):
    qk, x, mask, w_in, pe = _inputs(regions, n, seed=regions * 1000 + n)
    got = _run(qk, x, mask, w_in, pe, 1.0 / math.sqrt(64))
    want = _reference(qk, x, mask, w_in, pe, 1.0 / math.sqrt(64))
    assert torch.isfinite(got).all()
    assert (got - want).abs().max().item() < 6e-3


def test_masks_all_or_half():
    """Fully masked regions attend uniformly (softmax of equal -1e9 scores, as in the reference); a region whose first
    or second key half is fully masked takes everything from the other half."""
    regions, n = 6, 400
    qk, x, mask, w_in, pe = _inputs(regions, n, seed=5)
    mask[:] = False
    mask[0] = True
    mask[1, :208] = True
    mask[2, 208:] = True
    mask[3, :399] = True
    mask[4, 1:] = True
    got = _run(qk, x, mask, w_in, pe, 0.125)
    want = _reference(qk, x, mask, w_in, pe, 0.125)
    assert (got - want).abs().max().item() < 6e-3


@pytest.mark.parametrize("n", [80, 400])
def test_rising_scores_move_the_softmax_reference(n):
    """Scores that climb by hundreds along the keys: the running reference maximum has to move (and the P chunks
    already in tensor memory be rescaled) several times per row."""
    regions = 40
    g = torch.Generator().manual_seed(n)
    qk = torch.randn(regions * 2, 128, generator=g) * 0.1
    w_in = torch.zeros(128, 7)
    w_in[:, 0] = qk.mean(0) / qk.mean(0).norm() ** 2 * 3.0            # u[:, 0] ~ 3 for the mean row
    x = torch.rand(regions, n, 7, generator=g)
    ramp = torch.linspace(0.0, 60.0, n).view(1, n) * (1.0 + torch.rand(regions, 1, generator=g))
    x[:, :, 0] = ramp * torch.where(torch.rand(regions, 1, generator=g) < 0.5, 1.0, -1.0)   # rising or falling
    mask = torch.rand(regions, n, generator=g) < 0.2
    pe = sinusoid_table(n, 128)
    got = _run(qk, x, mask, w_in, pe, 1.0)
    want = _reference(qk, x, mask, w_in, pe, 1.0)
    u0 = (qk.double() @ w_in.double())[:, 0]
    assert (u0.abs() * 60 * 1.4427 > 64).any()                         # the scores really span > 2^64
    assert torch.isfinite(got).all()
    assert (got - want).abs().max().item() < 2e-2 * max(1.0, want.abs().max().item())


def test_eager_rescaling_matches(monkeypatch):
    """CHROMO_SQA_TAU=0 moves the reference on every new maximum; same result up to the repeated BF16 rounding of P."""
    qk, x, mask, w_in, pe = _inputs(130, 400, seed=77, qscale=2.0)
    lazy = _run(qk, x, mask, w_in, pe, 0.125)
    eager = _run(qk, x, mask, w_in, pe, 0.125, tau=0, monkeypatch=monkeypatch)
    want = _reference(qk, x, mask, w_in, pe, 0.125)
    assert not torch.equal(lazy, eager)
    assert (lazy - want).abs().max().item() < 6e-3
    assert (eager - want).abs().max().item() < 2e-2


def test_rejects_unsupported_shapes():
    lib = _lib.load()
    t = torch.zeros(1024, device="cuda")
    for regions, n in ((4, 402), (4, 404), (0, 80)):
        rc = lib.chromo_single_query_attention(regions, n, t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(),
                                               t.data_ptr(), 1.0, t.data_ptr(), t.data_ptr(), 10 ** 6, None)
        assert rc != 0 and lib.chromo_last_error()
