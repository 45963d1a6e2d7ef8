"""chromo_linear (the dense-projection kernel of every nn.Linear call site) in isolation:
strict-FP32 CUDA-core engine and BF16 tcgen05 engine against a torch FP32 reference."""
import pytest
import torch

from chromoformer_b200 import _lib

pytestmark = pytest.mark.gpu


def _linear(x, w, b, relu, flags):
    lib = _lib.load()
    z, m, k = x.shape
    n = w.shape[1]
    y = torch.empty(z, m, n, device="cuda")
    wp = w
    if flags & _lib.F_BF16:
        wp = torch.empty(z, n, k, dtype=torch.bfloat16, device="cuda")
        _lib.check(lib.chromo_pack_linear_weight(w.data_ptr(), wp.data_ptr(), n, k, z, n * k,
                                                 torch.cuda.current_stream().cuda_stream), "pack")
    _lib.check(lib.chromo_linear(x.data_ptr(), wp.data_ptr(), None if b is None else b.data_ptr(), y.data_ptr(),
                                 m, n, k, 1 if relu else 0, z, m * k, n * k, n, m * n, flags,
                                 torch.cuda.current_stream().cuda_stream), "chromo_linear")
    torch.cuda.synchronize()
    return y


def _ref(x, w, b, relu, bf16):
    if bf16:
        x, w = x.bfloat16().float(), w.bfloat16().float()
    y = torch.einsum("zmk,znk->zmn", x.double(), w.double())
    if b is not None:
        y = y + b.double().unsqueeze(1)
    return (torch.relu(y) if relu else y).float()


@pytest.mark.parametrize("m,n,k,z,relu", [(1000, 1024, 128, 3, False), (37, 128, 256, 2, True), (513, 2, 128, 1, False),
                                          (64, 256, 128, 3, True), (4096, 128, 384, 1, True)])
def test_linear_fp32(m, n, k, z, relu):
    g = torch.Generator().manual_seed(m + n)
    x = torch.randn(z, m, k, generator=g).cuda(); w = (torch.randn(z, n, k, generator=g) * 0.1).cuda()
    b = torch.randn(z, n, generator=g).cuda()
    y = _linear(x, w, b, relu, 0)
    assert (y - _ref(x, w, b, relu, False)).abs().max().item() < 2e-5 * k ** 0.5


@pytest.mark.parametrize("m,n,k,z,relu", [(1000, 1024, 128, 3, False), (128, 128, 128, 1, False), (37, 128, 256, 2, True),
                                          (4096, 256, 128, 3, True), (300, 80, 128, 1, False), (4096, 128, 384, 1, True),
                                          (129, 128, 400, 1, False), (256, 400, 128, 1, False)])
def test_linear_bf16_tcgen05(m, n, k, z, relu):
    g = torch.Generator().manual_seed(m + n)
    x = torch.randn(z, m, k, generator=g).cuda(); w = (torch.randn(z, n, k, generator=g) * 0.1).cuda()
    b = torch.randn(z, n, generator=g).cuda()
    y = _linear(x, w, b, relu, _lib.F_BF16)
    # exact BF16 products, FP32 accumulation: only the summation order differs from the reference
    err = (y - _ref(x, w, b, relu, True)).abs().max().item()
    assert err < 3e-5 * k ** 0.5, err
    # and it is a BF16-accurate version of the FP32 product
    assert (y - _ref(x, w, b, relu, False)).abs().max().item() < 0.05 * (k / 128) ** 0.5
