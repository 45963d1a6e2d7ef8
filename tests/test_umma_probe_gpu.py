"""Hardware-semantics probes for tcgen05 operand forms not yet used by the product kernels."""
import pytest
import torch

from chromoformer_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 5, 7])
@pytest.mark.parametrize("n,k", [(32, 128), (128, 32), (64, 64)])
def test_umma_operand_forms(mode, n, k):
    """mode bit 0: B MN-major in shared memory; bit 1: A in TMEM (tcgen05.st); bit 2: MN-major B with
    k-blocks N*16 B apart (the K-major bytes of B^T re-read as MN-major).  D = bf16(A) @ bf16(B)."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(n * 1000 + k)
    A = torch.randn(128, k, generator=g).cuda()
    B = torch.randn(k, n, generator=g).cuda()
    D = torch.full((128, n), float("nan"), device="cuda")
    _lib.check(lib.chromo_debug_umma_probe(mode, A.data_ptr(), B.data_ptr(), D.data_ptr(), n, k,
                                           torch.cuda.current_stream().cuda_stream), "probe")
    torch.cuda.synchronize()
    want = A.bfloat16().double() @ B.bfloat16().double()
    err = (D.double() - want).abs().max().item()
    assert err < 1e-3 * k ** 0.5, (mode, err)
