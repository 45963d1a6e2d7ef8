"""Hardware-semantics probes of tcgen05 operand forms (MN-major B in shared memory, A in TMEM).  The probe kernel is test
infrastructure: tests/csrc/umma_probe.cu -> tests/libchromo_probe.so (built by __graft_entry__.build()), not part of the
product library."""
import ctypes
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
_PROBE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libchromo_probe.so")


def _probe():
    lib = ctypes.CDLL(_PROBE)
    lib.chromo_debug_umma_probe.restype = ctypes.c_int32
    lib.chromo_debug_umma_probe.argtypes = [ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                            ctypes.c_int32, ctypes.c_void_p]
    return lib


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 5, 7])
@pytest.mark.parametrize("n,k", [(32, 128), (128, 32), (64, 64)])
def test_umma_operand_forms(mode, n, k):
    """mode bit 0: B MN-major in shared memory; bit 1: A in TMEM (tcgen05.st); bit 2: MN-major B with
    k-blocks N*16 B apart (the K-major bytes of B^T re-read as MN-major).  D = bf16(A) @ bf16(B)."""
    lib = _probe()
    g = torch.Generator().manual_seed(n * 1000 + k)
    A = torch.randn(128, k, generator=g).cuda()
    B = torch.randn(k, n, generator=g).cuda()
    D = torch.full((128, n), float("nan"), device="cuda")
    assert lib.chromo_debug_umma_probe(mode, A.data_ptr(), B.data_ptr(), D.data_ptr(), n, k,
                                       torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    want = A.bfloat16().double() @ B.bfloat16().double()
    err = (D.double() - want).abs().max().item()
    assert err < 1e-3 * k ** 0.5, (mode, err)
