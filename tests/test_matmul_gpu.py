"""chromo_matmul (umma_train.cu): the tensor-core contraction behind the data / weight gradients of chromo_backward, in all
four operand orientations, against float64 torch on BF16-rounded operands (the kernel rounds operands to BF16 and
accumulates in FP32: what is left is summation order, <= 1e-4 relative to the output scale)."""
import ctypes

import pytest
import torch

from chromoformer_b200 import _lib

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float64)


@pytest.mark.parametrize("M,N,K,a_t,b_t,acc,ksplit", [
    (576, 128, 1024, 0, 1, 1, 4),      # dX += dProj W_att           (Regulation, bsz 64)
    (576, 256, 128, 0, 1, 0, 1),       # dAtt = gR W_o
    (1024, 128, 576, 1, 1, 1, 5),      # dW_att += dProj^T X
    (128, 256, 512, 1, 1, 1, 4),       # dW_2 += g^T F               (Pairwise)
    (256, 128, 64, 1, 1, 1, 1),        # Embedding: 64 tokens, one partial chunk
    (200, 48, 100, 0, 0, 0, 1),        # ragged M / K, narrow N, both K-major (forward orientation)
    (132, 272, 40, 1, 0, 1, 1),        # N > 256 (17 tiles of 16), A transposed only
])
def test_matmul_orientations(M, N, K, a_t, b_t, acc, ksplit):
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn((K, M) if a_t else (M, K), device="cuda", generator=g)
    B = torch.randn((K, N) if b_t else (N, K), device="cuda", generator=g)
    C0 = torch.randn(M, N, device="cuda", generator=g)
    C = C0.clone()
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.chromo_matmul(A.data_ptr(), A.stride(0), a_t, B.data_ptr(), B.stride(0), b_t, C.data_ptr(), C.stride(0), M, N, K, acc,
                           ksplit, st)
    assert rc == 0, lib.chromo_last_error()
    torch.cuda.synchronize()
    a = _bf(A).t() if a_t else _bf(A)
    b = _bf(B) if b_t else _bf(B).t()
    want = a @ b + (C0.double() if acc else 0.0)
    err = (C.double() - want).abs().max().item()
    assert err < 1e-4 * max(1.0, want.abs().max().item()), err


def test_matmul_rejects_unsupported_shapes():
    lib = _lib.load()
    A = torch.randn(64, 64, device="cuda"); C = torch.zeros(64, 10, device="cuda")
    B = torch.randn(10, 64, device="cuda")
    assert lib.chromo_matmul(A.data_ptr(), 64, 0, B.data_ptr(), 64, 0, C.data_ptr(), 10, 64, 10, 64, 0, 1, None) < 0
    assert b"not supported" in lib.chromo_last_error()
