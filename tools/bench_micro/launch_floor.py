"""Per-launch cost of the training GEMM inside a CUDA graph: a chain of dependent chromo_matmul calls (each reads the
previous one's output), alone and interleaved with a trivial elementwise kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from chromoformer_b200 import _lib  # noqa: E402

lib = _lib.load()


def chain(M, N, K, n, interleave, b_t=0):
    A = [torch.randn(M, K, device="cuda") * 0.05 for _ in range(2)]
    B = torch.randn((K, N) if b_t else (N, K), device="cuda") * 0.05
    assert N == K

    def run():
        st = torch.cuda.current_stream().cuda_stream
        for i in range(n):
            src, dst = A[i & 1], A[(i + 1) & 1]
            rc = lib.chromo_matmul(src.data_ptr(), K, 0, B.data_ptr(), B.stride(0), b_t, dst.data_ptr(), N, M, N, K, 0, 1, st)
            assert rc == 0, lib.chromo_last_error()
            if interleave:
                dst.mul_(1.0)
    run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 / n * 1e3


for (M, N, K) in [(128, 128, 128), (576, 128, 128), (1728, 128, 128), (1728, 256, 256)]:
    for inter in (False, True):
        print(f"M={M} N={N} K={K} interleave={inter}: {chain(M, N, K, 100, inter):.2f} us per GEMM (+ elementwise)" )
# trivial kernel alone
x = torch.randn(1 << 16, device="cuda")
g = torch.cuda.CUDAGraph()
x.mul_(1.0)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    for _ in range(100):
        x.mul_(1.0)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.replay()
e1.record(); torch.cuda.synchronize()
print(f"elementwise alone: {e0.elapsed_time(e1) / 20 / 100 * 1e3:.2f} us per launch")
