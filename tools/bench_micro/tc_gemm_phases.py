"""Phase clocks of CTA 0 of the training GEMM (chromo_debug_trace): where the ~5 us of a one-tile launch go."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from chromoformer_b200 import _lib  # noqa: E402

lib = _lib.load()
buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
names = ["start", "tmem alloc + barrier init + sync", "A staged (chunk 0)", "B staged (chunk 0)", "sync", "all chunks issued",
         "accumulator ready", "epilogue done", "dealloc"]
for (M, N, K, b_t) in [(128, 128, 128, 0), (1728, 128, 128, 0), (1728, 256, 256, 1), (576, 128, 1024, 1)]:
    A = torch.randn(M, K, device="cuda") * 0.05
    B = torch.randn((K, N) if b_t else (N, K), device="cuda") * 0.05
    C = torch.empty(M, N, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(3):
        lib.chromo_debug_trace(buf.data_ptr() if rep == 2 else None)
        lib.chromo_matmul(A.data_ptr(), K, 0, B.data_ptr(), B.stride(0), b_t, C.data_ptr(), N, M, N, K, 0, 1, st)
        torch.cuda.synchronize()
    lib.chromo_debug_trace(None)
    t = buf[2048:2057].cpu().tolist()
    print(f"M={M} N={N} K={K} b_t={b_t}: total {t[8] - t[0]} cycles")
    for i in range(1, 9):
        print(f"   {t[i] - t[0]:7d} (+{t[i] - t[i - 1]:6d})  {names[i]}")
