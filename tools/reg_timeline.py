"""Timeline of one CTA of the fused Regulation kernel (chromo_debug_trace): cycles between the hand-offs of one layer.

    python tools/reg_timeline.py [layer=2] [genes=4096]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from chromoformer_b200 import ChromoformerClassifier, _lib, synthetic  # noqa: E402

NAMES = {1: "D x ready", 2: "D [q|k] issued", 3: "D [v|g] issued", 4: "D qkfree seen", 5: "D vgfree seen", 6: "D att ready",
         7: "D out-proj issued", 8: "D u ready", 9: "D ffn1 issued", 10: "D f ready A", 11: "D f ready B", 12: "D ffn2 issued",
         13: "D   chunk in smem", 14: "D   mmas queued", 15: "D   prev chunk done, refill issued",
         20: "S accq seen", 21: "S q,k staged", 22: "S group barrier", 23: "S ofree seen", 24: "S sr seen", 25: "S P stored",
         26: "S vr seen (O issue)", 30: "V accvg seen", 31: "V v stored", 32: "V or seen", 33: "V att stored",
         50: "E   acc loaded", 51: "E   partial sums stored", 52: "E   quarter barrier", 53: "E   U packed + stored",
         40: "E attention done", 41: "E barrier", 42: "E prm loaded", 43: "E acco seen", 44: "E u ready", 45: "E accf1 seen",
         46: "E f ready", 47: "E accf2 seen", 48: "E x ready"}


def main():
    layer = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    lib = _lib.load()
    dev = torch.device("cuda:0")
    model = ChromoformerClassifier(seed=123).cuda().eval()
    cfg = _lib.Config.from_buffer_copy(model._cfg)
    T = B * 9
    x = torch.randn(3, T, 128, device=dev)
    y = torch.empty_like(x)
    flags = _lib.F_BF16
    nws = _lib.check(lib.chromo_workspace_floats(ctypes.byref(cfg), B, flags), "ws")
    ws = torch.empty(nws, device=dev)
    batch = synthetic.make_batch(B, ragged=False, seed=0)
    im = [batch["interaction_masks"][b].to(dev).contiguous() for b in (2000, 500, 100)]
    imp = (ctypes.c_void_p * 3)(*[m.data_ptr() for m in im])
    fq = batch["interaction_freq"].to(dev).contiguous()
    st = torch.cuda.current_stream().cuda_stream

    def run(fl):
        _lib.check(lib.chromo_regulation_layer(ctypes.byref(cfg), model.flat_params.data_ptr(), -1, x.data_ptr(), y.data_ptr(),
                                               T * 128, imp, fq.data_ptr(), B, ws.data_ptr(), nws, fl, st), "reg")
    run(flags)
    run(flags | _lib.F_PACKED)
    only = set(int(x) for x in sys.argv[3].split(',')) if len(sys.argv) > 3 else None
    buf = torch.zeros(1536, dtype=torch.int64, device=dev)
    lib.chromo_debug_trace(ctypes.c_void_p(buf.data_ptr()))
    run(flags | _lib.F_PACKED)
    torch.cuda.synchronize()
    lib.chromo_debug_trace(None)
    ev = []
    for v in buf.cpu().tolist():
        if v:
            ev.append((v & 0xFFFFFFFFFFFF, v >> 48))
    ev.sort()
    t0 = ev[0][0]
    # layer boundaries = "D x ready" events
    starts = [t for t, e in ev if e == 1]
    print(f"{len(ev)} events, {len(starts)} layers, total {ev[-1][0] - t0} cycles; per layer:",
          [starts[i + 1] - starts[i] for i in range(len(starts) - 1)])
    lo = starts[layer]
    hi = starts[layer + 1] if layer + 1 < len(starts) else ev[-1][0] + 1
    prev = lo
    for t, e in ev:
        if lo <= t < hi and (only is None or e in only):
            print(f"{t - lo:7d}  (+{t - prev:5d})  {NAMES.get(e, e)}")
            prev = t


if __name__ == "__main__":
    main()
