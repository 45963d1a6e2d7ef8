"""Time the fused Regulation kernel alone (as bench.py's `roofline` does): ms per launch, TFLOP/s, fraction of the measured peak."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chromoformer_b200 import ChromoformerClassifier, _lib, synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = _lib.load()
dev = torch.device("cuda:0")
model = ChromoformerClassifier(seed=123).cuda().eval()
cfg = _lib.Config.from_buffer_copy(model._cfg)
T = B * 9
x = torch.randn(3, T, 128, device=dev); y = torch.empty_like(x)
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
for _ in range(5): run(flags | _lib.F_PACKED)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run(flags | _lib.F_PACKED)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
flops = 6 * 3.0 * T * 2 * (4 * 256 * 128 + 256 * 128 + 2 * 128 * 256 + 8 * 9 * 32 * 2)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"]
print(f"reg_layer_fused: {ms:.4f} ms/launch, {flops / ms / 1e9:.1f} TFLOP/s, frac {flops / ms / 1e9 / peak:.3f} of {peak}")
