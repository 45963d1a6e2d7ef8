"""ms per fused training step (TrainStep, Chromoformer-reg, bsz 64 dense) - the `train` leg of bench.py on its own.
Usage: python tools/train_time.py [bf16|fp32] [graph|eager]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from chromoformer_b200 import ChromoformerRegressor, _lib, synthetic  # noqa: E402
from chromoformer_b200.trainer import TrainStep  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
graph = (sys.argv[2] if len(sys.argv) > 2 else "graph") == "graph"
from chromoformer_b200.parallel import init_distributed  # noqa: E402

rank, local, world = init_distributed()
if world > 1:
    torch.cuda.set_device(local)
lib = _lib.load()
reg = ChromoformerRegressor(seed=123).cuda().train()
reg.precision = prec
tb = synthetic.make_batch(64, ragged=False, seed=100)
dev = {k: ({b: t.cuda() for b, t in v.items()} if isinstance(v, dict) else v.cuda()) for k, v in tb.items()}
target = dev["labels_reg"].view(-1, 1)
step = TrainStep(reg, lr=3e-5, regression=True, use_graph=graph)
lib.chromo_launch_counter(1)
step(dev, target)
launches = int(lib.chromo_launch_counter(1))
for _ in range(6):
    step(dev, target)
bufs = step.input_buffers(dev, target)
if bufs is not None and os.environ.get("TRAIN_TIME_COPY") is None:
    dev, target = bufs
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    step(dev, target)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
if world > 1:
    import torch.distributed as dist
    t = torch.tensor([ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
if rank == 0:
    print(f"world {world} overlap {step.overlap} {prec} {'graph' if graph else 'eager'}: {ms:.3f} ms/step = {world * 64 / ms * 1e3:.0f} samples/s, {launches} launches, loss {step.loss.item():.4f}")
