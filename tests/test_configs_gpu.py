"""Non-default shapes through every code path: other i_max, other window / bin counts (kernels fall back to the
generic variants), other feature counts; forward in both precisions and gradients, against the oracle."""
import pytest
import torch

from _util import KWS
from chromoformer_b200 import ChromoformerClassifier, ChromoformerRegressor, _lib, synthetic
from oracle import chromoformer_oracle as oracle

pytestmark = pytest.mark.gpu


def _sd(model):
    return {k: v.detach().cpu().clone() for k, v in model.named_parameters()}


@pytest.mark.parametrize("i_max,w_max,n_feats", [(4, 40000, 7), (8, 20000, 7), (8, 40000, 5), (3, 12000, 6)])
def test_other_shapes_forward_and_gradients(i_max, w_max, n_feats):
    model = ChromoformerRegressor(n_feats, 128, 128, dict(KWS[0]), dict(KWS[1]), dict(KWS[2]), seed=31)
    sd = _sd(model)
    batch = synthetic.make_batch(70, i_max=i_max, w_max=w_max, n_feats=n_feats, ragged=True, full_masks=True, seed=41)
    target = batch["labels_reg"].view(-1, 1)
    loss_o, logits_o, grads_o = oracle.forward_backward(sd, synthetic.forward_args(batch), target, True)
    model.cuda()
    args = synthetic.forward_args(batch, "cuda")
    model.eval()
    with torch.no_grad():
        model.precision = "fp32"
        y32 = model(*args).cpu()
        model.precision = "bf16"
        y16 = model(*args).cpu()
    assert (y32 - logits_o).abs().max().item() < 5e-5
    assert (y16 - logits_o).abs().max().item() < 1e-2
    model.precision = "fp32"
    model.train()
    out = model(*args)
    loss = torch.nn.functional.mse_loss(out, target.cuda())
    loss.backward()
    assert abs(loss.item() - loss_o.item()) < 1e-5
    # 5e-4 of each tensor's largest gradient.  FFN-1 weights / biases get 1e-2: a hidden unit whose pre-activation lies within
    # FP32 rounding distance of zero takes either branch of the ReLU (ours and the oracle's sums are ordered differently), and
    # one flipped (token, unit) moves that unit's row of dW_1 by a few 1e-3 of the tensor's maximum.
    for name, p in model.named_parameters():
        g = grads_o[name]
        if g is None:
            assert p.grad is None
            continue
        err = (p.grad.cpu() - g).abs().max().item() / max(g.abs().max().item(), 1e-7)
        assert err < (1e-2 if ".ff.l1." in name else 5e-4), (name, err)


def test_input_validation_errors():
    model = ChromoformerClassifier(seed=1).cuda().eval()
    batch = synthetic.make_batch(4, seed=0)
    args = synthetic.forward_args(batch, "cuda")
    with pytest.raises(ValueError):                       # CPU tensor handed to a CUDA model
        bad = list(args); bad[5] = bad[5].cpu()
        model(*bad)
    with pytest.raises(ValueError):                       # wrong feature count
        bad = list(args); bad[0] = {b: t[..., :6].contiguous() for b, t in args[0].items()}
        model(*bad)
    with pytest.raises(ValueError):                       # interaction mask of another i_max
        bad = list(args); bad[4] = {b: t[:, :, :8, :8].contiguous() for b, t in args[4].items()}
        model(*bad)
    two_layer = dict(KWS[0]); two_layer["n_layers"] = 2   # embed depth the pruned path does not cover
    with pytest.raises(_lib.ChromoLibError):
        ChromoformerClassifier(7, 128, 128, two_layer, dict(KWS[1]), dict(KWS[2]))
