"""CPU oracle for the Chromoformer hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A self-contained restatement, in plain FP32 PyTorch/numpy on the CPU, of the
algorithm the reference (dohlee/chromoformer) executes on the path this repo
accelerates.  It restates the reference AS WRITTEN (full n x n attention, all
query rows, all pCRE slots) — none of the pruning / re-association the CUDA
kernels use — so agreement between the two is a statement about the kernels.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The package
``chromoformer_b200`` never does.

Pinning (see ``tests/test_oracle_pinning.py`` and ``tests/golden/make_golden.py``):
  * ``net.py:558-568`` known answers  -3.1917 / -3.1917 / -0.1900,
  * ``demo/random_prediction.out`` (100 demo genes, untrained seed-123 weights),
  * outputs of the reference itself imported from /root/reference (logits,
    gradients, post-AdamW parameters, dataset items) committed under
    ``tests/golden/``.

Functions take a ``state_dict`` in the reference's dict layout.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

NEG = -1e9          # modules.py:72,184 (FP32 branch)
LN_EPS = 1e-5       # nn.LayerNorm default, modules.py:25,96,146


# ----------------------------------------------------------------------------- blocks
def position_table(n, d):
    """net.py:23-29 / net.py:97-103."""
    pe = torch.zeros(n, d)
    pos = torch.arange(0, n, 1).unsqueeze(1)
    k = torch.exp(-np.log(10000) * torch.arange(0, d, 2) / d)
    pe[:, 0::2] = torch.sin(pos * k)
    pe[:, 1::2] = torch.cos(pos * k)
    return pe


def _layer_norm(x, sd, pre):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + "weight"], sd[pre + "bias"], LN_EPS)


def _heads(t, n_heads):
    b, n, w = t.shape
    return t.reshape(b, n, n_heads, w // n_heads).transpose(1, 2)     # [b, h, n, dh]


def _attend(q, k, v, mask, bias, gamma_f):
    """modules.py:58-77 / 170-189: scale, then gamma*bias, then masked_fill(-1e9), softmax, PV."""
    dh = q.shape[-1]
    s = torch.matmul(q, k.transpose(-1, -2)) / (dh ** 0.5)
    if bias is not None:
        s = s + gamma_f.view(1, -1, 1, 1) * bias.unsqueeze(1)
    if mask is not None:
        s = s.masked_fill(mask, NEG)
    return torch.matmul(torch.softmax(s, dim=-1), v)


def self_attention_block(x, mask, bias, sd, pre, n_heads, gate):
    """MultiHeadAttention.forward, modules.py:28-88 (fused projection chunk order q,k,v,gate)."""
    proj = x @ sd[pre + "att.weight"].t()
    parts = torch.chunk(proj, 4 if gate else 3, dim=-1)
    q, k, v = (_heads(t, n_heads) for t in parts[:3])
    a = _attend(q, k, v, mask, bias, sd[pre + "gamma_f"])
    if gate:
        a = a * torch.sigmoid(_heads(parts[3], n_heads))
    a = a.transpose(1, 2).reshape(x.shape[0], x.shape[1], -1)
    return _layer_norm(x + a @ sd[pre + "ff.weight"].t() + sd[pre + "ff.bias"], sd, pre + "ln.")


def cross_attention_block(x_p, x_c, mask, sd, pre, n_heads):
    """PairwiseMultiHeadAttention.forward, modules.py:150-195 (bias is always None on this path)."""
    q = _heads(x_p @ sd[pre + "p_att.weight"].t(), n_heads)
    k, v = (_heads(t, n_heads) for t in torch.chunk(x_c @ sd[pre + "c_att.weight"].t(), 2, dim=-1))
    a = _attend(q, k, v, mask, None, None)
    a = a.transpose(1, 2).reshape(x_p.shape[0], x_p.shape[1], -1)
    return _layer_norm(x_p + a @ sd[pre + "ff.weight"].t() + sd[pre + "ff.bias"], sd, pre + "ln.")


def feed_forward_block(x, sd, pre):
    """FeedForward.forward, modules.py:100-101."""
    h = torch.relu(x @ sd[pre + "l1.weight"].t() + sd[pre + "l1.bias"])
    return _layer_norm(x + h @ sd[pre + "l2.weight"].t() + sd[pre + "l2.bias"], sd, pre + "ln.")


def _count_layers(sd, pre):
    n = 0
    while f"{pre}transformer.layers.{n}.ff.l1.weight" in sd:
        n += 1
    return n


# ------------------------------------------------------------------------ sub-models
def embedding_transformer(x, mask, sd, pre, n_heads):
    """EmbeddingTransformer.forward, net.py:31-59.  x [B,1,n,F], mask [B,1,1,n,n]."""
    bsz, n = mask.shape[0], mask.shape[4]
    m = mask.reshape(-1, 1, n, n)
    h = x.reshape(-1, n, x.shape[3]) @ sd[pre + "lin_proj.weight"].t()
    h = h + position_table(n, h.shape[-1]).unsqueeze(0)
    for l in range(_count_layers(sd, pre)):
        lp = f"{pre}transformer.layers.{l}."
        h = self_attention_block(h, m, None, sd, lp + "self_att.", n_heads, gate=False)
        h = feed_forward_block(h, sd, lp + "ff.")
    h = h.reshape(bsz, -1, n, h.shape[-1])
    return h, h[:, :, n // 2]


def pairwise_interaction_transformer(x_p_full, x_pcre, mask, sd, pre, n_heads):
    """PairwiseInteractionTransformer.forward, net.py:105-139."""
    bsz, n = mask.shape[0], mask.shape[4]
    slots = x_pcre.shape[1]
    m = mask.reshape(-1, 1, n, n)
    p = x_p_full.expand(-1, slots, -1, -1).reshape(-1, n, x_p_full.shape[3])
    p = p @ sd[pre + "lin_proj_p.weight"].t()
    c = x_pcre.reshape(-1, n, x_pcre.shape[3]) @ sd[pre + "lin_proj_pcre.weight"].t()
    c = c + position_table(n, c.shape[-1]).unsqueeze(0)
    for l in range(_count_layers(sd, pre)):
        lp = f"{pre}transformer.layers.{l}."
        p = cross_attention_block(p, c, m, sd, lp + "self_att.", n_heads)
        p = feed_forward_block(p, sd, lp + "ff.")
    return p.reshape(bsz, slots, n, -1)[:, :, n // 2]


def regulation_transformer(x, mask, bias, sd, pre, n_heads):
    """RegulationTransformer.forward, net.py:152-153 (gated, frequency-biased self-attention)."""
    for l in range(_count_layers(sd, pre)):
        lp = f"{pre}transformer.layers.{l}."
        x = self_attention_block(x, mask, bias, sd, lp + "self_att.", n_heads, gate=True)
        x = feed_forward_block(x, sd, lp + "ff.")
    return x


DEFAULT_HEADS = {"embed": 2, "pairwise_interaction": 2, "regulation": 8}


def chromoformer_forward(sd, promoter_feats, promoter_pad_masks, pcre_feats, pcre_pad_masks,
                         interaction_masks, interaction_freq, binsizes=(2000, 500, 100), heads=None):
    """ChromoformerBase.forward, net.py:332-380.  Inputs are dicts keyed by bin size."""
    heads = heads or DEFAULT_HEADS
    x_in, x_out = {}, {}
    for b in binsizes:
        full, tss = embedding_transformer(promoter_feats[b], promoter_pad_masks[b], sd, f"embed.{b}.",
                                          heads["embed"])
        pw = pairwise_interaction_transformer(full, pcre_feats[b], pcre_pad_masks[b], sd,
                                              f"pairwise_interaction.{b}.", heads["pairwise_interaction"])
        x_in[b] = torch.cat([tss, pw], dim=1)
    for b in binsizes:
        x_out[b] = regulation_transformer(x_in[b], interaction_masks[b], interaction_freq, sd,
                                          f"regulation.{b}.", heads["regulation"])
    z = torch.cat([x_out[b][:, 0] for b in binsizes], dim=1) + torch.cat([x_in[b][:, 0] for b in binsizes], dim=1)
    h = torch.relu(z @ sd["fc_head.0.weight"].t() + sd["fc_head.0.bias"])
    return h @ sd["fc_head.2.weight"].t() + sd["fc_head.2.bias"]


def init_state_dict(regression=False, seed=42, n_feats=7, d_emb=128, d_head=128, binsizes=(2000, 500, 100),
                    embed=(1, 2, 128), pairwise=(2, 2, 128, 256), regulation=(6, 8, 256, 256)):
    """Initial weights of `ChromoformerClassifier(seed)` / `ChromoformerRegressor(seed)` in the dict layout: the same
    sequence of nn.Linear / nn.LayerNorm constructor calls after torch.manual_seed(seed) as net.py:301-330 (regressor:
    net.py:413-428 builds the 2-logit head first, then the 1-output head) with modules.py:16-25,94-96,135-146 inside.
    Tuples are (n_layers, n_heads, [d_model,] d_ff)."""
    import torch.nn as nn
    torch.manual_seed(seed)
    sd = {}

    def put(prefix, module):
        for k, v in module.state_dict().items():
            sd[prefix + k] = v.detach().clone()

    def self_att(pre, heads, d_model, gate):
        sd[pre + "gamma_f"] = torch.ones(heads)
        put(pre + "w_bias.", nn.Linear(2, heads, bias=False))
        put(pre + "att.", nn.Linear(d_emb, (4 if gate else 3) * d_model, bias=False))
        put(pre + "ff.", nn.Linear(d_model, d_emb))
        put(pre + "ln.", nn.LayerNorm(d_emb))

    def ffn(pre, d_ff):
        put(pre + "l1.", nn.Linear(d_emb, d_ff))
        put(pre + "l2.", nn.Linear(d_ff, d_emb))
        put(pre + "ln.", nn.LayerNorm(d_emb))

    for b in binsizes:
        pre = f"embed.{b}."
        put(pre + "lin_proj.", nn.Linear(n_feats, d_emb, bias=False))
        for l in range(embed[0]):
            self_att(f"{pre}transformer.layers.{l}.self_att.", embed[1], d_emb, gate=False)
            ffn(f"{pre}transformer.layers.{l}.ff.", embed[2])
    for b in binsizes:
        pre = f"pairwise_interaction.{b}."
        put(pre + "ln.", nn.LayerNorm(d_emb))
        put(pre + "lin_proj_p.", nn.Linear(d_emb, d_emb, bias=False))
        put(pre + "lin_proj_pcre.", nn.Linear(n_feats, d_emb, bias=False))
        for l in range(pairwise[0]):
            ap = f"{pre}transformer.layers.{l}.self_att."
            sd[ap + "gamma_f"] = torch.ones(pairwise[1])
            put(ap + "p_att.", nn.Linear(d_emb, pairwise[2], bias=False))
            put(ap + "c_att.", nn.Linear(d_emb, 2 * pairwise[2], bias=False))
            put(ap + "ff.", nn.Linear(pairwise[2], d_emb))
            put(ap + "ln.", nn.LayerNorm(d_emb))
            ffn(f"{pre}transformer.layers.{l}.ff.", pairwise[3])
    for b in binsizes:
        pre = f"regulation.{b}."
        for l in range(regulation[0]):
            self_att(f"{pre}transformer.layers.{l}.self_att.", regulation[1], regulation[2], gate=True)
            ffn(f"{pre}transformer.layers.{l}.ff.", regulation[3])
    heads = [(nn.Linear(d_emb * 3, d_head), nn.Linear(d_head, 2))]
    if regression:
        heads.append((nn.Linear(d_emb * 3, d_head), nn.Linear(d_head, 1)))
    put("fc_head.0.", heads[-1][0])
    put("fc_head.2.", heads[-1][1])
    return sd


def legacy_to_dict_layout(sd):
    """Key renaming of misc/convert_weight.py:19-88 (embed2000 -> embed.2000 ...)."""
    out = {}
    for k, v in sd.items():
        head, _, rest = k.partition(".")
        for short, full in (("embed", "embed"), ("pw_int", "pairwise_interaction"), ("reg", "regulation")):
            if head.startswith(short) and head[len(short):].isdigit():
                k = f"{full}.{head[len(short):]}.{rest}"
                break
        out[k] = v
    return out


# ----------------------------------------------------------------------------- training
class bf16_operands:
    """Context manager: every `@` contraction of the restatement (the nn.Linear call sites, modules.py:38,48,100-101,
    159-160; the attention products use torch.matmul and stay FP32) sees its two operands rounded to BF16, with a
    straight-through gradient - the arithmetic contract of `precision = "bf16"` (BF16 operands, FP32 accumulation).  A
    yardstick for how far BF16 training gradients may sit from FP32 ones: the rounding itself, not a kernel's fault."""

    def __enter__(self):
        self._orig = torch.Tensor.__matmul__
        orig = self._orig

        def rnd(x):
            return x + (x.to(torch.bfloat16).to(x.dtype) - x).detach()

        torch.Tensor.__matmul__ = lambda a, b: orig(rnd(a), rnd(b))
        return self

    def __exit__(self, *exc):
        torch.Tensor.__matmul__ = self._orig
        return False


def loss_fn(logits, target, regression):
    """train.py:156,193: MSELoss / CrossEntropyLoss with mean reduction."""
    if regression:
        return torch.mean((logits - target) ** 2)
    return F.cross_entropy(logits, target)


def forward_backward(sd, batch, target, regression, binsizes=(2000, 500, 100)):
    """Loss and gradients of every tensor of ``sd`` by autograd over the restated forward
    (train.py:184-195).  Tensors the graph never touches get ``None``."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    logits = chromoformer_forward(leaves, *batch, binsizes=binsizes)
    loss = loss_fn(logits, target, regression)
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    return loss.detach(), logits.detach(), dict(zip(leaves.keys(), grads))


def adamw_step(p, g, m, v, step, lr=3e-5, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
    """torch.optim.AdamW single-tensor update (train.py:157,196), restated; returns new (p, m, v)."""
    p = p * (1.0 - lr * weight_decay)
    m = m + (g - m) * (1.0 - beta1)
    v = v * beta2 + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * (m / denom), m, v


# ---------------------------------------------------------------------------- input path
def bin_and_pad(x, bin_size, max_n_bins):
    """ChromoformerDataset._bin_and_pad, data.py:68-99.  x: float32 [F, L] numpy array."""
    length = x.shape[1]
    n_bins = math.ceil(length / bin_size)
    cols = [np.log(x[:, i * bin_size:(i + 1) * bin_size].astype(np.float32).mean(axis=1, dtype=np.float32)
                   + np.float32(1.0)) for i in range(n_bins)]
    left = math.ceil((max_n_bins - n_bins) / 2)
    right = math.floor((max_n_bins - n_bins) / 2)
    out = np.zeros((x.shape[0], max_n_bins), dtype=np.float32)
    if n_bins:
        out[:, left:left + n_bins] = np.stack(cols, axis=1)
    return out, left, n_bins, right


def region_representation(raw, bin_size, max_n_bins, strand="+", window=None):
    """_get_region_representation, data.py:101-113.  raw: fp16 [F, L]."""
    x = raw.astype(np.float32)
    if window is not None:
        x = x[:, 20000 - window // 2:20000 + window // 2]
    out, left, n_bins, right = bin_and_pad(x, bin_size, max_n_bins)
    if strand == "+":
        return out, left, n_bins, right
    return out[:, ::-1].copy(), right, n_bins, left


def build_item(promoter_raw, strand, pcre_raws, scores, i_max=8, binsizes=(2000, 500, 100), w_prom=40000,
               w_max=40000):
    """ChromoformerDataset.__getitem__, data.py:115-212, from in-memory raw regions."""
    k = len(pcre_raws)
    item = {"promoter_feats": {}, "promoter_pad_masks": {}, "pcre_feats": {}, "pcre_pad_masks": {},
            "interaction_masks": {}}
    freq = np.zeros((i_max + 1, i_max + 1), dtype=np.float32)
    for b in binsizes:
        n = w_max // b
        xp, lp, nb, _ = region_representation(promoter_raw, b, n, strand, window=w_prom)
        mp = np.ones((1, 1, n, n), dtype=bool)
        mp[0, 0, lp:lp + nb, lp:lp + nb] = False
        feats = np.zeros((i_max, n, xp.shape[0]), dtype=np.float32)
        masks = np.ones((i_max, 1, n, n), dtype=bool)
        for i, (raw, score) in enumerate(zip(pcre_raws, scores)):
            xc, lc, nc, _ = region_representation(raw, b, n)
            feats[i] = xc.T
            masks[i, 0, lp:lp + nb, lc:lc + nc] = False
            freq[0, i + 1] = score
        im = np.ones((1, i_max + 1, i_max + 1), dtype=bool)
        im[0, :k + 1, :k + 1] = False
        item["promoter_feats"][b] = xp.T[None].copy()
        item["promoter_pad_masks"][b] = mp
        item["pcre_feats"][b] = feats
        item["pcre_pad_masks"][b] = masks
        item["interaction_masks"][b] = im
    item["interaction_freq"] = freq
    return item
