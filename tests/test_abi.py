"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol the header
declares, validates configurations, and agrees with the Python modules on the parameter layout."""
import ctypes
import os
import re

import pytest
import torch

from _util import KWS, ROOT
from chromoformer_b200 import Chromoformer, ChromoformerClassifier, ChromoformerRegressor, _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "chromoformer_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(chromo_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/chromoformer_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS), "ctypes table and header disagree"
    assert lib.chromo_abi_version() == 1


def test_config_validation_errors():
    lib = _lib.load()
    m = ChromoformerClassifier()
    cfg = _lib.Config.from_buffer_copy(m._cfg)
    assert lib.chromo_param_total(ctypes.byref(cfg)) == m.flat_params.numel()
    cfg.d_emb = 64
    assert lib.chromo_param_total(ctypes.byref(cfg)) == -1
    assert b"d_emb" in lib.chromo_last_error()
    cfg = _lib.Config.from_buffer_copy(m._cfg)
    cfg.embed_layers = 2
    assert lib.chromo_workspace_floats(ctypes.byref(cfg), 4, 0) == -1
    with pytest.raises(_lib.ChromoLibError):
        _lib.check(lib.chromo_workspace_floats(ctypes.byref(cfg), 4, 0), "ws")
    cfg = _lib.Config.from_buffer_copy(m._cfg)
    assert lib.chromo_workspace_floats(ctypes.byref(cfg), 0, 0) == -1
    assert lib.chromo_workspace_floats(ctypes.byref(cfg), 64, _lib.F_TRAINING) > \
        lib.chromo_workspace_floats(ctypes.byref(cfg), 64, 0) > 0


def test_flat_layout_matches_state_dict():
    """370 tensors; the 36 grad-less ones sit behind the active prefix; parameters are
    views of the flat buffer and survive load_state_dict."""
    m = ChromoformerRegressor(7, 128, 128, dict(KWS[0]), dict(KWS[1]), dict(KWS[2]), seed=1)
    table = _lib.param_table(m._cfg)
    assert len(table) == 370 == len(list(m.parameters()))
    unused = [n for n, (off, _) in table.items() if off >= m.n_active]
    assert len(unused) == 36
    assert all(("w_bias" in n) or ("gamma_f" in n) or n.endswith(".ln.weight") or n.endswith(".ln.bias") for n in unused)
    assert not any(n.startswith("regulation") and "gamma_f" in n for n in unused)
    base = m.flat_params.data_ptr()
    for name, p in m.named_parameters():
        off, numel = table[m._lib_name(name)]
        assert p.data_ptr() == base + 4 * off and p.numel() == numel and off % 4 == 0
    other = ChromoformerRegressor(7, 128, 128, dict(KWS[0]), dict(KWS[1]), dict(KWS[2]), seed=2)
    m.load_state_dict(other.state_dict())
    assert m._flat_is_current()
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), other.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    total = sum(p.numel() for p in m.parameters())
    assert total == 5342543 and ChromoformerClassifier().flat_params.numel() >= 5342672


def test_legacy_and_dict_models_share_weights():
    a, b = Chromoformer(seed=3), ChromoformerClassifier(seed=3)
    assert torch.equal(a.flat_params, b.flat_params)
    assert "embed2000.lin_proj.weight" in a.state_dict() and "embed.2000.lin_proj.weight" in b.state_dict()


def test_no_cpu_fallback():
    m = ChromoformerClassifier()
    from chromoformer_b200 import synthetic
    batch = synthetic.make_batch(2, seed=0)
    with pytest.raises(_lib.ChromoLibError):
        m(*synthetic.forward_args(batch))
    with pytest.raises(NotImplementedError):
        m.embed["2000"](batch["promoter_feats"][2000], batch["promoter_pad_masks"][2000])
