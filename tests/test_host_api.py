"""CPU-side checks: the C-ABI library loads and exports every symbol include/lightglue_b200.h
declares, and the host-side mirror of the reference interface behaves like the reference's
(constructor / conf / state_dict names / error behaviour).  No compute calls (no GPU here)."""
import os
import re

import pytest
import torch

from lightglue_b200 import LightGlue, _cabi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    header = open(os.path.join(ROOT, "include", "lightglue_b200.h")).read()
    declared = re.findall(r"LG_API\s+[\w\s\*]+?\b(lg_\w+)\s*\(", header)
    assert sorted(declared) == sorted(_cabi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.lg_build_info()


def test_blob_size_matches_parameter_count():
    lib = _cabi.load()
    for d, so in ((256, False), (128, False), (128, True)):
        m = LightGlue(features=None, input_dim=d, add_scale_ori=so)
        assert lib.lg_weight_blob_floats(d, 4 if so else 2, 9) == sum(t.numel() for t in m._blob_tensors())
    assert sum(p.numel() for p in LightGlue(features=None).parameters()) == 11_851_601


def test_state_dict_uses_reference_key_names():
    m = LightGlue(features=None, input_dim=128)
    sd = synth.make_state_dict(input_dim=128)
    keys = set(m.state_dict().keys())
    assert set(sd.keys()) | {"confidence_thresholds"} == keys
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert missing == ["confidence_thresholds"] and not unexpected


def test_legacy_checkpoint_names_are_renamed():
    sd = {"self_attn.3.Wqkv.weight": torch.zeros(1), "cross_attn.0.to_qk.bias": torch.zeros(1), "posenc.Wr.weight": 0}
    out = LightGlue._rename_legacy(sd, 9)
    assert "transformers.3.self_attn.Wqkv.weight" in out and "transformers.0.cross_attn.to_qk.bias" in out


def test_constructor_contract():
    with pytest.raises(ValueError):
        LightGlue(features="orb")
    m = LightGlue(features=None, depth_confidence=-1, filter_threshold=0.2)
    assert m.conf.depth_confidence == -1 and m.conf.width_confidence == 0.99 and m.conf.filter_threshold == 0.2
    assert m.conf.n_layers == 9 and m.conf.num_heads == 4 and m.conf.descriptor_dim == 256
    assert LightGlue.pruning_keypoint_thresholds == {"cpu": -1, "mps": -1, "cuda": 1024, "flash": 1536}
    assert LightGlue.required_data_keys == ["image0", "image1"]
    assert torch.allclose(
        m.confidence_thresholds,
        torch.tensor([0.9000, 0.8641, 0.8411, 0.8264, 0.8169, 0.8108, 0.8069, 0.8045, 0.8029]), atol=5e-5,
    )
    m.compile()
    assert m.static_lengths == [256, 512, 768, 1024, 1280, 1536]


def test_forward_rejects_missing_keys_and_cpu_tensors():
    m = LightGlue(features=None)
    data, _ = synth.make_pair(16)
    with pytest.raises(AssertionError):
        m({"image0": data["image0"]})
    with pytest.raises(RuntimeError):  # no CPU path: fail loudly
        m(data)
