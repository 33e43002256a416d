"""CPU: host-side logic of the drop-in model -- module tree / state_dict contract, init parity with
the oracle, the no-fallback rule, gradient-arena layout, batch sharding."""
import pytest
import torch

import ecgmm
from ecgmm import lib
from ecgmm import model as M
from ecgmm.parallel import shard_batch
from oracle import model as om


class Cfg:
    num_classes = 2
    device = "cpu"


def test_state_dict_layout_and_init_match_reference_restatement():
    torch.manual_seed(123)
    dut = M.ECGMultimodalModel(Cfg)
    torch.manual_seed(123)
    ora = om.ECGMultimodalModel()
    sd, so = dut.state_dict(), ora.state_dict()
    assert len(sd) == 229 and list(sd) == list(so)
    assert sum(p.numel() for p in dut.parameters()) == 11_912_487
    for k in so:
        assert sd[k].shape == so[k].shape and sd[k].dtype == so[k].dtype, k
        assert torch.equal(sd[k], so[k]), f"same seed must give the same initial {k}"
    ora.load_state_dict(sd, strict=True)
    dut.load_state_dict(so, strict=True)


def test_g3_dims_and_aliases():
    m = M.ECGMultimodalModel(Cfg, dims=(512, 128, 32), clinical_features=2)
    assert (m.image_dim, m.signal_dim, m.clinical_dim, m.modal_dim) == (512, 128, 32, 512)
    assert m.image_encoder.fc.weight.shape == (512, 512)
    assert m.fusion_classifier[0].weight.shape == (128, 672)
    assert m.get_clinical_feature_dim() == 2
    assert ecgmm.MultimodalModel is M.ECGMultimodalModel
    for name in ("image_encoder", "image_norm", "signal_encoder", "signal_norm", "clinical_encoder", "clinical_norm",
                 "image_classifier", "signal_classifier", "clinical_classifier", "attention_fusion",
                 "fusion_classifier"):
        assert hasattr(m, name)
    assert hasattr(m.attention_fusion, "weights") and hasattr(m.attention_fusion, "norm")
    assert hasattr(m.signal_encoder, "initial") and hasattr(m.signal_encoder, "classifier")


def test_freeze_mode_parameter_count():
    m = M.ECGMultimodalModel(Cfg)
    for enc in (m.image_encoder, m.signal_encoder, m.clinical_encoder):  # train.py:35-40
        for p in enc.parameters():
            p.requires_grad = False
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 103_307


def test_signal_checkpoint_loads_into_standalone_encoder():
    import os

    from golden_util import GOLDEN_DIR

    net = M.ResNet1D_SE(1, 2)
    sd = torch.load(os.path.join(GOLDEN_DIR, "best_ptbxl.pth"), map_location="cpu")
    net.load_state_dict(sd, strict=True)
    m = M.ECGMultimodalModel(Cfg)
    res = m.load_pretrained_signal_encoder(os.path.join(GOLDEN_DIR, "best_ptbxl.pth"))
    assert set(res.missing_keys) == {"classifier.4.weight", "classifier.4.bias"}


def test_no_cpu_or_aten_fallback():
    m = M.ECGMultimodalModel(Cfg)
    with pytest.raises(lib.EcgmmError):
        m(torch.zeros(2, 3, 32, 32), torch.zeros(2, 100), torch.zeros(2, 24))
    with pytest.raises(lib.EcgmmError):
        m.image_encoder.conv1(torch.zeros(1, 3, 8, 8))
    with pytest.raises(lib.EcgmmError):
        m.image_encoder.layer1[0].bn1(torch.zeros(1, 64, 2, 2))
    with pytest.raises(lib.EcgmmError):
        m.fusion_classifier(torch.zeros(2, 768))


def test_grad_arena_reverse_execution_layout():
    enc = M.ResNet18()
    order = enc._exec_order_params()
    assert {id(p) for p in order} == {id(p) for p in enc.parameters()}
    G = M.GradArena(order, "cpu")
    assert G.offsets[id(enc.fc.bias)][0] == 0  # last executed parameter first
    end_l4 = G.end_of(enc.layer4[0].conv1.weight)
    end_l3 = G.end_of(enc.layer3[0].conv1.weight)
    assert 0 < end_l4 < end_l3 < G.total
    assert G.offsets[id(enc.conv1.weight)][0] + enc.conv1.weight.numel() <= G.total
    for off, n, _ in G.offsets.values():
        assert off % 4 == 0
    g = G(enc.conv1.weight)
    assert g.shape == enc.conv1.weight.shape and g.data_ptr() >= G.flat.data_ptr()
    assert float(G.flat.abs().sum()) == 0.0


def test_shard_batch():
    x, y = torch.arange(16).view(8, 2), torch.arange(8)
    parts = [shard_batch((x, y), r, 4) for r in range(4)]
    assert torch.equal(torch.cat([p[0] for p in parts]), x) and torch.equal(torch.cat([p[1] for p in parts]), y)
    with pytest.raises(ValueError):
        shard_batch((x,), 0, 3)


def test_arena_layout_is_reverse_execution_order_and_aligned():
    """Gradient arenas: parameters in REVERSE execution order (so finished prefixes can be all-reduced early), every
    slice 16-byte aligned, duplicates laid out once."""
    from ecgmm.model import ArenaLayout, GradArena

    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 8, 3, 64)]
    lay = ArenaLayout(ps)
    offs = [lay.offsets[id(p)][0] for p in ps]
    assert offs == sorted(offs, reverse=True) and offs[-1] == 0
    assert ArenaLayout(ps + [ps[1]]).total == lay.total  # a parameter listed twice is laid out once
    assert all(o % 4 == 0 for o in offs) and lay.total % 4 == 0
    G = GradArena(lay, torch.device("cpu"))
    assert G.flat.numel() == lay.total and float(G.flat.abs().sum()) == 0.0
    assert G(ps[0]).shape == ps[0].shape and G.end_of(ps[3]) == 64
    a, b = G(ps[2]), G(ps[2])
    assert a is not b and a.data_ptr() == b.data_ptr()  # fresh views every time (see test_param_grads_alias_the_arena)


def test_stage_caches_are_dropped_on_mode_and_device_changes():
    """The per-stage parameter list / arena layout caches must not survive train()/eval() or .to()."""
    m = ecgmm.ECGMultimodalModel(Cfg)
    enc = m.image_encoder
    ps = enc.cached_params()
    assert enc.cached_params() is ps and len(ps) == len(list(enc.parameters()))
    lay = enc.new_arena(torch.device("cpu")).offsets
    assert enc.new_arena(torch.device("cpu")).offsets is lay
    m.eval()
    assert enc.cached_params() is not ps
    lay2 = enc.new_arena(torch.device("cpu")).offsets
    assert lay2 is not lay
    m.to(torch.float32)
    assert enc.new_arena(torch.device("cpu")).offsets is not lay2
    head_ps = m._head.cached_params()
    m.train()
    assert m._head.cached_params() is not head_ps
    # the head borrows the model's modules: fc.weight of a replaced sub-module shows up after the next mode switch
    assert any(p is m.fusion_classifier.lin1.weight for p in m._head.cached_params())
