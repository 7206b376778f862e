"""CPU: the oracle restatement against (a) the golden vectors produced by the unmodified reference and
(b) the reference itself when /root/reference is present (build container)."""
import os

import pytest
import torch

from oracle import configs, ref_loader
from oracle import taskprompter_ref as TPR
from oracle import invpt_ref as IPR
from oracle.make_golden import sd_checksum

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["tp_tiny", "tp_tiny1", "tp_tiny_de"])
def test_taskprompter_oracle_vs_golden(name):
    fx = torch.load(os.path.join(GOLD, f"{name}.pt"), weights_only=False)
    cfg = configs.taskprompter(fx["cfg"])
    sd = TPR.init_state_dict(cfg, seed=fx["seed"])
    assert sd_checksum(sd) == fx["sd_sha256"], "deterministic initialiser drifted from the fixture"
    with torch.no_grad():
        out = TPR.forward(sd, cfg, fx["x"])
    for t, ref in fx["out"].items():
        # same algorithm, same fp32 library kernels, different op order: rounding-level agreement
        assert (out[t] - ref).abs().max() <= 2e-6 * ref.abs().max().clamp_min(1.0) + 2e-6, t
        if t in ("semseg", "human_parts"):
            assert torch.equal(out[t].argmax(1), ref.argmax(1))


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
@pytest.mark.parametrize("name", ["tp_tiny", "tp_tiny1", "tp_tiny_de"])
def test_taskprompter_oracle_vs_reference(name):
    cfg = configs.taskprompter(name)
    torch.manual_seed(0)
    model = ref_loader.build_taskprompter(cfg).eval()     # the reference's own initialisation
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.8, 1.2)
    x = torch.randn(2, 3, *cfg["img_size"])
    with torch.no_grad():
        ref = model(x)
        out = TPR.forward(model.state_dict(), cfg, x)
    for t in cfg["tasks"]:
        assert (out[t] - ref[t]).abs().max() <= 2e-6, t


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
@pytest.mark.parametrize("name", ["tp_tiny", "tp_tiny_de"])
def test_accelerate_shares_reference_state_dict(name):
    import mtt_b200  # noqa: F401
    from mtt_b200 import taskprompter as TP

    cfg = configs.taskprompter(name)
    ref = ref_loader.build_taskprompter(cfg).eval()
    mine = TP.accelerate(ref)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in a)


@pytest.mark.parametrize("name", ["ip_tiny", "ip_cfg1"])
def test_invpt_oracle_vs_golden(name):
    fx = torch.load(os.path.join(GOLD, f"{name}.pt"), weights_only=False)
    cfg = configs.invpt(fx["cfg"])
    sd = IPR.init_state_dict(cfg, seed=fx["seed"])
    assert sd_checksum(sd) == fx["sd_sha256"], "deterministic initialiser drifted from the fixture"
    with torch.no_grad():
        out = IPR.forward(sd, cfg, fx["x"])
    for t, ref in fx["out"].items():
        assert (out[t] - ref).abs().max() <= 5e-6 * ref.abs().max().clamp_min(1.0), t
        if fx["inter_preds"] is not None:
            assert (out["inter_preds"][t] - fx["inter_preds"][t]).abs().max() <= 5e-6, t


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
@pytest.mark.parametrize("name", ["ip_tiny", "ip_cfg1"])
def test_invpt_oracle_vs_reference(name):
    cfg = configs.invpt(name)
    torch.manual_seed(0)
    model = ref_loader.build_invpt(cfg).eval()
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.SyncBatchNorm)):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.8, 1.2)
    x = torch.randn(2, 3, *cfg["img_size"])
    with torch.no_grad():
        ref = model(x)
        out = IPR.forward(model.state_dict(), cfg, x)
    for t in cfg["tasks"]:
        assert (out[t] - ref[t]).abs().max() <= 5e-6, t
        assert (out["inter_preds"][t] - ref["inter_preds"][t]).abs().max() <= 5e-6, t


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_invpt_accelerate_shares_reference_state_dict():
    import mtt_b200  # noqa: F401
    from mtt_b200 import invpt as IP

    cfg = configs.invpt("ip_tiny")
    ref = ref_loader.build_invpt(cfg).eval()
    mine = IP.accelerate(ref)
    a, b = ref.state_dict(), mine.state_dict()
    assert set(a.keys()) == set(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in a)


# ---- Swin-backbone TaskPrompter: oracle only, no CUDA path yet -------------------------
@pytest.mark.parametrize("name", ["tps_tiny", "tps_tiny4"])
def test_swin_oracle_vs_golden(name):
    from oracle import taskprompter_swin_ref as SR

    fx = torch.load(os.path.join(GOLD, f"{name}.pt"), weights_only=False)
    cfg = configs.taskprompter_swin(fx["cfg"])
    sd = SR.init_state_dict(cfg, seed=fx["seed"])
    assert sd_checksum(sd) == fx["sd_sha256"], "deterministic initialiser drifted from the fixture"
    with torch.no_grad():
        out = SR.forward(sd, cfg, fx["x"])
    for t, ref in fx["out"].items():
        assert out[t].shape == ref.shape
        assert (out[t] - ref).abs().max() <= 5e-6 * ref.abs().max().clamp_min(1.0) + 2e-6, t
        if t == "semseg":
            assert torch.equal(out[t].argmax(1), ref.argmax(1))


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
@pytest.mark.parametrize("name", ["tps_tiny", "tps_tiny4", "tps_mid"])
def test_swin_oracle_vs_reference(name):
    """Against the unmodified reference with ITS OWN initialisation (plus non-trivial biases, bias tables and
    BatchNorm statistics): shifted and padded windows, 2x2 channel windows, patch merging of the logit maps."""
    from oracle import taskprompter_swin_ref as SR

    cfg = configs.taskprompter_swin(name)
    torch.manual_seed(0)
    model = ref_loader.build_taskprompter_swin(cfg).eval()
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.SyncBatchNorm)):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.8, 1.2)
        for k, v in model.named_parameters():
            if "relative_position_bias_table" in k:
                v.normal_(0, 0.5)
            elif k.endswith(".bias"):
                v.normal_(0, 0.05)
    ref_sd = model.state_dict()
    derived = {k for k in ref_sd if "relative_position_index" in k or "attn_mask" in k}
    shapes = SR.param_shapes(cfg)
    assert set(shapes) == set(ref_sd) - derived
    assert all(tuple(ref_sd[k].shape) == tuple(shapes[k]) for k in shapes)
    x = torch.randn(1 if name == "tps_mid" else 2, 3, *cfg["img_size"])
    with torch.no_grad():
        ref = model(x)
        out = SR.forward(ref_sd, cfg, x)
    for t in cfg["tasks"]:
        assert out[t].shape == ref[t].shape
        assert (out[t] - ref[t]).abs().max() <= 5e-6, t


def test_swin_window_tables_match_definitions():
    """The derived tables the oracle recomputes (instead of reading the reference's buffers)."""
    from oracle import taskprompter_swin_ref as SR

    idx = SR.relative_position_index(3)
    assert idx.shape == (9, 9) and idx.min() == 0 and idx.max() == 24 and idx[0, 0] == 12      # centre of a 5x5 table
    assert idx[0, 8] == 0 and idx[8, 0] == 24
    m = SR.shifted_window_mask(8, 8, 4, 2)
    assert m.shape == (4, 16, 16) and set(m.unique().tolist()) == {-100.0, 0.0}
    assert (m[0] == 0).all()                      # the top-left window holds one region only
    assert (m[3] != 0).any() and torch.equal(m[3], m[3].t())


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_swin_param_shapes_at_the_reference_config():
    """The reference's Cityscapes-3D Swin-B model (cs_swinB_taskprompter.yml: 1024x2048, img_ds_ratio 0.75, window 12,
    depths 2-2-18-2): every parameter name and shape the oracle expects equals the reference module's (274.6 M
    parameters; construction only, no forward)."""
    from oracle import taskprompter_swin_ref as SR

    cfg = configs.taskprompter_swin("tps_swinB")
    ref_sd = ref_loader.build_taskprompter_swin(cfg).state_dict()
    derived = {k for k in ref_sd if "relative_position_index" in k or "attn_mask" in k}
    shapes = SR.param_shapes(cfg)
    assert set(shapes) == set(ref_sd) - derived and len(shapes) == 732
    assert all(tuple(ref_sd[k].shape) == tuple(shapes[k]) for k in shapes)
    assert [SR.stage_geometry(cfg, i)[1] for i in range(4)] == [(192, 384), (96, 192), (48, 96), (24, 48)]
    assert [SR.level_resolution(cfg, i) for i in range(4)] == [(96, 192), (48, 96), (24, 48), (24, 48)]
