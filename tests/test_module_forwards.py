"""The reference's sub-module forward signatures on the fused kernels, against the oracle's taps:

  Block.forward(x, task_prompts) -> (x, (prompt_logits, raw_chan), task_prompts)     TP taskprompter.py:270-279
  TaskPrompter.forward(x)        -> (task_fea, {})                                   :392-422
  ConvHead / DEConvHead.forward  -> logits at the head's resolution                  :697, :712-715

Each case runs twice: on CPU with the kernels replaced by tests/emul_ops.py (host logic: packing, joint-stream
layout, workspaces) and, marked gpu, through the C ABI on the device.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import configs
from oracle import taskprompter_ref as TPR


def _model(name, seed=3):
    import mtt_b200  # noqa: F401
    from mtt_b200 import taskprompter as TP

    cfg = configs.taskprompter(name)
    sd = TPR.init_state_dict(cfg, seed=seed)
    m = TP.build_from_config(cfg, nsplit=2, use_graph=False).eval()
    m.load_state_dict(sd, strict=True)
    return TP, cfg, sd, m


def _rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()


def _run_cases(dev, tol):
    out = {}
    for name in ("tp_tiny", "tp_tiny_de"):
        TP, cfg, sd, m = _model(name)
        m = m.to(dev)
        g = torch.Generator().manual_seed(17)
        x = torch.randn(2, 3, *cfg["img_size"], generator=g)
        taps = {}
        with torch.no_grad():
            ref = TPR.forward(sd, cfg, x, taps=taps)
            # ---- TaskPrompter.forward
            fea, info = m.backbone(x.to(dev))
            assert info == {}
            for t in cfg["tasks"]:
                assert fea[t].shape == taps[f"task_fea.{t}"].shape
                assert _rel(fea[t], taps[f"task_fea.{t}"]) < tol, (name, "task_fea", t)
            # ---- heads on the oracle's task features
            for t in cfg["tasks"]:
                hd = TPR.deconv_head if cfg["head"] == "deconv" else TPR.conv_head
                want = hd(sd, t, taps[f"task_fea.{t}"])
                got = m.heads[t](taps[f"task_fea.{t}"].to(dev))
                assert got.shape == want.shape
                assert _rel(got, want) < tol, (name, "head", t)
            # ---- Block.forward: block 1 on the oracle's block-0 outputs
            xb, pb = taps["block0.x"].to(dev), taps["block0.prompts"].to(dev)
            x1, (lg, rc), p1 = m.backbone.blocks[1](xb, pb)
            rx, rp, rl, rrc = TPR.block_forward(sd, "backbone.blocks.1.", cfg, taps["block0.x"], taps["block0.prompts"], True)
            assert _rel(x1, rx) < tol and _rel(p1, rp) < tol, (name, "block")
            assert _rel(lg, rl) < tol and _rel(rc, rrc) < tol, (name, "block logits")
            m.backbone.blocks[1].emit_logits = False
            _, (lg2, rc2), _ = m.backbone.blocks[1](xb, pb)
            assert lg2 is None and rc2 is None
        out[name] = True
    return out


def _run_invpt_cases(dev, tol):
    import mtt_b200  # noqa: F401
    from mtt_b200 import invpt as IP
    from oracle import invpt_ref as IPR

    cfg = configs.invpt("ip_tiny")
    sd = IPR.init_state_dict(cfg, seed=5)
    m = IP.build_from_config(cfg, nsplit=2, use_graph=False).eval()
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    x = torch.randn(2, 3, *cfg["img_size"], generator=torch.Generator().manual_seed(23))
    taps = {}
    with torch.no_grad():
        ref = IPR.forward(sd, cfg, x, taps=taps)
        feats_ref = IPR.vit_forward(sd, cfg, x)
        # ---- VisionTransformer.forward (vit.py:332-361)
        last, feats = m.backbone(x.to(dev))
        assert len(feats) == 4 and feats[-1].shape == last.shape
        for a, b in zip(feats, feats_ref):
            assert a.shape == b.shape and _rel(a, b) < tol, "vit features"
        # ---- TransformerDecoder.forward on the oracle's features (transformer_decoder.py:69-98)
        x_dict, inter = m.multi_task_decoder([f.to(dev) for f in feats_ref])
        for t in cfg["tasks"]:
            assert _rel(x_dict[t], taps[f"x_dict.{t}"]) < tol, ("x_dict", t)
            assert _rel(inter[t], taps[f"inter.{t}"]) < tol, ("inter_pred", t)
            # ---- MLPHead.forward + the wrapper's resize reproduce the model output
            y = F.interpolate(m.heads[t](x_dict[t]).cpu(), x.shape[-2:], mode="bilinear")
            assert _rel(y, ref[t]) < tol, ("head", t)
        # ---- InvPT.forward on the oracle's intermediate tensors (invpt.py:502-544)
        xd = m.multi_task_decoder.invpt({t: taps[f"ms_feat.{t}"].to(dev) for t in cfg["tasks"]},
                                       {t: taps[f"inter.{t}"].to(dev) for t in cfg["tasks"]},
                                       [taps["back0"].to(dev), taps["back1"].to(dev), None, None])
        for t in cfg["tasks"]:
            assert _rel(xd[t], taps[f"x_dict.{t}"]) < tol, ("invpt", t)


def test_invpt_module_forwards_emulated(monkeypatch):
    import mtt_b200  # noqa: F401
    from mtt_b200 import taskprompter as TP, invpt as IP
    import emul_ops

    emul_ops.install(monkeypatch)
    monkeypatch.setattr(TP, "_check_input", lambda mod, x: None)
    monkeypatch.setattr(IP, "_check_input", lambda mod, x: None)
    _run_invpt_cases(torch.device("cpu"), 3e-4)


@pytest.mark.gpu
def test_invpt_module_forwards_gpu(cuda_dev):
    _run_invpt_cases(cuda_dev, 3e-4)


def test_module_forwards_emulated(monkeypatch):
    import mtt_b200  # noqa: F401
    from mtt_b200 import taskprompter as TP
    import emul_ops

    emul_ops.install(monkeypatch)
    monkeypatch.setattr(TP, "_check_input", lambda mod, x: None)
    _run_cases(torch.device("cpu"), 2e-4)


@pytest.mark.gpu
def test_module_forwards_gpu(cuda_dev):
    _run_cases(cuda_dev, 2e-4)


@pytest.mark.gpu
def test_weights_are_shared_between_plans_and_repacked_on_update(cuda_dev):
    """One packed copy per (module, device, mode) serves every plan; an in-place parameter update re-packs it."""
    TP, cfg, sd, m = _model("tp_tiny1", seed=4)
    m = m.to(cuda_dev)
    x = torch.randn(2, 3, *cfg["img_size"], device=cuda_dev)
    with torch.no_grad():
        a = {k: v.clone() for k, v in m(x).items()}
        pl2, pl1 = m.plan(2, cuda_dev), m.plan(1, cuda_dev)
        assert pl2.Wb[0] is pl1.Wb[0] and pl2.Wh[0] is pl1.Wh[0]
        m.heads[cfg["tasks"][0]].linear_pred.bias.add_(1.0)
        b = m(x)
    t0 = cfg["tasks"][0]
    assert torch.allclose(b[t0], a[t0] + 1.0, atol=1e-4)
