"""CPU tests of the host-side launch sequences (packing, offsets, level bookkeeping, BN folding) with
the kernels replaced by tests/emul_ops.py, against the oracle restatement."""
import pytest
import torch

from oracle import configs
from oracle import taskprompter_ref as TPR


@pytest.mark.parametrize("name,nsplit,tol", [("tp_tiny", 2, 2e-4), ("tp_tiny1", 2, 2e-4), ("tp_tiny", 1, 8e-2),
                                             ("tp_tiny_de", 2, 2e-4)])
def test_taskprompter_plan_matches_oracle(monkeypatch, name, nsplit, tol):
    import mtt_b200  # noqa: F401
    from mtt_b200 import taskprompter as TP
    import emul_ops

    emul_ops.install(monkeypatch)
    cfg = configs.taskprompter(name)
    sd = TPR.init_state_dict(cfg, seed=3)
    model = TP.build_from_config(cfg, nsplit=nsplit, use_graph=False).eval()
    missing = model.load_state_dict(sd, strict=True)
    torch.manual_seed(1)
    x = torch.randn(2, 3, *cfg["img_size"])
    with torch.no_grad():
        ref = TPR.forward(sd, cfg, x)
        got = model.plan(2, torch.device("cpu")).run(x, graph=False)
    for t in cfg["tasks"]:
        err = (got[t] - ref[t]).norm() / ref[t].norm()
        assert got[t].shape == ref[t].shape
        assert err < tol, f"{name} {t}: rel-L2 {err:.3e}"


def test_wrapper_refuses_cpu_and_training():
    import mtt_b200  # noqa: F401
    from mtt_b200 import taskprompter as TP

    cfg = configs.taskprompter("tp_tiny1")
    model = TP.build_from_config(cfg)
    with pytest.raises(NotImplementedError):
        model.train()(torch.zeros(1, 3, *cfg["img_size"]))
    with pytest.raises(RuntimeError):
        model.eval()(torch.zeros(1, 3, *cfg["img_size"]))


def test_state_dict_keys_match_oracle_names():
    import mtt_b200  # noqa: F401
    from mtt_b200 import taskprompter as TP

    for name in ("tp_tiny", "tp_tiny1", "tp_tiny_de"):
        cfg = configs.taskprompter(name)
        model = TP.build_from_config(cfg)
        sd = TPR.init_state_dict(cfg)
        mine = model.state_dict()
        assert set(mine) == set(sd)
        assert all(mine[k].shape == sd[k].shape for k in sd)


@pytest.mark.parametrize("name,nsplit,tol", [("ip_tiny", 2, 3e-4), ("ip_cfg1", 2, 3e-4)])
def test_invpt_plan_matches_oracle(monkeypatch, name, nsplit, tol):
    import mtt_b200  # noqa: F401
    from mtt_b200 import invpt as IP
    from oracle import invpt_ref as IPR
    import emul_ops

    emul_ops.install(monkeypatch)
    cfg = configs.invpt(name)
    sd = IPR.init_state_dict(cfg, seed=5)
    model = IP.build_from_config(cfg, nsplit=nsplit, use_graph=False).eval()
    model.load_state_dict(sd, strict=True)
    torch.manual_seed(2)
    x = torch.randn(2, 3, *cfg["img_size"])
    with torch.no_grad():
        ref = IPR.forward(sd, cfg, x)
        got = model.plan(2, torch.device("cpu")).run(x, graph=False)
    for t in cfg["tasks"]:
        err = (got[t] - ref[t]).norm() / ref[t].norm()
        assert err < tol, f"{name} {t}: rel-L2 {err:.3e}"
        err = (got["inter_preds"][t] - ref["inter_preds"][t]).norm() / ref["inter_preds"][t].norm()
        assert err < tol, f"{name} inter {t}: rel-L2 {err:.3e}"


def test_taskprompter_predict_matches_get_output(monkeypatch):
    """predict() = forward + get_output (TP/utils/utils.py:27-63) fused into the final resize."""
    import mtt_b200  # noqa: F401
    from mtt_b200 import taskprompter as TP
    from oracle import postproc_ref
    import emul_ops

    emul_ops.install(monkeypatch)
    cfg = configs.taskprompter("tp_tiny")
    sd = TPR.init_state_dict(cfg, seed=3)
    model = TP.build_from_config(cfg, nsplit=2, use_graph=False).eval()
    model.load_state_dict(sd, strict=True)
    torch.manual_seed(1)
    x = torch.randn(2, 3, *cfg["img_size"])
    with torch.no_grad():
        ref = TPR.forward(sd, cfg, x)
        got = model.plan(2, torch.device("cpu"), postproc=True).run(x, graph=False)
    for t in cfg["tasks"]:
        want = postproc_ref.get_output(ref[t], t)
        assert got[t].shape == want.shape and got[t].dtype == want.dtype, t
        if want.dtype == torch.int64:
            assert (got[t] == want).float().mean() > 0.995, t
        else:
            assert (got[t] - want).abs().max() <= 2e-3 * want.abs().max().clamp_min(1.0), t
