"""CPU: the loss restatement (oracle/loss_ref.py) and the backward pass it drives through the forward restatement,
against golden vectors made with the reference's own criterion and autograd through the unmodified reference model
(tests/golden/losses.pt, oracle/make_golden.py::make_losses). This pins the oracle for the training rows of SURVEY.md
section 8f (N1 backward, N3 loss reductions); there is no CUDA training path yet, so nothing here is marked gpu."""
import os

import pytest
import torch

from oracle import configs, loss_ref
from oracle import taskprompter_ref as TPR

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses.pt")


@pytest.fixture(scope="module")
def fx():
    return torch.load(GOLD, weights_only=False)


def test_task_losses_and_prediction_gradients(fx):
    c = fx["criterion"]
    preds = {t: c["preds"][t].clone().requires_grad_() for t in c["tasks"]}
    out = loss_ref.multi_task_loss(preds, c["labels"], c["tasks"], fx["weights"])
    for k, want in c["losses"].items():
        assert abs(float(out[k].detach()) - want) <= 2e-6 * max(1.0, abs(want)), (k, float(out[k].detach()), want)
    out["total"].backward()
    for t in c["tasks"]:
        d = (preds[t].grad - c["dpreds"][t]).abs().max().item()
        assert d <= 1e-7 + 1e-5 * c["dpreds"][t].abs().max().item(), (t, d)


def test_loss_edge_cases():
    # everything ignored: the reference divides by max(n_valid, 1) and returns 0
    pred = torch.randn(1, 4, 5, 6)
    assert float(loss_ref.cross_entropy(pred, torch.full((1, 1, 5, 6), 255.0))) == 0.0
    assert float(loss_ref.l1(torch.randn(1, 1, 5, 6), torch.full((1, 1, 5, 6), -1.0), ignore_index=-1)) == 0.0
    # HED weighting with no positive pixel (w == 1): the reference returns 0
    assert float(loss_ref.balanced_bce(torch.randn(1, 1, 4, 4), torch.zeros(1, 1, 4, 4))) == 0.0


def test_backward_through_the_forward_restatement_matches_reference_autograd(fx):
    m = fx["model"]
    cfg = configs.taskprompter(m["cfg"])
    sd = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running_" not in k else v)
          for k, v in TPR.init_state_dict(cfg, seed=m["seed"]).items()}
    out = TPR.forward(sd, cfg, m["x"])
    loss = loss_ref.multi_task_loss(out, m["labels"], cfg["tasks"], fx["weights"])
    for k, want in m["losses"].items():
        assert abs(float(loss[k].detach()) - want) <= 1e-5 * max(1.0, abs(want)), (k, float(loss[k].detach()), want)
    loss["total"].backward()
    assert set(m["grad_norm"]) == {k for k, v in sd.items() if v.requires_grad}
    # a parameter the outputs do not depend on (token_trans1 of the LAST block: its prompts are never read again) has
    # no gradient here and an all-zero one in the reference
    grad = {k: (sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])) for k in m["grad_norm"]}
    assert [k for k in m["grad_norm"] if sd[k].grad is None] == [k for k, n in m["grad_norm"].items() if n == 0.0]
    for k, want in m["grad_norm"].items():
        got = float(grad[k].norm())
        assert abs(got - want) <= 2e-4 * want + 1e-7, (k, got, want)
    for k, want in m["grad_full"].items():
        d = (grad[k] - want).abs().max().item()
        assert d <= 1e-6 + 2e-4 * want.abs().max().item(), (k, d)
    for k, want in m["grad_sum"].items():
        got = float(grad[k].double().sum())
        assert abs(got - want) <= 1e-5 + 2e-3 * m["grad_norm"][k], (k, got, want)


@pytest.mark.gpu
def test_device_losses_match_the_reference_criterion(cuda_dev, fx):
    """mtt_b200.losses (device reductions + gradient kernels) against the values and prediction gradients the
    REFERENCE criterion produced (golden part a), and on the all-ignored / no-positive edge cases."""
    import mtt_b200  # noqa: F401
    from mtt_b200 import losses as ML

    c = fx["criterion"]
    p = {"TASKS": {"NAMES": c["tasks"]}, "edge_w": 0.95, "ignore_index": 255, "ignore_invalid_area_depth": True,
         "loss_kwargs": {"loss_weights": fx["weights"]}}
    crit = ML.get_criterion(p)
    preds = {t: c["preds"][t].to(cuda_dev).requires_grad_() for t in c["tasks"]}
    labels = {t: c["labels"][t].to(cuda_dev) for t in c["tasks"]}
    out = crit(preds, labels, tasks=c["tasks"])
    for k, want in c["losses"].items():
        assert abs(float(out[k].detach()) - want) <= 3e-6 * max(1.0, abs(want)), (k, float(out[k].detach()), want)
    out["total"].backward()
    for t in c["tasks"]:
        d = (preds[t].grad.cpu() - c["dpreds"][t]).abs().max().item()
        assert d <= 1e-7 + 2e-5 * c["dpreds"][t].abs().max().item(), (t, d)
    # edge cases (the reference divides by max(n_valid, 1) / returns 0)
    z = lambda *s: torch.randn(*s, device=cuda_dev)
    assert float(ML.CrossEntropyLoss()(z(1, 4, 5, 6), torch.full((1, 1, 5, 6), 255.0))) == 0.0
    assert float(ML.L1Loss(ignore_index=-1)(z(1, 1, 5, 6), torch.full((1, 1, 5, 6), -1.0))) == 0.0
    assert float(ML.BalancedBinaryCrossEntropyLoss()(z(1, 1, 4, 4), torch.zeros(1, 1, 4, 4))) == 0.0
    # HED weighting and a large map (more elements than one pass of the grid) against the oracle restatement
    x, y = z(2, 1, 300, 400), (torch.rand(2, 1, 300, 400) < 0.2).float()
    y[torch.rand(2, 1, 300, 400) < 0.05] = 255.0
    want = loss_ref.balanced_bce(x.cpu(), y)
    assert abs(float(ML.BalancedBinaryCrossEntropyLoss()(x, y)) - float(want)) <= 3e-6 * float(want)
    xs, ys = z(2, 21, 300, 400), torch.randint(0, 21, (2, 1, 300, 400)).float()
    want = loss_ref.cross_entropy(xs.cpu(), ys)
    assert abs(float(ML.CrossEntropyLoss()(xs, ys)) - float(want)) <= 3e-6 * float(want)
