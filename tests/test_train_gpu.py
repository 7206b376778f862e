"""-m gpu: the training-step kernels (csrc/train_ops.cu) one by one against the torch restatements of tests/emul_ops.py
(whose adjoints come from torch autograd of the forward restatements), then the whole training step through
libmtt_sm100.so against the reference's train-mode goldens (tests/golden/train_*.pt, oracle/make_golden.py train):
train-mode outputs, losses (mtt_b200.losses kernels), every parameter gradient, BatchNorm running statistics,
clip_grad_norm_ + Adam."""
import os

import pytest
import torch

from oracle import configs
from oracle import taskprompter_ref as TPR

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


class _Emu:
    def __getitem__(self, name):
        import emul_ops
        from mtt_b200 import ops

        def call(*a, **k):
            mp = pytest.MonkeyPatch()
            emul_ops.install(mp)
            try:
                return getattr(ops, name)(*a, **k)
            finally:
                mp.undo()
        return call


@pytest.fixture(scope="module")
def both(cuda_dev):
    import mtt_b200  # noqa: F401
    from mtt_b200 import ops
    return ops, _Emu()


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def to_dev(ops, sp, dev):
    d = ops.Split(sp.rows, sp.cols, dev, sp.nsplit, ld=sp.ld, zero=True)
    d.buf.copy_(sp.buf)
    return d


def cpu_split(emu, x):
    return emu["split_f32"](x, 2)


def test_colsum_rows_and_mapping(both, cuda_dev):
    ops, emu = both
    torch.manual_seed(0)
    x = torch.randn(1000, 70)
    for kw in (dict(), dict(rows=12 * 5, in_group=5, src_group=83, src_offset=3)):
        a, b = torch.randn(70), None
        b = a.clone()
        ops.colsum(x.to(cuda_dev), (ad := a.to(cuda_dev)), accumulate=True, **kw)
        emu["colsum"](x, b, accumulate=True, **kw)
        assert relerr(ad, b) < 1e-5
    wide = torch.randn(3, 5000)                       # few rows, many columns (pos_embed gradient over the batch)
    o = torch.empty(5000, device=cuda_dev)
    ops.colsum(wide.to(cuda_dev), o)
    assert relerr(o, wide.sum(0)) < 1e-6


@pytest.mark.parametrize("rows,cols", [(37, 128), (1030, 1024), (5, 350)])
def test_layernorm_bwd(both, cuda_dev, rows, cols):
    ops, emu = both
    torch.manual_seed(1)
    x, dy, g = torch.randn(rows, cols) * 2 + 0.3, torch.randn(rows, cols), torch.randn(cols)
    dx0 = torch.randn(rows, cols)
    res = []
    for dev, f in ((cuda_dev, ops), ("cpu", None)):
        dx, dg, db = dx0.clone().to(dev), torch.ones(cols, device=dev), torch.ones(cols, device=dev)
        args = (x.to(dev), dy.to(dev), g.to(dev), 1e-6, dx, dg, db)
        (ops.layernorm_bwd if f else emu["layernorm_bwd"])(*args, accumulate_dx=True)
        res.append((dx, dg, db))
    for a, b in zip(*res):
        assert relerr(a, b) < 2e-5


def test_act_split_and_bwd(both, cuda_dev):
    ops, emu = both
    torch.manual_seed(2)
    pre, dy = torch.randn(77, 52) * 2, torch.randn(77, 52)
    for act in (0, 1, 2):
        s_g = ops.act_split(pre.to(cuda_dev), act)
        s_c = emu["act_split"](pre, act)
        assert relerr(s_g.float(), s_c.float()) < 1e-5
        dg, dc = torch.empty(77, 52, device=cuda_dev), torch.empty(77, 52)
        ops.act_bwd(pre.to(cuda_dev), dy.to(cuda_dev), act, dg)
        emu["act_bwd"](pre, dy, act, dc)
        assert relerr(dg, dc) < 1e-5
    d2 = dy.to(cuda_dev).clone()
    ops.act_bwd(pre.to(cuda_dev), d2, 1, d2)           # in place
    emu["act_bwd"](pre, dy, 1, dc)
    assert relerr(d2, dc) < 1e-5


def test_axpy_rows(both, cuda_dev):
    ops, _ = both
    torch.manual_seed(3)
    base, src, sc = torch.randn(50, 33), torch.randn(50, 33), torch.randn(50)
    d = torch.empty(50, 33, device=cuda_dev)
    ops.axpy_rows(base.to(cuda_dev), src.to(cuda_dev), sc.to(cuda_dev), d)
    assert relerr(d, base + sc[:, None] * src) < 1e-6
    ops.axpy_rows(None, src.to(cuda_dev), None, d)
    assert relerr(d, src) == 0.0


def test_transpose_planes(both, cuda_dev):
    ops, emu = both
    torch.manual_seed(4)
    a = cpu_split(emu, torch.randn(3 * 41, 72))
    ag = to_dev(ops, a, cuda_dev)
    for kw in (dict(B=3, R=41, Ccols=72), dict(B=3, R=5, Ccols=70, in_batch_rows=41),
               dict(B=3, R=5, Ccols=72, in_batch_rows=41, side_by_side=True), dict(B=1, R=123, Ccols=72)):
        g, c = ops.transpose_planes(ag, **kw), emu["transpose_planes"](a, **kw)
        assert (g.rows, g.cols, g.ld) == (c.rows, c.cols, c.ld)
        assert torch.equal(g.buf.cpu()[:, :, :c.cols], c.buf[:, :, :c.cols])


@pytest.mark.parametrize("act", [0, 1])
def test_batchnorm_train(both, cuda_dev, act):
    ops, emu = both
    torch.manual_seed(5)
    rows, cols = 3000, 45
    x, dy = torch.randn(rows, cols) * 1.5 + 0.2, torch.randn(rows, cols)
    gam, bet = torch.rand(cols) + 0.5, torch.randn(cols)
    out = []
    for dev, real in ((cuda_dev, True), ("cpu", False)):
        f = (lambda n: getattr(ops, n)) if real else (lambda n: emu[n])
        sums, mr = torch.empty(2 * cols, device=dev), torch.empty(2 * cols, device=dev)
        rm, rv = torch.zeros(cols, device=dev), torch.ones(cols, device=dev)
        xd, dyd, g, b = x.to(dev), dy.to(dev), gam.to(dev), bet.to(dev)
        f("bn_stats")(xd, sums)
        f("bn_finalize")(sums, rows, 1e-5, 0.1, mr, rm, rv)
        y = torch.empty(rows, cols, device=dev)
        ys = ops.Split(rows, cols, dev, 2, zero=True)
        f("bn_act")(xd, mr, g, b, act, out_f32=y, out_split=ys)
        s2 = torch.empty(2 * cols, device=dev)
        f("bn_bwd_reduce")(xd, dyd, mr, g, b, act, s2)
        dx = torch.empty(rows, cols, device=dev)
        f("bn_bwd_apply")(xd, dyd, mr, g, b, act, s2, rows, dx)
        out.append((sums, mr, rm, rv, y, ys.float(), s2, dx))
    for i, (a, b) in enumerate(zip(*out)):
        assert relerr(a, b) < 5e-5, i
    # against nn.BatchNorm2d in train mode + autograd
    bn = torch.nn.BatchNorm2d(cols).train()
    with torch.no_grad():
        bn.weight.copy_(gam)
        bn.bias.copy_(bet)
    xi = x.t().reshape(1, cols, rows, 1).clone().requires_grad_(True)
    z = bn(xi)
    z = torch.nn.functional.gelu(z) if act == 1 else z
    z.backward(dy.t().reshape(1, cols, rows, 1))
    assert relerr(out[0][7], xi.grad.reshape(cols, rows).t()) < 5e-4
    assert relerr(out[0][2], bn.running_mean) < 1e-5 and relerr(out[0][3], bn.running_var) < 1e-5


def test_attn_softmax_bwd(both, cuda_dev):
    ops, emu = both
    torch.manual_seed(6)
    B, H, N, T = 3, 2, 77, 3          # two query tiles, the second ragged; odd N
    BH, ld = B * H, 80
    S, dP, dr = torch.randn(BH * N, ld) * 3, torch.randn(BH * N, ld), torch.randn(BH, T, N)
    # delta from its definition rowdot(dO, O) with O = P V, dP = dO V^T (so the kernel output must equal the softmax adjoint)
    V, dO = torch.randn(BH, N, 64), torch.randn(BH, N, 64)
    P = (S[:, :N].reshape(BH, N, N) * 0.125).softmax(-1)
    dP[:, :N] = (dO @ V.transpose(1, 2)).reshape(BH * N, N)
    O = (P @ V).reshape(B, H, N, 64).permute(0, 2, 1, 3).reshape(B * N, H * 64)
    dOf = dO.reshape(B, H, N, 64).permute(0, 2, 1, 3).reshape(B * N, H * 64).contiguous()
    o_s = cpu_split(emu, O)
    res = []
    for dev, real in ((cuda_dev, True), ("cpu", False)):
        f = (lambda n: getattr(ops, n)) if real else (lambda n: emu[n])
        s, d = S.clone().to(dev), dP.clone().to(dev)
        delta = torch.empty(BH * N, device=dev)
        f("attn_delta")(dOf.to(dev), to_dev(ops, o_s, dev), delta, B=B, N=N, H=H, head_dim=64)
        ds, pt, dst = (ops.Split(BH * N, ld, dev, 2, zero=True) for _ in range(3))
        f("attn_softmax_bwd")(s, d, delta, BH=BH, N=N, scale=0.125, d_raw=dr.to(dev), T=T, ds=ds, pt=pt, dst=dst)
        assert torch.equal(s.cpu(), S) and torch.equal(d.cpu(), dP)          # inputs are read only
        res.append((delta, ds.float()[:, :N], pt.float()[:, :N], dst.float()[:, :N]))
    for a, b in zip(*res):
        assert relerr(a, b) < 3e-5
    # against torch autograd of softmax
    s_ = S[:, :N].reshape(BH, N, N).clone().requires_grad_(True)
    (s_ * 0.125).softmax(-1).backward(dP[:, :N].reshape(BH, N, N))
    want = s_.grad.clone()
    want[:, :T] += dr
    assert relerr(res[0][1], want.reshape(BH * N, N)) < 1e-4


@pytest.mark.parametrize("nchw", [True, False])
def test_bilinear_bwd(both, cuda_dev, nchw):
    ops, emu = both
    torch.manual_seed(7)
    B, h, w, C, H2, W2 = 2, 5, 7, 9, 20, 23
    dy = torch.randn(B, C, H2, W2) if nchw else torch.randn(B * H2 * W2, C)
    base = torch.randn(B * h * w, C)
    for acc in (False, True):
        g, c = base.clone().to(cuda_dev), base.clone()
        ops.bilinear_bwd(dy.to(cuda_dev), nchw=nchw, B=B, h=h, w=w, Cdim=C, H2=H2, W2=W2, dx=g, accumulate=acc)
        emu["bilinear_bwd"](dy, nchw=nchw, B=B, h=h, w=w, Cdim=C, H2=H2, W2=W2, dx=c, accumulate=acc)
        assert relerr(g, c) < 1e-5


@pytest.mark.parametrize("nh", [1, 2])
def test_gate_and_chan_logits_bwd(both, cuda_dev, nh):
    ops, emu = both
    torch.manual_seed(8)
    B, T, H, dh, gh, gw = 2, 3, 2, 64, 4, 6
    C, P = H * dh, gh * gw
    N = T + P
    x, lg, rc = torch.randn(B * N, C), torch.randn(B, H, T, N), torch.randn(B, T, C, nh, nh)
    dys, dyc = torch.randn(B * P, C), torch.randn(B * P, C)
    res = []
    for dev, real in ((cuda_dev, True), ("cpu", False)):
        dx, dl, drc = torch.ones(B * N, C, device=dev), torch.ones(B, H, T, N, device=dev), torch.ones(B, T, C, nh, nh, device=dev)
        (ops.gate_bwd if real else emu["gate_bwd"])(x.to(dev), N, T, lg.to(dev), rc.to(dev), 1, dys.to(dev), dyc.to(dev), dx, dl,
                                                    drc, B=B, T=T, N=N, H=H, Cdim=C, gh=gh, gw=gw, nh=nh, nw=nh)
        res.append((dx, dl, drc))
    for a, b in zip(*res):
        assert relerr(a, b) < 2e-5
    cp, d_rc = torch.randn(B * T, P), torch.randn(B, T, C, nh, nh)
    xn = cpu_split(emu, x)
    res = []
    for dev, real in ((cuda_dev, True), ("cpu", False)):
        dcp, dxn = torch.empty(B * T, P, device=dev), torch.ones(B * N, C, device=dev)
        (ops.chan_logits_bwd if real else emu["chan_logits_bwd"])(d_rc.to(dev), cp.to(dev), to_dev(ops, xn, dev), dcp, dxn, B=B,
                                                                  N=N, T=T, Cdim=C, gh=gh, gw=gw, nh=nh, nw=nh)
        res.append((dcp, dxn))
    for a, b in zip(*res):
        assert relerr(a, b) < 2e-5


def test_ctr_bwd(both, cuda_dev):
    ops, emu = both
    torch.manual_seed(9)
    B, T, H, P, f = 2, 3, 4, 24, 20
    N, M = T + P, 2 * 24
    dnew, Fm, lg = torch.randn(T, M, f), torch.randn(T, M, f), torch.randn(B, H, T, N)
    w0, b0, w2 = torch.randn(T, H, H), torch.randn(T, H), torch.randn(T, H)
    res = []
    for dev, real in ((cuda_dev, True), ("cpu", False)):
        outs = [torch.zeros(B, H, T, N, device=dev), torch.zeros(T, H, H, device=dev), torch.zeros(T, H, device=dev),
                torch.zeros(T, H, device=dev), torch.zeros(T, device=dev)]
        (ops.ctr_bwd if real else emu["ctr_bwd"])(dnew.to(dev), Fm.to(dev), lg.to(dev), w0.to(dev), b0.to(dev), w2.to(dev), *outs,
                                                  T=T, M=M, Cdim=f, ld=f, rows_per_batch=P, B=B, H=H, N=N)
        res.append(outs)
    for a, b in zip(*res):
        assert relerr(a, b) < 5e-5


def test_im2col_transposed_operands(both, cuda_dev):
    ops, emu = both
    torch.manual_seed(10)
    B, H, W, C = 2, 5, 7, 37
    x = torch.randn(B * H * W, C)
    g, c = ops.im2col3x3_t(x.to(cuda_dev), B=B, H=H, W=W, Cdim=C), emu["im2col3x3_t"](x, B=B, H=H, W=W, Cdim=C)
    assert torch.equal(g.buf.cpu()[:, :, :c.cols], c.buf[:, :, :c.cols])
    img = torch.randn(2, 3, 32, 48)
    g, c = ops.im2col_patch_t(img.to(cuda_dev), 16), emu["im2col_patch_t"](img, 16)
    assert torch.equal(g.buf.cpu()[:, :, :c.cols], c.buf[:, :, :c.cols])


def test_sumsq_and_adam(both, cuda_dev):
    ops, emu = both
    torch.manual_seed(11)
    n = 100003
    p, g = torch.randn(n), torch.randn(n) * 3
    res = []
    for dev, real in ((cuda_dev, True), ("cpu", False)):
        f = (lambda k: getattr(ops, k)) if real else (lambda k: emu[k])
        pd, gd, m, v, ss = p.clone().to(dev), g.to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros((), device=dev)
        for step in (1, 2):
            f("sumsq")(gd, ss)
            f("adam_step")(pd, gd, m, v, lr=1e-3, weight_decay=1e-2, step=step, gnorm_sq=ss, max_norm=10.0, grad_scale=0.5)
        res.append((ss, pd, m, v))
    for a, b in zip(*res):
        assert relerr(a, b) < 2e-5
    # torch.optim.Adam + clip_grad_norm_ on the same numbers
    q = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([q], lr=1e-3, weight_decay=1e-2)
    for _ in range(2):
        q.grad = g.clone() * 0.5
        torch.nn.utils.clip_grad_norm_([q], 10.0)
        opt.step()
    assert relerr(res[0][1], q.detach()) < 2e-5


# ---- the whole step ------------------------------------------------------------------------------------------------------
def _run_step(name, cuda_dev):
    import mtt_b200  # noqa: F401
    from mtt_b200 import losses
    from test_train import _build, _check_grads, _fixture

    fx = _fixture(name)
    cfg, model, ts = _build(fx, cuda_dev)
    ts.zero_grad()
    with torch.no_grad():
        out = ts.forward(fx["x"].to(cuda_dev), drop_rand=fx["masks"])
    s = fx["out_stride"]
    for t in cfg["tasks"]:
        ref = fx["out"][t]
        err = (out[t][..., ::s, ::s].cpu() - ref).norm() / ref.norm()
        assert err < 2e-4, f"train-mode forward {t}: rel-L2 {err:.3e}"
    p = dict(TASKS=dict(NAMES=list(cfg["tasks"])), edge_w=0.95, ignore_index=255, ignore_invalid_area_depth=True,
             loss_kwargs=dict(loss_weights={t: fx["weights"][t] for t in cfg["tasks"]}))
    crit = losses.get_criterion(p)
    leaves = {t: out[t].detach().requires_grad_(True) for t in cfg["tasks"]}
    lab = {t: v.to(cuda_dev) for t, v in fx["labels"].items()}
    loss = crit(leaves, lab, tasks=cfg["tasks"])
    for k, want in fx["losses"].items():
        assert abs(float(loss[k].detach()) - want) <= 2e-4 * max(1.0, abs(want)), (k, float(loss[k].detach()), want)
    grads = torch.autograd.grad(loss["total"], [leaves[t] for t in cfg["tasks"]])
    with torch.no_grad():
        ts.backward(dict(zip(cfg["tasks"], grads)))
    torch.cuda.synchronize()
    _check_grads(fx, ts, 2e-3)
    sdm = model.state_dict()
    for k, want in fx["running"].items():
        assert torch.allclose(sdm[k].cpu(), want, rtol=2e-4, atol=1e-6), k
    with torch.no_grad():
        ts.optimizer_step()
    assert abs(float(ts.gnorm.sqrt()) - fx["total_norm"]) <= 2e-3 * fx["total_norm"]
    before = TPR.init_state_dict(cfg, seed=fx["seed"])
    for k, want in fx["param_after"].items():
        if fx["grad_norm"][k] < 1e-4 * fx["total_norm"]:
            continue
        step_ref, step_got = want - before[k], ts.P_(k).detach().cpu() - before[k]
        close = (step_got - step_ref).abs() <= 1e-3 * step_ref.abs() + 2e-7
        assert close.float().mean() > 0.98, (k, close.float().mean().item())
    return ts


@pytest.mark.parametrize("name", ["tp_tiny", "tp_tiny1"])
def test_training_step_matches_reference(cuda_dev, name):
    _run_step(name, cuda_dev)


def test_training_step_full_width(cuda_dev):
    """tp_cfg4_d4 (ViT-L width, 512 x 512, 5 PASCAL tasks, N = 1029 tokens, batch 2).

    (a) Against the reference fixture: train-mode outputs and losses, and every parameter gradient's NORM. The L1-type
    normals loss has a sign() in its gradient: the pixels whose prediction sits within the forward tolerance (1e-4) of
    the label flip it, so ELEMENTWISE agreement with gradients the reference computed from ITS outputs stops at ~1e-2
    at this size (1.5 M loss terms) -- that is a property of the loss, not of the reverse pass.
    (b) The reverse pass itself, elementwise on every parameter: the same d loss / d prediction tensors go through
    TrainStep.backward and through torch autograd of the train-mode restatement (pinned to the reference on CPU by
    test_train.py::test_train_mode_restatement_is_pinned_to_the_reference) on the GPU in fp32 with TF32 off."""
    import mtt_b200  # noqa: F401
    from mtt_b200 import losses
    from test_train import _build, _fixture, _oracle_grads

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    fx = _fixture("tp_cfg4_d4")
    cfg, model, ts = _build(fx, cuda_dev)
    ts.zero_grad()
    with torch.no_grad():
        out = ts.forward(fx["x"].to(cuda_dev), drop_rand=fx["masks"])
    s = fx["out_stride"]
    for t in cfg["tasks"]:
        ref = fx["out"][t]
        err = (out[t][..., ::s, ::s].cpu() - ref).norm() / ref.norm()
        assert err < 2e-4, f"train-mode forward {t}: rel-L2 {err:.3e}"
    p = dict(TASKS=dict(NAMES=list(cfg["tasks"])), edge_w=0.95, ignore_index=255, ignore_invalid_area_depth=True,
             loss_kwargs=dict(loss_weights={t: fx["weights"][t] for t in cfg["tasks"]}))
    leaves = {t: out[t].detach().requires_grad_(True) for t in cfg["tasks"]}
    loss = losses.get_criterion(p)(leaves, {t: v.to(cuda_dev) for t, v in fx["labels"].items()}, tasks=cfg["tasks"])
    for k, want in fx["losses"].items():
        assert abs(float(loss[k].detach()) - want) <= 2e-4 * max(1.0, abs(want)), (k, float(loss[k].detach()), want)
    grads = dict(zip(cfg["tasks"], torch.autograd.grad(loss["total"], [leaves[t] for t in cfg["tasks"]])))
    with torch.no_grad():
        ts.backward(grads)
    torch.cuda.synchronize()
    floor = 1e-4 * fx["total_norm"]
    for k, want in fx["grad_norm"].items():                                            # (a)
        err = abs(ts.G_(k).norm().item() - want) / max(want, floor)
        assert err < (5e-2 if ".ctr_attn_conv." in k else 1e-2), (k, err)
    _, _, og = _oracle_grads(fx, cfg, cuda_dev, fx["x"], grad_out=grads)                # (b)
    bad = []
    for k, ref in og.items():
        err = (ts.G_(k) - ref).norm().item() / max(ref.norm().item(), floor)
        if not err < (1e-2 if ".ctr_attn_conv." in k else 2e-3):
            bad.append((k, err))
    assert not bad, f"{len(bad)} of {len(og)} parameter gradients off: {sorted(bad, key=lambda kv: -kv[1])[:8]}"


def test_torch_facing_step_delivers_the_same_gradients(cuda_dev):
    """TrainStep.apply: model outputs attached to autograd; loss.backward() fills param.grad (through autograd's own
    accumulation) with the gradients of the native reverse pass."""
    import mtt_b200  # noqa: F401
    from mtt_b200 import losses
    from test_train import _build, _fixture

    fx = _fixture("tp_tiny1")
    cfg, model, ts = _build(fx, cuda_dev)
    torch.manual_seed(5)
    out = ts.apply(fx["x"].to(cuda_dev))
    p = dict(TASKS=dict(NAMES=list(cfg["tasks"])), edge_w=0.95, ignore_index=255, ignore_invalid_area_depth=True,
             loss_kwargs=dict(loss_weights={t: fx["weights"][t] for t in cfg["tasks"]}))
    loss = losses.get_criterion(p)(out, {t: v.to(cuda_dev) for t, v in fx["labels"].items()}, tasks=cfg["tasks"])
    loss["total"].backward()
    got = {n: prm.grad.detach().clone() for n, prm in model.named_parameters()}
    assert all(g is not None and torch.isfinite(g).all() for g in got.values())
    nz = sum(float(g.abs().sum()) > 0 for g in got.values())
    assert nz >= len(got) - 2, f"only {nz} of {len(got)} parameters received a gradient"
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in got.values()))
    assert 0.2 * fx["total_norm"] < float(total) < 5 * fx["total_norm"]      # DropPath masks differ from the fixture's


def test_graph_replay_equals_eager_steps(cuda_dev):
    """TrainStep(use_graph=True) captures forward + criterion + reverse pass once and replays it: after three steps on
    changing batches the parameters equal those of the eagerly launched steps (DropPath off: both draw nothing)."""
    import mtt_b200  # noqa: F401
    from mtt_b200 import losses, taskprompter as TP
    from mtt_b200.train import TrainStep
    from oracle.make_golden import synthetic_labels

    cfg = configs.taskprompter("tp_tiny1")
    cfg["drop_path_rate"] = 0.0
    sd = TPR.init_state_dict(cfg, seed=5)
    p = dict(TASKS=dict(NAMES=list(cfg["tasks"])), edge_w=0.95, ignore_index=255, ignore_invalid_area_depth=True,
             loss_kwargs=dict(loss_weights={"semseg": 1.0, "edge": 50.0}))
    crit = losses.get_criterion(p)
    g = torch.Generator().manual_seed(9)
    batches = [(torch.randn(2, 3, *cfg["img_size"], generator=g), synthetic_labels(cfg["tasks"], cfg["num_output"], 2,
                                                                                  *cfg["img_size"], g)) for _ in range(3)]
    res = []
    for use_graph in (False, True):
        model = TP.build_from_config(cfg, use_graph=False)
        model.load_state_dict(sd)
        model.to(cuda_dev)
        ts = TrainStep(model, lr=1e-3, use_graph=use_graph)
        losses_seen = []
        with torch.no_grad():
            for x, y in batches:
                out = ts.step(x.to(cuda_dev), {t: v.to(cuda_dev) for t, v in y.items()}, crit)
                losses_seen.append(float(out["total"]))
        torch.cuda.synchronize()
        res.append((ts.grads.flat.clone(), losses_seen, {k: v.clone() for k, v in model.state_dict().items() if "running_" in k}))
    (pe, le, be), (pg, lg, bg) = res
    assert all(abs(a - b) <= 1e-4 * max(1.0, abs(a)) for a, b in zip(le, lg)), (le, lg)
    assert le[0] != le[1]                                        # the replays really saw different batches
    # the gradients of the third step (parameters themselves are not compared: a bias in front of a BatchNorm has a zero
    # gradient up to atomics-order noise, and Adam turns that noise into +-lr steps)
    assert relerr(pg, pe) < 1e-3
    for k in be:
        # the running MEAN in front of a BatchNorm contains the conv bias, whose zero gradient is noise that Adam turns into
        # +-lr steps (lr = 1e-3 here, 2 steps before the last forward)
        assert torch.allclose(be[k], bg[k], rtol=2e-3, atol=5e-3), k
