"""CPU, world_size 2 over gloo: the N>1 plumbing used by bench.py (barrier, max-over-ranks time, batch
sharding, output gather). The forward has no data-path collective, so this is all there is."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, torch
    sys.path.insert(0, %r)
    import mtt_b200
    from mtt_b200 import dist as D
    rank, world, local = D.setup("gloo")
    assert world == 2
    D.barrier(world)
    t = D.max_over_ranks(10.0 + rank, world)
    assert t == 11.0, t
    a, b = D.shard_batch(7, rank, world)
    assert (a, b) == ((0, 4) if rank == 0 else (4, 7))
    # the shards shard_batch actually produces: 7 images -> 4 + 3 (uneven)
    n = b - a
    out = {"semseg": torch.arange(a, b, dtype=torch.float32)[:, None].repeat(1, 3), "depth": torch.full((n, 1), 10.0 + rank)}
    g = D.gather_outputs(out, world)
    if rank == 0:
        assert g["semseg"].shape == (7, 3) and torch.equal(g["semseg"][:, 0], torch.arange(7.0))
        assert g["depth"].shape == (7, 1) and g["depth"][:4].eq(10).all() and g["depth"][4:].eq(11).all()
        print("OK")
    D.teardown(world)
""") % ROOT


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29671", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "OK" in out.stdout


TRAIN_WORKER = textwrap.dedent("""
    import os, sys, torch
    sys.path.insert(0, %r)
    sys.path.insert(0, os.path.join(%r, "tests"))
    import pytest
    import mtt_b200, emul_ops
    from mtt_b200 import dist as D, taskprompter as TP
    from mtt_b200.train import TrainStep
    from oracle import configs, taskprompter_ref as TPR
    rank, world, local = D.setup("gloo")
    emul_ops.install(pytest.MonkeyPatch())
    cfg = configs.taskprompter("tp_tiny1")
    cfg["drop_path_rate"] = 0.0
    sd = TPR.init_state_dict(cfg, seed=7)

    def step(pg, x, gout):
        m = TP.build_from_config(cfg, use_graph=False)
        m.load_state_dict(sd)
        ts = TrainStep(m, process_group=pg, bucket_mb=0.25, max_norm=5.0, lr=1e-3)
        ts.zero_grad()
        with torch.no_grad():
            out = ts.forward(x)
            ts.backward(gout)
            g = ts.grads.flat.clone()
            ts.optimizer_step()
        return out, g, ts.params.flat.clone(), m

    g_ = torch.Generator().manual_seed(3)
    x = torch.randn(4, 3, *cfg["img_size"], generator=g_)
    gout = {t: torch.randn(4, cfg["num_output"][t], *cfg["img_size"], generator=g_) for t in cfg["tasks"]}
    sl = slice(2 * rank, 2 * rank + 2)
    out2, g2, p2, m2 = step(torch.distributed.group.WORLD, x[sl], {t: v[sl] for t, v in gout.items()})
    out1, g1, p1, m1 = step(None, x, gout)           # the whole batch in one process
    for t in cfg["tasks"]:                           # SyncBatchNorm: each rank's outputs are rows of the full-batch outputs
        assert torch.allclose(out2[t], out1[t][sl], rtol=1e-4, atol=1e-5), t
    # the all-reduced gradient arena (a SUM over ranks) equals the full-batch gradients
    err = (g2 - g1).norm() / g1.norm()
    assert err < 1e-4, float(err)
    # running statistics were updated from the global batch; clip + Adam use the averaged gradient
    for (k1, b1), (k2, b2) in zip(m1.state_dict().items(), m2.state_dict().items()):
        if "running_" in k1:
            assert torch.allclose(b1, b2, rtol=1e-4, atol=1e-6), k1
    if rank == 0:
        print("OK", float(err))
    D.teardown(world)
""") % (ROOT, ROOT)


def test_two_rank_training_step_gloo(tmp_path):
    """world_size 2 over gloo, kernels emulated: SyncBatchNorm statistics and the bucketed gradient all-reduce of
    mtt_b200/train.py reproduce the single-process step on the concatenated batch."""
    script = tmp_path / "tw.py"
    script.write_text(TRAIN_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29673", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    assert "OK" in out.stdout
