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
