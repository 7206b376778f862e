"""CPU: the stream-K work list of the CTA-pair GEMM (csrc/gemm2_tc.cu, sk_schedule), computed by the library's own
schedule code through the mtt_debug_streamk_schedule hook -- no GPU, no kernel. Checked invariants are the ones the
kernel's protocol relies on: every k-block of every tile is computed exactly once; a pair contributes (a piece that
starts inside its tile) at most once and only as its FIRST piece, so a contribution never waits on anything; the
pairs that finish a tile follow its owner consecutively; no pair of a split launch is empty."""
import ctypes as C

import pytest


def _schedule(L, tiles, k_iters, pairs, policy=2):
    L.mtt_set_gemm_streamk(policy)          # 2 = split whenever legal (the schedule under test); 1 = automatic (default)
    try:
        return _pieces(L, tiles, k_iters, pairs)
    finally:
        L.mtt_set_gemm_streamk(1)


def _pieces(L, tiles, k_iters, pairs):
    out = []
    for p in range(pairs):
        buf = (C.c_int32 * (3 * 512))()
        n = L.mtt_debug_streamk_schedule(tiles, k_iters, pairs, p, buf, 512)
        assert n <= 512
        out.append([(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(n)])
    return out


@pytest.mark.parametrize("pairs", [74, 66, 2])
@pytest.mark.parametrize("tiles,k_iters", [(204, 16), (272, 16), (68, 64), (108, 16), (20, 64), (12, 37), (74, 16),
                                           (75, 16), (10, 144), (3, 4), (1, 64), (150, 1), (100, 9), (512, 50), (73, 5)])
def test_streamk_schedule_covers_every_k_block_once(tiles, k_iters, pairs):
    import mtt_b200  # noqa: F401
    from mtt_b200 import lib

    L = lib.load()
    sched = _schedule(L, tiles, k_iters, pairs)
    owner, cover, split = {}, {}, False
    for p, pieces in enumerate(sched):
        for i, (t, k0, k1) in enumerate(pieces):
            assert 0 <= t < tiles and 0 <= k0 < k1 <= k_iters, (p, pieces)
            if k0 > 0:
                assert i == 0, "a contribution must be the pair's first piece"
                split = True
            else:
                assert t not in owner
                owner[t] = p
            for k in range(k0, k1):
                assert (t, k) not in cover, "k-block computed twice"
                cover[(t, k)] = p
    assert len(cover) == tiles * k_iters
    assert set(owner) == set(range(tiles))
    if split:
        assert all(sched), "a split launch must give every pair work (the owner waits on its successors)"
        for t in range(tiles):       # the pairs that finish a tile are the owner's immediate successors, in k order
            ps = [cover[(t, k)] for k in range(k_iters)]
            assert ps == sorted(ps) and ps[0] == owner[t]
            assert sorted(set(ps)) == list(range(ps[0], ps[-1] + 1))
        load = [sum(k1 - k0 for _, k0, k1 in pieces) for pieces in sched]
        assert max(load) - min(load) <= k_iters // 4 + 1 or max(load) <= -(-tiles * k_iters // pairs) + 1


def test_streamk_schedule_backbone_shapes_are_balanced():
    """cfg4 bs 4 (M = 4116): qkv / fc1 / fc2 on 74 pairs finish within one k-block of the ideal share."""
    import mtt_b200  # noqa: F401
    from mtt_b200 import lib

    L = lib.load()
    for tiles, k_iters, plain in [(204, 16, 48), (272, 16, 64), (68, 64, 64)]:
        load = [sum(k1 - k0 for _, k0, k1 in pieces) for pieces in _schedule(L, tiles, k_iters, 74)]
        assert max(load) <= tiles * k_iters / 74 + 1 < plain


def test_streamk_automatic_policy():
    """Default policy: only single-partial-round, long-K problems that leave at least half of the pairs idle are split
    (profiles/r3_streamk.md): fc2 at batch 1 (20 tiles x 64 k-blocks) and the weight-gradient GEMM of proj (16 x 65) are,
    the multi-round backbone GEMMs of cfg4 bs 4, the 68-tile fc2 and the 48-tile dW of qkv (measured slower) are not."""
    import mtt_b200  # noqa: F401
    from mtt_b200 import lib

    L = lib.load()

    def split(tiles, k_iters):
        return any((k0, k1) != (0, k_iters) for pieces in _schedule(L, tiles, k_iters, 74, policy=1) for _, k0, k1 in pieces)

    assert split(16, 65) and split(20, 64) and split(12, 37)
    assert not split(204, 16) and not split(272, 16) and not split(68, 64) and not split(64, 65) and not split(48, 65)
    assert not split(20, 16)


def test_streamk_schedule_invariants_on_random_problems():
    """300 random (tiles, k-blocks, pairs) triples under the split-whenever-legal policy: exact cover, contributions only as
    first pieces, no empty pair in a split launch, per-pair load within one k-block of the ideal share for the split part."""
    import random

    import mtt_b200  # noqa: F401
    from mtt_b200 import lib

    L = lib.load()
    rng = random.Random(7)
    n_split = 0
    for _ in range(300):
        pairs = rng.choice([74, 74, 66, 72, 33, 8])
        tiles = rng.randint(1, 6 * pairs)
        k_iters = rng.choice([1, 2, 4, 9, 16, 37, 48, 64, 65, 144])
        sched = _schedule(L, tiles, k_iters, pairs)
        cover, split = set(), False
        for p, pieces in enumerate(sched):
            for i, (t, k0, k1) in enumerate(pieces):
                assert 0 <= t < tiles and 0 <= k0 < k1 <= k_iters
                assert k0 == 0 or i == 0
                split |= (k0, k1) != (0, k_iters)
                for k in range(k0, k1):
                    assert (t, k) not in cover
                    cover.add((t, k))
        assert len(cover) == tiles * k_iters
        if split:
            n_split += 1
            assert all(sched)
            r = tiles % pairs
            whole = tiles // pairs                      # whole tiles every pair runs after its stream-K range
            load = [sum(k1 - k0 for _, k0, k1 in pieces) for pieces in sched]
            assert max(load) <= whole * k_iters + -(-r * k_iters // pairs), (tiles, k_iters, pairs)
    assert n_split > 50
