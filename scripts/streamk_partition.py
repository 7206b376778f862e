"""Stream-K tail for the CTA-pair GEMM: the work partition as plain integer logic, with
its invariants checked and its effect modelled -- a design prototype, nothing here is used by the kernels yet.

Scheme ("data-parallel + stream-K hybrid"): with T output tiles, P CTA pairs and K k-iterations per tile, the first
floor(T / P) * P tiles are processed whole, one after the other, as today. The R = T mod P left-over tiles are cut
along K: pair c processes k-iterations [c * R * K // P, (c + 1) * R * K // P) of the linearised (tile, k) space. A
pair's share is shorter than one tile, so it touches at most two tiles. The segment that contains a tile's LAST
k-iteration is that tile's finisher: it waits for the other segments' fp32 partial accumulators (written to a
workspace slot, published with a release flag), adds them and runs the normal epilogue.

    python scripts/streamk_partition.py            # checks + table for the cfg4 backbone shapes
"""
import math

PAIRS = 74


def partition(T, K, P=PAIRS):
    """-> (dp_tiles per pair, {pair: [(tile, k0, k1, finisher, n_partials_to_wait, partial_slot)]})."""
    full = T // P
    R = T - full * P
    segs = {c: [] for c in range(P)}
    if R == 0:
        return full, segs
    contributors = {}
    for c in range(P):
        lo, hi = c * R * K // P, (c + 1) * R * K // P
        while lo < hi:
            t = lo // K
            k0 = lo - t * K
            k1 = min(K, k0 + (hi - lo))
            contributors.setdefault(t, []).append((c, k0, k1))
            lo += k1 - k0
    for t, parts in contributors.items():
        slot = 0
        for (c, k0, k1) in parts:
            fin = k1 == K
            segs[c].append((full * P + t, k0, k1, fin, len(parts) - 1 if fin else 0, None if fin else slot))
            if not fin:
                slot += 1
    return full, segs


def check(T, K, P=PAIRS):
    full, segs = partition(T, K, P)
    R = T - full * P
    cover = {}
    for c, ss in segs.items():
        assert len(ss) <= 2, "a pair's stream-K share spans at most two tiles"
        for (t, k0, k1, fin, nwait, slot) in ss:
            assert 0 <= k0 < k1 <= K
            for k in range(k0, k1):
                assert (t, k) not in cover
                cover[(t, k)] = c
    assert len(cover) == R * K, "every (tile, k) of the tail exactly once"
    for t in range(full * P, T):
        fins = [s for ss in segs.values() for s in ss if s[0] == t and s[3]]
        parts = [s for ss in segs.values() for s in ss if s[0] == t and not s[3]]
        assert len(fins) == 1 and fins[0][4] == len(parts)
        assert sorted(s[5] for s in parts) == list(range(len(parts)))
    loads = [sum(k1 - k0 for (_, k0, k1, _, _, _) in ss) for ss in segs.values()]
    assert not loads or max(loads) - min(loads) <= 1 or R == 0
    return full, R, (max(loads) if loads else 0)


def model(M, N, K_elems, name):
    tiles = math.ceil(M / 256) * math.ceil(N / 256)
    K = math.ceil(K_elems / 64)
    full, R, share = check(tiles, K)
    now = math.ceil(tiles / PAIRS) * K                      # k-iterations on the critical path today
    sk = full * K + share                                   # with the stream-K tail
    fix = 2.4 if R else 0        # reading <= 2 partial 128x256 fp32 tiles per CTA from L2 + flags: ~2 us = 2.4 k-iterations of 0.85 us
    print(f"| {name} | {tiles} | {tiles / PAIRS:.2f} | {K} | {now} | {sk} (+{fix:.1f} fix-up) | {100 * (1 - (sk + fix) / now):.1f} % |")
    return now, sk + fix


if __name__ == "__main__":
    import random
    random.seed(0)
    for _ in range(300):
        check(random.randint(1, 400), random.randint(1, 80), random.choice([1, 2, 37, 74]))
    print("invariants hold on 300 random (tiles, k-iterations, pairs) cases\n")
    print("cfg4 backbone linears, M = 4116, CTA-pair 256x256 tiles on 74 pairs (k-iteration = 64 elements of K):\n")
    print("| GEMM | tiles | waves | k-iters / tile | critical path today | with stream-K tail | saved |")
    print("|---|---:|---:|---:|---:|---|---:|")
    tot_now = tot_sk = 0.0
    for name, N, Kd, w in (("qkv", 3072, 1024, 1.0), ("fc1", 4096, 1024, 1.0), ("fc2", 1024, 4096, 1.0)):
        a, b = model(4116, N, Kd, name)
        tot_now += a
        tot_sk += b
    print(f"\nAll three: {100 * (1 - tot_sk / tot_now):.1f} % of their time = {100 * (1 - tot_sk / tot_now) * 0.45:.1f} % of the "
          "cfg4 step (they are ~45 % of it, profiles/r1s_launch_shares.md).")
