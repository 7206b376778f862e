"""Schedule model of the tcgen05 GEMM / conv launches of one forward (no GPU needed): for every mtt_gemm launch the
plan makes, the tile grid the kernel-variant heuristic picks (gemm_host.cu), the number of waves on 148 SMs / 74 CTA
pairs, and the fraction of issued MMA work that is useful (rows / columns / waves actually needed). Quantifies what a
stream-K tail can recover.

    python scripts/gemm_schedule_model.py tp_cfg4 4 > profiles/gemm_schedule_tp_cfg4.md
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

import mtt_b200  # noqa: E402,F401
from mtt_b200 import configs, ops  # noqa: E402

SMS = 148


def variant(M, N, K, conv):
    """gemm_host.cu's automatic choice: 1 = single CTA 128x128, 2 = CTA pair 256x256."""
    n256 = (N + 255) // 256 * 256
    if conv:
        k_eff = K * 9
        pair_tiles = ((M + 255) // 256) * (n256 // 256)
        return 2 if (k_eff >= 2048 and pair_tiles >= 48 and N > 256) else 1
    wide = N >= 512 and (n256 - N) * 8 <= n256
    return 2 if (wide and M > 128 and (N >= 2048 or K >= 2048)) else 1


def model(M, N, K, conv):
    v = variant(M, N, K, conv)
    tm, tn, slots = (256, 256, SMS // 2) if v == 2 else (128, 128, SMS)
    tiles_m, tiles_n = math.ceil(M / tm), math.ceil(N / tn)
    tiles = tiles_m * tiles_n
    waves = tiles / slots
    # issued MMA area vs useful area (the ragged last N tile is narrowed to a multiple of 16 by the kernels)
    n_issued = (tiles_n - 1) * tn + math.ceil((N - (tiles_n - 1) * tn) / 16) * 16
    pad = (tiles_m * tm * n_issued) / (M * N)
    quant = math.ceil(waves) / waves
    return v, tiles, waves, pad, quant


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "tp_cfg4"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    import emul_ops

    recs = []

    class Rec:
        def setattr(self, obj, attr, val):
            setattr(obj, attr, val)

    emul_ops.install(Rec())          # CPU stand-ins for every op; we only need the call sequence
    real = ops.gemm

    def spy(a, w, **kw):
        M = kw.get("M") or a.rows
        N = kw.get("N") or w.rows
        K = kw.get("K") or a.cols
        recs.append((M, N, K, kw.get("conv") is not None and kw["conv"][3] == 3))
    ops.gemm = spy
    for fn in ("layernorm", "attention", "im2col_patch", "broadcast_rows", "chan_logits", "gate_split", "ctr_weights",
               "ctr_mix", "bilinear", "bilinear_postproc", "split_rows", "layernorm_seg", "zero_insert", "dwconv3x3_s2",
               "avgpool", "invpt_attention", "bilinear_sum3"):
        if hasattr(ops, fn):
            setattr(ops, fn, lambda *a, **k: None)
    if name.startswith("tp_"):
        from mtt_b200 import taskprompter as TP
        cfg = configs.taskprompter(name)
    else:
        from mtt_b200 import invpt as TP
        cfg = configs.invpt(name)
    with torch.device("meta"):
        pass
    m = TP.build_from_config(cfg, use_graph=False).eval()
    pl = m.plan(B, torch.device("cpu"))
    pl._launch(torch.zeros(B, 3, *cfg["img_size"]))
    ops.gemm = real
    agg = {}
    for r in recs:
        agg[r] = agg.get(r, 0) + 1
    print(f"# GEMM / conv schedule of one {name} forward, batch {B} (model: scripts/gemm_schedule_model.py)\n")
    print("`flops` = algorithmic 2 M N K (x9 for 3x3 convs) summed over the launches of that shape; `padding` = issued / useful "
          "MMA area (M rounded up to the tile, ragged N narrowed to 16); `wave loss` = ceil(waves) / waves, the time a "
          "stream-K tail could recover.\n")
    print("| launches | M | N | K | conv3x3 | kernel | tiles | waves | padding | wave loss | GFLOP |")
    print("|---:|---:|---:|---:|---|---|---:|---:|---:|---:|---:|")
    tot = lost = bb = bb_lost = 0.0
    for (M, N, K, conv), n in sorted(agg.items(), key=lambda kv: -kv[1] * kv[0][0] * kv[0][1] * kv[0][2] * (9 if kv[0][3] else 1)):
        v, tiles, waves, pad, quant = model(M, N, K, conv)
        fl = 2.0 * M * N * K * (9 if conv else 1) * n / 1e9
        tot += fl
        lost += fl * (pad * quant - 1)
        if min(N, K) >= 1024 and M > 1024:       # the backbone's qkv / proj / fc1 / fc2: alone on the GPU when they run
            bb += fl
            bb_lost += fl * (pad * quant - 1)
        print(f"| {n} | {M} | {N} | {K} | {'yes' if conv else ''} | {'pair 256x256' if v == 2 else '1-CTA 128x128'} | {tiles} | "
              f"{waves:.2f} | {pad:.3f} | {quant:.3f} | {fl:.1f} |")
    print(f"\nTotal {tot:.0f} GFLOP algorithmic; FLOP-weighted schedule overhead (padding x wave loss - 1) = {100 * lost / tot:.1f} % "
          "of the tensor work, assuming every wave costs one full tile time (the single-wave decoder launches overlap each other "
          "on side streams, so their wave loss is an upper bound).")
    if bb:
        print(f"\nBackbone linears alone (N, K >= 1024; they run with the GPU to themselves): {bb:.0f} GFLOP, schedule overhead "
              f"{100 * bb_lost / bb:.1f} % = what row padding (M = {recs[0][0] if recs else 0} ...) and wave quantisation cost.")


if __name__ == "__main__":
    main()
