"""Stream-K schedule of the CTA-pair GEMM, per shape: one-pair-per-tile (policy 0) against the split (policy 2) on the
same box, CUDA events around 20 launches each, L2 flushed between launches. Development aid (profiles/r3_streamk.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mtt_b200  # noqa: F401
from mtt_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ws = ops.streamk_workspace(dev)
SHAPES = [("qkv bs4", 4116, 3072, 1024), ("fc1 bs4", 4116, 4096, 1024), ("fc2 bs4", 4116, 1024, 4096),
          ("proj bs4 (pair)", 4116, 1024, 1024), ("qkv bs1", 1029, 3072, 1024), ("fc2 bs1", 1029, 1024, 4096),
          ("dW qkv", 3072, 1024, 4116), ("dW fc1", 4096, 1024, 4116), ("dW fc2", 1024, 4096, 4116),
          ("dA qkv", 4116, 1024, 3072), ("dW proj", 1024, 1024, 4116)]


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


ops.set_gemm_variant(2)
print(f"{'shape':18s} {'tiles':>5s} {'k-blk':>5s} {'plain us':>9s} {'split us':>9s} {'ratio':>6s}")
for name, M, N, K in SHAPES:
    a = ops.split_f32(torch.randn(M, K, device=dev), 2)
    w = ops.split_f32(torch.randn(N, K, device=dev) * 0.02, 2)
    bias = torch.randn(N, device=dev)
    of = torch.zeros(M, N, device=dev)
    t = {}
    for pol in (0, 2):
        ops.set_gemm_streamk(pol)
        t[pol] = timeit(lambda: ops.gemm(a, w, bias=bias, out_f32=of, sk_ws=ws))
    tiles = ((M + 255) // 256) * ((N + 255) // 256)
    print(f"{name:18s} {tiles:5d} {(K + 63) // 64:5d} {t[0]:9.1f} {t[2]:9.1f} {t[2] / t[0]:6.2f}")
ops.set_gemm_streamk(1)
ops.set_gemm_variant(0)
