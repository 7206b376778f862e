"""The x4 NHWC bilinear resize of the cfg4 decoder (taskprompter.py:420: 32x32 -> 128x128, 350 channels, bs 4, split output)
and a pure-write / copy reference on the same box. With `ncu -k regex:bilinear --set full` for the capture; run plainly it
prints CUDA-event timings (L2 flushed between launches)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mtt_b200  # noqa: F401
from mtt_b200 import ops

dev = torch.device("cuda:0")
B, h, w, C, H2, W2 = 4, 32, 32, 350, 128, 128
x = torch.randn(B * h * w, C + 2, device=dev)
out = ops.Split(B * H2 * W2, C, dev, 2)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


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


nbytes = out.buf.numel() * 2
t = timeit(lambda: ops.bilinear(x, x.stride(0), B, h, w, C, H2, W2, out_split=out))
print(f"bilinear_nhwc x4 -> split [{B * H2 * W2}, {C}]: {t:.1f} us, {nbytes / 1e6:.1f} MB written = {nbytes / t / 1e6:.2f} TB/s")
big = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
src = torch.randn(nbytes // 4, device=dev)
t = timeit(lambda: big.fill_(1.0))
print(f"torch fill of the same {nbytes / 1e6:.1f} MB: {t:.1f} us = {nbytes / t / 1e6:.2f} TB/s written")
t = timeit(lambda: big.copy_(src))
print(f"torch copy of the same size: {t:.1f} us = {2 * nbytes / t / 1e6:.2f} TB/s read + written")
g1 = torch.empty(1 << 28, dtype=torch.float32, device=dev)
t = timeit(lambda: g1.fill_(2.0), iters=5)
print(f"torch fill of 1 GiB: {t:.1f} us = {(1 << 30) / t / 1e6:.2f} TB/s written")
