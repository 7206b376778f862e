"""Clock-stamp timeline of the warp-specialised attention kernel (development aid): per key block, what the MMA warp
and softmax warp 0 of two co-resident CTAs were doing, in SM cycles.

    python scripts/attn_trace.py [B H N]      # default 4 16 1029 (cfg4)

Softmax stamps per block: a = before the wait for S_j, b = S_j ready, c = P_j computed, d = P_j published.
MMA stamps per block: x = P_g ready, y = PV_g issued, z = S_{g+2} issued."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
import mtt_b200
from mtt_b200 import ops, lib

B, H, N = [int(x) for x in sys.argv[1:4]] if len(sys.argv) > 3 else (4, 16, 1029)
if os.environ.get("MTT_ATTN_VARIANT"):
    ops.set_attention_variant(int(os.environ["MTT_ATTN_VARIANT"]))
dev = torch.device("cuda:0")
C = H * 64
qkv = ops.split_f32(torch.randn(B * N, 3 * C, device=dev), 2)
out = ops.Split(B * N, C, dev, 2)
lg = torch.empty(B, H, 5, N, device=dev)
for _ in range(3):
    ops.attention(qkv, out, B=B, N=N, H=H, scale=0.125, prompt_logits=lg, T=5)
torch.cuda.synchronize()
buf = torch.zeros(4096, dtype=torch.int32, device=dev)
L = lib.load()
L.mtt_set_attention_trace(ctypes.c_void_p(buf.data_ptr()))
ops.attention(qkv, out, B=B, N=N, H=H, scale=0.125, prompt_logits=lg, T=5)
torch.cuda.synchronize()
L.mtt_set_attention_trace(ctypes.c_void_p(0))
t = buf.cpu().numpy().astype("int64") & 0xFFFFFFFF
nkv = (N + 63) // 64
for cta in range(2):
    sm_, mm_ = t[(cta * 2) * 1024:(cta * 2 + 1) * 1024], t[(cta * 2 + 1) * 1024:(cta * 2 + 2) * 1024]
    print(f"---- CTA slot {cta}: smid {sm_[0]} / {mm_[0]}")
    s = sm_[1:]
    m = mm_[1:]
    nblk = 0
    while 4 * nblk + 3 < 956 and s[4 * nblk + 3] != 0:
        nblk += 1
    t0 = s[0]
    d = lambda a, b: int((b - a) & 0xFFFFFFFF)
    print("blk |  softmax: start(rel)  wait_S  compute  publish | mma: P_ready(rel)  PV_issue  S_issue")
    tot = {"wait": 0, "comp": 0, "pub": 0, "gap": 0}
    for k in range(nblk):
        a, b, c_, dd = s[4 * k:4 * k + 4]
        x, y, z = m[3 * k:3 * k + 3]
        gap = d(s[4 * k - 1], a) if k else 0
        print(f"{k:3d} | {d(t0, a):8d} {d(a, b):7d} {d(b, c_):8d} {d(c_, dd):8d}  (gap {gap:5d}) | {d(t0, x):8d} {d(x, y):8d} {d(y, z):8d}"
              + ("   <- item boundary" if (k + 1) % nkv == 0 else ""))
        tot["wait"] += d(a, b); tot["comp"] += d(b, c_); tot["pub"] += d(c_, dd); tot["gap"] += gap
    span = d(t0, s[4 * nblk - 1])
    ep = sm_[960:960 + 60].reshape(20, 3)
    for it in range(20):
        if ep[it, 2]:
            print(f"  item {it}: epilogue waits PV_last {d(ep[it, 0], ep[it, 1])} cycles, O/l + stores {d(ep[it, 1], ep[it, 2])}")
    print(f"blocks {nblk}, span {span} cycles = {span / max(nblk, 1):.0f} / block; softmax warp 0: wait_S {tot['wait']} "
          f"compute {tot['comp']} publish {tot['pub']} between-block {tot['gap']}")
