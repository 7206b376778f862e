"""clock64 timeline of one CTA of the simple attention kernel (debug aid)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mtt_b200
from mtt_b200 import ops, lib
dev = torch.device("cuda:0")
B, H, N, T = 4, 16, 1029, 5
C = H * 64
qkv = ops.split_f32(torch.randn(B * N, 3 * C, device=dev), 2)
out = ops.Split(B * N, C, dev, 2)
lg = torch.empty(B, H, T, N, device=dev)
ops.set_attention_variant(1)
L = lib.load()
L.mtt_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
buf = torch.zeros(768, dtype=torch.int64, device=dev)
for _ in range(3):
    ops.attention(qkv, out, B=B, N=N, H=H, scale=0.125, prompt_logits=lg, T=T)
torch.cuda.synchronize()
L.mtt_debug_set_attn_trace(ctypes.c_void_p(buf.data_ptr()))
ops.attention(qkv, out, B=B, N=N, H=H, scale=0.125, prompt_logits=lg, T=T)
torch.cuda.synchronize()
L.mtt_debug_set_attn_trace(ctypes.c_void_p(0))
t = buf.cpu().tolist()
for name, off in (("thread0", 0), ("thread40", 256), ("thread200", 512)):
    ev = [x for x in t[off:off + 256] if x]
    base = ev[0]
    print(name, "events", len(ev), "total cycles", ev[-1] - base)
    d = [ev[i + 1] - ev[i] for i in range(len(ev) - 1)]
    per = 9 if off == 0 else 6
    for it in range(0, min(len(d), per * 4), per):
        print("   ", d[it:it + per])
