"""One fused-attention launch at the cfg4 shape (B=4, H=16, N=1029) for ncu:
    ncu --set full --import-source on -k regex:attention -c 1 python scripts/ncu_attn.py [nsplit]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mtt_b200
from mtt_b200 import ops

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
B, H, N, T = 4, 16, 1029, 5
C = H * 64
qkv = ops.split_f32(torch.randn(B * N, 3 * C, device=dev), ns)
out = ops.Split(B * N, C, dev, ns)
lg = torch.empty(B, H, T, N, device=dev)
for rep in range(2):
    if rep == 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ops.attention(qkv, out, B=B, N=N, H=H, scale=0.125, prompt_logits=lg, T=T)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
