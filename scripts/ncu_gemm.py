"""The four ViT-L block GEMMs (qkv, proj, fc1, fc2 at M = 4*1029) once each, for ncu captures:
    ncu --set full -k regex:gemm -c 4 python scripts/ncu_gemm.py <variant> <nsplit>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mtt_b200
from mtt_b200 import ops

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
ops.set_gemm_variant(variant)
M = 4 * 1029
shapes = [(3072, 1024, ops.ACT_NONE, False), (1024, 1024, ops.ACT_NONE, True), (4096, 1024, ops.ACT_GELU, False),
          (1024, 4096, ops.ACT_NONE, True)]
jobs = []
for N, K, act, res in shapes:
    a = ops.split_f32(torch.randn(M, K, device=dev), ns)
    w = ops.split_f32(torch.randn(N, K, device=dev) * 0.02, ns)
    bias = torch.randn(N, device=dev)
    of = torch.zeros(M, N, device=dev) if res else None
    osp = None if res else ops.Split(M, N, dev, ns)
    jobs.append((a, w, bias, act, of, osp))
for rep in range(2):
    if rep == 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    for a, w, bias, act, of, osp in jobs:
        ops.gemm(a, w, bias=bias, act=act, residual=of, out_f32=of, out_split=osp)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
