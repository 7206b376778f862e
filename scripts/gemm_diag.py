"""Where does the GEMM time go?  Run under MTT_GEMM_DEBUG = 0 (normal), 1 (no TMA loads), 2 (no epilogue
stores), 3 (neither) for each kernel variant.  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mtt_b200
from mtt_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from microbench import bench_gemm

M = 4 * 1029
for v in (1, 2):
    ops.set_gemm_variant(v)
    for ns in (2, 1):
        print(f"[debug={os.environ.get('MTT_GEMM_DEBUG', '0')} variant={v}]", end=" ")
        bench_gemm(M, 3072, 1024, ns)
        print(f"[debug={os.environ.get('MTT_GEMM_DEBUG', '0')} variant={v}]", end=" ")
        bench_gemm(M, 1024, 4096, ns, res=True)
