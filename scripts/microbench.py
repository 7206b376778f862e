"""Per-kernel timing on one B200 (CUDA events, L2 flushed between iterations). Development aid;
the judged numbers come from bench.py."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mtt_b200
from mtt_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def bench_gemm(M, N, K, nsplit, act=ops.ACT_NONE, res=False):
    a = ops.split_f32(torch.randn(M, K, device=dev), nsplit)
    w = ops.split_f32(torch.randn(N, K, device=dev) * 0.02, nsplit)
    bias = torch.randn(N, device=dev)
    of = torch.zeros(M, N, device=dev) if res else None
    osp = None if res else ops.Split(M, N, dev, nsplit)
    fn = lambda: ops.gemm(a, w, bias=bias, act=act, residual=of, out_f32=of, out_split=osp)
    ms = timeit(fn)
    fl = 2.0 * M * N * K
    print(f"gemm M={M} N={N} K={K} nsplit={nsplit} act={act} res={res}: {ms*1e3:.1f} us  "
          f"{fl/ms/1e9:.1f} TFLOP/s algorithmic, {fl*(3 if nsplit==2 else 1)/ms/1e9:.1f} TFLOP/s bf16-MMA")


def bench_attn(B, H, N, nsplit, T=5):
    C = H * 64
    qkv = ops.split_f32(torch.randn(B * N, 3 * C, device=dev), nsplit)
    out = ops.Split(B * N, C, dev, nsplit)
    lg = torch.empty(B, H, T, N, device=dev)
    fn = lambda: ops.attention(qkv, out, B=B, N=N, H=H, scale=0.125, prompt_logits=lg, T=T)
    ms = timeit(fn)
    fl = 4.0 * B * H * N * N * 64
    print(f"attn B={B} H={H} N={N} nsplit={nsplit}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s algorithmic, "
          f"{fl*(3 if nsplit==2 else 1)/ms/1e9:.1f} TFLOP/s bf16-MMA")


def bench_conv(B, H, W, Cin, Cout, nsplit, act=ops.ACT_GELU):
    from mtt_b200 import pack
    x = ops.split_f32(torch.randn(B * H * W, Cin, device=dev), nsplit)
    w = pack.pack_conv_weight(torch.randn(Cout, Cin, 3, 3, device=dev) * 0.02, nsplit)
    bias = torch.randn(Cout, device=dev)
    osp = ops.Split(B * H * W, Cout, dev, nsplit)
    fn = lambda: ops.gemm(x, w, N=Cout, K=Cin, bias=bias, act=act, out_split=osp, conv=(B, H, W, 3, 1))
    ms = timeit(fn)
    fl = 2.0 * B * H * W * Cout * 9 * Cin
    print(f"conv3x3 B={B} {H}x{W} {Cin}->{Cout} nsplit={nsplit}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s algorithmic, "
          f"{fl*(3 if nsplit==2 else 1)/ms/1e9:.1f} TFLOP/s bf16-MMA")


if __name__ == "__main__":
    M = 4 * 1029
    if len(sys.argv) > 1 and sys.argv[1] == "conv":
        for v in (1, 2, 3):
            ops.set_gemm_variant(v)
            print("---- gemm variant", v)
            bench_conv(4, 128, 128, 350, 350, 2)      # TaskPrompter ConvHead.mt_proj (taskprompter.py:691), cfg4
            bench_conv(4, 32, 32, 350, 350, 2)        # fea_fuse 3x3 (:362), cfg4
            bench_conv(4, 128, 128, 576, 576, 2)      # InvPT mt_proj (invpt.py:493), cfg3
            bench_conv(4, 112, 144, 768, 768, 2)      # ConvHead, cfg2
            bench_conv(4, 28, 36, 768, 768, 2)        # fea_fuse 3x3, cfg2 (48 pair tiles)
            bench_conv(4, 16, 16, 1024, 1024, 2, act=ops.ACT_RELU)   # InvPT ConvBlock (16 pair tiles)
            bench_conv(4, 16, 16, 1024, 512, 2, act=ops.ACT_RELU)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "attn":
        for v in [int(x) for x in os.environ.get("ATTN_VARIANTS", "3,5").split(",")]:
            ops.set_attention_variant(v)
            print("---- attention variant", v)
            for ns in (2, 1):
                bench_attn(4, 16, 1029, ns)
            bench_attn(4, 12, 1012, 2, T=4)
            bench_attn(1, 16, 8195, 2, T=3)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        for v in (1, 2, 3):
            ops.set_gemm_variant(v)
            print("---- gemm variant", v)
            for ns in (2, 1):
                bench_gemm(M, 3072, 1024, ns)
                bench_gemm(M, 1024, 1024, ns, res=True)
                bench_gemm(M, 4096, 1024, ns, act=ops.ACT_GELU)
                bench_gemm(M, 1024, 4096, ns, res=True)
                bench_gemm(4096, 304, 1024, ns)
                bench_gemm(4096, 352, 608, ns)
        sys.exit(0)
    for ns in (2, 1):
        bench_gemm(M, 3072, 1024, ns)
        bench_gemm(M, 1024, 1024, ns, res=True)
        bench_gemm(M, 4096, 1024, ns, act=ops.ACT_GELU)
        bench_gemm(M, 1024, 4096, ns, res=True)
        bench_attn(4, 16, 1029, ns)
    bench_attn(1, 16, 8195, 2, T=3)
