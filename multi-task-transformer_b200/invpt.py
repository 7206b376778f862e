"""InvPT (ViT backbone + inverted-pyramid multi-task decoder) with the reference's nn.Module
boundaries and a fused sm_100a forward.

Module classes and parameter names mirror the reference so that its checkpoints load unchanged:

  VisionTransformer / Block / Attention        InvPT/models/transformers/vit.py:172-351
  TransformerDecoder / ConvBlock / MLPHead     InvPT/models/transformers/transformer_decoder.py:18-131
  InvPT / InvPTStage / InvPTBlock / SelfAttention / UpEmbed   InvPT/models/transformers/invpt.py:19-544
  TransformerNet                               InvPT/models/transformer_net.py:12-38

The modules only own parameters; all arithmetic runs in libmtt_sm100.so through `ops` (see
taskprompter.py for the conventions). Eval-mode only (SyncBatchNorm = running statistics, DropPath =
identity). Not reproduced because the reference never consumes them: scale_embed[2]'s output and
`norm_mt`.
"""
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import ops
from .pack import fold_bn, pack_conv_weight, pack_linear_weight
from .taskprompter import PARITY, Mlp, PatchEmbed, _trunc_normal_


# --------------------------------------------------------------------------------------------
# parameter containers
# --------------------------------------------------------------------------------------------
class VitAttention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias=True):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class VitBlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=True):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = VitAttention(dim, num_heads, qkv_bias)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class VisionTransformer(nn.Module):
    """vit.py:227-330 (same constructor arguments that matter for the forward)."""

    def __init__(self, select_list, img_size=(224, 224), patch_size=16, in_chans=3, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=True, **_unused):
        super().__init__()
        if isinstance(img_size, int):
            img_size = (img_size, img_size)
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.patch_size = patch_size
        self.in_chans = in_chans
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.blocks = nn.Sequential(*[VitBlock(embed_dim, num_heads, mlp_ratio, qkv_bias) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.select_list = list(select_list)
        _trunc_normal_(self.pos_embed, std=.02)
        _trunc_normal_(self.cls_token, std=.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, std=.02)
                nn.init.zeros_(m.bias)


class ConvBlock(nn.Module):
    """transformer_decoder.py:100-122: conv3x3 (no bias) -> BN -> ReLU."""

    def __init__(self, inplanes, planes):
        super().__init__()
        self.conv = nn.Conv2d(inplanes, planes, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)


class MLPHead(nn.Module):
    """transformer_decoder.py:124-131."""

    def __init__(self, in_channels, num_classes):
        super().__init__()
        self.linear_pred = nn.Conv2d(in_channels, num_classes, kernel_size=1)


class UpEmbed(nn.Module):
    """invpt.py:19-43; Sequential indices 1,2,4,5 carry the parameters."""

    def __init__(self, in_chans, embed_dim):
        super().__init__()
        self.proj = nn.Sequential(
            nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False),
            nn.Conv2d(in_chans, embed_dim, 3, padding=2, stride=1, bias=False, dilation=2),
            nn.BatchNorm2d(embed_dim), nn.ReLU(inplace=True),
            nn.Conv2d(embed_dim, embed_dim, 3, padding=2, stride=1, bias=False, dilation=2),
            nn.BatchNorm2d(embed_dim), nn.ReLU(inplace=True))


class _DwBn(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.conv = nn.Conv2d(dim, dim, 3, padding=1, stride=2, bias=False, groups=dim)
        self.bn = nn.BatchNorm2d(dim)


class SelfAttention(nn.Module):
    """invpt.py:68-164 (q_method 'dw_bn', kv_method 'avg')."""

    def __init__(self, fea_no, dim, num_heads, stride_kv):
        super().__init__()
        self.fea_no, self.dim, self.num_heads, self.stride_kv = fea_no, dim, num_heads, stride_kv
        self.scale = dim ** -0.5
        self.conv_proj_q = nn.ModuleList([_DwBn(dim) for _ in range(fea_no)])
        self.proj_q = nn.Linear(dim, dim)
        self.proj_k = nn.Linear(dim, dim)
        self.proj_v = nn.Linear(dim, dim)
        self.proj = nn.Linear(dim, dim)
        self.fuse_attn = nn.Conv2d(num_heads * 2, num_heads, 1)


class InvPTBlock(nn.Module):
    def __init__(self, task_no, dim, num_heads, stride_kv, mlp_ratio=4.):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.attn = SelfAttention(task_no, dim, num_heads, stride_kv)


class InvPTStage(nn.Module):
    def __init__(self, task_no, stage_idx, in_chans, embed_dim, num_heads, stride_kv):
        super().__init__()
        self.stage_idx = stage_idx
        self.patch_embed = None if stage_idx == 0 else nn.ModuleList(
            [UpEmbed(in_chans, embed_dim) for _ in range(task_no)])
        self.blocks = nn.ModuleList([InvPTBlock(task_no, embed_dim, num_heads, stride_kv)])
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)


class InvPT(nn.Module):
    """invpt.py:419-500."""

    def __init__(self, p, in_chans, ori_embed_dim):
        super().__init__()
        tasks = list(p.TASKS.NAMES)
        T = len(tasks)
        dims = [in_chans, in_chans // 2, in_chans // 4]
        self.dims = dims
        self.norm_mts = nn.ModuleList()
        self.redu_chan = nn.ModuleList()
        self.invpt_stages = nn.ModuleList()
        prev = in_chans
        for i in range(3):
            self.invpt_stages.append(InvPTStage(T, i, prev, dims[i], 2, 2 ** (i + 1)))
            prev = dims[i]
            self.norm_mts.append(nn.LayerNorm(dims[i] * T))
            self.redu_chan.append(nn.ModuleList([nn.Conv2d(dims[i], in_chans, 1) for _ in range(T)]))
        self.norm_mt = nn.LayerNorm(T * dims[2])   # never called by the reference forward
        self.mt_proj = nn.ModuleDict({t: nn.Sequential(nn.Conv2d(in_chans, in_chans, 3, padding=1),
                                                       nn.BatchNorm2d(in_chans), nn.ReLU(True)) for t in tasks})
        self.mix_proj = nn.ModuleDict({t: nn.Sequential(nn.Conv2d(ori_embed_dim + p.TASKS.NUM_OUTPUT[t], in_chans, 1))
                                       for t in tasks})


class TransformerDecoder(nn.Module):
    """transformer_decoder.py:18-67."""

    def __init__(self, p):
        super().__init__()
        self.embed_dim = p.embed_dim
        d0 = p.embed_dim + p.PRED_OUT_NUM_CONSTANT
        tasks = list(p.TASKS.NAMES)
        C = p.backbone_channels[-1]
        self.intermediate_head = nn.ModuleDict({t: nn.Conv2d(p.embed_dim, p.TASKS.NUM_OUTPUT[t], 1) for t in tasks})
        self.invpt = InvPT(p, in_chans=d0, ori_embed_dim=p.embed_dim)
        self.preliminary_decoder = nn.ModuleDict(
            {t: nn.Sequential(ConvBlock(C, C), ConvBlock(C, p.embed_dim)) for t in tasks})
        self.scale_embed = nn.ModuleList([
            nn.ConvTranspose2d(p.backbone_channels[0], d0 // 4, kernel_size=3, stride=2, padding=1, output_padding=1),
            nn.Conv2d(p.backbone_channels[1], d0 // 2, 3, padding=1),
            nn.Conv2d(p.backbone_channels[2], d0, 3, padding=1),
            None])


class TransformerNet(nn.Module):
    """transformer_net.py:12-38. forward(x[B,3,H,W]) -> {task: [B,n_out,H,W], 'inter_preds': {task: ...}}."""

    def __init__(self, p, backbone, backbone_channels, heads, nsplit=PARITY, use_graph=True):
        super().__init__()
        self.p = p
        self.tasks = list(p.TASKS.NAMES)
        self.backbone = backbone
        self.multi_task_decoder = TransformerDecoder(p)
        self.heads = heads
        self.nsplit = nsplit
        self.use_graph = use_graph
        self._plans = {}

    def _param_version(self):
        return sum(int(q._version) for q in self.parameters()) + sum(int(b._version) for b in self.buffers())

    def plan(self, batch, device, postproc=False):
        key = (int(batch), str(device), int(self.nsplit), bool(postproc))
        ver = self._param_version()
        pl = self._plans.get(key)
        if pl is None or pl.version != ver:
            pl = _Plan(self, batch, device, self.nsplit, postproc=postproc)
            pl.version = ver
            self._plans[key] = pl
        return pl

    def _check(self, x):
        if self.training:
            raise NotImplementedError("mtt_b200 InvPT: fused forward is eval-only; backward kernels are not built "
                                      "yet (SURVEY.md section 8f N1)")
        if not x.is_cuda:
            raise RuntimeError("mtt_b200 has no CPU path: input must be a CUDA tensor on an sm_100a device")

    def forward(self, x):
        self._check(x)
        return self.plan(x.shape[0], x.device).run(x, graph=self.use_graph)

    def predict(self, x):
        """forward + the reference's `get_output` post-processing (InvPT/utils/utils.py:18-48) fused into the
        final resize of every task head (no full-resolution logits, no inter_preds)."""
        self._check(x)
        return self.plan(x.shape[0], x.device, postproc=True).run(x, graph=self.use_graph)


# --------------------------------------------------------------------------------------------
# the fused forward
# --------------------------------------------------------------------------------------------
class _Plan:
    def __init__(self, net, B, device, nsplit, postproc=False):
        ops._L.check(ops._L.load().mtt_device_check(), "mtt_device_check")
        self.postproc = postproc
        bb, dec, p = net.backbone, net.multi_task_decoder, net.p
        inv = dec.invpt
        self.B, self.dev, self.ns = B, device, nsplit
        self.tasks = list(net.tasks)
        self.T = T = len(self.tasks)
        self.C = C = bb.embed_dim
        self.H = bb.num_heads
        assert C // self.H == 64, "attention kernel is built for head_dim 64"
        self.gh, self.gw = bb.patch_embed.grid_size
        self.P = P = self.gh * self.gw
        self.N = N = 1 + P
        self.patch = bb.patch_size
        self.img = (self.gh * self.patch, self.gw * self.patch)
        self.select = list(bb.select_list)
        self.depth = len(bb.blocks)
        self.E = E = p.embed_dim
        self.dims = dims = list(inv.dims)
        down = p.mtt_resolution_downsample_rate
        self.h0, self.w0 = self.gh // down, self.gw // down
        self.th, self.tw = self.h0 * 8, self.w0 * 8
        self.n_out = [p.TASKS.NUM_OUTPUT[t] for t in self.tasks]
        self.graph = None
        self.static_in = None
        ns = nsplit
        h0, w0 = self.h0, self.w0

        def f32(t):
            return t.detach().to(device=device, dtype=torch.float32).contiguous()

        W = SimpleNamespace()
        W.pe_w = pack_linear_weight(f32(bb.patch_embed.proj.weight), ns)
        W.pe_b = f32(bb.patch_embed.proj.bias)
        W.pos = f32(bb.pos_embed)[0, 1:]
        W.cls = (f32(bb.cls_token)[0] + f32(bb.pos_embed)[0, :1]).contiguous()      # vit.py:334-339, row 0
        W.blocks = []
        for blk in bb.blocks:
            w = SimpleNamespace()
            w.n1w, w.n1b, w.n2w, w.n2b = f32(blk.norm1.weight), f32(blk.norm1.bias), f32(blk.norm2.weight), f32(blk.norm2.bias)
            w.eps = blk.norm1.eps
            w.qkv, w.qkv_b = pack_linear_weight(f32(blk.attn.qkv.weight), ns), f32(blk.attn.qkv.bias)
            w.proj, w.proj_b = pack_linear_weight(f32(blk.attn.proj.weight), ns), f32(blk.attn.proj.bias)
            w.fc1, w.fc1_b = pack_linear_weight(f32(blk.mlp.fc1.weight), ns), f32(blk.mlp.fc1.bias)
            w.fc2, w.fc2_b = pack_linear_weight(f32(blk.mlp.fc2.weight), ns), f32(blk.mlp.fc2.bias)
            W.blocks.append(w)
        W.nw, W.nb, W.neps = f32(bb.norm.weight), f32(bb.norm.bias), bb.norm.eps
        # ConvTranspose2d(k3,s2,p1,op1) == zero-insert + conv3x3(pad 1) with the spatially flipped,
        # in/out-transposed kernel
        wt = f32(dec.scale_embed[0].weight)                                           # [Cin, Cout, 3, 3]
        W.se0 = pack_conv_weight(wt.flip(2, 3).permute(1, 0, 2, 3).contiguous(), ns)
        W.se0_b = f32(dec.scale_embed[0].bias)
        W.se1, W.se1_b = pack_conv_weight(f32(dec.scale_embed[1].weight), ns), f32(dec.scale_embed[1].bias)
        W.tasks = []
        for t in self.tasks:
            tw = SimpleNamespace()
            pd = dec.preliminary_decoder[t]
            w0_, b0_ = fold_bn(f32(pd[0].conv.weight), None, pd[0].bn1)
            w1_, b1_ = fold_bn(f32(pd[1].conv.weight), None, pd[1].bn1)
            tw.pd0, tw.pd0_b = pack_conv_weight(w0_, ns), b0_.contiguous()
            tw.pd1, tw.pd1_b = pack_conv_weight(w1_, ns), b1_.contiguous()
            tw.ih, tw.ih_b = pack_linear_weight(f32(dec.intermediate_head[t].weight), ns), f32(dec.intermediate_head[t].bias)
            tw.mix, tw.mix_b = pack_linear_weight(f32(inv.mix_proj[t][0].weight), ns), f32(inv.mix_proj[t][0].bias)
            wm, bm = fold_bn(f32(inv.mt_proj[t][0].weight), f32(inv.mt_proj[t][0].bias), inv.mt_proj[t][1])
            tw.mt, tw.mt_b = pack_conv_weight(wm, ns), bm.contiguous()
            hd = net.heads[t]
            tw.lp, tw.lp_b = pack_linear_weight(f32(hd.linear_pred.weight), ns), f32(hd.linear_pred.bias)
            W.tasks.append(tw)
        W.stages = []
        for i, st in enumerate(inv.invpt_stages):
            sw = SimpleNamespace()
            if i > 0:
                sw.up = []
                for k in range(T):
                    pr = st.patch_embed[k].proj
                    wa, ba = fold_bn(f32(pr[1].weight), None, pr[2])
                    wb, bb_ = fold_bn(f32(pr[4].weight), None, pr[5])
                    sw.up.append((pack_conv_weight(wa, ns), ba.contiguous(), pack_conv_weight(wb, ns), bb_.contiguous()))
            blk = st.blocks[0]
            sw.n1w, sw.n1b, sw.n2w, sw.n2b = f32(blk.norm1.weight), f32(blk.norm1.bias), f32(blk.norm2.weight), f32(blk.norm2.bias)
            sw.eps = blk.norm1.eps
            qw, qb = [], []
            for k in range(T):
                cw, cb = fold_bn(f32(blk.attn.conv_proj_q[k].conv.weight), None, blk.attn.conv_proj_q[k].bn)
                qw.append(cw.reshape(dims[i], 9))
                qb.append(cb)
            sw.dw_w, sw.dw_b = torch.stack(qw).contiguous(), torch.stack(qb).contiguous()
            for nm in ("proj_q", "proj_k", "proj_v", "proj"):
                lin = getattr(blk.attn, nm)
                setattr(sw, nm, pack_linear_weight(f32(lin.weight), ns))
                setattr(sw, nm + "_b", f32(lin.bias))
            sw.fuse_w = f32(blk.attn.fuse_attn.weight).reshape(2, 4).contiguous()
            sw.fuse_b = f32(blk.attn.fuse_attn.bias)
            sw.fc1, sw.fc1_b = pack_linear_weight(f32(blk.mlp.fc1.weight), ns), f32(blk.mlp.fc1.bias)
            sw.fc2, sw.fc2_b = pack_linear_weight(f32(blk.mlp.fc2.weight), ns), f32(blk.mlp.fc2.bias)
            sw.nmw, sw.nmb, sw.nmeps = f32(inv.norm_mts[i].weight), f32(inv.norm_mts[i].bias), inv.norm_mts[i].eps
            if i > 0:
                sw.redu = [(pack_linear_weight(f32(inv.redu_chan[i][k].weight), ns), f32(inv.redu_chan[i][k].bias))
                           for k in range(T)]
            W.stages.append(sw)
        self.W = W

        # ---- workspace
        S = lambda r, c, **kw: ops.Split(r, c, device, ns, **kw)
        z = lambda *s: torch.zeros(*s, device=device, dtype=torch.float32)
        self.cols = S(B * P, self.patch * self.patch * bb.in_chans)
        self.xs = z(B * N, C)
        self.xn = S(B * N, C)
        self.qkv = S(B * N, 3 * C)
        self.ao = S(B * N, C)
        self.hid = S(B * N, 4 * C)
        self.zi = S(B * 4 * P, C)
        self.f1 = S(B * P, C)
        self.back0 = z(B * 4 * P, dims[2])
        self.back1 = z(B * P, dims[1])
        self.xfin = z(B * P, C)
        self.x0 = S(B * h0 * w0, C)
        self.p1 = [S(B * h0 * w0, C) for _ in range(T)]   # per task: the task chains run on side streams
        self.cat = [S(B * h0 * w0, E + n, zero=True) for n in self.n_out]
        self.inter = [z(B * h0 * w0, ops.round_up(n, 4)) for n in self.n_out]
        self.st = []
        for i in range(3):
            h, w = h0 * 2 ** i, w0 * 2 ** i
            Ci = dims[i]
            s = SimpleNamespace(h=h, w=w, C=Ci)
            s.kvs = 2 ** (i + 1)
            s.kh, s.kw = -(-h // s.kvs), -(-w // s.kvs)
            s.Lq = T * (h // 2) * (w // 2)
            s.Tk = T * s.kh * s.kw
            s.xj = z(B * T * h * w, Ci)
            s.xn32 = z(B * T * h * w, Ci)
            s.qin = S(B * s.Lq, Ci)
            s.kvin = S(B * s.Tk, Ci)
            s.q32, s.k32, s.v32 = z(B * s.Lq, Ci), z(B * s.Tk, Ci), z(B * s.Tk, Ci)
            s.score = z(B, 2, s.Lq, s.Tk) if i < 2 else None
            s.ao = S(B * s.Lq, Ci)
            s.a32 = z(B * s.Lq, Ci)
            s.xn = S(B * T * h * w, Ci)
            s.hid = S(B * T * h * w, 4 * Ci)
            if i == 0:
                s.ln32 = z(T * B * h * w, Ci)
            else:
                s.ln = S(T * B * h * w, Ci)
                s.rc32 = [z(B * h * w, dims[0]) for _ in range(T)]
                s.ue0 = [S(B * h * w, dims[i - 1]) for _ in range(T)]
                s.ue1 = [S(B * h * w, Ci) for _ in range(T)]
            self.st.append(s)
        tt = B * self.th * self.tw
        self.mss = [S(tt, dims[0]) for _ in range(T)]
        self.hm = [S(tt, dims[0]) for _ in range(T)]
        self.side = None
        self.pred = [z(tt, ops.round_up(n, 4)) for n in self.n_out]
        oh, ow = self.img
        if not postproc:
            self.out = {t: z(B, n, oh, ow) for t, n in zip(self.tasks, self.n_out)}
            self.out_inter = {t: z(B, n, oh, ow) for t, n in zip(self.tasks, self.n_out)}
        else:
            self.out, self.out_inter = {}, None
            for t in self.tasks:
                kind = ops.POSTPROC_KIND[t]
                shape = {0: (B, oh, ow), 1: (B, oh, ow), 2: (B, oh, ow), 3: (B, oh, ow, 3), 4: (B, oh, ow, 1)}[kind]
                self.out[t] = torch.zeros(shape, device=device, dtype=torch.int64 if kind == 0 else torch.float32)

    # ------------------------------------------------------------------------------------------
    def _vit_block(self, w):
        B, N = self.B, self.N
        ops.layernorm(self.xs, w.n1w, w.n1b, w.eps, out_split=self.xn)                     # vit.py:213
        ops.gemm(self.xn, w.qkv, bias=w.qkv_b, out_split=self.qkv)                         # :186
        ops.attention(self.qkv, self.ao, B=B, N=N, H=self.H, scale=64 ** -0.5)             # :189-193
        ops.gemm(self.ao, w.proj, bias=w.proj_b, residual=self.xs, out_f32=self.xs)        # :194,:213
        ops.layernorm(self.xs, w.n2w, w.n2b, w.eps, out_split=self.xn)                     # :214
        ops.gemm(self.xn, w.fc1, bias=w.fc1_b, act=ops.ACT_GELU, out_split=self.hid)
        ops.gemm(self.hid, w.fc2, bias=w.fc2_b, residual=self.xs, out_f32=self.xs)

    def _par(self, fn):
        """Run fn(k) for every task k, each on its own side stream forked from / joined to the current
        stream (the per-task chains are independent and individually too small to fill 148 SMs)."""
        T = self.T
        if self.dev.type != "cuda" or getattr(self, "serial", False):   # serial: one stream (per-kernel timing)
            for k in range(T):
                fn(k)
            return
        if self.side is None:
            self.side = [torch.cuda.Stream(device=self.dev) for _ in range(T)]
        main = torch.cuda.current_stream()
        for k in range(T):
            self.side[k].wait_stream(main)
            with torch.cuda.stream(self.side[k]):
                fn(k)
        for k in range(T):
            main.wait_stream(self.side[k])

    def _stage(self, i):
        """InvPTStage + InvPTBlock + multi-scale aggregation for stage i (invpt.py:400-417,290-312,522-539)."""
        B, T, W = self.B, self.T, self.W
        s, sw = self.st[i], W.stages[i]
        h, w, Ci = s.h, s.w, s.C
        hw = h * w
        if i > 0:
            sp = self.st[i - 1]
            skip = self.back1 if i == 1 else self.back0
            def up_embed(k):
                wa, ba, wb, bb_ = sw.up[k]
                ops.bilinear(sp.xj, sp.xj.stride(0), B, sp.h, sp.w, sp.C, h, w, out_split=s.ue0[k],
                             in_batch_rows=T * sp.h * sp.w, in_row_offset=k * sp.h * sp.w)            # UpEmbed :32
                ops.gemm(s.ue0[k], wa, N=Ci, K=sp.C, bias=ba, act=ops.ACT_RELU, out_split=s.ue1[k],
                         conv=(B, h, w, 3, 2))                                                        # :33-35
                ops.gemm(s.ue1[k], wb, N=Ci, K=Ci, bias=bb_, act=ops.ACT_RELU, residual=skip, res_row_mod=B * hw,
                         out_f32=s.xj, regroup=(hw, T * hw, k * hw), conv=(B, h, w, 3, 2))            # :36-38,:406-411
            self._par(up_embed)
        # ---- InvPTBlock
        ops.layernorm(s.xj, sw.n1w, sw.n1b, sw.eps, out_f32=s.xn32)                                    # :298
        ops.dwconv3x3_s2(s.xn32, sw.dw_w, sw.dw_b, s.qin, B=B, T=T, h=h, w=w, Cdim=Ci)                 # :171-173
        ops.avgpool(s.xn32, s.kvin, BT=B * T, h=h, w=w, Cdim=Ci, s=s.kvs)                              # :175-187
        ops.gemm(s.qin, sw.proj_q, bias=sw.proj_q_b, out_f32=s.q32)                                    # :200
        ops.gemm(s.kvin, sw.proj_k, bias=sw.proj_k_b, out_f32=s.k32)                                   # :201
        ops.gemm(s.kvin, sw.proj_v, bias=sw.proj_v_b, out_f32=s.v32)                                   # :202
        prev = self.st[i - 1].score if i > 0 else None
        ops.invpt_attention(s.q32, s.k32, s.v32, s.ao, B=B, Lq=s.Lq, Tk=s.Tk, Cdim=Ci, scale=Ci ** -0.5,
                            prev_score=prev, T=T, qh=h // 2, qw=w // 2, fuse_w=sw.fuse_w, fuse_b=sw.fuse_b,
                            score_out=s.score)                                                          # :204-236
        ops.gemm(s.ao, sw.proj, bias=sw.proj_b, out_f32=s.a32)                                         # :238
        qhw = (h // 2) * (w // 2)
        self._par(lambda k: ops.bilinear(                                                              # :299-306
            s.a32, s.a32.stride(0), B, h // 2, w // 2, Ci, h, w, out_f32=s.xj, accumulate=True,
            in_batch_rows=T * qhw, in_row_offset=k * qhw, out_batch_rows=T * hw, out_row_offset=k * hw))
        ops.layernorm(s.xj, sw.n2w, sw.n2b, sw.eps, out_split=s.xn)                                    # :307
        ops.gemm(s.xn, sw.fc1, bias=sw.fc1_b, act=ops.ACT_GELU, out_split=s.hid)
        ops.gemm(s.hid, sw.fc2, bias=sw.fc2_b, residual=s.xj, out_f32=s.xj)
        # ---- joint-channel LayerNorm over all tasks, per-task slices to the common resolution
        ops.layernorm_seg(s.xj, sw.nmw, sw.nmb, sw.nmeps, rows=B * hw, cols=Ci, S=T, in_group=hw, src_group=T * hw,
                          seg_stride=hw, out_f32=s.ln32 if i == 0 else None,
                          out_split=None if i == 0 else s.ln, out_seg_stride=B * hw)                   # :524-526
        d0 = self.dims[0]

        def aggregate(k):
            # per-task slice -> (i > 0: 1x1 redu_chan) kept at its own resolution; after the last stage the three
            # maps are resized to 8h0 x 8w0, summed and written ONCE as the split operand of mt_proj (:528-543)
            if i > 0:
                rw, rb = sw.redu[k]
                ops.gemm(s.ln, rw, M=B * hw, bias=rb, out_f32=s.rc32[k], a_row_offset=k * B * hw)      # :535-536
            if i == 2:
                s0_, s1_ = self.st[0], self.st[1]
                ops.bilinear_sum3([(s0_.ln32, s0_.h, s0_.w, 0, k * B * s0_.h * s0_.w),
                                   (s1_.rc32[k], s1_.h, s1_.w, 0, 0),
                                   (s.rc32[k], h, w, 0, 0)],
                                  self.mss[k], B=B, Cdim=d0, H2=self.th, W2=self.tw)                   # :537-539
                self._head(k)
        self._par(aggregate)

    def _head(self, k):
        B, W = self.B, self.W
        tw = W.tasks[k]
        d0 = self.dims[0]
        ops.gemm(self.mss[k], tw.mt, N=d0, K=d0, bias=tw.mt_b, act=ops.ACT_RELU, out_split=self.hm[k],
                 conv=(B, self.th, self.tw, 3, 1))                                                     # invpt.py:541-543
        n = self.n_out[k]
        ops.gemm(self.hm[k], tw.lp, bias=tw.lp_b, out_f32=self.pred[k][:, :n], N=n)                    # MLPHead
        if self.postproc:
            ops.bilinear_postproc(self.pred[k], self.pred[k].stride(0), B, self.th, self.tw, n, self.img[0],
                                  self.img[1], ops.POSTPROC_KIND[self.tasks[k]], self.out[self.tasks[k]])
        else:
            ops.bilinear(self.pred[k], self.pred[k].stride(0), B, self.th, self.tw, n, self.img[0], self.img[1],
                         out_nchw=self.out[self.tasks[k]])                                             # transformer_net.py:35

    def _launch(self, img):
        B, N, P, C, T, W = self.B, self.N, self.P, self.C, self.T, self.W
        h0, w0, E = self.h0, self.w0, self.E
        ops.im2col_patch(img, self.patch, self.cols)
        ops.gemm(self.cols, W.pe_w, bias=W.pe_b, residual=W.pos, res_row_mod=P, out_f32=self.xs,
                 regroup=(P, N, 1))                                                                    # vit.py:333,339
        ops.broadcast_rows(W.cls, self.xs, B, N)                                                       # :334-339
        for idx, w in enumerate(W.blocks):
            self._vit_block(w)
            if idx + 1 in self.select:
                which = self.select.index(idx + 1)
                if which == 0:      # scale_embed[0]: ConvTranspose2d as zero-insert + flipped 3x3 conv
                    ops.zero_insert(self.xs, self.zi, B=B, h=self.gh, w=self.gw, Cdim=C, src_group=N, src_offset=1)
                    ops.gemm(self.zi, W.se0, N=self.dims[2], K=C, bias=W.se0_b, out_f32=self.back0,
                             conv=(B, 2 * self.gh, 2 * self.gw, 3, 1))                                 # transformer_decoder.py:63,80
                elif which == 1:    # scale_embed[1]
                    ops.split_rows(self.xs, self.f1, rows=B * P, cols=C, in_group=P, src_group=N, src_offset=1)
                    ops.gemm(self.f1, W.se1, N=self.dims[1], K=C, bias=W.se1_b, out_f32=self.back1,
                             conv=(B, self.gh, self.gw, 3, 1))                                         # :64,:80
                # which == 2: scale_embed[2]'s output is never consumed by the reference
        ops.layernorm_seg(self.xs, W.nw, W.nb, W.neps, rows=B * P, cols=C, S=1, in_group=P, src_group=N,
                          src_offset=1, out_f32=self.xfin)                                             # vit.py:348-349
        ops.bilinear(self.xfin, C, B, self.gh, self.gw, C, h0, w0, out_split=self.x0)                  # transformer_decoder.py:85-86
        s0 = self.st[0]
        hw0 = h0 * w0
        def prelim(k):
            tw = W.tasks[k]
            n = self.n_out[k]
            ops.gemm(self.x0, tw.pd0, N=C, K=C, bias=tw.pd0_b, act=ops.ACT_RELU, out_split=self.p1[k],
                     conv=(B, h0, w0, 3, 1))                                                           # ConvBlock 1
            ops.gemm(self.p1[k], tw.pd1, N=E, K=C, bias=tw.pd1_b, act=ops.ACT_RELU, out_split=self.cat[k],
                     conv=(B, h0, w0, 3, 1))                                                           # ConvBlock 2
            ops.gemm(self.cat[k], tw.ih, K=E, bias=tw.ih_b, out_f32=self.inter[k][:, :n], N=n,
                     out_split=self.cat[k], out_col_offset=E)                                          # :94; invpt.py:511
            ops.gemm(self.cat[k], tw.mix, bias=tw.mix_b, out_f32=s0.xj, regroup=(hw0, T * hw0, k * hw0))  # invpt.py:512
            if not self.postproc:
                ops.bilinear(self.inter[k], self.inter[k].stride(0), B, h0, w0, n, self.img[0], self.img[1],
                             out_nchw=self.out_inter[self.tasks[k]])                                   # transformer_net.py:36
        self._par(prelim)
        for i in range(3):
            self._stage(i)

    def run(self, x, graph=True):
        if tuple(x.shape[1:]) != (3, *self.img) or x.dtype != torch.float32:
            raise ValueError(f"expected fp32 input [B,3,{self.img[0]},{self.img[1]}], got {tuple(x.shape)} {x.dtype}")
        if not graph:
            self._launch(x.contiguous())
        else:
            if self.static_in is None:
                self.static_in = torch.empty_like(x, memory_format=torch.contiguous_format)
            self.static_in.copy_(x, non_blocking=True)
            if self.graph is None:
                self._launch(self.static_in)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch(self.static_in)
                self.graph = g
            self.graph.replay()
        out = dict(self.out)
        if self.out_inter is not None:
            out["inter_preds"] = dict(self.out_inter)
        return out


def build_from_config(cfg, nsplit=PARITY, use_graph=True):
    """cfg: dict as in oracle/configs.invpt() (mirrors IP/utils/common_config.py:15-21,39-51)."""
    H, Wd = cfg["img_size"]
    gh, gw = H // cfg["patch"], Wd // cfg["patch"]
    p = SimpleNamespace(TASKS=SimpleNamespace(NAMES=list(cfg["tasks"]), NUM_OUTPUT=dict(cfg["num_output"])),
                        embed_dim=cfg["embed_dim"], PRED_OUT_NUM_CONSTANT=cfg["pred_const"],
                        mtt_resolution_downsample_rate=cfg["down"], backbone_channels=[cfg["C"]] * 4,
                        spatial_dim=[[gh, gw]] * 4)
    p.final_embed_dim = cfg["embed_dim"] + cfg["pred_const"]
    bb = VisionTransformer(cfg["select"], img_size=(H, Wd), patch_size=cfg["patch"], embed_dim=cfg["C"],
                           depth=cfg["depth"], num_heads=cfg["heads"])
    heads = nn.ModuleDict({t: MLPHead(p.final_embed_dim, cfg["num_output"][t]) for t in cfg["tasks"]})
    return TransformerNet(p, bb, p.backbone_channels, heads, nsplit=nsplit, use_graph=use_graph)


def accelerate(ref_model, nsplit=PARITY, use_graph=True):
    """Drop-in: build the fused InvPT from a REFERENCE TransformerNet instance (InvPT/models/transformer_net.py),
    sharing its parameters through an exact `load_state_dict(strict=True)`."""
    p = ref_model.multi_task_decoder.p
    bb = ref_model.backbone
    mine_bb = VisionTransformer(list(bb.select_list), img_size=tuple(bb.patch_embed.img_size),
                                patch_size=bb.patch_embed.patch_size[0], embed_dim=bb.embed_dim,
                                depth=len(bb.blocks), num_heads=bb.blocks[0].attn.num_heads)
    heads = nn.ModuleDict({t: MLPHead(ref_model.heads[t].linear_pred.weight.shape[1],
                                      ref_model.heads[t].linear_pred.weight.shape[0]) for t in ref_model.tasks})
    m = TransformerNet(p, mine_bb, p.backbone_channels, heads, nsplit=nsplit, use_graph=use_graph)
    m.load_state_dict(ref_model.state_dict(), strict=True)
    return m.eval()
