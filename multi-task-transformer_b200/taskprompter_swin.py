"""TaskPrompter with the Swin backbone (the reference's Cityscapes-3D model family) -- SURVEY.md section 8f N2.

Module classes, constructor arguments and parameter / buffer names mirror
TaskPrompter/models/transformers/taskprompter_swin.py (TP below), so reference checkpoints load unchanged:

  WindowAttention       TP:120-212    relative_position_bias_table, relative_position_index (buffer), qkv, proj
  SwinTransformerBlock  TP:215-405    norm1, attn, norm2, mlp, attn_mask (buffer), chan_q, chan_kv, token_trans, chan_proj,
                                      token_trans1
  PatchMerging          TP:408-481    reduction, norm, process_chan_attn, task_prompts_up, spa_attn_ds
  BasicLayer            TP:484-540    blocks, downsample
  TaskPrompterSwin      TP:542-774    patch_embed (+ norm), task_prompts, fea_fuse / fea_decode_spa / fea_decode_chan,
                                      multi_scale_fuse, layers, norm
used through taskprompter.TaskPrompterWrapper (models/taskprompter_wrapper.py:9-40) with ConvHead / DEConvHead.

The modules own parameters; the forward is `_SwinPlan`: packed weights + a fixed workspace + one launch sequence, captured
in a CUDA graph. What runs where:
  * every Linear / 1x1 / 3x3 convolution on the tcgen05 GEMM (mtt_gemm, the named block operators);
  * window partition with cyclic shift and zero padding, with the T task prompts replicated in front of every window:
    one gather kernel writing the joint window stream [B * nW * (T + ws^2), C] (TP:326-340, :177-181);
  * window attention with relative-position bias and shift mask on the patch x patch part (TP:183-204): one kernel per
    block over (window, head), exporting the un-scaled prompt-row logits (TP:189, :351-354);
  * window reverse / un-shift / crop, the residual add and the window-average of the prompt outputs (TP:210, :343-360):
    one scatter kernel; the prompt-row logits land directly in the [B, heads, T, T + H*W] layout the gating kernel of the
    ViT path reads;
  * channel attention between the prompts and the channels of the attention output (TP:372-396): chan_kv as a GEMM over
    the transposed map, the T x C logits / softmax / mixing in one small kernel;
  * PatchMerging (TP:430-472): 2x2 gather, LayerNorm, GEMM; the learned stride-2 3x3 down-sampling of the logit maps and
    the channel-logit up-projection as small direct kernels.
Exact algebraic re-orderings (results equal up to fp32 rounding): the bilinear x2 of the gated maps (TP:747-748) is applied
AFTER the 1x1 decode convs and fea_fuse[0] instead of before -- bilinear resampling and per-pixel linear maps commute --
which runs those convolutions at a quarter of the pixels.
Unsupported (raises): the '3ddet' task (FCOS3D head, needs mmdet3d), absolute position embedding (ape=True; no reference
config uses it). Eval mode only.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import ops
from .taskprompter import (PARITY, Mlp, PatchEmbed, _cached, _dev_ctx, _f32, _HeadSpace, _launch_head, _pack_head,
                           _Streams, _trunc_normal_, _version)

STRIDES = (8, 16, 32, 32)          # utils/common_config.py:37: level il lives at 1/STRIDES[il] of the image


# --------------------------------------------------------------------------------------------
# parameter containers (names = reference state_dict keys)
# --------------------------------------------------------------------------------------------
def relative_position_index(ws):
    """[ws*ws, ws*ws] index into the (2ws-1)^2 bias table (TP:146-157)."""
    ys, xs = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
    ys, xs = ys.flatten(), xs.flatten()
    dy = ys[:, None] - ys[None, :] + ws - 1
    dx = xs[:, None] - xs[None, :] + ws - 1
    return dy * (2 * ws - 1) + dx


def shifted_window_mask(Hp, Wp, ws, shift):
    """[nW, ws*ws, ws*ws]: 0 where two tokens of a cyclically shifted window come from the same image region, -100
    otherwise (TP:276-290)."""
    region = torch.zeros(Hp, Wp)
    cuts = lambda n: [(0, n - ws), (n - ws, n - shift), (n - shift, n)]
    k = 0
    for (y0, y1) in cuts(Hp):
        for (x0, x1) in cuts(Wp):
            region[y0:y1, x0:x1] = k
            k += 1
    win = region.reshape(Hp // ws, ws, Wp // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = win[:, None, :] - win[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


class WindowAttention(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window_size - 1) ** 2, num_heads))
        self.register_buffer("relative_position_index", relative_position_index(window_size))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        _trunc_normal_(self.relative_position_bias_table, std=.02)


class SwinTransformerBlock(nn.Module):
    def __init__(self, last_block, p, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4.,
                 qkv_bias=True):
        super().__init__()
        self.LAST_BLOCK_FLAG = last_block
        self.dim, self.input_resolution, self.num_heads = dim, tuple(input_resolution), num_heads
        self.window_size, self.shift_size = window_size, shift_size
        if min(self.input_resolution) <= self.window_size:       # TP:243-246
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        H, W = self.input_resolution
        ws = self.window_size
        self.padded = (H + (ws - H % ws) % ws, W + (ws - W % ws) % ws)
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, ws, num_heads, qkv_bias)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        mask = shifted_window_mask(self.padded[0], self.padded[1], ws, self.shift_size) if self.shift_size > 0 else None
        self.register_buffer("attn_mask", mask)
        ce = p.chan_embed_dim
        self.chan_q = nn.Linear(ce, ce, bias=qkv_bias)
        self.chan_kv = nn.Linear(H * W, ce * 2, bias=qkv_bias)
        self.token_trans = nn.Linear(dim, ce)
        if not last_block:
            self.chan_proj = nn.Linear(ce, ce)
            self.token_trans1 = nn.Linear(ce, dim)


class PatchMerging(nn.Module):
    def __init__(self, p, num_heads, input_resolution, dim):
        super().__init__()
        self.input_resolution, self.dim = tuple(input_resolution), dim
        T = len(p.TASKS.NAMES) * p.prompt_len
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)
        self.process_chan_attn = nn.Linear(dim, 2 * dim, bias=False)
        self.task_prompts_up = nn.Linear(dim, 2 * dim, bias=False)
        self.spa_attn_ds = nn.Conv2d(num_heads * T, num_heads * T, kernel_size=3, padding=1, stride=2)


class BasicLayer(nn.Module):
    def __init__(self, last_layer, p, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4., qkv_bias=True,
                 downsample=False):
        super().__init__()
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(last_layer and i == depth - 1, p, dim, input_resolution, num_heads, window_size,
                                 0 if i % 2 == 0 else window_size // 2, mlp_ratio, qkv_bias) for i in range(depth)])
        self.downsample = PatchMerging(p, num_heads, input_resolution, dim) if downsample else None


class SwinPatchEmbed(PatchEmbed):
    """timm PatchEmbed with norm_layer=LayerNorm (TP:592-595, patch_norm=True)."""

    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__(img_size, patch_size, in_chans, embed_dim)
        self.norm = nn.LayerNorm(embed_dim)


class TaskPrompterSwin(nn.Module):
    """TP:542-666 (same constructor arguments that matter for the forward; `p` needs TASKS.NAMES, prompt_len,
    chan_embed_dim, chan_nheads, img_ds_ratio, level_embed_dim, final_embed_dim, backbone_channels, ori_spatial_dim)."""

    def __init__(self, p, img_size=224, patch_size=4, in_chans=3, embed_dim=96, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4., qkv_bias=True, ape=False, **_unused):
        super().__init__()
        if ape:
            raise NotImplementedError("mtt_b200 TaskPrompterSwin: absolute position embedding is not supported")
        if isinstance(img_size, int):
            img_size = (img_size, img_size)
        tasks = list(p.TASKS.NAMES)
        if "3ddet" in tasks:
            raise NotImplementedError("mtt_b200 TaskPrompterSwin: the '3ddet' task needs the FCOS3D head (mmdet3d): "
                                      "SURVEY.md 8f N4")
        self.p = p
        self.num_layers = len(depths)
        self.embed_dim, self.depths, self.heads, self.window_size = embed_dim, tuple(depths), tuple(num_heads), window_size
        self.patch_size, self.in_chans = patch_size, in_chans
        self.img_ds_ratio = p.img_ds_ratio
        self.full_img_size = tuple(img_size)
        self.resolution = [[int(s[0] * self.img_ds_ratio), int(s[1] * self.img_ds_ratio)] for s in p.ori_spatial_dim]
        ds_size = [int(s * self.img_ds_ratio) for s in img_size]                         # TP:590
        self.patch_embed = SwinPatchEmbed(ds_size, patch_size, in_chans, embed_dim)
        self.patch_grid = self.patch_embed.grid_size
        for i in range(self.num_layers - 1):              # PatchMerging halves each map (TP taskprompter_swin.py:438
            gh, gw = self.patch_grid[0] // 2 ** i, self.patch_grid[1] // 2 ** i     # asserts "x size (H*W) are not even")
            if gh % 2 or gw % 2:
                raise ValueError(f"TaskPrompterSwin: the stage-{i} token map {gh} x {gw} (image {tuple(img_size)} x ratio "
                                 f"{self.img_ds_ratio} / patch {patch_size}) must be even on both axes for PatchMerging; the "
                                 "reference asserts the same")
        cnh = int(round(math.sqrt(p.chan_nheads)))
        for i in range(self.num_layers):                  # the channel gate of level i is a cnh x cnh grid of windows over its map
            gh, gw = self.patch_grid[0] // 2 ** i, self.patch_grid[1] // 2 ** i     # (TP taskprompter_swin.py:738-763)
            if cnh * cnh != p.chan_nheads or gh % cnh or gw % cnh:
                raise ValueError(f"TaskPrompterSwin: chan_nheads={p.chan_nheads} must be a perfect square whose root divides "
                                 f"every level's token map (level {i}: {gh} x {gw}); the reference fails on such sizes too")
        assert p.prompt_len == 1, "prompt_len != 1 is unsupported (as in the reference's channel branch)"
        self.prompts_len = len(tasks) * p.prompt_len
        self.task_prompts = nn.Parameter(torch.ones(self.prompts_len, embed_dim))
        _trunc_normal_(self.task_prompts, mean=1., std=1.)
        Lv, f = p.level_embed_dim, p.final_embed_dim
        self.fea_fuse, self.fea_decode_spa, self.fea_decode_chan = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for il in range(self.num_layers):
            cur = p.backbone_channels[il]
            self.fea_fuse.append(nn.ModuleDict({t: nn.Sequential(
                nn.Conv2d(Lv * 2, f, 1), nn.Conv2d(f, f, 3, padding=1), nn.BatchNorm2d(f), nn.GELU(),
                nn.Conv2d(f, f, 3, padding=1)) for t in tasks}))
            self.fea_decode_spa.append(nn.ModuleDict({t: nn.Sequential(nn.Conv2d(cur, Lv, 1)) for t in tasks}))
            self.fea_decode_chan.append(nn.ModuleDict({t: nn.Sequential(nn.Conv2d(cur, Lv, 1)) for t in tasks}))
        self.multi_scale_fuse = nn.ModuleDict({t: nn.Conv2d(f, f, 3, padding=1) for t in tasks})
        self.layers = nn.Sequential(*[
            BasicLayer(i == self.num_layers - 1, p, embed_dim * 2 ** i,
                       (self.patch_grid[0] // 2 ** i, self.patch_grid[1] // 2 ** i), depths[i], num_heads[i], window_size,
                       mlp_ratio, qkv_bias, downsample=i < self.num_layers - 1) for i in range(self.num_layers)])
        self.norm = nn.LayerNorm(embed_dim * 2 ** (self.num_layers - 1))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        raise RuntimeError("TaskPrompterSwin runs fused inside TaskPrompterWrapper.forward; call the wrapper")


# --------------------------------------------------------------------------------------------
# packed weights
# --------------------------------------------------------------------------------------------
def _lin(mod, device, ns):
    w = ops.pack_weight(_f32(mod.weight, device).reshape(mod.weight.shape[0], -1), ns)
    b = _f32(mod.bias, device) if mod.bias is not None else None
    return w, b


def _pack_swin_block(blk, device, ns):
    def build():
        f = lambda t: _f32(t, device)
        w = SimpleNamespace()
        w.n1w, w.n1b, w.n2w, w.n2b = f(blk.norm1.weight), f(blk.norm1.bias), f(blk.norm2.weight), f(blk.norm2.bias)
        w.eps = blk.norm1.eps
        w.qkv, w.qkv_b = _lin(blk.attn.qkv, device, ns)
        w.proj, w.proj_b = _lin(blk.attn.proj, device, ns)
        ws = blk.window_size
        L = ws * ws
        idx = blk.attn.relative_position_index.reshape(-1).to(device)
        # bias[h, query, key] (TP:193-195) and the shift mask [nW, query, key], both stored TRANSPOSED for the kernel
        w.biasT = f(blk.attn.relative_position_bias_table)[idx].reshape(L, L, blk.num_heads).permute(2, 1, 0).contiguous()
        w.maskT = f(blk.attn_mask).transpose(1, 2).contiguous() if blk.attn_mask is not None else None
        w.fc1, w.fc1_b = _lin(blk.mlp.fc1, device, ns)
        w.fc2, w.fc2_b = _lin(blk.mlp.fc2, device, ns)
        w.cq, w.cq_b = _lin(blk.chan_q, device, ns)
        w.ckv, w.ckv_b = _lin(blk.chan_kv, device, ns)
        w.tt, w.tt_b = _lin(blk.token_trans, device, ns)
        w.last = blk.LAST_BLOCK_FLAG
        if not w.last:
            w.cp, w.cp_b = _lin(blk.chan_proj, device, ns)
            w.tt1, w.tt1_b = _lin(blk.token_trans1, device, ns)
        return w
    return _cached(blk, ("pack", device, ns), build)


def _pack_merge(dsm, device, ns):
    def build():
        f = lambda t: _f32(t, device)
        w = SimpleNamespace()
        w.red, _ = _lin(dsm.reduction, device, ns)
        w.nw, w.nb, w.eps = f(dsm.norm.weight), f(dsm.norm.bias), dsm.norm.eps
        w.pca = f(dsm.process_chan_attn.weight)                                            # [2C, C] fp32 (small kernel)
        w.tpu, _ = _lin(dsm.task_prompts_up, device, ns)
        w.ds_w, w.ds_b = f(dsm.spa_attn_ds.weight), f(dsm.spa_attn_ds.bias)               # [HT, HT, 3, 3]
        return w
    return _cached(dsm, ("pack", device, ns), build)


def _pack_swin_decoder(bb, tasks, device, ns):
    def build():
        f = lambda t: _f32(t, device)
        p = bb.p
        Lv, ff = p.level_embed_dim, p.final_embed_dim
        Lv_pad = ops.round_up(Lv, 8)
        W = SimpleNamespace(levels=[], msf=[])
        for il in range(bb.num_layers):
            lv = []
            for t in tasks:
                tw = SimpleNamespace()
                tw.spa, tw.spa_b = _lin(bb.fea_decode_spa[il][t][0], device, ns)
                tw.chan, tw.chan_b = _lin(bb.fea_decode_chan[il][t][0], device, ns)
                fu = bb.fea_fuse[il][t]
                w0 = f(fu[0].weight).reshape(ff, 2 * Lv)
                w0p = torch.zeros(ff, 2 * Lv_pad, device=device)      # K laid out like the `cat` buffer
                w0p[:, :Lv] = w0[:, :Lv]
                w0p[:, Lv_pad:Lv_pad + Lv] = w0[:, Lv:]
                tw.f0, tw.f0_b = ops.pack_weight(w0p, ns), f(fu[0].bias)
                tw.f1, tw.f1_b = ops.pack_conv_weight(f(fu[1].weight), fu[1].bias, fu[2], ns)   # conv3x3 + eval BN
                tw.f4, tw.f4_b = ops.pack_conv_weight(f(fu[4].weight), fu[4].bias, None, ns)    # 3x3 here (TP:630)
                lv.append(tw)
            W.levels.append(lv)
        for t in tasks:
            W.msf.append(ops.pack_conv_weight(f(bb.multi_scale_fuse[t].weight), bb.multi_scale_fuse[t].bias, None, ns))
        return W
    return _cached(bb, ("decoder", device, ns, tuple(tasks)), build)


# --------------------------------------------------------------------------------------------
# the fused forward
# --------------------------------------------------------------------------------------------
class _SwinPlan:
    def __init__(self, bb, heads, tasks, target, B, device, nsplit, mode="full"):
        ops._L.check(ops._L.load().mtt_device_check(), "mtt_device_check")
        if mode not in ("full",):
            raise NotImplementedError("mtt_b200 TaskPrompterSwin: only the wrapper forward is built (no predict())")
        device = torch.device(device)
        self.bb, self.heads, self.tasks, self.target = bb, heads, list(tasks), target
        self.B, self.dev, self.ns, self.mode = B, device, nsplit, mode
        self.T = T = len(self.tasks)
        p = bb.p
        self.ce = ce = p.chan_embed_dim
        self.r = int(round(math.sqrt(ce)))
        self.nh = self.nw = int(round(math.sqrt(p.chan_nheads)))
        assert self.r * self.r == ce and self.r % self.nh == 0
        self.Lv, self.f = p.level_embed_dim, p.final_embed_dim
        self.Lv_pad, self.f_ld = ops.round_up(self.Lv, 8), ops.round_up(self.f, 8)
        self.img = bb.full_img_size
        self.ds_img = tuple(bb.patch_embed.img_size)
        self.graph, self.static_in = None, None
        self.streams = _Streams(device, max(T, 2))
        ns = nsplit
        S = lambda r, c, **kw: ops.Split(r, c, device, ns, **kw)
        z = lambda *s: torch.zeros(*s, device=device, dtype=torch.float32)
        with _dev_ctx(device):
            self._pack()
            E, patch = bb.embed_dim, bb.patch_size
            gh, gw = bb.patch_grid
            self.img_ds = z(B, bb.in_chans, *self.ds_img) if self.ds_img != self.img else None
            self.cols = S(B * gh * gw, patch * patch * bb.in_chans)
            self.x0 = z(B * gh * gw, E)
            # ---- per stage
            self.st = []
            for i, layer in enumerate(bb.layers):
                C = E * 2 ** i
                H, W = gh // 2 ** i, gw // 2 ** i
                blk0 = layer.blocks[0]
                ws, heads_i = blk0.window_size, blk0.num_heads
                Hp, Wp = blk0.padded
                nW = (Hp // ws) * (Wp // ws)
                L, wl = H * W, ws * ws
                rows_w = B * nW * (T + wl)
                s = SimpleNamespace(C=C, H=H, W=W, L=L, ws=ws, heads=heads_i, Hp=Hp, Wp=Wp, nW=nW, wl=wl, rows_w=rows_w)
                s.x = z(B * L, C)
                s.p = z(B * T, C)
                s.xn32, s.pn32 = z(B * L, C), z(B * T, C)
                s.ps, s.chan_ps = S(B * T, C), S(B * T, ce)
                s.sw = S(rows_w, C)
                s.qkv = S(rows_w, 3 * C)
                s.ao = S(rows_w, C)
                s.raw = z(B * nW, heads_i, T, wl)
                s.o32 = z(rows_w, C)
                s.xa32 = z(B * L, C)
                s.logits = z(B, heads_i, T, T + L)                 # prompt-row logits in the gating kernel's layout
                s.q32 = z(B * T, ce)
                s.xat = S(B * C, L, zero=True)
                s.kv32 = z(B * C, 2 * ce)
                # chan_kv is Linear(H*W -> 2 ce) over B*C rows: few output tiles, very long K. Split K so that about one
                # wave of 148 CTAs is busy (K chunks are multiples of 64, at least 512 columns each)
                tiles = -(-(B * C) // 128) * -(-(2 * ce) // 128)
                s.kchunks = max(1, min(32, 148 // tiles, L // 512))
                s.kv_part = z(s.kchunks, B * C, 2 * ce) if s.kchunks > 1 else None
                s.co32, s.cos = z(B * T, ce), S(B * T, ce)
                s.t1 = S(B * T, ce)
                s.rc = z(B, T, C, self.nh, self.nw)
                hid = layer.blocks[0].mlp.fc1.out_features
                s.ws_mlp = ops.workspace(ops.workspace_bytes(ops._L.OP_LN_MLP_RESIDUAL, rows=B * L, Cdim=C, hidden=hid,
                                                             nsplit=ns), device)
                s.ws_mlp_p = ops.workspace(ops.workspace_bytes(ops._L.OP_LN_MLP_RESIDUAL, rows=B * T, Cdim=C, hidden=hid,
                                                               nsplit=ns), device)
                if layer.downsample is not None:
                    s.m32 = z(B * L // 4, 4 * C)
                    s.ms = S(B * L // 4, 4 * C)
                    s.logits_ds = z(B, heads_i, T, T + L // 4)
                    s.rc_up = z(B, T, 2 * C, self.nh, self.nw)
                self.st.append(s)
            last = self.st[-1]
            self.xfin = z(B * last.L, last.C)
            # ---- decoder levels: level il lives on the map AFTER stage il's merging (the last level: the final norm)
            self.lv = []
            for il in range(bb.num_layers):
                h, w = bb.resolution[il]
                Cl = p.backbone_channels[il]
                d = SimpleNamespace(h=h, w=w, P=h * w, C=Cl, heads=self.st[il].heads)
                d.ws_gate = ops.workspace(ops.workspace_bytes(ops._L.OP_GATED_CONV1X1, rows=B * d.P, Cdim=Cl, nsplit=ns, T=T),
                                          device)
                d.cat = [S(B * d.P, 2 * self.Lv_pad, zero=True) for _ in range(T)]
                d.g32 = [z(B * d.P, self.f_ld) for _ in range(T)]
                d.up = [S(B * 4 * d.P, self.f, zero=True) for _ in range(T)]
                d.mid = [S(B * 4 * d.P, self.f, zero=True) for _ in range(T)]
                d.out32 = None if il == 0 else [z(B * 4 * d.P, self.f_ld) for _ in range(T)]
                self.lv.append(d)
            h0, w0 = 2 * self.lv[0].h, 2 * self.lv[0].w
            self.fh, self.fw = h0, w0
            self.acc = [z(B * h0 * w0, self.f_ld) for _ in range(T)]
            self.accs = [S(B * h0 * w0, self.f, zero=True) for _ in range(T)]
            self.hs = [_HeadSpace(hw, B, h0, w0, device, ns) for hw in self.Wh]
            oh, ow = self.target if self.target is not None else self.img
            self.out_hw = (oh, ow)
            self.out = {t: z(B, hw.n_out, oh, ow) for t, hw in zip(self.tasks, self.Wh)}

    def _pack(self):
        bb, dev, ns = self.bb, self.dev, self.ns
        f = lambda t: _f32(t, dev)

        def stem():
            W = SimpleNamespace()
            W.pe_w = ops.pack_weight(f(bb.patch_embed.proj.weight).reshape(bb.embed_dim, -1), ns)
            W.pe_b = f(bb.patch_embed.proj.bias)
            W.pnw, W.pnb, W.pneps = f(bb.patch_embed.norm.weight), f(bb.patch_embed.norm.bias), bb.patch_embed.norm.eps
            W.prompts = f(bb.task_prompts)
            W.nw, W.nb, W.neps = f(bb.norm.weight), f(bb.norm.bias), bb.norm.eps
            return W
        self.Ws = _cached(bb, ("stem", dev, ns), stem)
        self.Wb = [[_pack_swin_block(blk, dev, ns) for blk in layer.blocks] for layer in bb.layers]
        self.Wm = [_pack_merge(layer.downsample, dev, ns) if layer.downsample is not None else None for layer in bb.layers]
        self.Wd = _pack_swin_decoder(bb, self.tasks, dev, ns)
        self.Wh = [_pack_head(self.heads[t], dev, ns) for t in self.tasks]
        self.version = _version(bb) + _version(self.heads)

    @property
    def serial(self):
        return self.streams.serial

    @serial.setter
    def serial(self, v):
        self.streams.serial = bool(v)

    # ------------------------------------------------------------------------------------------
    def _block(self, s, w, blk):
        """One SwinTransformerBlock with task prompts (TP:310-405) on stage buffers s.x [B*L, C] / s.p [B*T, C]."""
        B, T, C, ce = self.B, self.T, s.C, self.ce
        shift = blk.shift_size
        ops.layernorm(s.x, w.n1w, w.n1b, w.eps, out_f32=s.xn32)                                    # :322
        ops.layernorm(s.p, w.n1w, w.n1b, w.eps, out_f32=s.pn32)                                    # :317
        ops.split_f32(s.p, self.ns, out=s.ps)
        ops.gemm(s.ps, w.tt, bias=w.tt_b, out_split=s.chan_ps)                                     # :319 token_trans
        ops.swin_window_gather(s.xn32, s.pn32, s.sw, B=B, H=s.H, W=s.W, Cdim=C, T=T, ws=s.ws, shift=shift)   # :326-340
        ops.gemm(s.sw, w.qkv, bias=w.qkv_b, out_split=s.qkv)                                       # :183
        ops.swin_window_attention(s.qkv, s.ao, s.raw, w.biasT, w.maskT, BW=B * s.nW, nW=s.nW, T=T, L=s.wl,
                                  heads=s.heads, scale=(C // s.heads) ** -0.5)                      # :185-206
        ops.gemm(s.ao, w.proj, bias=w.proj_b, out_f32=s.o32)                                       # :207
        ops.swin_window_scatter(s.o32, s.raw, s.xa32, s.x, s.p, s.logits, B=B, H=s.H, W=s.W, Cdim=C, T=T, ws=s.ws,
                                shift=shift, heads=s.heads, last=w.last)                            # :210, :343-360, :399
        # The prompt path -- channel attention between the prompts and the channels of the attention output (:372-396),
        # then the prompt update and the prompts' own MLP (:403-404) -- is a chain of small launches on B*T rows that
        # only needs xa: it runs on a side stream next to the MLP of the B*H*W patch rows (:400) and joins at the end.
        def prompt_path():
            ops.gemm(s.chan_ps, w.cq, bias=w.cq_b, out_f32=s.q32)
            ops.transpose_split(s.xa32, s.xat, B=B, L=s.L, Cdim=C)
            if s.kchunks > 1:   # C rows x (H*W) columns: a handful of M tiles with thousands of K blocks -> split K
                ops.gemm_splitk(s.xat, w.ckv, s.kv_part, s.kv32, K=s.L, bias=w.ckv_b, chunks=s.kchunks)
            else:
                ops.gemm(s.xat, w.ckv, K=s.L, bias=w.ckv_b, out_f32=s.kv32)
            ops.swin_chan_attention(s.q32, s.kv32, s.co32, s.cos, s.rc, B=B, T=T, Cdim=C, ce=ce, nh=self.nh, nw=self.nw)
            if not w.last:
                ops.gemm(s.cos, w.cp, bias=w.cp_b, out_split=s.t1)                                 # chan_proj
                ops.gemm(s.t1, w.tt1, bias=w.tt1_b, residual=s.p, out_f32=s.p)                     # token_trans1; :403
                ops.ln_mlp_residual(s.p, w.n2w, w.n2b, w.eps, w.fc1, w.fc1_b, w.fc2, w.fc2_b, s.ws_mlp_p)   # :404

        self.streams.par([prompt_path,
                          lambda: ops.ln_mlp_residual(s.x, w.n2w, w.n2b, w.eps, w.fc1, w.fc1_b, w.fc2, w.fc2_b, s.ws_mlp)])

    def _merge(self, i):
        """PatchMerging (TP:430-472): stage i -> the inputs of stage i + 1 and of decoder level i."""
        B, T = self.B, self.T
        s, n, w = self.st[i], self.st[i + 1], self.Wm[i]
        ops.swin_merge_gather(s.x, s.m32, B=B, H=s.H, W=s.W, Cdim=s.C)                             # :441-447
        ops.layernorm(s.m32, w.nw, w.nb, w.eps, out_split=s.ms)
        ops.gemm(s.ms, w.red, out_f32=n.x)                                                         # :449-450
        ops.conv3x3_s2_maps(s.logits, w.ds_w, w.ds_b, s.logits_ds, B=B, Cin=s.heads * T, H=s.H, W=s.W,
                            in_stride=T + s.L, in_offset=T, out_stride=T + s.L // 4, out_offset=T)  # :458-460
        ops.swin_chan_up(s.rc, w.pca, s.rc_up, BT=B * T, Cdim=s.C, nwin=self.nh * self.nw)         # :463-466
        ops.split_f32(s.p, self.ns, out=s.ps)
        ops.gemm(s.ps, w.tpu, out_f32=n.p)                                                         # :469

    def _level(self, il, x_src, logits, rc):
        """cal_task_feature (TP:721-774) at level il on X = x_src [B*P, C]."""
        B, T, d = self.B, self.T, self.lv[il]
        lvw = self.Wd.levels[il]
        ops.gated_conv1x1(x_src, d.P, 0, logits, rc,
                          [(tw.spa, tw.spa_b, tw.chan, tw.chan_b, d.cat[ti]) for ti, tw in enumerate(lvw)],
                          self.Lv, self.Lv_pad, d.ws_gate, B=B, T=T, N=T + d.P, H=d.heads, Cdim=d.C, gh=d.h, gw=d.w,
                          nh=self.nh, nw=self.nw)                                                   # :736-751 (1x1 first)
        ops.gemm_grouped([(d.cat[ti], tw.f0, dict(bias=tw.f0_b, out_f32=d.g32[ti][:, :self.f], N=self.f))
                          for ti, tw in enumerate(lvw)])                                           # fea_fuse[0]
        self.streams.par([lambda ti=ti: ops.bilinear(d.g32[ti], self.f_ld, B, d.h, d.w, self.f, 2 * d.h, 2 * d.w,
                                                     out_split=d.up[ti]) for ti in range(T)])     # :747-748 (moved)
        ops.gemm_grouped([(d.up[ti], tw.f1, dict(N=self.f, K=self.f, bias=tw.f1_b, act=ops.ACT_GELU, out_split=d.mid[ti],
                                                 conv=(B, 2 * d.h, 2 * d.w, 3, 1))) for ti, tw in enumerate(lvw)])
        dst = self.acc if il == 0 else d.out32
        ops.gemm_grouped([(d.mid[ti], tw.f4, dict(N=self.f, K=self.f, bias=tw.f4_b, out_f32=dst[ti][:, :self.f],
                                                  conv=(B, 2 * d.h, 2 * d.w, 3, 1))) for ti, tw in enumerate(lvw)])
        if il > 0:                                                                                  # TP:713-716
            self.streams.par([lambda ti=ti: ops.bilinear(d.out32[ti], self.f_ld, B, 2 * d.h, 2 * d.w, self.f, self.fh,
                                                         self.fw, out_f32=self.acc[ti][:, :self.f], accumulate=True)
                              for ti in range(T)])

    def _head_chain(self, ti, t, hw, hs):
        B = self.B
        oh, ow = self.out_hw
        wm, bm = self.Wd.msf[ti]
        ops.split_f32(self.acc[ti][:, :self.f], self.ns, out=self.accs[ti])
        ops.gemm(self.accs[ti], wm, N=self.f, K=self.f, bias=bm, out_split=hs.up, conv=(B, self.fh, self.fw, 3, 1))  # :717
        _launch_head(hs, hw)
        ops.bilinear(hs.pred, hs.pred.stride(0), B, hs.ph, hs.pw, hw.n_out, oh, ow, out_nchw=self.out[t])  # wrapper :35

    def _launch(self, img):
        B, T, bb, W = self.B, self.T, self.bb, self.Ws
        if self.img_ds is not None:                                                                 # TP:676-677
            h, w = self.img
            ops.bilinear(img.view(B * bb.in_chans * h * w, 1), 1, B * bb.in_chans, h, w, 1, self.ds_img[0], self.ds_img[1],
                         out_nchw=self.img_ds.view(B * bb.in_chans, 1, *self.ds_img))
            img = self.img_ds
        s0 = self.st[0]
        ops.im2col_patch(img, bb.patch_size, self.cols)
        ops.gemm(self.cols, W.pe_w, bias=W.pe_b, out_f32=self.x0)                                  # TP:679
        ops.layernorm(self.x0, W.pnw, W.pnb, W.pneps, out_f32=s0.x)                                # patch_embed.norm
        ops.broadcast_rows(W.prompts, s0.p, B, T)                                                  # :685
        n_stage = len(self.st)
        for i, s in enumerate(self.st):
            for j, blk in enumerate(bb.layers[i].blocks):
                self._block(s, self.Wb[i][j], blk)
            if i < n_stage - 1:
                self._merge(i)
                n = self.st[i + 1]
                self._level(i, n.x, s.logits_ds, s.rc_up)                                          # :702-707
        last = self.st[-1]
        ops.layernorm(last.x, W.nw, W.nb, W.neps, out_f32=self.xfin)                               # :709
        self._level(n_stage - 1, self.xfin, last.logits, last.rc)
        self.streams.par([lambda ti=ti, t=t, hw=hw, hs=hs: self._head_chain(ti, t, hw, hs)
                          for ti, (t, hw, hs) in enumerate(zip(self.tasks, self.Wh, self.hs))])

    def run(self, x, graph=True):
        if tuple(x.shape[1:]) != (self.bb.in_chans, *self.img) or x.dtype != torch.float32:
            raise ValueError(f"expected fp32 input [B,3,{self.img[0]},{self.img[1]}], got {tuple(x.shape)} {x.dtype}")
        with _dev_ctx(self.dev):
            if _version(self.bb) + _version(self.heads) != self.version:
                self._pack()
                self.graph = None
            if not graph:
                self._launch(x.contiguous())
                return dict(self.out)
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
            return dict(self.out)

    def launches_per_forward(self):
        with _dev_ctx(self.dev):
            n0 = ops.launch_count()
            self._launch(self.static_in if self.static_in is not None else
                         torch.zeros(self.B, 3, *self.img, device=self.dev))
            return ops.launch_count() - n0


def build_from_config(cfg, nsplit=PARITY, use_graph=True):
    """cfg: dict as in configs.taskprompter_swin() (mirrors TP/utils/common_config.py:34-41,64-90)."""
    from .taskprompter import ConvHead, DEConvHead, TaskPrompterWrapper
    h, w = cfg["img_size"]
    E = cfg["embed_dim"]
    p = SimpleNamespace(TASKS=SimpleNamespace(NAMES=list(cfg["tasks"]), NUM_OUTPUT=dict(cfg["num_output"])),
                        prompt_len=cfg.get("prompt_len", 1), chan_embed_dim=cfg["chan_embed_dim"],
                        chan_nheads=cfg["chan_nheads"], img_ds_ratio=cfg["img_ds_ratio"],
                        level_embed_dim=cfg["level_embed_dim"], final_embed_dim=cfg["f"],
                        backbone_channels=[2 * E, 4 * E, 8 * E, 8 * E],
                        ori_spatial_dim=[[h // st, w // st] for st in STRIDES])
    if "dd_label_map_size" in cfg:
        p.dd_label_map_size = tuple(cfg["dd_label_map_size"])
    bb = TaskPrompterSwin(p, img_size=(h, w), patch_size=cfg["patch"], embed_dim=E, depths=tuple(cfg["depths"]),
                          num_heads=tuple(cfg["heads"]), window_size=cfg["window"])
    head_cls = DEConvHead if cfg.get("head", "conv") == "deconv" else ConvHead
    heads = nn.ModuleDict({t: head_cls(cfg["f"], cfg["num_output"][t]) for t in cfg["tasks"]})
    return TaskPrompterWrapper(p, bb, heads, nsplit=nsplit, use_graph=use_graph)
