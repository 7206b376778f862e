"""CPU restatement of the InvPT forward (ViT backbone + InvPT decoder) -- TEST INFRASTRUCTURE.

Functional, state_dict-driven restatement of the reference forward in eval() mode (SyncBatchNorm =
BatchNorm with running statistics, DropPath = identity). Parameter names follow the reference
TransformerNet state_dict (`backbone.*`, `multi_task_decoder.*`, `heads.*`).

Pinned against the unmodified reference by tests/test_oracle.py (where /root/reference exists) and
against tests/golden/ip_*.pt everywhere.

Reference lines (relative to InvPT/):
  VisionTransformer.forward_features  models/transformers/vit.py:332-351 (Attention :184-196, Block :211-215)
  TransformerDecoder.forward          models/transformers/transformer_decoder.py:69-98 (ConvBlock :116-122)
  InvPT.forward                       models/transformers/invpt.py:502-544
  InvPTStage.forward                  models/transformers/invpt.py:400-417 (UpEmbed :32-43)
  InvPTBlock.forward                  models/transformers/invpt.py:290-312
  SelfAttention.forward               models/transformers/invpt.py:166-241
  TransformerNet.forward              models/transformer_net.py:22-38
"""
import math

import torch
import torch.nn.functional as F

D = "multi_task_decoder."


def _p(sd, name, x):
    return sd[name].to(x.dtype)


def _lin(x, sd, name):
    return F.linear(x, _p(sd, name + ".weight", x), _p(sd, name + ".bias", x))


def _ln(x, sd, name, eps):
    return F.layer_norm(x, (x.shape[-1],), _p(sd, name + ".weight", x), _p(sd, name + ".bias", x), eps)


def _conv(x, sd, name, **kw):
    b = sd.get(name + ".bias")
    return F.conv2d(x, _p(sd, name + ".weight", x), None if b is None else b.to(x.dtype), **kw)


def _bn(x, sd, name, eps=1e-5):
    return F.batch_norm(x, _p(sd, name + ".running_mean", x), _p(sd, name + ".running_var", x),
                        _p(sd, name + ".weight", x), _p(sd, name + ".bias", x), False, 0.0, eps)


def _tok(x):   # [B,C,h,w] -> [B,hw,C]
    return x.flatten(2).transpose(1, 2)


def _map(x, h, w):   # [B,hw,C] -> [B,C,h,w]
    return x.transpose(1, 2).reshape(x.shape[0], -1, h, w)


def stage_dims(cfg):
    d0 = cfg["embed_dim"] + cfg["pred_const"]
    return [d0, d0 // 2, d0 // 4]


def grid_of(cfg):
    return cfg["img_size"][0] // cfg["patch"], cfg["img_size"][1] // cfg["patch"]


def vit_forward(sd, cfg, img):
    """vit.py:332-351: returns the 4 multi-scale token maps [B,P,C] (cls dropped)."""
    C, Hh = cfg["C"], cfg["heads"]
    dh = C // Hh
    x = F.conv2d(img, _p(sd, "backbone.patch_embed.proj.weight", img), _p(sd, "backbone.patch_embed.proj.bias", img),
                 stride=cfg["patch"])
    x = _tok(x)
    B = x.shape[0]
    x = torch.cat([_p(sd, "backbone.cls_token", x).expand(B, -1, -1), x], dim=1)      # :334-336
    x = x + _p(sd, "backbone.pos_embed", x)                                            # :339
    feats = []
    for i in range(cfg["depth"]):
        pre = f"backbone.blocks.{i}."
        s = _ln(x, sd, pre + "norm1", 1e-6)
        N = s.shape[1]
        qkv = _lin(s, sd, pre + "attn.qkv").reshape(B, N, 3, Hh, dh).permute(2, 0, 3, 1, 4)
        attn = ((qkv[0] @ qkv[1].transpose(-2, -1)) * dh ** -0.5).softmax(-1)          # :189-190
        o = (attn @ qkv[2]).transpose(1, 2).reshape(B, N, C)
        x = x + _lin(o, sd, pre + "attn.proj")                                         # :213
        hdn = F.gelu(_lin(_ln(x, sd, pre + "norm2", 1e-6), sd, pre + "mlp.fc1"))
        x = x + _lin(hdn, sd, pre + "mlp.fc2")                                         # :214
        if i + 1 in cfg["select"]:
            feats.append(x[:, 1:])                                                     # :345-346
    feats.append(_ln(x, sd, "backbone.norm", 1e-6)[:, 1:])                             # :348-349
    return feats


def self_attention(sd, pre, x_maps, heads, kv_stride, prev_score):
    """invpt.py:193-241 for one InvPTBlock. x_maps: list over tasks of LN1'd maps [B,C,h,w].
    Returns (attention output tokens [B, T*(h/2)(w/2), C], fused pre-softmax score)."""
    T = len(x_maps)
    B, C, h, w = x_maps[0].shape
    q_in, kv_in = [], []
    for i, xm in enumerate(x_maps):
        q = F.conv2d(xm, _p(sd, f"{pre}conv_proj_q.{i}.conv.weight", xm), None, stride=2, padding=1, groups=C)  # :125-137
        q_in.append(_tok(_bn(q, sd, f"{pre}conv_proj_q.{i}.bn")))
        kv_in.append(_tok(F.avg_pool2d(xm, kv_stride, kv_stride, 0, ceil_mode=True)))                            # :139-147
    q = _lin(torch.cat(q_in, 1), sd, pre + "proj_q")
    kvt = torch.cat(kv_in, 1)
    k = _lin(kvt, sd, pre + "proj_k")
    v = _lin(kvt, sd, pre + "proj_v")
    d = C // heads
    sp = lambda t: t.reshape(B, -1, heads, d).transpose(1, 2)                         # 'b t (h d) -> b h t d'
    q, k, v = sp(q), sp(k), sp(v)
    score = (q @ k.transpose(-2, -1)) * C ** -0.5                                      # :92,:204 full-dim scale
    if prev_score is not None:                                                         # :207-229
        sh, sw = h // 4, w // 4
        Tk = prev_score.shape[-1]
        ups = []
        for i in range(T):
            s = prev_score[:, :, sh * sw * i: sh * sw * (i + 1), :]                    # [B,hd,sh*sw,Tk]
            s = s.permute(0, 1, 3, 2).reshape(B * heads, Tk, sh, sw)
            s = F.interpolate(s, scale_factor=2, mode="bilinear", align_corners=False)
            ups.append(s.reshape(B, heads, Tk, -1).permute(0, 1, 3, 2))
        both = torch.cat([score, torch.cat(ups, dim=2)], dim=1)                        # [B,2*hd,Lq,Tk]
        score = _conv(both, sd, pre + "fuse_attn")
    o = score.softmax(-1) @ v                                                          # :232-235
    o = o.transpose(1, 2).reshape(B, -1, C)
    return _lin(o, sd, pre + "proj"), score


def invpt_block(sd, pre, x_list, heads, kv_stride, prev_score):
    """invpt.py:290-312."""
    T = len(x_list)
    B, C, h, w = x_list[0].shape
    res = torch.cat([_tok(x) for x in x_list], dim=1)                                  # [B, T*hw, C]
    xn = _ln(res, sd, pre + "norm1", 1e-5)
    maps = [_map(xn[:, h * w * i: h * w * (i + 1)], h, w) for i in range(T)]
    a, score = self_attention(sd, pre + "attn.", maps, heads, kv_stride, prev_score)
    sh, sw = h // 2, w // 2
    up = []
    for i in range(T):
        m = _map(a[:, sh * sw * i: sh * sw * (i + 1)], sh, sw)
        up.append(_tok(F.interpolate(m, size=(h, w), mode="bilinear", align_corners=False)))   # :299-305
    x = res + torch.cat(up, dim=1)
    x = x + _lin(F.gelu(_lin(_ln(x, sd, pre + "norm2", 1e-5), sd, pre + "mlp.fc1")), sd, pre + "mlp.fc2")
    return [_map(x[:, h * w * i: h * w * (i + 1)], h, w) for i in range(T)], score


def up_embed(sd, pre, x):
    """UpEmbed (invpt.py:32-43): bilinear x2 -> [3x3 dilated(2) conv, BN, ReLU] x 2."""
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    x = F.relu(_bn(_conv(x, sd, pre + "proj.1", padding=2, dilation=2), sd, pre + "proj.2"))
    return F.relu(_bn(_conv(x, sd, pre + "proj.4", padding=2, dilation=2), sd, pre + "proj.5"))


def forward(sd, cfg, img, taps=None):
    """TransformerNet.forward (transformer_net.py:22-38): {task: [B,n_out,H,W], 'inter_preds': {...}}."""
    tasks = cfg["tasks"]
    gh, gw = grid_of(cfg)
    dims = stage_dims(cfg)
    feats = vit_forward(sd, cfg, img)
    # ---- TransformerDecoder.forward (transformer_decoder.py:69-98)
    maps = [_map(f, gh, gw) for f in feats]
    back = [
        F.conv_transpose2d(maps[0], _p(sd, D + "scale_embed.0.weight", img), _p(sd, D + "scale_embed.0.bias", img),
                           stride=2, padding=1, output_padding=1),                                       # :63
        _conv(maps[1], sd, D + "scale_embed.1", padding=1),                                              # :64
    ]   # scale_embed[2]'s output is never consumed (stage 0 has no patch_embed, invpt.py:401-412)
    h0, w0 = gh // cfg["down"], gw // cfg["down"]
    x = F.interpolate(maps[3], size=(h0, w0), mode="bilinear", align_corners=False)                      # :85-86
    x_list, inter = [], {}
    for t in tasks:
        y = x
        for j in range(2):                                                                               # ConvBlock x2
            pre = f"{D}preliminary_decoder.{t}.{j}."
            y = F.relu(_bn(_conv(y, sd, pre + "conv", padding=1), sd, pre + "bn1"))
        inter[t] = _conv(y, sd, f"{D}intermediate_head.{t}")                                             # :94
        if taps is not None:
            taps[f"ms_feat.{t}"], taps[f"inter.{t}"] = y, inter[t]
        x_list.append(_conv(torch.cat([y, inter[t]], 1), sd, f"{D}invpt.mix_proj.{t}.0"))                # invpt.py:509-513
    if taps is not None:
        taps["back0"], taps["back1"] = back
    # ---- InvPT.forward (invpt.py:516-543)
    th, tw = h0 * 8, w0 * 8
    ms = {t: 0 for t in tasks}
    score = None
    for i in range(3):
        if i > 0:                                                                                        # InvPTStage :401-412
            skip = back[1] if i == 1 else back[0]
            x_list = [up_embed(sd, f"{D}invpt.invpt_stages.{i}.patch_embed.{k}.", xm) + skip
                      for k, xm in enumerate(x_list)]
        x_list, score = invpt_block(sd, f"{D}invpt.invpt_stages.{i}.blocks.0.", x_list, 2, 2 ** (i + 1), score)
        h, w = x_list[0].shape[2:]
        xc = torch.cat([_tok(xm) for xm in x_list], dim=2)                                               # :524-525 channel cat
        xc = _map(_ln(xc, sd, f"{D}invpt.norm_mts.{i}", 1e-5), h, w)                                     # :526
        for k, t in enumerate(tasks):
            tx = xc[:, dims[i] * k: dims[i] * (k + 1)]
            if i > 0:
                tx = _conv(tx, sd, f"{D}invpt.redu_chan.{i}.{k}")                                        # :535-536
            ms[t] = ms[t] + F.interpolate(tx, size=(th, tw), mode="bilinear", align_corners=False)       # :537-539
        if taps is not None:
            taps[f"stage{i}.score"] = score
            for k, t in enumerate(tasks):
                taps[f"stage{i}.x.{t}"] = x_list[k]
    out = {}
    for t in tasks:
        pre = f"{D}invpt.mt_proj.{t}."
        y = F.relu(_bn(_conv(ms[t], sd, pre + "0", padding=1), sd, pre + "1"))                           # :541-543
        if taps is not None:
            taps[f"x_dict.{t}"] = y
        y = _conv(y, sd, f"heads.{t}.linear_pred")                                                       # MLPHead
        out[t] = F.interpolate(y, img.shape[-2:], mode="bilinear")                                       # transformer_net.py:35
    out["inter_preds"] = {t: F.interpolate(v, img.shape[-2:], mode="bilinear") for t, v in inter.items()}
    return out


def init_state_dict(cfg, seed=0, dtype=torch.float32):
    """Random parameters with the reference's names and shapes (deterministic CPU generator); BatchNorm
    running statistics are randomised so eval-mode BN folding is exercised."""
    g = torch.Generator().manual_seed(seed)
    C, tasks = cfg["C"], cfg["tasks"]
    T = len(tasks)
    gh, gw = grid_of(cfg)
    P = gh * gw
    E = cfg["embed_dim"]
    dims = stage_dims(cfg)
    sd = {}

    def tn(*shape, std=0.02):
        return (torch.randn(*shape, generator=g) * std).clamp_(-2, 2).to(dtype)

    def lin(name, o, i):
        sd[name + ".weight"] = tn(o, i)
        sd[name + ".bias"] = tn(o)

    def ln(name, n):
        sd[name + ".weight"] = (1 + 0.1 * torch.randn(n, generator=g)).to(dtype)
        sd[name + ".bias"] = (0.1 * torch.randn(n, generator=g)).to(dtype)

    def conv(name, o, i, k, bias=True, groups=1, wname=".weight"):
        bound = 1.0 / math.sqrt((i // groups) * k * k)
        sd[name + wname] = ((torch.rand(o, i // groups, k, k, generator=g) * 2 - 1) * bound).to(dtype)
        if bias:
            sd[name + ".bias"] = ((torch.rand(o, generator=g) * 2 - 1) * bound).to(dtype)

    def bn(name, n):
        sd[name + ".weight"] = (1 + 0.1 * torch.randn(n, generator=g)).to(dtype)
        sd[name + ".bias"] = (0.1 * torch.randn(n, generator=g)).to(dtype)
        sd[name + ".running_mean"] = (0.1 * torch.randn(n, generator=g)).to(dtype)
        sd[name + ".running_var"] = (1 + 0.2 * torch.rand(n, generator=g)).to(dtype)
        sd[name + ".num_batches_tracked"] = torch.tensor(0)

    sd["backbone.cls_token"] = tn(1, 1, C)
    sd["backbone.pos_embed"] = tn(1, P + 1, C)
    conv("backbone.patch_embed.proj", C, 3, cfg["patch"])
    for i in range(cfg["depth"]):
        pre = f"backbone.blocks.{i}."
        ln(pre + "norm1", C)
        lin(pre + "attn.qkv", 3 * C, C)
        lin(pre + "attn.proj", C, C)
        ln(pre + "norm2", C)
        lin(pre + "mlp.fc1", 4 * C, C)
        lin(pre + "mlp.fc2", C, 4 * C)
    ln("backbone.norm", C)
    for t in tasks:
        conv(f"{D}intermediate_head.{t}", cfg["num_output"][t], E, 1)
    for i in range(3):
        ln(f"{D}invpt.norm_mts.{i}", dims[i] * T)
    for i in range(3):
        for k in range(T):
            conv(f"{D}invpt.redu_chan.{i}.{k}", dims[0], dims[i], 1)
    for i in range(3):
        if i > 0:
            for k in range(T):
                pre = f"{D}invpt.invpt_stages.{i}.patch_embed.{k}.proj."
                conv(pre + "1", dims[i], dims[i - 1], 3, bias=False)
                bn(pre + "2", dims[i])
                conv(pre + "4", dims[i], dims[i], 3, bias=False)
                bn(pre + "5", dims[i])
        pre = f"{D}invpt.invpt_stages.{i}.blocks.0."
        ln(pre + "norm1", dims[i])
        ln(pre + "norm2", dims[i])
        lin(pre + "mlp.fc1", 4 * dims[i], dims[i])
        lin(pre + "mlp.fc2", dims[i], 4 * dims[i])
        for k in range(T):
            conv(f"{pre}attn.conv_proj_q.{k}.conv", dims[i], dims[i], 3, bias=False, groups=dims[i])
            bn(f"{pre}attn.conv_proj_q.{k}.bn", dims[i])
        for nm in ("proj_q", "proj_k", "proj_v", "proj"):
            lin(pre + "attn." + nm, dims[i], dims[i])
        conv(pre + "attn.fuse_attn", 2, 4, 1)
    ln(f"{D}invpt.norm_mt", dims[2] * T)
    for t in tasks:
        conv(f"{D}invpt.mt_proj.{t}.0", dims[0], dims[0], 3)
        bn(f"{D}invpt.mt_proj.{t}.1", dims[0])
    for t in tasks:
        conv(f"{D}invpt.mix_proj.{t}.0", dims[0], E + cfg["num_output"][t], 1)
    for t in tasks:
        for j, (ci, co) in enumerate([(C, C), (C, E)]):
            pre = f"{D}preliminary_decoder.{t}.{j}."
            conv(pre + "conv", co, ci, 3, bias=False)
            bn(pre + "bn1", co)
    sd[D + "scale_embed.0.weight"] = tn(C, dims[2], 3, 3, std=0.05)   # ConvTranspose2d weight [in, out, k, k]
    sd[D + "scale_embed.0.bias"] = tn(dims[2], std=0.05)
    conv(D + "scale_embed.1", dims[1], C, 3)
    conv(D + "scale_embed.2", dims[0], C, 3)
    for t in tasks:
        conv(f"heads.{t}.linear_pred", cfg["num_output"][t], dims[0], 1)
    return sd
