"""CPU restatement of the Swin-backbone TaskPrompter forward -- TEST INFRASTRUCTURE (there is
no CUDA path for it yet, this pins the oracle the kernels will be checked against).

Written from the algorithm of TaskPrompter/models/transformers/taskprompter_swin.py (TP = that file), as pure
functions over a state dict with the reference's parameter names; index tables (relative-position index, shifted-
window mask) are recomputed from the config rather than read from the reference's buffers.

  TaskPrompterSwin.forward          TP:674-718        forward()
  SwinTransformerBlock.forward      TP:310-405        block()
  WindowAttention.forward           TP:167-212        window_attention()
  PatchMerging.forward              TP:430-472        patch_merging()
  TaskPrompterSwin.cal_task_feature TP:721-774        task_features()
  TaskPrompterWrapper.forward       models/taskprompter_wrapper.py:22-40
  ConvHead / DEConvHead             models/transformers/taskprompter.py:688-715 (oracle.taskprompter_ref)

Pinned by tests/test_oracle.py::test_swin_* against the unmodified reference (<= 5e-6) and against
tests/golden/tps_*.pt."""
import math

import torch
import torch.nn.functional as F

from oracle.taskprompter_ref import _bn, _conv, conv_head, deconv_head

STRIDES = (8, 16, 32, 32)          # utils/common_config.py:37: level il lives at 1/STRIDES[il] of the image


def _lin(x, sd, name):
    b = sd.get(name + ".bias")
    return F.linear(x, sd[name + ".weight"].to(x.dtype), None if b is None else b.to(x.dtype))


def _ln(x, sd, name, eps=1e-5):     # nn.LayerNorm default eps (the Swin file does not override it)
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype), eps)


def _mlp(sd, pre, x):
    return _lin(F.gelu(_lin(x, sd, pre + "mlp.fc1")), sd, pre + "mlp.fc2")


def stage_geometry(cfg, i):
    """(dim, (H, W), window, shifts per block) of stage i (TP:633-646, :245-249, :508-514)."""
    ratio = cfg["img_ds_ratio"]
    gh = int(cfg["img_size"][0] * ratio) // cfg["patch"]
    gw = int(cfg["img_size"][1] * ratio) // cfg["patch"]
    H, W = gh // 2 ** i, gw // 2 ** i
    ws = cfg["window"]
    clipped = min(H, W) <= ws
    if clipped:
        ws = min(H, W)
    shifts = [0 if (j % 2 == 0 or clipped) else cfg["window"] // 2 for j in range(cfg["depths"][i])]
    return cfg["embed_dim"] * 2 ** i, (H, W), ws, shifts


def level_resolution(cfg, il):
    r = cfg["img_ds_ratio"]
    return int(cfg["img_size"][0] // STRIDES[il] * r), int(cfg["img_size"][1] // STRIDES[il] * r)   # TP:587-589


def relative_position_index(ws):
    """[ws*ws, ws*ws] index into the (2ws-1)^2 bias table (TP:146-157): (dy + ws-1) * (2ws-1) + (dx + ws-1)."""
    ys, xs = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
    ys, xs = ys.flatten(), xs.flatten()
    dy = ys[:, None] - ys[None, :] + ws - 1
    dx = xs[:, None] - xs[None, :] + ws - 1
    return dy * (2 * ws - 1) + dx


def shifted_window_mask(Hp, Wp, ws, shift):
    """[nW, ws*ws, ws*ws]: 0 where two tokens of a (cyclically shifted) window come from the same image region,
    -100 otherwise (TP:276-293)."""
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


def to_windows(x, ws):      # [B, Hp, Wp, C] -> [B * nW, ws*ws, C], windows row-major, tokens row-major (TP:90-101)
    B, Hp, Wp, C = x.shape
    return x.reshape(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def from_windows(w, ws, B, Hp, Wp):   # inverse of to_windows (TP:104-117)
    C = w.shape[-1]
    return w.reshape(B, Hp // ws, Wp // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)


def window_attention(sd, pre, xw, prompts_n, heads, ws, mask):
    """xw [B*nW, ws*ws, C] normed window tokens, prompts_n [B, T, C] normed prompts (replicated into every window and
    put FIRST). Returns window outputs [B*nW, ws*ws, C], the un-scaled logits of the prompt queries against the window
    tokens [B*nW, heads, T, ws*ws], and the window-averaged prompt outputs [B, T, C] (TP:167-212)."""
    BW, L, C = xw.shape
    B, T, _ = prompts_n.shape
    nW = BW // B
    s = torch.cat([prompts_n[:, None].expand(B, nW, T, C).reshape(BW, T, C), xw], dim=1)
    N = T + L
    dh = C // heads
    qkv = _lin(s, sd, pre + "qkv").reshape(BW, N, 3, heads, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    raw = q @ k.transpose(-2, -1)                                       # :189, consumed un-scaled downstream
    logits = raw * dh ** -0.5
    bias = sd[pre + "relative_position_bias_table"].to(xw.dtype)[relative_position_index(ws).reshape(-1).to(xw.device)]
    bias = bias.reshape(L, L, heads).permute(2, 0, 1)                   # [heads, L, L]
    extra = bias[None]
    if mask is not None:                                                # window w of every image gets mask[w]
        extra = extra + mask.to(xw).repeat(B, 1, 1)[:, None]          # device and dtype of the activations
    patch_part = logits[:, :, T:, T:] + extra                           # :196 / :201: only patch x patch entries
    logits = torch.cat([logits[:, :, :T, :],
                        torch.cat([logits[:, :, T:, :T], patch_part], dim=-1)], dim=2)
    o = (logits.softmax(dim=-1) @ v).transpose(1, 2).reshape(BW, N, C)
    o = _lin(o, sd, pre + "proj")
    prompts_out = o[:, :T].reshape(B, nW, T, C).mean(dim=1)              # :210
    return o[:, T:], raw[:, :, :T, T:], prompts_out


def block(sd, pre, cfg, x, prompts, dim, res, ws, shift, heads, last_block):
    """One SwinTransformerBlock with task prompts (TP:310-405). x [B, H*W, C], prompts [B, T, C]. Returns x, prompts,
    raw_spa [B, heads, T, H, W] and raw_chan [B, T, C, nh, nw] (both un-scaled, pre-softmax)."""
    H, W = res
    B, L, C = x.shape
    T = prompts.shape[1]
    prompts_n = _ln(prompts, sd, pre + "norm1")
    chan_p = _lin(prompts, sd, pre + "token_trans")                    # from the UN-normalised prompts (:319)
    xn = _ln(x, sd, pre + "norm1").reshape(B, H, W, C)
    pad_b, pad_r = (ws - H % ws) % ws, (ws - W % ws) % ws
    xn = F.pad(xn, (0, 0, 0, pad_r, 0, pad_b))                          # zeros AFTER the norm (:328-331)
    Hp, Wp = H + pad_b, W + pad_r
    if shift:
        xn = torch.roll(xn, shifts=(-shift, -shift), dims=(1, 2))
    mask = shifted_window_mask(Hp, Wp, ws, shift) if shift else None
    ow, raw, prompts_sp = window_attention(sd, pre + "attn.", to_windows(xn, ws), prompts_n, heads, ws, mask)
    xa = from_windows(ow, ws, B, Hp, Wp)
    # prompt-row logits back onto the (padded, shifted) map: [B*nW, heads, T, ws*ws] -> [B, heads, T, Hp, Wp] (:351-354)
    raw = raw.reshape(B, Hp // ws, Wp // ws, heads, T, ws, ws).permute(0, 3, 4, 1, 5, 2, 6).reshape(B, heads, T, Hp, Wp)
    if shift:
        xa = torch.roll(xa, shifts=(shift, shift), dims=(1, 2))
        raw = torch.roll(raw, shifts=(shift, shift), dims=(3, 4))
    xa = xa[:, :H, :W].reshape(B, L, C)
    raw_spa = raw[..., :H, :W]

    # channel attention between the prompts and the CHANNELS of the attention output (:372-396)
    ce = cfg["chan_embed_dim"]
    r = int(round(math.sqrt(ce)))
    nh = nw = int(round(math.sqrt(cfg["chan_nheads"])))
    wh, ww = r // nh, r // nw
    q = _lin(chan_p, sd, pre + "chan_q")                                # [B, T, ce]
    kv = _lin(xa.transpose(1, 2), sd, pre + "chan_kv").reshape(B, C, 2, ce)
    k, v = kv[:, :, 0], kv[:, :, 1]                                     # [B, C, ce]

    def grid(t):   # [B, n, ce] with ce = (nh wh nw ww) -> [B, nh*nw, n, wh*ww]
        n = t.shape[1]
        return t.reshape(B, n, nh, wh, nw, ww).permute(0, 2, 4, 1, 3, 5).reshape(B, nh * nw, n, wh * ww)

    qg, kg, vg = grid(q), grid(k), grid(v)
    raw_c = qg @ kg.transpose(-2, -1)                                   # [B, nh*nw, T, C], un-scaled
    chan_out = (raw_c * ce ** -0.5).softmax(dim=-1) @ vg               # [B, nh*nw, T, wh*ww]
    chan_out = chan_out.reshape(B, nh, nw, T, wh, ww).permute(0, 3, 1, 4, 2, 5).reshape(B, T, ce)
    raw_chan = raw_c.reshape(B, nh, nw, T, C).permute(0, 3, 4, 1, 2)    # [B, T, C, nh, nw]

    x = x + xa                                                          # :399 (drop_path = identity in eval)
    x = x + _mlp(sd, pre, _ln(x, sd, pre + "norm2"))
    if last_block:                                                      # :402: the very last block stops here
        return x, prompts_sp, raw_spa, raw_chan
    new_p = prompts_sp + _lin(_lin(chan_out, sd, pre + "chan_proj"), sd, pre + "token_trans1")
    prompts = prompts + new_p
    prompts = prompts + _mlp(sd, pre, _ln(prompts, sd, pre + "norm2"))
    return x, prompts, raw_spa, raw_chan


def patch_merging(sd, pre, x, prompts, raw_spa, raw_chan, res):
    """TP:430-472: 2x2 token merge (order: (0,0), (1,0), (0,1), (1,1)) + LN + Linear(4C -> 2C); the prompt-row logit
    maps are halved by a learned stride-2 3x3 conv over (head, task) channels, the channel logits and the prompts by
    bias-free Linear(C -> 2C)."""
    H, W = res
    B, L, C = x.shape
    g = x.reshape(B, H, W, C)
    m = torch.cat([g[:, 0::2, 0::2], g[:, 1::2, 0::2], g[:, 0::2, 1::2], g[:, 1::2, 1::2]], dim=-1)
    x = _lin(_ln(m.reshape(B, -1, 4 * C), sd, pre + "norm"), sd, pre + "reduction")
    _, heads, T, _, _ = raw_spa.shape
    raw_spa = F.conv2d(raw_spa.reshape(B, heads * T, H, W), sd[pre + "spa_attn_ds.weight"].to(x.dtype),
                       sd[pre + "spa_attn_ds.bias"].to(x.dtype), stride=2, padding=1)
    raw_spa = raw_spa.reshape(B, heads, T, H // 2, W // 2)
    raw_chan = _lin(raw_chan.permute(0, 1, 3, 4, 2), sd, pre + "process_chan_attn").permute(0, 1, 4, 2, 3)
    return x, _lin(prompts, sd, pre + "task_prompts_up"), raw_spa, raw_chan


def task_features(sd, cfg, x, raw_spa, raw_chan, il):
    """TP:721-774 at level il: X [B, C, h, w]; per task the spatially gated and the channel-gated copy of X (gate =
    1 + raw logit), each bilinearly doubled and decoded by a 1x1 conv, then fused."""
    h, w = level_resolution(cfg, il)
    B, L, C = x.shape
    X = x.transpose(1, 2).reshape(B, C, h, w)
    heads = raw_spa.shape[1]
    nh, nw = raw_chan.shape[-2:]
    out = {}
    for ti, t in enumerate(cfg["tasks"]):
        g_spa = raw_spa[:, :, ti].repeat_interleave(C // heads, dim=1)              # head hd gates its channel block
        ys = X * (1 + g_spa)
        g_chan = raw_chan[:, ti].repeat_interleave(h // nh, dim=2).repeat_interleave(w // nw, dim=3)   # [B, C, h, w]
        yc = X * (1 + g_chan)
        if t != "3ddet":
            ys = F.interpolate(ys, scale_factor=2, mode="bilinear", align_corners=False)
            yc = F.interpolate(yc, scale_factor=2, mode="bilinear", align_corners=False)
        fs = _conv(ys, sd, f"backbone.fea_decode_spa.{il}.{t}.0")
        fc = _conv(yc, sd, f"backbone.fea_decode_chan.{il}.{t}.0")
        pre = f"backbone.fea_fuse.{il}.{t}."
        y = _conv(_conv(torch.cat([fs, fc], dim=1), sd, pre + "0"), sd, pre + "1", padding=1)
        out[t] = _conv(F.gelu(_bn(y, sd, pre + "2")), sd, pre + "4", padding=1)
    return out


def backbone_forward(sd, cfg, img):
    """TaskPrompterSwin.forward (TP:674-718): {task: [B, f, 2*h0, 2*w0]}."""
    if cfg["img_ds_ratio"] != 1:
        img = F.interpolate(img, scale_factor=cfg["img_ds_ratio"], mode="bilinear", align_corners=False)
    x = F.conv2d(img, sd["backbone.patch_embed.proj.weight"].to(img.dtype),
                 sd["backbone.patch_embed.proj.bias"].to(img.dtype), stride=cfg["patch"])
    x = _ln(x.flatten(2).transpose(1, 2), sd, "backbone.patch_embed.norm")
    B = x.shape[0]
    prompts = sd["backbone.task_prompts"].to(x.dtype)[None].expand(B, -1, -1)
    levels = []
    n_stage = len(cfg["depths"])
    for i in range(n_stage):
        dim, res, ws, shifts = stage_geometry(cfg, i)
        for j, shift in enumerate(shifts):
            last = i == n_stage - 1 and j == len(shifts) - 1
            x, prompts, raw_spa, raw_chan = block(sd, f"backbone.layers.{i}.blocks.{j}.", cfg, x, prompts, dim, res, ws,
                                                  shift, cfg["heads"][i], last)
        if i < n_stage - 1:
            x, prompts, raw_spa, raw_chan = patch_merging(sd, f"backbone.layers.{i}.downsample.", x, prompts, raw_spa,
                                                          raw_chan, res)
            levels.append(task_features(sd, cfg, x, raw_spa, raw_chan, i))
    levels.append(task_features(sd, cfg, _ln(x, sd, "backbone.norm"), raw_spa, raw_chan, n_stage - 1))
    out = {}
    for t in cfg["tasks"]:
        maps = [lv[t] for lv in levels]
        if t == "3ddet":
            out[t] = maps
            continue
        size = maps[0].shape[-2:]
        acc = 0
        for m_ in maps:                                                  # python sum(): 0 + m0 + m1 + ...
            acc = acc + F.interpolate(m_, size, mode="bilinear")
        out[t] = _conv(acc, sd, f"backbone.multi_scale_fuse.{t}", padding=1)
    return out


def forward(sd, cfg, img):
    """TaskPrompterWrapper.forward: {task: [B, n_out, H, W]} at dd_label_map_size or the input size."""
    feats = backbone_forward(sd, cfg, img)
    size = tuple(cfg["dd_label_map_size"]) if "dd_label_map_size" in cfg else img.shape[-2:]
    head = deconv_head if cfg.get("head", "conv") == "deconv" else conv_head
    return {t: F.interpolate(head(sd, t, feats[t]), size, mode="bilinear") for t in cfg["tasks"]}


# ------------------------------------------------------------------------------------------------------------------
def param_shapes(cfg):
    """Ordered {name: shape} of every PARAMETER and BatchNorm buffer of the reference model for this config (index /
    mask buffers are derived data and not listed)."""
    E, T = cfg["embed_dim"], len(cfg["tasks"]) * cfg["prompt_len"]
    Lv, f, ce = cfg["level_embed_dim"], cfg["f"], cfg["chan_embed_dim"]
    bc = [2 * E, 4 * E, 8 * E, 8 * E]
    s = {}

    def lin(n, o, i, bias=True):
        s[n + ".weight"] = (o, i)
        if bias:
            s[n + ".bias"] = (o,)

    def conv(n, o, i, k):
        s[n + ".weight"] = (o, i, k, k)
        s[n + ".bias"] = (o,)

    def norm(n, c):
        s[n + ".weight"] = (c,)
        s[n + ".bias"] = (c,)

    def bn(n, c):
        norm(n, c)
        s[n + ".running_mean"] = (c,)
        s[n + ".running_var"] = (c,)
        s[n + ".num_batches_tracked"] = ()

    s["backbone.task_prompts"] = (T, E)
    conv("backbone.patch_embed.proj", E, 3, cfg["patch"])
    norm("backbone.patch_embed.norm", E)
    for il in range(4):
        for t in cfg["tasks"]:
            p = f"backbone.fea_fuse.{il}.{t}."
            conv(p + "0", f, 2 * Lv, 1)
            conv(p + "1", f, f, 3)
            bn(p + "2", f)
            conv(p + "4", f, f, 3)
    for il in range(4):
        for t in cfg["tasks"]:
            conv(f"backbone.fea_decode_spa.{il}.{t}.0", Lv, bc[il], 1)
    for il in range(4):
        for t in cfg["tasks"]:
            conv(f"backbone.fea_decode_chan.{il}.{t}.0", Lv, bc[il], 1)
    for t in cfg["tasks"]:
        if t != "3ddet":
            conv(f"backbone.multi_scale_fuse.{t}", f, f, 3)
    n_stage = len(cfg["depths"])
    for i in range(n_stage):
        dim, (H, W), ws, shifts = stage_geometry(cfg, i)
        for j in range(len(shifts)):
            p = f"backbone.layers.{i}.blocks.{j}."
            norm(p + "norm1", dim)
            s[p + "attn.relative_position_bias_table"] = ((2 * ws - 1) ** 2, cfg["heads"][i])
            lin(p + "attn.qkv", 3 * dim, dim)
            lin(p + "attn.proj", dim, dim)
            norm(p + "norm2", dim)
            lin(p + "mlp.fc1", 4 * dim, dim)
            lin(p + "mlp.fc2", dim, 4 * dim)
            lin(p + "chan_q", ce, ce)
            lin(p + "chan_kv", 2 * ce, H * W)
            lin(p + "token_trans", ce, dim)
            if not (i == n_stage - 1 and j == len(shifts) - 1):
                lin(p + "chan_proj", ce, ce)
                lin(p + "token_trans1", dim, ce)
        if i < n_stage - 1:
            p = f"backbone.layers.{i}.downsample."
            lin(p + "reduction", 2 * dim, 4 * dim, bias=False)
            norm(p + "norm", 4 * dim)
            lin(p + "process_chan_attn", 2 * dim, dim, bias=False)
            lin(p + "task_prompts_up", 2 * dim, dim, bias=False)
            conv(p + "spa_attn_ds", cfg["heads"][i] * T, cfg["heads"][i] * T, 3)
    norm("backbone.norm", 8 * E)
    for t in cfg["tasks"]:
        if cfg.get("head", "conv") == "deconv":
            h2 = f // 2
            s[f"heads.{t}.mt_proj.0.weight"] = (f, h2, 2, 2)
            s[f"heads.{t}.mt_proj.0.bias"] = (h2,)
            bn(f"heads.{t}.mt_proj.1", h2)
            conv(f"heads.{t}.mt_proj.3", h2, h2, 3)
            bn(f"heads.{t}.mt_proj.4", h2)
            conv(f"heads.{t}.linear_pred", cfg["num_output"][t], h2, 1)
        else:
            conv(f"heads.{t}.mt_proj.0", f, f, 3)
            bn(f"heads.{t}.mt_proj.1", f)
            conv(f"heads.{t}.linear_pred", cfg["num_output"][t], f, 1)
    return s


def init_state_dict(cfg, seed=0, dtype=torch.float32):
    """Deterministic random parameters with the reference's names and shapes (fixtures carry only inputs/outputs and
    a checksum). Scales are chosen so that logits, gates and BatchNorm folding are all exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(cfg).items():
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            sd[name] = torch.tensor(0)
        elif leaf == "running_var":
            sd[name] = (1 + 0.2 * torch.rand(shape, generator=g)).to(dtype)
        elif leaf == "running_mean":
            sd[name] = (0.1 * torch.randn(shape, generator=g)).to(dtype)
        elif name == "backbone.task_prompts":
            sd[name] = (1 + torch.randn(shape, generator=g)).clamp_(-2, 2).to(dtype)
        elif len(shape) == 1 and leaf == "weight":                       # norm scales
            sd[name] = (1 + 0.1 * torch.randn(shape, generator=g)).to(dtype)
        elif leaf == "bias":
            sd[name] = (0.05 * torch.randn(shape, generator=g)).to(dtype)
        elif leaf == "relative_position_bias_table":
            sd[name] = (0.5 * torch.randn(shape, generator=g)).to(dtype)
        else:                                                            # linear / conv weights: fan-in scaled
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[name] = (torch.randn(shape, generator=g) / math.sqrt(fan_in)).to(dtype)
    return sd
