"""CPU restatement of the TaskPrompter (ViT) forward -- TEST INFRASTRUCTURE (the parity oracle).

Functional, state_dict-driven restatement of the reference forward in eval() mode, written from the
algorithm, not from the reference's code structure: the per-head / per-window
Python loops of the reference become broadcasts, dead code (`chan_x`, the softmaxed copies of the
logit maps) is omitted. Works in the dtype of `x` (fp32 like the reference, or fp64 for a tighter
yard-stick). Parameter names follow the reference state_dict (TaskPrompterWrapper: `backbone.*`,
`heads.*`) so a reference checkpoint drives it directly.

Pinned against the unmodified reference by tests/test_oracle_vs_reference.py (run where
/root/reference exists) and against tests/golden/*.pt everywhere.

Reference lines (relative to TaskPrompter/):
  Attention.forward      models/transformers/taskprompter.py:195-254
  Block.forward          models/transformers/taskprompter.py:270-279
  TaskPrompter.forward   models/transformers/taskprompter.py:392-422
  cal_task_feature       models/transformers/taskprompter.py:424-487
  ConvHead.forward       models/transformers/taskprompter.py:688-698
  DEConvHead.forward     models/transformers/taskprompter.py:700-715
  wrapper forward        models/taskprompter_wrapper.py:22-40
  PatchEmbed / Mlp       timm==0.5.4 (pinned TaskPrompter/README.md:75), restated in oracle/shim
"""
import math

import torch
import torch.nn.functional as F


def _lin(x, sd, name):
    return F.linear(x, sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype))


def _ln(x, sd, name, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype), eps)


def _conv(x, sd, name, padding=0):
    b = sd.get(name + ".bias")
    return F.conv2d(x, sd[name + ".weight"].to(x.dtype), None if b is None else b.to(x.dtype), padding=padding)


_TRAIN = None      # train_mode(): {"rand": iterator of uniform [B,1,1] draws or None} while restating model.train()


class train_mode:
    """Context: the restatement follows model.train() -- BatchNorm2d normalises with batch statistics (running statistics
    are not tracked here) and DropPath (timm 0.5.4 drop_path, taskprompter.py:265,273-277; rates linspace(0, rate, depth),
    :320) scales each residual branch by floor(keep + u) / keep per sample, u drawn from `rand` (an iterable of [B,1,1]
    tensors replaying the reference's torch.rand calls) or from torch.rand."""

    def __init__(self, drop_path_rate=0.15, rand=None):
        self.state = {"rate": drop_path_rate, "rand": iter(rand) if rand is not None else None}

    def __enter__(self):
        global _TRAIN
        _TRAIN = self.state

    def __exit__(self, *a):
        global _TRAIN
        _TRAIN = None


def _drop_path(x, idx, depth):
    if _TRAIN is None:
        return x
    dp = float(torch.linspace(0, _TRAIN["rate"], depth)[idx])
    if dp == 0.0:
        return x
    keep = 1 - dp
    u = next(_TRAIN["rand"]).to(x) if _TRAIN["rand"] is not None else torch.rand((x.shape[0], 1, 1), dtype=x.dtype, device=x.device)
    return x / keep * torch.floor(keep + u)


def _bn(x, sd, name, eps=1e-5):
    # eval-mode BatchNorm2d: running statistics (taskprompter.py:362 BatchNorm2d, :692); batch statistics under train_mode
    if _TRAIN is not None:
        return F.batch_norm(x, None, None, sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype), True, 0.0, eps)
    return F.batch_norm(x, sd[name + ".running_mean"].to(x.dtype), sd[name + ".running_var"].to(x.dtype),
                        sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype), False, 0.0, eps)


def grid_of(cfg):
    return cfg["img_size"][0] // cfg["patch"], cfg["img_size"][1] // cfg["patch"]


def block_forward(sd, pre, cfg, x, prompts, want_logits, idx=0):
    """One TaskPrompter block (taskprompter.py:270-279 with Attention :195-254).

    x [B,P,C] patches, prompts [B,T,C]. Returns x, prompts and, if want_logits, the two tensors
    cal_task_feature consumes: R_prompt [B,H,T,N] = raw q.k^T rows of the prompt queries, and
    Rc [B,T,C,nh,nw] = raw channel logits."""
    B, P, C = x.shape
    T = prompts.shape[1]
    H = cfg["heads"]
    dh = C // H
    xn = _ln(x, sd, pre + "norm1")                 # :272 norm1(x)
    pn = _ln(prompts, sd, pre + "norm1")           # :272 norm1(task_prompts)
    s = torch.cat([pn, xn], dim=1)                 # :199 prompts first
    N = T + P
    qkv = _lin(s, sd, pre + "attn.qkv").reshape(B, N, 3, H, dh).permute(2, 0, 3, 1, 4)  # :201
    q, k, v = qkv[0], qkv[1], qkv[2]
    raw = q @ k.transpose(-2, -1)                  # :204 un-scaled logits
    attn = (raw * dh ** -0.5).softmax(dim=-1)      # :205-206
    o = (attn @ v).transpose(1, 2).reshape(B, N, C)  # :210
    o = _lin(o, sd, pre + "attn.proj")             # :212
    o_p, o_x = o[:, :T], o[:, T:]                  # :214 sep_prompt
    cp = _lin(pn, sd, pre + "attn.token_trans")    # :219  [B,T,P]
    o_p = o_p + _lin(cp, sd, pre + "attn.token_trans1")  # :250
    R_prompt = Rc = None
    if want_logits:
        R_prompt = raw[:, :, :T, :]
        gh, gw = grid_of(cfg)
        nh = nw = int(round(math.sqrt(cfg["chan_nheads"])))   # :233
        wh, ww = gh // nh, gw // nw
        # Rc[b,t,c,i,j] = sum_{pixel in window (i,j)} cp[b,t,pixel] * xn[b,pixel,c]   (:236-240,:246)
        cpw = cp.reshape(B, T, nh, wh, nw, ww)
        xw = xn.reshape(B, nh, wh, nw, ww, C)
        Rc = torch.einsum("btihjw,bihjwc->btcij", cpw, xw)
    D = cfg["depth"]
    x = x + _drop_path(o_x, idx, D)                 # :273 (drop_path = identity in eval)
    x = x + _drop_path(_mlp(sd, pre, _ln(x, sd, pre + "norm2")), idx, D)        # :274
    prompts = prompts + _drop_path(o_p, idx, D)     # :276
    prompts = prompts + _drop_path(_mlp(sd, pre, _ln(prompts, sd, pre + "norm2")), idx, D)  # :277
    return x, prompts, R_prompt, Rc


def _mlp(sd, pre, x):
    # timm Mlp: fc1 -> GELU(erf) -> fc2
    return _lin(F.gelu(_lin(x, sd, pre + "mlp.fc1")), sd, pre + "mlp.fc2")


def cal_task_feature(sd, cfg, x, R_prompt, Rc, il):
    """Spatial-channel task prompting + optional cross-task reweighting (taskprompter.py:424-487).
    x [B,P,C] (un-normed at intermediate levels, LN_final(x) at the last). Returns {task: [B,f,gh,gw]}."""
    B, P, C = x.shape
    gh, gw = grid_of(cfg)
    H = cfg["heads"]
    dh = C // H
    T = len(cfg["tasks"])
    X = x.transpose(1, 2).reshape(B, C, gh, gw)     # :427
    nh = nw = int(round(math.sqrt(cfg["chan_nheads"])))
    wh, ww = gh // nh, gw // nw
    fea = {}
    for t, task in enumerate(cfg["tasks"]):
        # spatial gate: channel c uses head c // dh's prompt->patch logit (:436-446)
        g = R_prompt[:, :, t, T:].reshape(B, H, 1, gh, gw).expand(B, H, dh, gh, gw).reshape(B, C, gh, gw)
        s_fea = _conv(X * (1 + g), sd, f"backbone.fea_decode_spa.{il}.{task}.0")      # :447
        # channel gate: per (channel, window) scalar (:452-467)
        gc = Rc[:, t]                                                                  # [B,C,nh,nw]
        gc = gc.reshape(B, C, nh, 1, nw, 1).expand(B, C, nh, wh, nw, ww).reshape(B, C, gh, gw)
        c_fea = _conv(X * (1 + gc), sd, f"backbone.fea_decode_chan.{il}.{task}.0")    # :468
        y = torch.cat([s_fea, c_fea], dim=1)                                           # :471
        pre = f"backbone.fea_fuse.{il}.{task}."
        y = _conv(y, sd, pre + "0")                  # 1x1 (2e -> f)
        y = _conv(y, sd, pre + "1", padding=1)       # 3x3 (f -> f)
        y = F.gelu(_bn(y, sd, pre + "2"))            # BN, GELU
        y = _conv(y, sd, pre + "4")                  # 1x1 (f -> f)
        fea[task] = y
    if cfg["use_ctr"]:                               # :478-485
        new = {}
        for t, task in enumerate(cfg["tasks"]):
            a = R_prompt[:, :, t:t + 1, :T]          # [B,H,1,T] prompt<->prompt affinity
            pre = f"backbone.ctr_attn_conv.{il}.{task}."
            w = _conv(F.gelu(_conv(a, sd, pre + "0")), sd, pre + "2")   # [B,1,1,T]
            new[task] = sum(w[:, :, :, j:j + 1] * fea[tt] for j, tt in enumerate(cfg["tasks"]))
        fea = new
    return fea


def backbone_forward(sd, cfg, img, taps=None):
    """TaskPrompter.forward (taskprompter.py:392-422). Returns {task: [B,f,4gh,4gw]}."""
    dt = img.dtype
    B = img.shape[0]
    T = len(cfg["tasks"])
    x = _conv_stride(img, sd, "backbone.patch_embed.proj", cfg["patch"])     # timm PatchEmbed
    x = x.flatten(2).transpose(1, 2)
    x = x + sd["backbone.pos_embed"].to(dt)[:, 1:]                           # :394 (cls slot skipped)
    prompts = sd["backbone.task_prompts"].to(dt)[None].expand(B, -1, -1)     # :397
    select = list(cfg["select"])
    acc = {t: 0 for t in cfg["tasks"]}
    R_prompt = Rc = None
    for idx in range(cfg["depth"]):
        want = (idx + 1 in select) or (idx == cfg["depth"] - 1)
        x, prompts, R_prompt_i, Rc_i = block_forward(sd, f"backbone.blocks.{idx}.", cfg, x, prompts, want, idx)
        if want:
            R_prompt, Rc = R_prompt_i, Rc_i
        if taps is not None:
            taps[f"block{idx}.x"] = x
            taps[f"block{idx}.prompts"] = prompts
        if idx + 1 in select:                                                # :406-411
            il = sum(1 for s in select if idx >= s - 1) - 1
            cur = cal_task_feature(sd, cfg, x, R_prompt, Rc, il)
            if taps is not None:
                taps[f"level{il}.R_prompt"] = R_prompt
                taps[f"level{il}.Rc"] = Rc
                for t in cfg["tasks"]:
                    taps[f"level{il}.fea.{t}"] = cur[t]
            for t in cfg["tasks"]:
                acc[t] = acc[t] + cur[t]
    x = _ln(x, sd, "backbone.norm")                                          # :413
    cur = cal_task_feature(sd, cfg, x, R_prompt, Rc, 3)                       # :416-417 (logits of the last block)
    out = {}
    for t in cfg["tasks"]:
        if taps is not None:
            taps[f"level3.fea.{t}"] = cur[t]
        out[t] = F.interpolate(acc[t] + cur[t], scale_factor=4, mode="bilinear")   # :419-420
    return out


def _conv_stride(x, sd, name, stride):
    return F.conv2d(x, sd[name + ".weight"].to(x.dtype), sd[name + ".bias"].to(x.dtype), stride=stride)


def conv_head(sd, task, x):
    """ConvHead.forward (taskprompter.py:688-698): 3x3 conv -> BN -> GELU -> 1x1 conv."""
    pre = f"heads.{task}."
    y = _conv(x, sd, pre + "mt_proj.0", padding=1)
    y = F.gelu(_bn(y, sd, pre + "mt_proj.1"))
    return _conv(y, sd, pre + "linear_pred")


def deconv_head(sd, task, x):
    """DEConvHead.forward (taskprompter.py:700-715): ConvTranspose2d(k2,s2) -> BN -> GELU -> 3x3 conv -> BN -> GELU
    -> 1x1 conv, at twice the input resolution."""
    pre = f"heads.{task}."
    y = F.conv_transpose2d(x, sd[pre + "mt_proj.0.weight"].to(x.dtype), sd[pre + "mt_proj.0.bias"].to(x.dtype),
                           stride=2)
    y = F.gelu(_bn(y, sd, pre + "mt_proj.1"))
    y = F.gelu(_bn(_conv(y, sd, pre + "mt_proj.3", padding=1), sd, pre + "mt_proj.4"))
    return _conv(y, sd, pre + "linear_pred")


def forward(sd, cfg, img, taps=None):
    """TaskPrompterWrapper.forward (models/taskprompter_wrapper.py:22-40): {task: [B,n_out,H,W]}."""
    feats = backbone_forward(sd, cfg, img, taps)
    out = {}
    for t in cfg["tasks"]:
        if taps is not None:
            taps[f"task_fea.{t}"] = feats[t]
        head = deconv_head if cfg.get("head", "conv") == "deconv" else conv_head      # utils/common_config.py:64-70
        y = head(sd, t, feats[t])
        # wrapper :34-38: every task is resized to the input size (or dd_label_map_size) EXCEPT '3ddet'
        out[t] = y if t == "3ddet" else F.interpolate(y, tuple(cfg.get("dd_label_map_size", img.shape[-2:])), mode="bilinear")
    return out


def init_state_dict(cfg, seed=0, dtype=torch.float32):
    """Random parameters with the reference's names, shapes and (approximately) its init statistics
    (taskprompter.py:343-344,373,496-522; conv = PyTorch default). Used when the reference itself is
    not importable (GPU box). BatchNorm running stats are randomised so that BN folding is exercised."""
    g = torch.Generator().manual_seed(seed)
    C, T = cfg["C"], len(cfg["tasks"])
    gh, gw = grid_of(cfg)
    P = gh * gw
    e, f, H = cfg["e"], cfg["f"], cfg["heads"]
    sd = {}

    def tn(*shape, std=0.02, mean=0.0):
        return (torch.randn(*shape, generator=g) * std).clamp_(-2, 2).add_(mean).to(dtype)

    def lin(name, out_f, in_f, bias_std=0.02):
        sd[name + ".weight"] = tn(out_f, in_f)
        sd[name + ".bias"] = tn(out_f, std=bias_std)

    def ln(name, n):
        sd[name + ".weight"] = (1 + 0.1 * torch.randn(n, generator=g)).to(dtype)
        sd[name + ".bias"] = (0.1 * torch.randn(n, generator=g)).to(dtype)

    def conv(name, o, i, k, bias=True):
        bound = 1.0 / math.sqrt(i * k * k)
        sd[name + ".weight"] = ((torch.rand(o, i, k, k, generator=g) * 2 - 1) * bound).to(dtype)
        if bias:
            sd[name + ".bias"] = ((torch.rand(o, generator=g) * 2 - 1) * bound).to(dtype)

    def bn(name, n):
        sd[name + ".weight"] = (1 + 0.1 * torch.randn(n, generator=g)).to(dtype)
        sd[name + ".bias"] = (0.1 * torch.randn(n, generator=g)).to(dtype)
        sd[name + ".running_mean"] = (0.1 * torch.randn(n, generator=g)).to(dtype)
        sd[name + ".running_var"] = (1 + 0.2 * torch.rand(n, generator=g)).to(dtype)
        sd[name + ".num_batches_tracked"] = torch.tensor(0)

    conv("backbone.patch_embed.proj", C, 3, cfg["patch"])
    sd["backbone.pos_embed"] = tn(1, P + 1, C)
    sd["backbone.task_prompts"] = (torch.randn(T, C, generator=g)).clamp_(-3, 3).add_(1.0).clamp_(-2, 2).to(dtype)
    for i in range(cfg["depth"]):
        pre = f"backbone.blocks.{i}."
        ln(pre + "norm1", C)
        lin(pre + "attn.qkv", 3 * C, C)
        lin(pre + "attn.proj", C, C)
        lin(pre + "attn.token_trans", P, C)
        lin(pre + "attn.token_trans1", C, P)
        ln(pre + "norm2", C)
        lin(pre + "mlp.fc1", 4 * C, C)
        lin(pre + "mlp.fc2", C, 4 * C)
    ln("backbone.norm", C)
    for il in range(4):
        for t in cfg["tasks"]:
            pre = f"backbone.fea_fuse.{il}.{t}."
            conv(pre + "0", f, 2 * e, 1)
            conv(pre + "1", f, f, 3)
            bn(pre + "2", f)
            conv(pre + "4", f, f, 1)
            if cfg["use_ctr"]:
                conv(f"backbone.ctr_attn_conv.{il}.{t}.0", H, H, 1)
                conv(f"backbone.ctr_attn_conv.{il}.{t}.2", 1, H, 1)
            conv(f"backbone.fea_decode_spa.{il}.{t}.0", e, C, 1)
            conv(f"backbone.fea_decode_chan.{il}.{t}.0", e, C, 1)
    for t in cfg["tasks"]:
        if cfg.get("head", "conv") == "deconv":                       # DEConvHead (:700-715)
            h2 = f // 2
            conv(f"heads.{t}.mt_proj.0", f, h2, 2, bias=False)        # ConvTranspose2d weight is [in, out, kh, kw]
            sd[f"heads.{t}.mt_proj.0.bias"] = ((torch.rand(h2, generator=g) * 2 - 1) * 0.1).to(dtype)
            bn(f"heads.{t}.mt_proj.1", h2)
            conv(f"heads.{t}.mt_proj.3", h2, h2, 3)
            bn(f"heads.{t}.mt_proj.4", h2)
            conv(f"heads.{t}.linear_pred", cfg["num_output"][t], h2, 1)
        else:
            conv(f"heads.{t}.mt_proj.0", f, f, 3)
            bn(f"heads.{t}.mt_proj.1", f)
            conv(f"heads.{t}.linear_pred", cfg["num_output"][t], f, 1)
    return sd
