"""Weight pre-packing: fp32 parameters -> split-bf16 operand planes in the layouts the kernels read.

Packing is parameter preprocessing (done once per parameter version, outside the timed hot path);
it uses torch only for permutes / BatchNorm folding of the parameters themselves and the library's
mtt_split_f32 for the cast.
"""
import torch

from . import ops


def pack_linear_weight(w, nsplit):
    """nn.Linear / 1x1 conv weight [N, K] (or [N, K, 1, 1]) -> Split [N, K] (K-major, ld multiple of 8)."""
    w2 = w.detach().reshape(w.shape[0], -1).float().contiguous()
    return ops.split_f32(w2, nsplit)


def pack_conv_weight(w, nsplit):
    """Conv2d weight [N, Cin, kh, kw] -> Split [N, kh*kw*cin_pad], tap-major, Cin zero-padded to 64."""
    N, Cin, kh, kw = w.shape
    cin_pad = ops.round_up(Cin, 64)
    wt = torch.zeros(N, kh * kw, cin_pad, dtype=torch.float32, device=w.device)
    wt[:, :, :Cin] = w.detach().float().permute(0, 2, 3, 1).reshape(N, kh * kw, Cin)
    return ops.split_f32(wt.reshape(N, kh * kw * cin_pad), nsplit)


def fold_bn(w, b, bn):
    """Fold an eval-mode BatchNorm2d into the preceding conv: returns (w', b') fp32.
    y = (conv(x) - mean) / sqrt(var + eps) * gamma + beta."""
    s = (bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps))
    w2 = w.detach().float() * s.reshape(-1, *([1] * (w.dim() - 1)))
    b0 = b.detach().float() if b is not None else torch.zeros_like(s)
    b2 = (b0 - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
    return w2, b2
