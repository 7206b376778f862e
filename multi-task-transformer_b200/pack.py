"""Weight pre-packing helpers kept for the kernel tests and micro-benchmarks: thin wrappers over the C entry points
mtt_pack_weight / mtt_pack_conv_weight (include/mtt_b200.h), plus the eval-BatchNorm folding identity in plain torch
(tests/test_pack_identities.py checks the C packing against it). The models pack through ops.pack_* directly."""
import torch

from . import ops


def pack_linear_weight(w, nsplit):
    """nn.Linear / 1x1 conv weight [N, K] (or [N, K, 1, 1]) -> Split [N, K] (K-major, ld multiple of 8)."""
    return ops.pack_weight(w.detach().reshape(w.shape[0], -1).float().contiguous(), nsplit)


def pack_conv_weight(w, nsplit):
    """Conv2d weight [N, Cin, kh, kw] -> Split [N, kh*kw*cin_pad], tap-major, Cin zero-padded to 64."""
    return ops.pack_conv_weight(w.detach().float().contiguous(), None, None, nsplit)[0]


def fold_bn(w, b, bn):
    """Fold an eval-mode BatchNorm2d into the preceding conv: returns (w', b') fp32.
    y = (conv(x) - mean) / sqrt(var + eps) * gamma + beta."""
    s = (bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps))
    w2 = w.detach().float() * s.reshape(-1, *([1] * (w.dim() - 1)))
    b0 = b.detach().float() if b is not None else torch.zeros_like(s)
    b2 = (b0 - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
    return w2, b2
