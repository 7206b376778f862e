"""CPU: the algebraic identities the weight packing relies on, checked in plain PyTorch (no kernels involved).

* eval BatchNorm folds into the preceding convolution (pack.fold_bn; TP taskprompter.py:362,691, IP :33-38,:113,:493);
* ConvTranspose2d(k2, s2, p0) == 3x3 convolution (pad 1) of the zero-inserted input with tap (1-dy, 1-dx) holding
  W[:, :, dy, dx]^T (taskprompter.py _Plan, DEConvHead.mt_proj[0], reference taskprompter.py:704);
* ConvTranspose2d(k3, s2, p1, op1) == 3x3 convolution (pad 1) of the zero-inserted input with the spatially flipped,
  in/out-transposed kernel (invpt.py _Plan, scale_embed[0], reference transformer_decoder.py:63)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def zero_insert(x):
    B, C, h, w = x.shape
    z = torch.zeros(B, C, 2 * h, 2 * w, dtype=x.dtype)
    z[:, :, ::2, ::2] = x
    return z


def test_fold_bn_equals_conv_then_eval_batchnorm():
    import mtt_b200  # noqa: F401
    from mtt_b200.pack import fold_bn

    torch.manual_seed(0)
    for bias in (True, False):
        conv = nn.Conv2d(7, 5, 3, padding=1, bias=bias).double()
        bn = nn.BatchNorm2d(5).double().eval()
        with torch.no_grad():
            bn.weight.normal_(1, 0.2)
            bn.bias.normal_(0, 0.2)
            bn.running_mean.normal_(0, 0.3)
            bn.running_var.uniform_(0.5, 1.5)
        x = torch.randn(2, 7, 6, 9, dtype=torch.float64)
        w, b = fold_bn(conv.weight.detach(), conv.bias.detach() if bias else None, bn)
        got = F.conv2d(x, w.double(), b.double(), padding=1)
        want = bn(conv(x))
        assert (got - want).abs().max() < 1e-5          # fold_bn returns fp32


def test_deconv_k2s2_as_conv3x3_on_zero_inserted_input():
    torch.manual_seed(1)
    Cin, Cout = 6, 4
    wt = torch.randn(Cin, Cout, 2, 2, dtype=torch.float64)          # ConvTranspose2d weight layout [in, out, kh, kw]
    bias = torch.randn(Cout, dtype=torch.float64)
    x = torch.randn(2, Cin, 5, 7, dtype=torch.float64)
    want = F.conv_transpose2d(x, wt, bias, stride=2)
    w3 = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64)
    for dy in range(2):
        for dx in range(2):
            w3[:, :, 1 - dy, 1 - dx] = wt[:, :, dy, dx].t()
    got = F.conv2d(zero_insert(x), w3, bias, padding=1)
    assert got.shape == want.shape and (got - want).abs().max() < 1e-12


def test_deconv_k3s2p1op1_as_flipped_conv3x3_on_zero_inserted_input():
    torch.manual_seed(2)
    Cin, Cout = 5, 3
    wt = torch.randn(Cin, Cout, 3, 3, dtype=torch.float64)
    bias = torch.randn(Cout, dtype=torch.float64)
    x = torch.randn(2, Cin, 4, 6, dtype=torch.float64)
    want = F.conv_transpose2d(x, wt, bias, stride=2, padding=1, output_padding=1)
    got = F.conv2d(zero_insert(x), wt.flip(2, 3).permute(1, 0, 2, 3), bias, padding=1)
    assert got.shape == want.shape and (got - want).abs().max() < 1e-12
