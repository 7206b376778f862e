"""The reference's inference-time image pre-processing (TP/inference.py:93-115,127-133).

CPU: the oracle restatement (oracle/preproc_ref.py) against the golden vectors produced with the reference's own
Normalize / ToTensor classes and cv2.resize (tests/golden/preproc.pt, oracle/make_golden.py::make_preproc).
-m gpu: mtt_preprocess_image through the C ABI against the oracle and the golden vectors.
Tolerance: the pipeline is a handful of fp32 operations per pixel on values of magnitude <= 2.7; the CUDA kernel
follows the reference's operation order without FMA contraction, so agreement is 2 ulp (<= 1e-6 absolute)."""
import os

import numpy as np
import pytest
import torch

from oracle import preproc_ref as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preproc.pt")
TOL = 1e-6


def test_preproc_oracle_vs_golden():
    fx = torch.load(GOLD, weights_only=False)
    assert len(fx["cases"]) >= 3
    for c in fx["cases"]:
        got = P.infer_transform(c["bgr_u8"].numpy(), c["out_hw"])
        ref = c["out"].numpy()
        assert got.shape == ref.shape and got.dtype == np.float32
        assert np.abs(got - ref).max() <= TOL, np.abs(got - ref).max()


def test_preproc_oracle_identity_resize_is_exact_normalisation():
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (20, 31, 3), dtype=np.uint8)
    got = P.infer_transform(img, (20, 31))[0]
    want = ((img[:, :, ::-1].astype(np.float32) / np.float32(255.0)) - np.asarray(P.MEAN, np.float32)) \
        / np.asarray(P.STD, np.float32)
    assert np.array_equal(got, want.transpose(2, 0, 1))


@pytest.mark.gpu
def test_preprocess_kernel_vs_golden_and_oracle(cuda_dev):
    from mtt_b200 import ops

    fx = torch.load(GOLD, weights_only=False)
    for c in fx["cases"]:
        got = ops.preprocess_image(c["bgr_u8"].to(cuda_dev), c["out_hw"])
        torch.cuda.synchronize()
        assert (got.cpu() - c["out"]).abs().max().item() <= TOL
    # inference sizes: PASCAL-Context image -> 512 x 512, a batch of two, and an up-scale to the Cityscapes input
    rng = np.random.default_rng(11)
    for (B, h, w, H, W) in [(1, 375, 500, 512, 512), (2, 281, 500, 512, 512), (1, 512, 1024, 1024, 2048)]:
        img = rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)
        got = ops.preprocess_image(torch.from_numpy(img).to(cuda_dev), (H, W))
        torch.cuda.synchronize()
        ref = np.concatenate([P.infer_transform(img[b], (H, W)) for b in range(B)])
        assert tuple(got.shape) == (B, 3, H, W)
        assert np.abs(got.cpu().numpy() - ref).max() <= TOL
    # RGB input (bgr=False) = the same image with the channels pre-swapped
    img = rng.integers(0, 256, (40, 60, 3), dtype=np.uint8)
    a = ops.preprocess_image(torch.from_numpy(img).to(cuda_dev), (64, 64), bgr=True)
    b = ops.preprocess_image(torch.from_numpy(np.ascontiguousarray(img[:, :, ::-1])).to(cuda_dev), (64, 64), bgr=False)
    assert torch.equal(a, b)


@pytest.mark.gpu
def test_preprocess_rejects_bad_arguments(cuda_dev):
    from mtt_b200 import ops

    img = torch.zeros(8, 8, 3, dtype=torch.uint8, device=cuda_dev)
    with pytest.raises(RuntimeError):
        ops.preprocess_image(img, (8, 8), std=(0.0, 1.0, 1.0))
