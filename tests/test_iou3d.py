"""Rotated BEV IoU / NMS: the oracle restatement against analytic known answers (CPU), and the device
kernels against the oracle (gpu): pairwise overlap / IoU within 1e-5 of the box area scale, NMS keep lists EXACT
(index work) on boxes whose pairwise IoUs stay clear of the threshold by more than the kernels' rounding."""
import math

import numpy as np
import pytest
import torch

from oracle import iou3d_ref as R


def test_oracle_known_answers():
    sq = [0., 0., 2., 2., 0.]
    assert abs(R.box_overlap(sq, sq) - 4.0) < 1e-5 and abs(R.iou_bev(sq, sq) - 1.0) < 1e-5
    assert abs(R.box_overlap(sq, [1., 1., 3., 3., 0.]) - 1.0) < 1e-5                 # axis-aligned: 1 x 1
    assert abs(R.iou_bev(sq, [1., 1., 3., 3., 0.]) - 1.0 / 7.0) < 1e-5
    assert R.box_overlap(sq, [5., 5., 6., 6., 0.3]) == 0.0                           # disjoint
    # unit square rotated by 45 degrees, centred in a 2 x 2 square: fully inside (diagonal sqrt(2) < 2)
    assert abs(R.box_overlap(sq, [0.5, 0.5, 1.5, 1.5, math.pi / 4]) - 1.0) < 1e-5
    # 2 x 2 square rotated by 45 degrees about the same centre: regular octagon, area 8 (sqrt(2) - 1)
    assert abs(R.box_overlap(sq, [0., 0., 2., 2., math.pi / 4]) - 8.0 * (math.sqrt(2) - 1)) < 1e-4
    # a rotation by 90 degrees of a 4 x 2 box about its centre overlaps the original in the central 2 x 2
    assert abs(R.box_overlap([0., 0., 4., 2., 0.], [0., 0., 4., 2., math.pi / 2]) - 4.0) < 1e-4
    # rotation by pi leaves a box unchanged
    assert abs(R.iou_bev([0., 0., 4., 2., 0.3], [0., 0., 4., 2., 0.3 + math.pi]) - 1.0) < 1e-4
    assert abs(R.iou_normal([0., 0., 2., 2., 1.0], [1., 0., 3., 2., -2.0]) - 1.0 / 3.0) < 1e-6   # angle ignored
    # greedy sweep: box 1 is suppressed by 0, box 2 survives (suppressed only by 1, which is gone)
    chain = [[0., 0., 2., 2., 0.], [0.6, 0., 2.6, 2., 0.], [1.4, 0., 3.4, 2., 0.]]
    assert R.nms(chain, 0.5) == [0, 2]


def _random_boxes(n, seed, extent=20.0):
    rng = np.random.default_rng(seed)
    cx, cy = rng.uniform(0, extent, n), rng.uniform(0, extent, n)
    w, h = rng.uniform(1.0, 6.0, n), rng.uniform(1.0, 6.0, n)
    ang = rng.uniform(-math.pi, math.pi, n)
    return np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2, ang], 1).astype(np.float32)


@pytest.mark.gpu
def test_pairwise_iou_and_overlap(cuda_dev):
    import mtt_b200  # noqa: F401
    from mtt_b200 import iou3d

    A, B = _random_boxes(37, 1), _random_boxes(53, 2)
    A[0] = B[0]                                           # an identical pair
    B[1] = [A[1][0], A[1][1], A[1][2], A[1][3], A[1][4] + math.pi / 2]
    ta, tb = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    ov, iou = iou3d.boxes_overlap_bev(ta, tb).cpu().numpy(), iou3d.boxes_iou_bev(ta, tb).cpu().numpy()
    rov, riou = R.pairwise(A, B, R.box_overlap), R.pairwise(A, B, R.iou_bev)
    assert np.abs(ov - rov).max() < 2e-4 and np.abs(iou - riou).max() < 2e-5
    assert abs(iou[0, 0] - 1.0) < 1e-5
    assert iou3d.boxes_iou_bev(ta[:0], tb).shape == (0, 53)


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,thresh", [(1, 3, 0.5), (64, 4, 0.3), (65, 5, 0.1), (300, 6, 0.25), (1000, 7, 0.5)])
def test_nms_keep_lists_are_exact(cuda_dev, n, seed, thresh):
    import mtt_b200  # noqa: F401
    from mtt_b200 import iou3d

    boxes = _random_boxes(n, seed, extent=12.0 + n ** 0.5)
    scores = np.random.default_rng(seed + 100).permutation(n).astype(np.float32)     # distinct scores
    order = np.argsort(-scores)
    tb, ts = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
    for rotated in (True, False):
        sb = boxes[order]
        m = min(n, 300)                                   # the oracle's pairwise pass is O(n^2) python
        tbm, tsm = tb[torch.from_numpy(order[:m].copy()).cuda()], ts[torch.from_numpy(order[:m].copy()).cuda()]
        fn = R.iou_bev if rotated else R.iou_normal
        mat = R.pairwise(sb[:m], sb[:m], fn)
        # a pair whose IoU sits within rounding distance of the threshold could legitimately flip between two correct
        # implementations: move the threshold to the nearest value that is clear of every pairwise IoU by > 1e-4
        th = thresh
        while (np.abs(mat - th) < 1e-4).any():
            th += 2.5e-4
        want = R.nms(sb[:m], th, rotated, iou=mat)
        got = (iou3d.nms_gpu(tbm, tsm, th) if rotated else iou3d.nms_normal_gpu(tbm, tsm, th)).cpu().numpy()
        sub = order[:m]
        assert got.tolist() == np.argsort(-scores[sub])[want].tolist()
    # the full set: properties (kept boxes are mutually below the threshold; every dropped box is covered by a kept one)
    keep = iou3d.nms_gpu(tb, ts, thresh).cpu().numpy()
    assert len(set(keep.tolist())) == len(keep) and (np.diff(scores[keep]) < 0).all()
    kb = tb[torch.from_numpy(keep).cuda()]
    kk = iou3d.boxes_iou_bev(kb, kb).cpu().numpy()
    np.fill_diagonal(kk, 0)
    assert (kk <= thresh + 1e-5).all()
    dropped = np.setdiff1d(np.arange(n), keep)
    if len(dropped):
        cov = iou3d.boxes_iou_bev(tb[torch.from_numpy(dropped).cuda()], kb).cpu().numpy()
        assert (cov.max(axis=1) > thresh - 1e-5).all()
    assert iou3d.nms_gpu(tb, ts, thresh, pre_maxsize=10, post_max_size=3).numel() <= 3
