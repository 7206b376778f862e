"""CPU restatement of the reference's rotated-box BEV overlap / IoU and NMS -- TEST INFRASTRUCTURE.

Follows the ALGORITHM of TaskPrompter/detection_toolbox/iou3d/src/iou3d_kernel.cu in float32 (numpy scalars), so
that it can serve as the yard-stick for the device kernels, which compute the same polygon by half-plane clipping:
  box_overlap   :124-241  corners rotated about the centre (:104-113), the 16 edge-edge intersections (:70-102), the
                          corners of one box inside the other (:47-68, margin 1e-5), angular sort about the centroid
                          (:115-122, :204-215), shoelace area (:226-240)
  iou_bev       :243-251  overlap / max(sa + sb - overlap, 1e-8)
  iou_normal    :332-340  axis-aligned IoU of the [x1,y1,x2,y2] part
  nms           iou3d_kernel.cu:285-330 (suppression matrix, strict '>') + iou3d.cpp:131-143 (greedy sweep in order)

PARITY PIN: the reference ships no test vectors for this extension and its only implementation is CUDA (it cannot run
in the CPU-only build container), so this restatement is pinned by analytic known answers instead
(tests/test_iou3d.py: axis-aligned pairs, identical boxes, a square rotated by 45 degrees inside / across another
square, disjoint boxes) -- "parity unpinned by reference outputs" for this row.
"""
import math

import numpy as np

F = np.float32
EPS = F(1e-8)


def _rot(cx, cy, c, s, x, y):
    dx, dy = F(x - cx), F(y - cy)
    return F(dx * c + dy * s + cx), F(-dx * s + dy * c + cy)


def _corners(b):
    x1, y1, x2, y2, a = [F(v) for v in b]
    cx, cy = F((x1 + x2) / F(2)), F((y1 + y2) / F(2))
    c, s = F(math.cos(a)), F(math.sin(a))
    return [_rot(cx, cy, c, s, x, y) for x, y in ((x1, y1), (x2, y1), (x2, y2), (x1, y2))]


def _cross3(p1, p2, p0):
    return F((p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1]))


def _intersection(p1, p0, q1, q0):
    if not (min(p0[0], p1[0]) <= max(q0[0], q1[0]) and min(q0[0], q1[0]) <= max(p0[0], p1[0]) and
            min(p0[1], p1[1]) <= max(q0[1], q1[1]) and min(q0[1], q1[1]) <= max(p0[1], p1[1])):
        return None
    s1, s2 = _cross3(q0, p1, p0), _cross3(p1, q1, p0)
    s3, s4 = _cross3(p0, q1, q0), _cross3(q1, p1, q0)
    if not (s1 * s2 > 0 and s3 * s4 > 0):
        return None
    s5 = _cross3(q1, p1, p0)
    if abs(s5 - s1) > EPS:
        return F((s5 * q0[0] - s1 * q1[0]) / (s5 - s1)), F((s5 * q0[1] - s1 * q1[1]) / (s5 - s1))
    a0, b0, c0 = p0[1] - p1[1], p1[0] - p0[0], p0[0] * p1[1] - p1[0] * p0[1]
    a1, b1, c1 = q0[1] - q1[1], q1[0] - q0[0], q0[0] * q1[1] - q1[0] * q0[1]
    D = a0 * b1 - a1 * b0
    return F((b0 * c1 - b1 * c0) / D), F((a1 * c0 - a0 * c1) / D)


def _in_box(b, p):
    x1, y1, x2, y2, a = [F(v) for v in b]
    cx, cy = F((x1 + x2) / F(2)), F((y1 + y2) / F(2))
    c, s = F(math.cos(-a)), F(math.sin(-a))
    rx, ry = _rot(cx, cy, c, s, p[0], p[1])
    m = F(1e-5)
    return rx > x1 - m and rx < x2 + m and ry > y1 - m and ry < y2 + m


def box_overlap(a, b):
    ca, cb = _corners(a), _corners(b)
    ca.append(ca[0])
    cb.append(cb[0])
    pts = []
    for i in range(4):
        for j in range(4):
            p = _intersection(ca[i + 1], ca[i], cb[j + 1], cb[j])
            if p is not None:
                pts.append(p)
    for k in range(4):
        if _in_box(a, cb[k]):
            pts.append(cb[k])
        if _in_box(b, ca[k]):
            pts.append(ca[k])
    if len(pts) < 3:
        return F(0)
    cx = F(sum(p[0] for p in pts) / F(len(pts)))
    cy = F(sum(p[1] for p in pts) / F(len(pts)))
    pts.sort(key=lambda p: math.atan2(p[1] - cy, p[0] - cx))
    area = F(0)
    for k in range(len(pts) - 1):
        ax, ay = pts[k][0] - pts[0][0], pts[k][1] - pts[0][1]
        bx, by = pts[k + 1][0] - pts[0][0], pts[k + 1][1] - pts[0][1]
        area = F(area + (ax * by - ay * bx))
    return F(abs(area) / F(2))


def iou_bev(a, b):
    sa = F((F(a[2]) - F(a[0])) * (F(a[3]) - F(a[1])))
    sb = F((F(b[2]) - F(b[0])) * (F(b[3]) - F(b[1])))
    so = box_overlap(a, b)
    return F(so / max(F(sa + sb - so), EPS))


def iou_normal(a, b):
    w = max(F(min(a[2], b[2]) - max(a[0], b[0])), F(0))
    h = max(F(min(a[3], b[3]) - max(a[1], b[1])), F(0))
    inter = F(w * h)
    sa = F((F(a[2]) - F(a[0])) * (F(a[3]) - F(a[1])))
    sb = F((F(b[2]) - F(b[0])) * (F(b[3]) - F(b[1])))
    return F(inter / max(F(sa + sb - inter), EPS))


def pairwise(A, B, fn=iou_bev):
    return np.array([[fn(a, b) for b in B] for a in A], dtype=np.float32).reshape(len(A), len(B))


def nms(boxes_sorted, thresh, rotated=True, iou=None):
    """Greedy NMS over boxes already sorted by descending score; returns the kept indices. `iou`: optional precomputed
    [N,N] matrix (to share the pairwise work between checks)."""
    n = len(boxes_sorted)
    fn = iou_bev if rotated else iou_normal
    removed = np.zeros(n, dtype=bool)
    keep = []
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        for j in range(i + 1, n):
            v = iou[i, j] if iou is not None else fn(boxes_sorted[i], boxes_sorted[j])
            if v > thresh:
                removed[j] = True
    return keep
