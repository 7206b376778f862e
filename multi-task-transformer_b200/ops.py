"""Tensor-level wrappers over the C ABI (include/mtt_b200.h).

PyTorch is used for device memory and streams only: every function here enqueues one library
kernel on ``torch.cuda.current_stream()`` and returns. Nothing in this module computes with torch
ops on the hot path.
"""
import ctypes as C

import torch

from . import lib as _L

ACT_NONE, ACT_GELU, ACT_RELU = _L.ACT_NONE, _L.ACT_GELU, _L.ACT_RELU


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def round_up(x, m):
    return (x + m - 1) // m * m


class Split:
    """An fp32-valued [rows, cols] matrix carried as bf16 planes hi (= bf16(x)) and lo (= bf16(x - hi)).

    Storage is one [nsplit, rows, ld] bf16 tensor; ``ld`` (elements) is a multiple of 8 so that TMA
    row strides are 16-byte aligned."""

    __slots__ = ("buf", "rows", "cols", "ld", "nsplit")

    def __init__(self, rows, cols, device, nsplit=2, ld=None, zero=False):
        self.rows, self.cols, self.nsplit = int(rows), int(cols), int(nsplit)
        self.ld = int(ld) if ld is not None else round_up(self.cols, 8)
        alloc = torch.zeros if zero else torch.empty
        self.buf = alloc((self.nsplit, self.rows, self.ld), dtype=torch.bfloat16, device=device)

    @property
    def hi(self):
        return self.buf[0]

    @property
    def lo(self):
        return self.buf[1] if self.nsplit == 2 else None

    def float(self):
        """Reconstruct the fp32 values (testing / debugging only)."""
        x = self.buf[0, :, : self.cols].float()
        if self.nsplit == 2:
            x = x + self.buf[1, :, : self.cols].float()
        return x


def split_f32(x, nsplit=2, cols_pad=None, out=None):
    """fp32 [rows, cols] (last dim contiguous) -> Split. Columns [cols, cols_pad) are zero."""
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    cols_pad = cols if cols_pad is None else cols_pad
    if out is None:
        out = Split(rows, cols_pad, x.device, nsplit, ld=round_up(cols_pad, 8))
    lib = _L.load()
    rc = lib.mtt_split_f32(_ptr(x), x.stride(0), _ptr(out.hi), _ptr(out.lo), out.ld, rows, cols,
                           cols_pad, _stream())
    _L.check(rc, "mtt_split_f32")
    return out


def layernorm(x, gamma, beta, eps, out_f32=None, out_split=None):
    """x fp32 [rows, cols] -> out_f32 (fp32 tensor) and/or out_split (Split)."""
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    lib = _L.load()
    rc = lib.mtt_layernorm(
        _ptr(x), x.stride(0), _ptr(gamma), _ptr(beta), float(eps),
        _ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0,
        _ptr(out_split.hi) if out_split is not None else None,
        _ptr(out_split.lo) if out_split is not None else None,
        out_split.ld if out_split is not None else 0, rows, cols, _stream())
    _L.check(rc, "mtt_layernorm")


def gemm(a, w, *, M=None, N=None, K=None, bias=None, act=ACT_NONE, residual=None, res_row_mod=0,
         out_f32=None, out_split=None, out_col_offset=0, regroup=None, conv=None, a_row_offset=0,
         a_gather=None, w_col_offset=0, a_col_offset=0, w_row_offset=0, out_row_offset=0, sk_ws=None):
    """D = act(A @ W^T + bias) + residual on the tcgen05 GEMM.

    a: Split [M, K] (or NHWC activation [B*H*W, C] when ``conv=(B, H, W, ksize, dil)``);
    w: Split [N, K] (conv: [N, ksize*ksize*cin_pad]); outputs: fp32 tensor and/or Split.
    regroup=(in_group, out_group, out_offset[, row_stride]) scatters output rows; a_gather=(group_rows,
    group_stride) gathers A rows in groups (M must be given); w_col_offset selects a K-slice of a wider packed W.
    sk_ws: optional stream-K workspace (streamk_workspace(device)); see mtt_gemm_desc.sk_ws."""
    d = _L.GemmDesc()
    _fill_gemm_desc(d, a, w, M, N, K, bias, act, residual, res_row_mod, out_f32, out_split, out_col_offset, regroup, conv,
                    a_row_offset, a_gather, w_col_offset, a_col_offset, w_row_offset, out_row_offset)
    if sk_ws is not None:
        d.sk_ws, d.sk_ws_bytes = sk_ws.data_ptr(), sk_ws.numel()
    rc = _L.load().mtt_gemm(C.byref(d), _stream())
    _L.check(rc, "mtt_gemm")


def gemm_grouped(calls):
    """calls: [(a, w, kwargs)] with gemm()'s arguments -- problems of identical geometry that differ only in their
    tensors -- as ONE persistent launch (mtt_gemm_grouped)."""
    arr = (_L.GemmDesc * len(calls))()
    for d, (a, w, kw) in zip(arr, calls):
        k = dict(M=None, N=None, K=None, bias=None, act=ACT_NONE, residual=None, res_row_mod=0, out_f32=None,
                 out_split=None, out_col_offset=0, regroup=None, conv=None, a_row_offset=0, a_gather=None, w_col_offset=0,
                 a_col_offset=0, w_row_offset=0, out_row_offset=0)
        k.update(kw)
        _fill_gemm_desc(d, a, w, **k)
    rc = _L.load().mtt_gemm_grouped(arr, len(calls), _stream())
    _L.check(rc, "mtt_gemm_grouped")


def gemm_splitk(a, w, partial, out_f32, *, K, bias=None, chunks):
    """out = A @ W^T + bias for a skinny A with a very long K: `chunks` K-slices run as the problems of ONE grouped
    launch into partial [chunks, M, N] (fp32), then a fixed-order reduction adds them and the bias. K-slices are
    multiples of 64 (the GEMM's K block)."""
    M, N = a.rows, w.rows
    step = round_up(-(-K // chunks), 64)
    calls, k0 = [], 0
    while k0 < K:
        kk = min(step, K - k0)
        calls.append((a, w, dict(M=M, N=N, K=kk, out_f32=partial[len(calls)], a_col_offset=k0, w_col_offset=k0)))
        k0 += kk
    gemm_grouped(calls) if len({c[2]["K"] for c in calls}) == 1 else [gemm(c[0], c[1], **c[2]) for c in calls]
    rc = _L.load().mtt_sum_partials(_ptr(partial), len(calls), M, N, partial.stride(-2), _ptr(bias), _ptr(out_f32),
                                    out_f32.stride(-2), _stream())
    _L.check(rc, "mtt_sum_partials")


def _fill_gemm_desc(d, a, w, M, N, K, bias, act, residual, res_row_mod, out_f32, out_split, out_col_offset, regroup, conv,
                    a_row_offset, a_gather, w_col_offset, a_col_offset=0, w_row_offset=0, out_row_offset=0):
    """w_row_offset / out_row_offset: first row of the B operand / of the split output (pointer offsets): lets one
    buffer hold the operands of several (batch, head) problems of a grouped launch."""
    nsplit = min(a.nsplit, w.nsplit)
    aoff = 2 * (a_row_offset * a.ld + a_col_offset)
    d.a_hi, d.a_lo, d.lda = a.hi.data_ptr() + aoff, (a.lo.data_ptr() + aoff if nsplit == 2 else 0), a.ld
    woff = 2 * (w_row_offset * w.ld + w_col_offset)
    d.b_hi, d.b_lo, d.ldb = w.hi.data_ptr() + woff, (w.lo.data_ptr() + woff if nsplit == 2 else 0), w.ld
    d.M = a.rows if M is None else M
    d.N = w.rows if N is None else N
    d.K = a.cols if K is None else K
    d.nsplit = nsplit
    if conv is not None:
        d.mode = 1
        d.B, d.H, d.W, d.ksize, d.dil = conv
    else:
        d.mode = 0
    d.bias = bias.data_ptr() if bias is not None else 0
    d.act = act
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(-1) == 1
        d.residual, d.ldr = residual.data_ptr(), residual.stride(-2)
    d.res_row_mod = res_row_mod
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.stride(-1) == 1
        d.out_f32, d.ldo_f32 = out_f32.data_ptr(), out_f32.stride(-2)
    if out_split is not None:
        ooff = 2 * (out_row_offset * out_split.ld + out_col_offset)
        d.out_hi = out_split.hi.data_ptr() + ooff
        d.out_lo = (out_split.lo.data_ptr() + ooff) if out_split.nsplit == 2 else 0
        d.ldo_bf = out_split.ld
        if nsplit == 2 and out_split.nsplit != 2:
            raise ValueError("nsplit=2 GEMM needs a 2-plane split output")
    if regroup is not None:
        d.in_group, d.out_group, d.out_offset = regroup[:3]
        d.out_row_stride = regroup[3] if len(regroup) > 3 else 1
    if a_gather is not None:    # (group_rows, group_stride): logical row (g, i) at physical row g*stride + i
        d.a_group_rows, d.a_group_stride = a_gather


def attention(qkv, out, *, B, N, H, scale, prompt_logits=None, T=0):
    """Fused softmax(q k^T * scale) v over [B, N] tokens; qkv Split [B*N, 3*H*64], out Split [B*N, H*64].
    prompt_logits: optional fp32 [B, H, T, N] receiving raw q.k^T of the first T query rows."""
    d = _L.AttnDesc()
    nsplit = min(qkv.nsplit, out.nsplit)
    assert qkv.ld == 3 * H * 64 and out.ld == H * 64
    d.qkv_hi, d.qkv_lo = qkv.hi.data_ptr(), (qkv.lo.data_ptr() if nsplit == 2 else 0)
    d.out_hi, d.out_lo = out.hi.data_ptr(), (out.lo.data_ptr() if nsplit == 2 else 0)
    if prompt_logits is not None:
        assert prompt_logits.dtype == torch.float32 and prompt_logits.is_contiguous()
        assert tuple(prompt_logits.shape) == (B, H, T, N)
        d.prompt_logits = prompt_logits.data_ptr()
    d.B, d.N, d.H, d.T, d.nsplit, d.scale = B, N, H, T, nsplit, float(scale)
    rc = _L.load().mtt_attention(C.byref(d), _stream())
    _L.check(rc, "mtt_attention")


def streamk_workspace(device):
    """A zero-filled stream-K workspace for gemm(sk_ws=...): one per stream that issues such launches."""
    with torch.cuda.device(device):
        return workspace(_L.load().mtt_gemm_streamk_bytes(), device)


def set_gemm_streamk(mode):
    """Stream-K policy of the CTA-pair GEMM: 0 off, 1 automatic (default), 2 whenever legal (tuning / testing knob; env
    MTT_GEMM_STREAMK)."""
    _L.load().mtt_set_gemm_streamk(int(mode))


def set_gemm_variant(v):
    """0 auto, 1 single-CTA 128x128, 2 CTA-pair 256x256, 3 CTA-pair 256x128 (tuning / testing knob)."""
    _L.load().mtt_set_gemm_variant(int(v))


def set_attention_variant(v):
    """0 = default attention kernel; other values select development variants when built (tuning / testing knob)."""
    _L.load().mtt_set_attention_variant(int(v))


def profile_begin():
    """Start per-launch timing of the tensor-core kernels (see mtt_profile_begin)."""
    _L.check(_L.load().mtt_profile_begin(), "mtt_profile_begin")


def profile_end(max_recs=4096):
    """Synchronise and return [(kind, M, N, K, ms, flops)] for every mtt_gemm (kind 0) / mtt_attention (kind 1) launch
    since profile_begin()."""
    arr = (_L.ProfileRec * max_recs)()
    n = C.c_int32(0)
    _L.check(_L.load().mtt_profile_end(arr, max_recs, C.byref(n)), "mtt_profile_end")
    return [(r.kind, r.M, r.N, r.K, r.ms, r.flops) for r in arr[:min(n.value, max_recs)]]


def launch_count(reset=False):
    lib = _L.load()
    n = lib.mtt_launch_count()
    if reset:
        lib.mtt_launch_count_reset()
    return n


def im2col_patch(img, patch, out):
    """img fp32 NCHW -> out Split [B*P, Cin*patch*patch]."""
    assert img.dtype == torch.float32 and img.is_contiguous()
    B, Cin, H, W = img.shape
    rc = _L.load().mtt_im2col_patch(_ptr(img), B, Cin, H, W, patch, _ptr(out.hi), _ptr(out.lo), out.ld,
                                    _stream())
    _L.check(rc, "mtt_im2col_patch")


def broadcast_rows(src, dst, B, group_rows):
    """dst[(b*group_rows + t), :] = src[t, :] for t < T; dst fp32 [B*group_rows, ld]."""
    T, Cc = src.shape
    assert src.is_contiguous() and src.dtype == torch.float32 and dst.dtype == torch.float32
    rc = _L.load().mtt_broadcast_rows(_ptr(src), _ptr(dst), B, T, Cc, group_rows, dst.stride(0), _stream())
    _L.check(rc, "mtt_broadcast_rows")


def chan_logits(cp, xn, out, *, B, N, T, Cdim, gh, gw, nh, nw):
    rc = _L.load().mtt_chan_logits(_ptr(cp), _ptr(xn.hi), _ptr(xn.lo), xn.ld, B, N, T, Cdim, gh, gw, nh, nw,
                                   _ptr(out), _stream())
    _L.check(rc, "mtt_chan_logits")


def gate_split(x, x_group_rows, x_row_offset, prompt_logits, chan_lg, task, ys, yc, *, B, T, N, H, Cdim, gh,
               gw, nh, nw, ntasks=1, task_stride=0):
    """Gate the patch map for tasks [task, task + ntasks) in one pass; ys / yc are the Splits of the FIRST task, task k's
    planes live k * task_stride elements further on (one task: plain Splits)."""
    assert x.dtype == torch.float32 and x.stride(-1) == 1 and ys.ld == yc.ld
    rc = _L.load().mtt_gate_split(_ptr(x), x.stride(-2), x_group_rows, x_row_offset, _ptr(prompt_logits),
                                  _ptr(chan_lg), task, ntasks, B, T, N, H, Cdim, gh, gw, nh, nw, _ptr(ys.hi),
                                  _ptr(ys.lo), _ptr(yc.hi), _ptr(yc.lo), ys.ld, task_stride, _stream())
    _L.check(rc, "mtt_gate_split")


def ctr_weights(prompt_logits, w0, b0, w2, b2, out, *, B, H, T, N):
    rc = _L.load().mtt_ctr_weights(_ptr(prompt_logits), B, H, T, N, _ptr(w0), _ptr(b0), _ptr(w2), _ptr(b2),
                                   _ptr(out), _stream())
    _L.check(rc, "mtt_ctr_weights")


def ctr_mix(F, w, acc, *, T, M, Cdim, ld, rows_per_batch, accumulate):
    rc = _L.load().mtt_ctr_mix(_ptr(F), _ptr(w), _ptr(acc), T, M, Cdim, ld, rows_per_batch,
                               1 if accumulate else 0, _stream())
    _L.check(rc, "mtt_ctr_mix")


def bilinear(x, ld_in, B, h, w, Cdim, H2, W2, *, out_f32=None, out_split=None, out_nchw=None,
             accumulate=False, in_batch_rows=0, in_row_offset=0, out_batch_rows=0, out_row_offset=0):
    """x: NHWC fp32 [B*h*w, ld_in] -> NHWC fp32 / NHWC Split / NCHW fp32 [B,C,H2,W2]."""
    rc = _L.load().mtt_bilinear(
        _ptr(x), ld_in, B, h, w, Cdim, H2, W2, _ptr(out_f32),
        out_f32.stride(-2) if out_f32 is not None else 0,
        _ptr(out_split.hi) if out_split is not None else None,
        _ptr(out_split.lo) if out_split is not None else None,
        out_split.ld if out_split is not None else 0, _ptr(out_nchw), 1 if accumulate else 0,
        in_batch_rows, in_row_offset, out_batch_rows, out_row_offset, _stream())
    _L.check(rc, "mtt_bilinear")


POSTPROC_KIND = {"semseg": 0, "human_parts": 0, "edge": 1, "sal": 2, "normals": 3, "depth": 4}


def bilinear_postproc(x, ld_in, B, h, w, Cdim, H2, W2, kind, out):
    """Bilinear resize fused with get_output's post-processing (TP/utils/utils.py:27-63); `out` is int64
    [B,H2,W2] for kind 0, fp32 otherwise."""
    i64 = out if kind == 0 else None
    f32 = None if kind == 0 else out
    assert out.is_contiguous() and out.dtype == (torch.int64 if kind == 0 else torch.float32)
    rc = _L.load().mtt_bilinear_postproc(_ptr(x), ld_in, B, h, w, Cdim, H2, W2, kind, _ptr(i64), _ptr(f32),
                                         _stream())
    _L.check(rc, "mtt_bilinear_postproc")


IMAGENET_MEAN = (0.485, 0.456, 0.406)   # TP/inference.py:99,107
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess_image(img_u8, out_hw, *, bgr=True, mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None):
    """The reference's inference pre-processing (TP/inference.py:93-115,127-133) on the device: img_u8 is a CUDA
    uint8 tensor [h,w,3] or [B,h,w,3] as cv2.imread returns it (BGR); returns fp32 [B,3,H,W] normalised and
    bilinearly resized (cv2 INTER_LINEAR), the tensor `model(x)` takes."""
    import ctypes
    if img_u8.dim() == 3:
        img_u8 = img_u8.unsqueeze(0)
    assert img_u8.is_cuda and img_u8.dtype == torch.uint8 and img_u8.is_contiguous() and img_u8.shape[-1] == 3
    B, h, w, _ = img_u8.shape
    H, W = out_hw
    if out is None:
        out = torch.empty(B, 3, H, W, device=img_u8.device, dtype=torch.float32)
    assert out.is_contiguous() and tuple(out.shape) == (B, 3, H, W) and out.dtype == torch.float32
    m3 = (ctypes.c_float * 3)(*mean)
    s3 = (ctypes.c_float * 3)(*std)
    rc = _L.load().mtt_preprocess_image(_ptr(img_u8), B, h, w, int(bool(bgr)), m3, s3, _ptr(out), H, W, _stream())
    _L.check(rc, "mtt_preprocess_image")
    return out


def bilinear_sum3(srcs, out, *, B, Cdim, H2, W2):
    """out (Split [B*H2*W2, C]) = sum_i bilinear(src_i -> H2 x W2); srcs: list of up to three
    (tensor fp32 [rows, ld], h, w, batch_rows, row_offset)."""
    arr = (_L.BilinearSrc * len(srcs))()
    for i, (t, h, w, brows, roff) in enumerate(srcs):
        assert t.dtype == torch.float32 and t.stride(-1) == 1
        arr[i].in_, arr[i].ld_in, arr[i].h, arr[i].w = t.data_ptr(), t.stride(-2), h, w
        arr[i].batch_rows, arr[i].row_offset = brows, roff
    rc = _L.load().mtt_bilinear_sum3(arr, len(srcs), B, Cdim, H2, W2, _ptr(out.hi), _ptr(out.lo), out.ld,
                                     _stream())
    _L.check(rc, "mtt_bilinear_sum3")


def split_rows(x, out, *, rows, cols, in_group=0, src_group=0, src_offset=0):
    """Gather fp32 rows of x (row r at (r // in_group) * src_group + src_offset + r % in_group) -> Split."""
    assert x.dtype == torch.float32 and x.stride(-1) == 1
    rc = _L.load().mtt_split_rows(_ptr(x), x.stride(-2), in_group, src_group, src_offset, _ptr(out.hi),
                                  _ptr(out.lo), out.ld, rows, cols, _stream())
    _L.check(rc, "mtt_split_rows")


def layernorm_seg(x, gamma, beta, eps, *, rows, cols, S=1, in_group=0, src_group=0, src_offset=0,
                  seg_stride=0, out_f32=None, out_split=None, out_seg_stride=0):
    assert x.dtype == torch.float32 and x.stride(-1) == 1
    rc = _L.load().mtt_layernorm_seg(
        _ptr(x), x.stride(-2), in_group, src_group, src_offset, seg_stride, S, _ptr(gamma), _ptr(beta),
        float(eps), _ptr(out_f32), out_f32.stride(-2) if out_f32 is not None else 0,
        _ptr(out_split.hi) if out_split is not None else None,
        _ptr(out_split.lo) if out_split is not None else None,
        out_split.ld if out_split is not None else 0, out_seg_stride, rows, cols, _stream())
    _L.check(rc, "mtt_layernorm_seg")


def zero_insert(x, out, *, B, h, w, Cdim, src_group, src_offset):
    assert x.dtype == torch.float32 and x.stride(-1) == 1
    rc = _L.load().mtt_zero_insert(_ptr(x), x.stride(-2), src_group, src_offset, B, h, w, Cdim, _ptr(out.hi),
                                   _ptr(out.lo), out.ld, _stream())
    _L.check(rc, "mtt_zero_insert")


def dwconv3x3_s2(x, weight, bias, out, *, B, T, h, w, Cdim):
    assert x.dtype == torch.float32 and x.stride(-1) == 1 and weight.is_contiguous() and bias.is_contiguous()
    rc = _L.load().mtt_dwconv3x3_s2(_ptr(x), x.stride(-2), B, T, h, w, Cdim, _ptr(weight), _ptr(bias),
                                    _ptr(out.hi), _ptr(out.lo), out.ld, _stream())
    _L.check(rc, "mtt_dwconv3x3_s2")


def avgpool(x, out, *, BT, h, w, Cdim, s):
    assert x.dtype == torch.float32 and x.stride(-1) == 1
    rc = _L.load().mtt_avgpool(_ptr(x), x.stride(-2), BT, h, w, Cdim, s, _ptr(out.hi), _ptr(out.lo), out.ld,
                               _stream())
    _L.check(rc, "mtt_avgpool")


def invpt_fuse_softmax(raw, P, *, B, Lq, Tk, scale, prev_score=None, T=0, qh=0, qw=0, fuse_w=None, fuse_b=None,
                       score_out=None):
    """The step between the two GEMMs of InvPT's cross-task attention (invpt.py:204-236): scale, cross-scale fusion with
    the up-sampled previous score, score export, softmax -> P Split [B*2*Lq, Tk]."""
    assert raw.is_contiguous() and raw.dtype == torch.float32
    if prev_score is not None:
        assert prev_score.is_contiguous() and fuse_w.is_contiguous()
    if score_out is not None:
        assert score_out.is_contiguous()
    rc = _L.load().mtt_invpt_fuse_softmax(_ptr(raw), B, Lq, Tk, float(scale), _ptr(prev_score), T, qh, qw, _ptr(fuse_w),
                                          _ptr(fuse_b), _ptr(score_out), _ptr(P.hi), _ptr(P.lo), P.ld, _stream())
    _L.check(rc, "mtt_invpt_fuse_softmax")


# ------------------------------------------------------------------------------------------------
# named operators of SURVEY.md section 8(b) (block_ops.cu) and the packing / layout entry points
# ------------------------------------------------------------------------------------------------
def _shape(rows=0, Cdim=0, hidden=0, nsplit=2, B=0, N=0, H=0, T=0):
    s = _L.Shape()
    s.rows, s.C, s.hidden, s.nsplit, s.B, s.N, s.H, s.T = rows, Cdim, hidden, nsplit, B, N, H, T
    return s


def _weight(w, nsplit):
    x = _L.Weight()
    x.hi, x.lo, x.ld = w.hi.data_ptr(), (w.lo.data_ptr() if nsplit == 2 else 0), w.ld
    return x


def workspace_bytes(op, **shape):
    return int(_L.load().mtt_workspace_bytes(op, C.byref(_shape(**shape))))


def workspace(nbytes, device):
    """A 256-byte aligned, ZERO-FILLED byte buffer (torch's caching allocator aligns to 512). Zero because the stream-K
    flag words inside the LayerNorm-fronted operators' workspaces must start at zero (include/mtt_b200.h, sk_ws)."""
    return torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def ws_split_view(ws, byte_offset, rows, cols, nsplit):
    """The Split living at `byte_offset` of a workspace, laid out as block_ops.cu does: [nsplit][rows][pad8(cols)]."""
    ld = round_up(cols, 8)
    n = nsplit * rows * ld
    sp = Split.__new__(Split)
    sp.rows, sp.cols, sp.ld, sp.nsplit = rows, cols, ld, nsplit
    sp.buf = ws[byte_offset: byte_offset + 2 * n].view(torch.bfloat16).view(nsplit, rows, ld)
    return sp


def ln_qkv(x, gamma, beta, eps, wqkv, bias, qkv, ws):
    """qkv = LN(x) @ Wqkv^T + b (taskprompter.py:272,:199,:201); LN(x) stays in `ws` as split planes."""
    rows, Cd = x.shape
    ns = min(wqkv.nsplit, qkv.nsplit)
    rc = _L.load().mtt_ln_qkv(_ptr(x), x.stride(0), _ptr(gamma), _ptr(beta), float(eps), C.byref(_weight(wqkv, ns)),
                              _ptr(bias), _ptr(qkv.hi), _ptr(qkv.lo), qkv.ld, C.byref(_shape(rows, Cd, nsplit=ns)),
                              _ptr(ws), ws.numel(), _stream())
    _L.check(rc, "mtt_ln_qkv")


def proj_residual(ao, wproj, bias, x):
    """x += ao @ Wproj^T + b in place (taskprompter.py:212,:273,:276)."""
    rows, Cd = x.shape
    ns = min(ao.nsplit, wproj.nsplit)
    rc = _L.load().mtt_proj_residual(_ptr(ao.hi), _ptr(ao.lo), ao.ld, C.byref(_weight(wproj, ns)), _ptr(bias), _ptr(x),
                                     x.stride(0), C.byref(_shape(rows, Cd, nsplit=ns)), _stream())
    _L.check(rc, "mtt_proj_residual")


def ln_mlp_residual(x, gamma, beta, eps, w1, b1, w2, b2, ws):
    """x += fc2(gelu(fc1(LN(x)))) in place (taskprompter.py:274,:277)."""
    rows, Cd = x.shape
    ns = min(w1.nsplit, w2.nsplit)
    rc = _L.load().mtt_ln_mlp_residual(_ptr(x), x.stride(0), _ptr(gamma), _ptr(beta), float(eps),
                                       C.byref(_weight(w1, ns)), _ptr(b1), C.byref(_weight(w2, ns)), _ptr(b2),
                                       C.byref(_shape(rows, Cd, hidden=w1.rows, nsplit=ns)), _ptr(ws), ws.numel(),
                                       _stream())
    _L.check(rc, "mtt_ln_mlp_residual")


def gated_conv1x1(x, x_group_rows, x_row_offset, prompt_logits, chan_lg, tasks, e, chan_col, ws, *, B, T, N, H, Cdim,
                  gh, gw, nh, nw):
    """Spatial + channel gating of ALL tasks of a level and their 2 * len(tasks) 1x1 decode convs (taskprompter.py:
    436-471) -- one gating launch, one grouped GEMM launch. tasks: [(w_spa, b_spa, w_chan, b_chan, cat Split)]."""
    cat0 = tasks[0][4]
    ns = min(tasks[0][0].nsplit, cat0.nsplit)
    arr = (_L.GatedTask * len(tasks))()
    for g, (w_spa, b_spa, w_chan, b_chan, cat) in zip(arr, tasks):
        assert cat.ld == cat0.ld
        g.w_spa, g.b_spa, g.w_chan, g.b_chan = _weight(w_spa, ns), b_spa.data_ptr(), _weight(w_chan, ns), b_chan.data_ptr()
        g.cat_hi, g.cat_lo = cat.hi.data_ptr(), (cat.lo.data_ptr() if ns == 2 else 0)
    rc = _L.load().mtt_gated_conv1x1(_ptr(x), x.stride(-2), x_group_rows, x_row_offset, _ptr(prompt_logits),
                                     _ptr(chan_lg), len(tasks), arr, gh, gw, nh, nw, e, cat0.ld, chan_col,
                                     C.byref(_shape(0, Cdim, nsplit=ns, B=B, N=N, H=H, T=T)), _ptr(ws), ws.numel(),
                                     _stream())
    _L.check(rc, "mtt_gated_conv1x1")


def conv3x3_bn_act(a, w3, b3, Cin, Cout, act, *, B, H, W, dil=1, mid=None, w_head=None, b_head=None, n_out=0,
                   out_f32=None, ws=None):
    """3x3 conv + folded BN + act on an NHWC Split (optionally followed by the fused 1x1 head -> out_f32)."""
    ns = min(a.nsplit, w3.nsplit)
    rc = _L.load().mtt_conv3x3_bn_act(
        _ptr(a.hi), _ptr(a.lo), a.ld, B, H, W, Cin, dil, C.byref(_weight(w3, ns)), _ptr(b3), Cout, act,
        _ptr(mid.hi) if mid is not None else None, _ptr(mid.lo) if mid is not None else None,
        mid.ld if mid is not None else 0, C.byref(_weight(w_head, ns)) if w_head is not None else None, _ptr(b_head),
        n_out, _ptr(out_f32), out_f32.stride(-2) if out_f32 is not None else 0, ns, _ptr(ws),
        ws.numel() if ws is not None else 0, _stream())
    _L.check(rc, "mtt_conv3x3_bn_act")


def pack_weight(w, nsplit):
    """fp32 [N, K] -> Split [N, K] through mtt_pack_weight."""
    assert w.dtype == torch.float32 and w.dim() == 2 and w.stride(1) == 1
    N, K = w.shape
    out = Split(N, round_up(K, 8), w.device, nsplit)
    out.cols = K
    rc = _L.load().mtt_pack_weight(_ptr(w), w.stride(0), N, K, nsplit, _ptr(out.hi), _ptr(out.lo), out.ld, _stream())
    _L.check(rc, "mtt_pack_weight")
    return out


def pack_conv_weight(w, bias, bn, nsplit, transposed=False):
    """Conv2d [N,Cin,k,k] (or ConvTranspose2d [Cin,N,k,k], transposed=True) + optional eval BatchNorm -> (Split
    [N, k*k*cin_pad] tap-major, folded bias fp32 [N]) through mtt_pack_conv_weight."""
    assert w.dtype == torch.float32 and w.is_contiguous() and w.dim() == 4 and w.shape[2] == w.shape[3]
    N, Cin = (w.shape[1], w.shape[0]) if transposed else (w.shape[0], w.shape[1])
    k = w.shape[2]
    cin_pad = round_up(Cin, 64)
    out = Split(N, k * k * cin_pad, w.device, nsplit)
    bias_out = torch.empty(N, dtype=torch.float32, device=w.device)
    scale = torch.empty(N, dtype=torch.float32, device=w.device)
    f = lambda t: t.detach().to(device=w.device, dtype=torch.float32).contiguous()
    if bn is not None:
        g, b, m, v, eps = f(bn.weight), f(bn.bias), f(bn.running_mean), f(bn.running_var), float(bn.eps)
    else:
        g = b = m = v = None
        eps = 0.0
    bias = f(bias) if bias is not None else None
    rc = _L.load().mtt_pack_conv_weight(_ptr(w), _ptr(bias), _ptr(g), _ptr(b), _ptr(m), _ptr(v), eps, N, Cin, k,
                                        1 if transposed else 0, nsplit, _ptr(out.hi), _ptr(out.lo), out.ld,
                                        _ptr(bias_out), _ptr(scale), _stream())
    _L.check(rc, "mtt_pack_conv_weight")
    return out, bias_out


def nchw_to_nhwc_split(x, out, col_offset=0):
    """x fp32 NCHW [B,C,H,W] -> columns [col_offset, col_offset + C) of out Split [B*H*W, >= C] (module-boundary
    forwards take NCHW like the reference)."""
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
    B, Cd, H, W = x.shape
    hi = C.c_void_p(out.hi.data_ptr() + 2 * col_offset)
    lo = C.c_void_p(out.lo.data_ptr() + 2 * col_offset) if out.nsplit == 2 else C.c_void_p(0)
    rc = _L.load().mtt_nchw_to_nhwc_split(_ptr(x), B, Cd, H, W, hi, lo, out.ld, _stream())
    _L.check(rc, "mtt_nchw_to_nhwc_split")


def nhwc_to_nchw(x, ld_in, B, Cd, H, W, out):
    assert x.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, Cd, H, W)
    rc = _L.load().mtt_nhwc_to_nchw(_ptr(x), ld_in, B, Cd, H, W, _ptr(out), _stream())
    _L.check(rc, "mtt_nhwc_to_nchw")


# ------------------------------------------------------------------------------------------------
# Swin-backbone TaskPrompter kernels (swin.cu)
# ------------------------------------------------------------------------------------------------
def swin_window_gather(xn, pn, out, *, B, H, W, Cdim, T, ws, shift):
    """LN1 outputs xn [B*H*W, C], pn [B*T, C] (fp32) -> out: split joint window stream [B*nW*(T + ws*ws), C]."""
    rc = _L.load().mtt_swin_window_gather(_ptr(xn), xn.stride(0), _ptr(pn), pn.stride(0), B, H, W, Cdim, T, ws, shift,
                                          _ptr(out.hi), _ptr(out.lo), out.ld, _stream())
    _L.check(rc, "mtt_swin_window_gather")


def swin_window_attention(qkv, out, raw, biasT, maskT, *, BW, nW, T, L, heads, scale):
    """Window attention with prompts; biasT [heads, L, L] / maskT [nW, L, L] are stored transposed ([.., key, query])."""
    Cd = out.cols
    assert biasT.is_contiguous() and (maskT is None or maskT.is_contiguous()) and raw.is_contiguous()
    rc = _L.load().mtt_swin_window_attention(_ptr(qkv.hi), _ptr(qkv.lo), qkv.ld, BW, nW, T, L, heads, Cd // heads,
                                             float(scale), _ptr(biasT), _ptr(maskT), _ptr(out.hi), _ptr(out.lo), out.ld,
                                             _ptr(raw), _stream())
    _L.check(rc, "mtt_swin_window_attention")


def swin_window_scatter(o32, raw, xa, x, p, logits, *, B, H, W, Cdim, T, ws, shift, heads, last):
    """proj output on the joint stream -> xa, x += xa, p += mean prompt rows (unless last), raw -> logits map."""
    rc = _L.load().mtt_swin_window_scatter(_ptr(o32), o32.stride(0), _ptr(raw), B, H, W, Cdim, T, ws, shift, heads,
                                           0 if last else 1, _ptr(xa), xa.stride(0), _ptr(x), x.stride(0), _ptr(p),
                                           p.stride(0), _ptr(logits), _stream())
    _L.check(rc, "mtt_swin_window_scatter")


def transpose_split(x, out, *, B, L, Cdim):
    """x fp32 [B*L, C] -> out Split [B*C, L] (per-image transpose)."""
    rc = _L.load().mtt_transpose_split(_ptr(x), x.stride(0), B, L, Cdim, _ptr(out.hi), _ptr(out.lo), out.ld, _stream())
    _L.check(rc, "mtt_transpose_split")


def swin_chan_attention(q, kv, co32, cos, rc_out, *, B, T, Cdim, ce, nh, nw):
    rc = _L.load().mtt_swin_chan_attention(_ptr(q), q.stride(0), _ptr(kv), kv.stride(0), B, T, Cdim, ce, nh, nw,
                                           _ptr(co32), co32.stride(0), _ptr(cos.hi), _ptr(cos.lo), cos.ld, _ptr(rc_out),
                                           _stream())
    _L.check(rc, "mtt_swin_chan_attention")


def swin_merge_gather(x, out, *, B, H, W, Cdim):
    rc = _L.load().mtt_swin_merge_gather(_ptr(x), x.stride(0), B, H, W, Cdim, _ptr(out), out.stride(0), _stream())
    _L.check(rc, "mtt_swin_merge_gather")


def conv3x3_s2_maps(x, w, b, out, *, B, Cin, H, W, in_stride, in_offset, out_stride, out_offset):
    assert w.is_contiguous() and x.is_contiguous() and out.is_contiguous()
    rc = _L.load().mtt_conv3x3_s2_maps(_ptr(x), _ptr(w), _ptr(b), B, Cin, w.shape[0], H, W, in_stride, in_offset,
                                       out_stride, out_offset, _ptr(out), _stream())
    _L.check(rc, "mtt_conv3x3_s2_maps")


def swin_chan_up(rc_in, w, out, *, BT, Cdim, nwin):
    assert rc_in.is_contiguous() and w.is_contiguous() and out.is_contiguous()
    rc = _L.load().mtt_swin_chan_up(_ptr(rc_in), _ptr(w), BT, Cdim, w.shape[0], nwin, _ptr(out), _stream())
    _L.check(rc, "mtt_swin_chan_up")


# ---- training step (csrc/train_ops.cu): fp32 [rows, cols] tensors, last dim contiguous ----------------------------------
def _ld(t):
    assert t.dtype == torch.float32 and t.stride(-1) == 1
    return t.stride(-2) if t.dim() >= 2 else t.shape[-1]


def colsum(x, out, accumulate=False, *, rows=None, in_group=0, src_group=0, src_offset=0):
    """out[c] (+)= sum_r x[row(r), c]; row mapping as split_rows (in_group = 0: the first `rows` rows)."""
    rows = x.shape[0] if rows is None else rows
    _L.check(_L.load().mtt_colsum(_ptr(x), _ld(x), rows, x.shape[1], in_group, src_group, src_offset, _ptr(out),
                                  int(accumulate), _stream()), "mtt_colsum")


def layernorm_bwd(x, dy, gamma, eps, dx, dgamma, dbeta, *, accumulate_dx=False):
    """dx (+)= d LN(x) . dy; dgamma / dbeta are ACCUMULATED (None: skip the parameter gradients)."""
    rows, cols = x.shape
    ws = torch.empty(2 * rows, dtype=torch.float32, device=x.device)
    _L.check(_L.load().mtt_layernorm_bwd(_ptr(x), _ld(x), _ptr(dy), _ld(dy), _ptr(gamma), float(eps), rows, cols, _ptr(dx),
                                         _ld(dx), int(accumulate_dx), _ptr(dgamma), _ptr(dbeta), _ptr(ws), _stream()),
             "mtt_layernorm_bwd")


def act_split(pre, act, out=None, nsplit=2):
    rows, cols = pre.shape
    if out is None:
        out = Split(rows, cols, pre.device, nsplit)
    _L.check(_L.load().mtt_act_split(_ptr(pre), _ld(pre), rows, cols, act, _ptr(out.hi), _ptr(out.lo), out.ld, _stream()),
             "mtt_act_split")
    return out


def act_bwd(pre, dy, act, dx):
    rows, cols = pre.shape
    _L.check(_L.load().mtt_act_bwd(_ptr(pre), _ld(pre), _ptr(dy), _ld(dy), rows, cols, act, _ptr(dx), _ld(dx), _stream()),
             "mtt_act_bwd")


def axpy_rows(base, src, row_scale, dst):
    """dst = base + row_scale[:, None] * src (base / row_scale may be None)."""
    rows, cols = src.shape
    _L.check(_L.load().mtt_axpy_rows(_ptr(base), _ld(base) if base is not None else 0, _ptr(src), _ld(src),
                                     _ptr(row_scale), rows, cols, _ptr(dst), _ld(dst), _stream()), "mtt_axpy_rows")


def transpose_planes(a, *, B=1, R=None, Ccols=None, in_batch_rows=None, out=None, side_by_side=False):
    """Split [B*R(+), C] -> Split: image b (rows b*in_batch_rows ... + R of `a`) transposed to a [C, R] block; blocks are
    stacked by rows ([B*C, ld >= R]) or, side_by_side, along the columns of one [C, B*R] matrix."""
    R = a.rows // B if R is None else R
    Ccols = a.cols if Ccols is None else Ccols
    in_batch_rows = R if in_batch_rows is None else in_batch_rows
    if out is None:
        rows, cols = (Ccols, B * R) if side_by_side else (B * Ccols, R)
        out = Split(rows, cols, a.hi.device, a.nsplit)     # pad columns are never read (TMA bounds)
    _L.check(_L.load().mtt_transpose_planes(_ptr(a.hi), _ptr(a.lo), a.ld, in_batch_rows, B, R, Ccols, _ptr(out.hi),
                                            _ptr(out.lo), out.ld, R if side_by_side else 0, _stream()),
             "mtt_transpose_planes")
    return out


def bn_stats(x, sums):
    _L.check(_L.load().mtt_bn_stats(_ptr(x), _ld(x), x.shape[0], x.shape[1], _ptr(sums), _stream()), "mtt_bn_stats")


def bn_finalize(sums, count, eps, momentum, mean_rstd, running_mean=None, running_var=None):
    _L.check(_L.load().mtt_bn_finalize(_ptr(sums), float(count), sums.numel() // 2, float(eps), float(momentum),
                                       _ptr(mean_rstd), _ptr(running_mean), _ptr(running_var), _stream()), "mtt_bn_finalize")


def bn_act(x, mean_rstd, gamma, beta, act, *, out_f32=None, out_split=None):
    rows, cols = x.shape
    _L.check(_L.load().mtt_bn_act(_ptr(x), _ld(x), rows, cols, _ptr(mean_rstd), _ptr(gamma), _ptr(beta), act,
                                  _ptr(out_f32), _ld(out_f32) if out_f32 is not None else 0,
                                  _ptr(out_split.hi) if out_split is not None else None,
                                  _ptr(out_split.lo) if out_split is not None else None,
                                  out_split.ld if out_split is not None else 0, _stream()), "mtt_bn_act")


def bn_bwd_reduce(x, dy, mean_rstd, gamma, beta, act, sums):
    rows, cols = x.shape
    _L.check(_L.load().mtt_bn_bwd_reduce(_ptr(x), _ld(x), _ptr(dy), _ld(dy), rows, cols, _ptr(mean_rstd), _ptr(gamma),
                                         _ptr(beta), act, _ptr(sums), _stream()), "mtt_bn_bwd_reduce")


def bn_bwd_apply(x, dy, mean_rstd, gamma, beta, act, sums, count, dx):
    rows, cols = x.shape
    _L.check(_L.load().mtt_bn_bwd_apply(_ptr(x), _ld(x), _ptr(dy), _ld(dy), rows, cols, _ptr(mean_rstd), _ptr(gamma),
                                        _ptr(beta), act, _ptr(sums), float(count), _ptr(dx), _ld(dx), _stream()),
             "mtt_bn_bwd_apply")


def attn_delta(dO, o, delta, *, B, N, H, head_dim):
    """delta [B*H*N] = rowdot(dO, O) per head: dO fp32 [B*N, H*head_dim], o = the forward's attention output (Split)."""
    _L.check(_L.load().mtt_attn_delta(_ptr(dO), _ld(dO), _ptr(o.hi), _ptr(o.lo), o.ld, B, N, H, head_dim, _ptr(delta),
                                      _stream()), "mtt_attn_delta")


def attn_softmax_bwd(S, dP, delta, *, BH, N, scale, d_raw, T, ds, pt=None, dst=None):
    """S, dP fp32 [BH*N, ld] (read only), delta fp32 [BH*N] -> Splits ds [BH*N queries, >= N], and optionally pt = P^T,
    dst = dS^T [BH*N keys, >= N queries] (same ld as ds)."""
    assert (pt is None) == (dst is None) and (pt is None or pt.ld == dst.ld == ds.ld)
    _L.check(_L.load().mtt_attn_softmax_bwd(_ptr(S), _ptr(dP), _ptr(delta), _ld(S), BH, N, float(scale), _ptr(d_raw), T, _ptr(ds.hi),
                                            _ptr(ds.lo), _ptr(pt.hi) if pt is not None else None,
                                            _ptr(pt.lo) if pt is not None else None,
                                            _ptr(dst.hi) if dst is not None else None,
                                            _ptr(dst.lo) if dst is not None else None, ds.ld, _stream()),
             "mtt_attn_softmax_bwd")


def bilinear_bwd(dy, *, nchw, B, h, w, Cdim, H2, W2, dx, accumulate=False):
    _L.check(_L.load().mtt_bilinear_bwd(_ptr(dy), 0 if nchw else _ld(dy), int(nchw), B, h, w, Cdim, H2, W2, _ptr(dx),
                                        _ld(dx), int(accumulate), _stream()), "mtt_bilinear_bwd")


def gate_bwd(x, x_group_rows, x_row_offset, prompt_logits, chan_lg, task, dys, dyc, dx, d_prompt_logits, d_chan_lg, *, B,
             T, N, H, Cdim, gh, gw, nh, nw):
    _L.check(_L.load().mtt_gate_bwd(_ptr(x), _ld(x), x_group_rows, x_row_offset, _ptr(prompt_logits), _ptr(chan_lg), task,
                                    B, T, N, H, Cdim, gh, gw, nh, nw, _ptr(dys), _ptr(dyc), _ld(dys), _ptr(dx), _ld(dx),
                                    _ptr(d_prompt_logits), _ptr(d_chan_lg), _stream()), "mtt_gate_bwd")


def chan_logits_bwd(d_rc, cp, xn, dcp, dxn, *, B, N, T, Cdim, gh, gw, nh, nw):
    _L.check(_L.load().mtt_chan_logits_bwd(_ptr(d_rc), _ptr(cp), _ptr(xn.hi), _ptr(xn.lo), xn.ld, B, N, T, Cdim, gh, gw, nh,
                                           nw, _ptr(dcp), _ptr(dxn), _ld(dxn), _stream()), "mtt_chan_logits_bwd")


def ctr_bwd(dnew, F, prompt_logits, w0, b0, w2, d_prompt_logits, dw0, db0, dw2, db2, *, T, M, Cdim, ld, rows_per_batch, B,
            H, N):
    ws = torch.empty(B * T * T, dtype=torch.float32, device=dnew.device)
    _L.check(_L.load().mtt_ctr_bwd(_ptr(dnew), _ptr(F), T, M, Cdim, ld, rows_per_batch, _ptr(prompt_logits), B, H, N,
                                   _ptr(w0), _ptr(b0), _ptr(w2), _ptr(ws), _ptr(d_prompt_logits), _ptr(dw0), _ptr(db0),
                                   _ptr(dw2), _ptr(db2), _stream()), "mtt_ctr_bwd")


def im2col3x3_t(x, *, B, H, W, Cdim, nsplit=2):
    """NHWC fp32 [B*H*W, C] -> Split [C*9, B*H*W]: rows (c, ky, kx)."""
    P = B * H * W
    out = Split(Cdim * 9, P, x.device, nsplit)
    _L.check(_L.load().mtt_im2col3x3_t(_ptr(x), _ld(x), B, H, W, Cdim, _ptr(out.hi), _ptr(out.lo), out.ld, _stream()),
             "mtt_im2col3x3_t")
    return out


def im2col_patch_t(img, patch, nsplit=2):
    """NCHW fp32 image -> Split [Cin*patch*patch, B*gh*gw]."""
    B, Cin, H, W = img.shape
    cols = B * (H // patch) * (W // patch)
    out = Split(Cin * patch * patch, cols, img.device, nsplit)
    _L.check(_L.load().mtt_im2col_patch_t(_ptr(img), B, Cin, H, W, patch, _ptr(out.hi), _ptr(out.lo), out.ld, _stream()),
             "mtt_im2col_patch_t")
    return out


def sumsq(g, out, accumulate=False):
    _L.check(_L.load().mtt_sumsq(_ptr(g), g.numel(), _ptr(out), int(accumulate), _stream()), "mtt_sumsq")


def adam_step(p, g, m, v, *, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, step, gnorm_sq=None, max_norm=0.0,
              grad_scale=1.0):
    _L.check(_L.load().mtt_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), float(lr), float(betas[0]),
                                     float(betas[1]), float(eps), float(weight_decay), int(step), _ptr(gnorm_sq),
                                     float(max_norm), float(grad_scale), _stream()), "mtt_adam_step")
