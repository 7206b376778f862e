// InvPT decoder kernels (reference: InvPT/models/transformers/invpt.py, transformer_decoder.py).
// The decoder's GEMMs / 3x3 convolutions run on gemm_tc.cu; this file holds the HBM / latency-bound
// pieces between them: row gathers, segmented (joint-channel) LayerNorm, depthwise-conv and
// average-pool token reductions, zero-insertion for the transposed convolution, and the small-KV
// cross-task attention with cross-scale score fusion.
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace mtt {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// physical row of logical row r: (r / in_group) * src_group + src_off + r % in_group
__device__ __forceinline__ long long map_row(long long r, long long in_group, long long src_group,
                                             long long src_off) {
  return in_group > 0 ? (r / in_group) * src_group + src_off + r % in_group : r + src_off;
}

// ------------------------------------------------------------------------------------------------
// fp32 rows (gathered) -> split planes (dense).  e.g. x[:, 1:] of the ViT stream (vit.py:345-346).
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ in, long long ld_in, long long in_group, long long src_group,
                  long long src_off, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                  long long ld_out, long long rows, int cols) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* src = in + map_row(row, in_group, src_group, src_off) * ld_in;
  for (int c = lane * 2; c < cols; c += 64) {
    const bool two = c + 1 < cols;
    uint32_t h, l;
    split_pack2(src[c], two ? src[c + 1] : 0.f, h, l);
    if (two && !(ld_out & 1)) {
      *reinterpret_cast<uint32_t*>(hi + row * ld_out + c) = h;
      if (lo) *reinterpret_cast<uint32_t*>(lo + row * ld_out + c) = l;
    } else {
      hi[row * ld_out + c] = __ushort_as_bfloat16((unsigned short)(h & 0xFFFF));
      if (lo) lo[row * ld_out + c] = __ushort_as_bfloat16((unsigned short)(l & 0xFFFF));
      if (two) {
        hi[row * ld_out + c + 1] = __ushort_as_bfloat16((unsigned short)(h >> 16));
        if (lo) lo[row * ld_out + c + 1] = __ushort_as_bfloat16((unsigned short)(l >> 16));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Segmented LayerNorm: a logical row is S segments of `cols` channels living in S physical rows
// (segment s at map_row(r) + s*seg_stride); statistics over all S*cols values, gamma/beta [S*cols].
// S = 1 with a row gather is the ViT's final norm on x[:, 1:]; S = T is InvPT's joint-channel
// LayerNorm over all tasks' tokens (invpt.py:524-526).  Segment s of logical row r is written to
// output row s*out_seg_stride + r.
__global__ void __launch_bounds__(256)
layernorm_seg_kernel(const float* __restrict__ in, long long ld_in, long long in_group, long long src_group,
                     long long src_off, long long seg_stride, int S, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, float* __restrict__ out_f32, long long ld_f32,
                     __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, long long ld_bf,
                     long long out_seg_stride, long long rows, int cols) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* base = in + map_row(row, in_group, src_group, src_off) * ld_in;
  const float n = (float)S * (float)cols;
  float s = 0.f;
  for (int k = 0; k < S; ++k) {
    const float* src = base + (long long)k * seg_stride * ld_in;
    for (int c = lane; c < cols; c += 32) s += src[c];
  }
  const float mean = warp_sum_f(s) / n;
  float ss = 0.f;
  for (int k = 0; k < S; ++k) {
    const float* src = base + (long long)k * seg_stride * ld_in;
    for (int c = lane; c < cols; c += 32) {
      const float a = src[c] - mean;
      ss += a * a;
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum_f(ss) / n + eps);
  for (int k = 0; k < S; ++k) {
    const float* src = base + (long long)k * seg_stride * ld_in;
    const long long orow = (long long)k * out_seg_stride + row;
    for (int c = lane; c < cols; c += 32) {
      const float y = (src[c] - mean) * rstd * gamma[k * cols + c] + beta[k * cols + c];
      if (out_f32) out_f32[orow * ld_f32 + c] = y;
      if (out_hi) {
        __nv_bfloat16 h, l;
        split_bf16(y, h, l);
        out_hi[orow * ld_bf + c] = h;
        if (out_lo) out_lo[orow * ld_bf + c] = l;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Zero insertion (stride-2 up-sampling with zeros) so that ConvTranspose2d(k3, s2, p1, op1)
// (transformer_decoder.py:63) becomes a plain 3x3 convolution with the flipped kernel.
__global__ void __launch_bounds__(256)
zero_insert_kernel(const float* __restrict__ in, long long ld_in, long long src_group, long long src_off,
                   int h, int w, int C, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                   long long ld_out) {
  const long long opix = blockIdx.x;  // (b * 2h + Y) * 2w + X
  const int X = (int)(opix % (2 * w)), Y = (int)((opix / (2 * w)) % (2 * h));
  const int b = (int)(opix / ((long long)4 * w * h));
  const bool src_ok = !(X & 1) && !(Y & 1);
  const float* src = in + ((long long)b * src_group + src_off + (long long)(Y >> 1) * w + (X >> 1)) * ld_in;
  for (int c = threadIdx.x * 2; c < C; c += blockDim.x * 2) {
    uint32_t hh = 0, ll = 0;
    if (src_ok) split_pack2(src[c], src[c + 1], hh, ll);
    *reinterpret_cast<uint32_t*>(hi + opix * ld_out + c) = hh;
    if (lo) *reinterpret_cast<uint32_t*>(lo + opix * ld_out + c) = ll;
  }
}

// ------------------------------------------------------------------------------------------------
// Per-task depthwise 3x3 stride-2 conv (+ folded eval BatchNorm) producing the Q tokens
// (invpt.py:125-137,171-173).  in: fp32 joint tokens [B, T*h*w, C]; out: split joint [B, T*(h/2)(w/2), C].
__global__ void __launch_bounds__(256)
dwconv_s2_kernel(const float* __restrict__ in, long long ld_in, int T, int h, int w, int C,
                 const float* __restrict__ wgt /*[T,C,9]*/, const float* __restrict__ bias /*[T,C]*/,
                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ld_out) {
  const int oh = h / 2, ow = w / 2;
  const long long orow = blockIdx.x;  // ((b*T + k) * oh + oy) * ow + ox
  const int ox = (int)(orow % ow), oy = (int)((orow / ow) % oh);
  const long long bk = orow / ((long long)ow * oh);
  const int k = (int)(bk % T);
  const float* base = in + bk * (long long)h * w * ld_in;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = bias[k * C + c];
    const float* wc = wgt + ((long long)k * C + c) * 9;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= w) continue;
        acc = fmaf(wc[ky * 3 + kx], base[((long long)iy * w + ix) * ld_in + c], acc);
      }
    }
    __nv_bfloat16 hh, ll;
    split_bf16(acc, hh, ll);
    hi[orow * ld_out + c] = hh;
    if (lo) lo[orow * ld_out + c] = ll;
  }
}

// Per-task average pooling, kernel = stride = s, ceil_mode, no padding (invpt.py:139-147).
__global__ void __launch_bounds__(256)
avgpool_kernel(const float* __restrict__ in, long long ld_in, int h, int w, int C, int s, int oh, int ow,
               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ld_out) {
  const long long orow = blockIdx.x;  // (bk * oh + oy) * ow + ox
  const int ox = (int)(orow % ow), oy = (int)((orow / ow) % oh);
  const long long bk = orow / ((long long)ow * oh);
  const float* base = in + bk * (long long)h * w * ld_in;
  const int y0 = oy * s, x0 = ox * s;
  const int y1 = min(y0 + s, h), x1 = min(x0 + s, w);
  const float inv = 1.0f / (float)((y1 - y0) * (x1 - x0));
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) acc += base[((long long)y * w + x) * ld_in + c];
    __nv_bfloat16 hh, ll;
    split_bf16(acc * inv, hh, ll);
    hi[orow * ld_out + c] = hh;
    if (lo) lo[orow * ld_out + c] = ll;
  }
}

// ------------------------------------------------------------------------------------------------
// InvPT cross-task attention (invpt.py:204-236), 2 heads, Tk <= 512 keys shared by all tasks:
//   s[h]     = (q_h . k_h) * C^-1/2
//   prev_up  = bilinear x2 (per task, over the query grid) of the previous stage's fused score
//   f[o]     = Wf[o,0] s[0] + Wf[o,1] s[1] + Wf[o,2] prev_up[0] + Wf[o,3] prev_up[1] + bf[o]
//   score_out = f (pre-softmax, consumed by the next stage);  out = softmax(f) v
constexpr int kQT = 16;       // queries per block
constexpr int kIAThreads = 256;
constexpr int kMaxTk = 512;

__device__ __forceinline__ void up2_coord(int d, int in_size, int& i0, int& i1, float& l1) {
  float s = 0.5f * (d + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

__global__ void __launch_bounds__(kIAThreads)
invpt_attn_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                  long long ldq, long long ldk, int Lq, int Tk, int C, float scale,
                  const float* __restrict__ prev, int T, int qh, int qw, const float* __restrict__ wf,
                  const float* __restrict__ bf, float* __restrict__ score_out,
                  __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, long long ldo) {
  extern __shared__ float sm[];
  float* qs = sm;                      // [kQT][C]
  float* ks = qs + kQT * C;            // [Tk][33]
  float* sc = ks + (size_t)Tk * 33;    // [2][kQT][Tk]
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * kQT;
  const int nq = min(kQT, Lq - l0);
  const int d = C / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < kQT * C; i += kIAThreads) {
    const int qi = i / C, c = i % C;
    qs[i] = qi < nq ? q[((long long)b * Lq + l0 + qi) * ldq + c] : 0.f;
  }
  // ---- raw scores, one head at a time; thread owns keys t = tid, tid + 256
  for (int hd = 0; hd < 2; ++hd) {
    float acc[2][kQT];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int qi = 0; qi < kQT; ++qi) acc[u][qi] = 0.f;
    for (int c0 = hd * d; c0 < (hd + 1) * d; c0 += 32) {
      const int cw = min(32, (hd + 1) * d - c0);
      __syncthreads();
      for (int t = warp; t < Tk; t += kIAThreads / 32)
        ks[t * 33 + lane] = lane < cw ? k[((long long)b * Tk + t) * ldk + c0 + lane] : 0.f;
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int t = tid + u * kIAThreads;
        if (t < Tk) {
          for (int j = 0; j < cw; ++j) {
            const float kv = ks[t * 33 + j];
#pragma unroll
            for (int qi = 0; qi < kQT; ++qi) acc[u][qi] = fmaf(qs[qi * C + c0 + j], kv, acc[u][qi]);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int t = tid + u * kIAThreads;
      if (t < Tk)
#pragma unroll
        for (int qi = 0; qi < kQT; ++qi) sc[(hd * kQT + qi) * Tk + t] = acc[u][qi] * scale;
    }
  }
  __syncthreads();
  // ---- cross-scale fusion (1x1 conv over [cur h0, cur h1, prev h0, prev h1]) and score export
  if (prev != nullptr) {
    const int sh = qh / 2, sw = qw / 2;
    const int Lp = T * sh * sw;
    for (int i = tid; i < nq * Tk; i += kIAThreads) {
      const int qi = i / Tk, t = i % Tk;
      const int l = l0 + qi;
      const int task = l / (qh * qw), r = l % (qh * qw);
      int y0, y1, x0, x1;
      float ly, lx;
      up2_coord(r / qw, sh, y0, y1, ly);
      up2_coord(r % qw, sw, x0, x1, lx);
      float pu[2];
#pragma unroll
      for (int hd = 0; hd < 2; ++hd) {
        const float* pb = prev + (((long long)b * 2 + hd) * Lp + (long long)task * sh * sw) * Tk + t;
        const float p00 = pb[(long long)(y0 * sw + x0) * Tk], p01 = pb[(long long)(y0 * sw + x1) * Tk];
        const float p10 = pb[(long long)(y1 * sw + x0) * Tk], p11 = pb[(long long)(y1 * sw + x1) * Tk];
        pu[hd] = (1.f - ly) * ((1.f - lx) * p00 + lx * p01) + ly * ((1.f - lx) * p10 + lx * p11);
      }
      const float s0 = sc[(0 * kQT + qi) * Tk + t], s1 = sc[(1 * kQT + qi) * Tk + t];
      const float f0 = wf[0] * s0 + wf[1] * s1 + wf[2] * pu[0] + wf[3] * pu[1] + bf[0];
      const float f1 = wf[4] * s0 + wf[5] * s1 + wf[6] * pu[0] + wf[7] * pu[1] + bf[1];
      sc[(0 * kQT + qi) * Tk + t] = f0;
      sc[(1 * kQT + qi) * Tk + t] = f1;
    }
    __syncthreads();
  }
  if (score_out != nullptr) {
    for (int i = tid; i < 2 * nq * Tk; i += kIAThreads) {
      const int hd = i / (nq * Tk), r = i % (nq * Tk);
      const int qi = r / Tk, t = r % Tk;
      score_out[(((long long)b * 2 + hd) * Lq + l0 + qi) * Tk + t] = sc[(hd * kQT + qi) * Tk + t];
    }
  }
  // ---- softmax over keys, one warp per (head, query) row
  for (int rowi = warp; rowi < 2 * kQT; rowi += kIAThreads / 32) {
    float* row = sc + (long long)rowi * Tk;
    float mx = -INFINITY;
    for (int t = lane; t < Tk; t += 32) mx = fmaxf(mx, row[t]);
    mx = warp_max_f(mx);
    float sum = 0.f;
    for (int t = lane; t < Tk; t += 32) {
      const float e = expf(row[t] - mx);
      row[t] = e;
      sum += e;
    }
    const float inv = 1.0f / warp_sum_f(sum);
    for (int t = lane; t < Tk; t += 32) row[t] *= inv;
  }
  __syncthreads();
  // ---- out[qi, c] = sum_t p[head(c)][qi][t] * v[t, c]; thread owns channels
  for (int c = tid; c < C; c += kIAThreads) {
    const int hd = c / d;
    float acc[kQT];
#pragma unroll
    for (int qi = 0; qi < kQT; ++qi) acc[qi] = 0.f;
    const float* vb = v + (long long)b * Tk * ldk + c;
    const float* pr = sc + (long long)hd * kQT * Tk;
    for (int t = 0; t < Tk; ++t) {
      const float vv = vb[(long long)t * ldk];
#pragma unroll
      for (int qi = 0; qi < kQT; ++qi) acc[qi] = fmaf(pr[qi * Tk + t], vv, acc[qi]);
    }
    for (int qi = 0; qi < nq; ++qi) {
      __nv_bfloat16 hh, ll;
      split_bf16(acc[qi], hh, ll);
      const long long orow = (long long)b * Lq + l0 + qi;
      out_hi[orow * ldo + c] = hh;
      if (out_lo) out_lo[orow * ldo + c] = ll;
    }
  }
}

}  // namespace mtt

using namespace mtt;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int mtt_split_rows(const float* in, int64_t ld_in, int64_t in_group, int64_t src_group,
                              int64_t src_offset, void* out_hi, void* out_lo, int64_t ld_out, int64_t rows,
                              int32_t cols, mtt_stream_t stream) {
  if (!in || !out_hi || rows <= 0 || cols <= 0 || ld_out < cols)
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_split_rows: bad arguments");
  split_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, STREAM>>>(
      in, ld_in, in_group, src_group, src_offset, static_cast<__nv_bfloat16*>(out_hi),
      static_cast<__nv_bfloat16*>(out_lo), ld_out, rows, cols);
  return check_launch("mtt_split_rows");
}

extern "C" int mtt_layernorm_seg(const float* in, int64_t ld_in, int64_t in_group, int64_t src_group,
                                 int64_t src_offset, int64_t seg_stride, int32_t S, const float* gamma,
                                 const float* beta, float eps, float* out_f32, int64_t ld_f32, void* out_hi,
                                 void* out_lo, int64_t ld_bf, int64_t out_seg_stride, int64_t rows,
                                 int32_t cols, mtt_stream_t stream) {
  if (!in || !gamma || !beta || rows <= 0 || cols <= 0 || S <= 0 || (!out_f32 && !out_hi))
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_layernorm_seg: bad arguments");
  layernorm_seg_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, STREAM>>>(
      in, ld_in, in_group, src_group, src_offset, seg_stride, S, gamma, beta, eps, out_f32, ld_f32,
      static_cast<__nv_bfloat16*>(out_hi), static_cast<__nv_bfloat16*>(out_lo), ld_bf, out_seg_stride, rows,
      cols);
  return check_launch("mtt_layernorm_seg");
}

extern "C" int mtt_zero_insert(const float* in, int64_t ld_in, int64_t src_group, int64_t src_offset,
                               int32_t B, int32_t h, int32_t w, int32_t C, void* out_hi, void* out_lo,
                               int64_t ld_out, mtt_stream_t stream) {
  if (!in || !out_hi || B <= 0 || h <= 0 || w <= 0 || C <= 0 || (C & 1) || (ld_out & 1))
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_zero_insert: bad arguments");
  zero_insert_kernel<<<(unsigned)((long long)B * 4 * h * w), 256, 0, STREAM>>>(
      in, ld_in, src_group, src_offset, h, w, C, static_cast<__nv_bfloat16*>(out_hi),
      static_cast<__nv_bfloat16*>(out_lo), ld_out);
  return check_launch("mtt_zero_insert");
}

extern "C" int mtt_dwconv3x3_s2(const float* in, int64_t ld_in, int32_t B, int32_t T, int32_t h, int32_t w,
                                int32_t C, const float* weight, const float* bias, void* out_hi,
                                void* out_lo, int64_t ld_out, mtt_stream_t stream) {
  if (!in || !weight || !bias || !out_hi || B <= 0 || T <= 0 || h < 2 || w < 2 || (h & 1) || (w & 1))
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_dwconv3x3_s2: bad arguments (h=%d w=%d must be even)", h, w);
  const long long rows = (long long)B * T * (h / 2) * (w / 2);
  const int threads = C >= 256 ? 256 : (C >= 128 ? 128 : 64);
  dwconv_s2_kernel<<<(unsigned)rows, threads, 0, STREAM>>>(in, ld_in, T, h, w, C, weight, bias,
                                                          static_cast<__nv_bfloat16*>(out_hi),
                                                          static_cast<__nv_bfloat16*>(out_lo), ld_out);
  return check_launch("mtt_dwconv3x3_s2");
}

extern "C" int mtt_avgpool(const float* in, int64_t ld_in, int32_t BT, int32_t h, int32_t w, int32_t C,
                           int32_t s, void* out_hi, void* out_lo, int64_t ld_out, mtt_stream_t stream) {
  if (!in || !out_hi || BT <= 0 || h <= 0 || w <= 0 || s <= 0)
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_avgpool: bad arguments");
  const int oh = (h + s - 1) / s, ow = (w + s - 1) / s;
  const int threads = C >= 256 ? 256 : (C >= 128 ? 128 : 64);
  avgpool_kernel<<<(unsigned)((long long)BT * oh * ow), threads, 0, STREAM>>>(
      in, ld_in, h, w, C, s, oh, ow, static_cast<__nv_bfloat16*>(out_hi),
      static_cast<__nv_bfloat16*>(out_lo), ld_out);
  return check_launch("mtt_avgpool");
}

extern "C" int mtt_invpt_attention(const mtt_invpt_attn_desc* d, mtt_stream_t stream) {
  if (!d || !d->q || !d->k || !d->v || !d->out_hi || d->B <= 0 || d->Lq <= 0 || d->Tk <= 0 ||
      d->Tk > kMaxTk || d->C <= 0 || (d->C & 1))
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_invpt_attention: bad arguments (Tk=%d C=%d)", d ? d->Tk : -1,
                     d ? d->C : -1);
  if (d->prev_score && (d->T <= 0 || d->qh <= 0 || d->qw <= 0 || (d->qh & 1) || (d->qw & 1) ||
                        d->T * d->qh * d->qw != d->Lq || !d->fuse_w || !d->fuse_b))
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_invpt_attention: bad fusion geometry");
  const size_t smem = ((size_t)kQT * d->C + (size_t)d->Tk * 33 + (size_t)2 * kQT * d->Tk) * sizeof(float);
  if (smem > 220 * 1024) return set_error(MTT_ERR_BAD_SHAPE, "mtt_invpt_attention: shared memory %zu", smem);
  static bool attr[kMaxDevices] = {};  // per device
  const int dev_ = current_device();
  if (!attr[dev_]) {
    cudaFuncSetAttribute(invpt_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    attr[dev_] = true;
  }
  dim3 grid((d->Lq + kQT - 1) / kQT, d->B);
  invpt_attn_kernel<<<grid, kIAThreads, smem, STREAM>>>(
      d->q, d->k, d->v, d->ldq, d->ldk, d->Lq, d->Tk, d->C, d->scale, d->prev_score, d->T, d->qh, d->qw,
      d->fuse_w, d->fuse_b, d->score_out, static_cast<__nv_bfloat16*>(d->out_hi),
      static_cast<__nv_bfloat16*>(d->out_lo), d->ldo);
  return check_launch("mtt_invpt_attention");
}
