// Shared between the 1-CTA (gemm_tc.cu) and 2-CTA (gemm2_tc.cu) tcgen05 GEMM / implicit-conv kernels:
// tile constants, the kernel parameter block, the A-tile TMA issue and the fused epilogue.
#pragma once
#include "host_common.h"
#include "ptx.cuh"

namespace mtt {

constexpr int BM = 128;  // rows of A (output pixels / tokens) per CTA
constexpr int BK = 64;   // K per pipeline stage: one 128-byte swizzle row of bf16
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;
constexpr uint32_t kTileBytes = BM * BK * 2;  // 128 x 64 bf16 = 16 KB

struct GemmParams {
  int M, N;
  int num_kb, taps, ksize, dil, mode;
  int NB, H, W, TW, TH, tiles_x, tiles_y;
  int tiles_m;  // number of 128-row sub-tiles
  int tiles_n;  // number of N tiles (of the kernel's BN)
  int cin_pad;
  int k_last_steps;  // 16-wide MMA k-steps holding data in the last k-block of a tap (1..4); the rest is zero padding
  uint32_t a_box_bytes;
  const float* bias;
  int act;
  const float* residual;
  long long ldr;
  int res_row_mod;
  float* out_f32;
  long long ldo_f32;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  long long ldo_bf;
  int in_group, out_group, out_offset;
  int out_row_stride;  // >= 1: stride of the in-group row index (see mtt_gemm_desc)
  int vec_ok;
  int vec32_ok;  // every epilogue pointer / stride allows 32-byte accesses
  int a_groups_per_tile;  // >0: A rows are gathered in groups through a rank-3 tensor map
  int debug;              // profiling aid (env MTT_GEMM_DEBUG): bit0 = skip TMA loads, bit1 = skip global stores
  // stream-K tail of the CTA-pair kernel (gemm2_tc.cu): the first sk_tiles tiles are split along K over ALL pairs
  int sk_tiles;           // 0: every tile is computed by one pair
  float* sk_part;         // [pair][cta rank][128 rows][256 columns] fp32 partial accumulators
  unsigned int* sk_flags; // [pair][cta rank][epilogue warp]: 1 = that warp's slice of the partial is published
};

// Stream-K workspace of the CTA-pair kernel for `pairs` CTA pairs: flags first (zero before the first launch and
// left zero by every launch), then one 256 x 256 fp32 partial tile per pair.
constexpr size_t kSkFlagBytes = 16384;  // up to 256 pairs x 2 CTAs x 8 warps x 4 bytes
constexpr size_t sk_workspace_bytes(int pairs) { return kSkFlagBytes + (size_t)pairs * 256 * 256 * 4; }

// Grouped launch (mtt_gemm_grouped): up to kMaxGroup problems of IDENTICAL geometry (M, N, K, mode, conv shape, nsplit,
// activation, row regrouping) that differ only in their operand / bias / residual / output pointers run as ONE
// persistent launch; tile index = problem * tiles_per_problem + tile. The per-task decoder chains of TaskPrompter
// (5 tasks x {spatial, channel} 1x1 convs, fea_fuse) are single-wave launches (96 tiles on 148 SMs) on their own.
constexpr int kMaxGroup = 32;
struct GroupProblem {
  const float* bias;
  const float* residual;
  float* out_f32;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
};
struct GemmGroup {
  int count;              // 0: not grouped (the kernel uses its four direct tensor-map parameters)
  int tiles_per_problem;
  GroupProblem prob[kMaxGroup];
};
struct GemmGroupMaps {
  CUtensorMap m[kMaxGroup][4];  // A hi, A lo, B hi, B lo per problem
};

// Host: validates the descriptor, fills GemmParams (tiles_n left to the caller) and encodes the four
// tensor maps; the B map's box has `b_box_rows` rows of N.
int gemm_prepare(const mtt_gemm_desc* d, int b_box_rows, GemmParams& p, CUtensorMap maps[4]);

// ---- A tile (128 rows x 64 K) of sub-tile `ms`, k-block kb, filter tap (dy, dx) -----------------
// kMode: 0 = load (1-CTA), 1 = load (CTA pair, signals the leader's barrier), 2 = L2 prefetch only
// (an L2 prefetch running kPrefetchAhead k-blocks ahead of the smem ring was measured and did not help:
// 62.4 -> 63.5 us on the qkv GEMM, so the producers do not issue it)
template <int NSPLIT, int kMode>
__device__ __forceinline__ void load_a_tile(const GemmParams& p, const CUtensorMap* tmA_hi,
                                            const CUtensorMap* tmA_lo, uint8_t* sa, uint64_t* bar, int ms,
                                            int kb, int dy, int dx) {
  auto ld2 = [&](void* dst, const CUtensorMap* m, int c0, int c1) {
    if (kMode == 2) tma_prefetch_2d(m, c0, c1);
    else if (kMode == 1) tma_load_2d_cg2(dst, m, bar, c0, c1);
    else tma_load_2d(dst, m, bar, c0, c1);
  };
  auto ld3 = [&](void* dst, const CUtensorMap* m, int c0, int c1, int c2) {
    if (kMode == 2) tma_prefetch_3d(m, c0, c1, c2);
    else if (kMode == 1) tma_load_3d_cg2(dst, m, bar, c0, c1, c2);
    else tma_load_3d(dst, m, bar, c0, c1, c2);
  };
  auto ld4 = [&](void* dst, const CUtensorMap* m, int c0, int c1, int c2, int c3) {
    if (kMode == 2) tma_prefetch_4d(m, c0, c1, c2, c3);
    else if (kMode == 1) tma_load_4d_cg2(dst, m, bar, c0, c1, c2, c3);
    else tma_load_4d(dst, m, bar, c0, c1, c2, c3);
  };
  if (p.mode == 0 && p.a_groups_per_tile > 0) {
    ld3(sa, tmA_hi, kb * BK, 0, ms * p.a_groups_per_tile);
    if (NSPLIT == 2) ld3(sa + kTileBytes, tmA_lo, kb * BK, 0, ms * p.a_groups_per_tile);
  } else if (p.mode == 0) {
    ld2(sa, tmA_hi, kb * BK, ms * BM);
    if (NSPLIT == 2) ld2(sa + kTileBytes, tmA_lo, kb * BK, ms * BM);
  } else {
    const int per_img = p.tiles_x * p.tiles_y;
    const int cb = ms / per_img;
    const int r = ms - cb * per_img;
    const int cy = (r / p.tiles_x) * p.TH + dy;
    const int cx = (r % p.tiles_x) * p.TW + dx;
    ld4(sa, tmA_hi, kb * BK, cx, cy, cb);
    if (NSPLIT == 2) ld4(sa + kTileBytes, tmA_lo, kb * BK, cx, cy, cb);
  }
}

// ---- where does row `row` of sub-tile `ms` go? --------------------------------------------------
struct RowInfo {
  long long mo;  // output row (after regrouping)
  long long mr;  // residual row
  bool ok;
};
__device__ __forceinline__ RowInfo row_info(const GemmParams& p, int ms, int row) {
  RowInfo r;
  long long m;
  if (p.mode == 0) {
    m = (long long)ms * BM + row;
    r.ok = m < p.M;
    if (p.a_groups_per_tile > 0) r.ok = r.ok && ms < p.tiles_m;
  } else {
    const int per_img = p.tiles_x * p.tiles_y;
    const int cb = ms / per_img;
    const int t = ms - cb * per_img;
    const int y = (t / p.tiles_x) * p.TH + row / p.TW;
    const int x = (t % p.tiles_x) * p.TW + row % p.TW;
    r.ok = (cb < p.NB) && (row < p.TW * p.TH) && (y < p.H) && (x < p.W);
    m = ((long long)cb * p.H + y) * p.W + x;
  }
  r.mo = m;
  if (p.in_group > 0) r.mo = (m / p.in_group) * p.out_group + p.out_offset + (m % p.in_group) * p.out_row_stride;
  r.mr = (p.res_row_mod > 0) ? (m % p.res_row_mod) : r.mo;
  return r;
}

// ---- fused epilogue for 32 consecutive accumulator columns starting at n0 -----------------------
__device__ __forceinline__ void epilogue_store32(const GemmParams& p, const uint32_t (&rr)[32], int n0,
                                                 const RowInfo& ri) {
  if (n0 >= p.N) return;
  if (p.vec32_ok && n0 + 32 <= p.N) {
    // full chunk, 32-byte aligned everywhere: 256-bit loads / stores (one full sector per lane per access)
    float v[32];
    if (p.bias) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const U32x8 b = ld_global_nc_v8(p.bias + n0 + g * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[g * 8 + j] = __uint_as_float(rr[g * 8 + j]) + __uint_as_float(b.v[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
    }
    if (p.act == MTT_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
    } else if (p.act == MTT_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.residual) {
      const float* rp = p.residual + ri.mr * p.ldr + n0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const U32x8 a = ld_global_v8(rp + g * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[g * 8 + j] += __uint_as_float(a.v[j]);
      }
    }
    if (p.out_f32) {
      float* op = p.out_f32 + ri.mo * p.ldo_f32 + n0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        U32x8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = __float_as_uint(v[g * 8 + j]);
        st_global_v8(op + g * 8, o);
      }
    }
    if (p.out_hi) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        U32x8 h, l;
#pragma unroll
        for (int j = 0; j < 8; ++j) split_pack2(v[g * 16 + 2 * j], v[g * 16 + 2 * j + 1], h.v[j], l.v[j]);
        st_global_v8(p.out_hi + ri.mo * p.ldo_bf + n0 + g * 16, h);
        if (p.out_lo) st_global_v8(p.out_lo + ri.mo * p.ldo_bf + n0 + g * 16, l);
      }
    }
    return;
  }
#pragma unroll
  for (int j8 = 0; j8 < 32; j8 += 8) {
    const int n = n0 + j8;
    if (n >= p.N) break;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(rr[j8 + j]);
    if (p.vec_ok && n + 8 <= p.N) {
      if (p.bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (p.act == MTT_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
      } else if (p.act == MTT_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.residual) {
        const float4* rp = reinterpret_cast<const float4*>(p.residual + ri.mr * p.ldr + n);
        const float4 a0 = rp[0], a1 = rp[1];
        v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
        v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
      }
      if (p.out_f32) {
        float4* op = reinterpret_cast<float4*>(p.out_f32 + ri.mo * p.ldo_f32 + n);
        op[0] = make_float4(v[0], v[1], v[2], v[3]);
        op[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
      if (p.out_hi) {
        uint4 h, l;
        split_pack2(v[0], v[1], h.x, l.x);
        split_pack2(v[2], v[3], h.y, l.y);
        split_pack2(v[4], v[5], h.z, l.z);
        split_pack2(v[6], v[7], h.w, l.w);
        *reinterpret_cast<uint4*>(p.out_hi + ri.mo * p.ldo_bf + n) = h;
        if (p.out_lo) *reinterpret_cast<uint4*>(p.out_lo + ri.mo * p.ldo_bf + n) = l;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (n + j >= p.N) break;
        float x = v[j];
        if (p.bias) x += __ldg(p.bias + n + j);
        if (p.act == MTT_ACT_GELU) x = gelu_erf(x);
        else if (p.act == MTT_ACT_RELU) x = fmaxf(x, 0.f);
        if (p.residual) x += p.residual[ri.mr * p.ldr + n + j];
        if (p.out_f32) p.out_f32[ri.mo * p.ldo_f32 + n + j] = x;
        if (p.out_hi) {
          __nv_bfloat16 h, l;
          split_bf16(x, h, l);
          p.out_hi[ri.mo * p.ldo_bf + n + j] = h;
          if (p.out_lo) p.out_lo[ri.mo * p.ldo_bf + n + j] = l;
        }
      }
    }
  }
}

// Entry points of the two kernels (defined in gemm_tc.cu / gemm2_tc.cu)
int launch_gemm_1cta(const mtt_gemm_desc* d, cudaStream_t stream);
int launch_gemm_1cta_grouped(const mtt_gemm_desc* d, int count, cudaStream_t stream);
int launch_gemm_2cta_grouped(const mtt_gemm_desc* d, int count, cudaStream_t stream);
int launch_gemm_2cta(const mtt_gemm_desc* d, int bn2, cudaStream_t stream);
void set_gemm_streamk(int mode);
int streamk_tiles(int tiles, int k_iters, int pairs);
int streamk_schedule_host(int tiles, int k_iters, int pairs, int pair, int* out, int max_pieces);

}  // namespace mtt
