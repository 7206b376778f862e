// Fused multi-head attention for the joint [prompts; patches] token sequence (head_dim 64).
//
// Replaces, per (batch, head), the eager sequence of TP/models/transformers/taskprompter.py:204-210
//   raw = q @ k^T ; attn = softmax(raw * scale) ; x = attn @ v
// (and IP/models/transformers/vit.py:189-193) without materialising the [B,H,N,N] maps, and emits
// the only part of `raw` the reference ever consumes: the un-scaled logits of the first T (prompt)
// query rows (taskprompter.py:436-437 spatial gates, :482 cross-task reweighting).
//
// One CTA = one 128-row query tile of one (b, h); keys/values stream in blocks of 128 by TMA.
//   S = Q K^T          tcgen05.mma SS, fp32 accumulator in TMEM (128 columns)
//   P = exp2(S*c - m)  registers; written back IN PLACE over S as packed bf16 hi/lo planes
//   O += P V           tcgen05.mma with A = P from TMEM, B = V from smem (MN-major), TMEM 64 columns
//   O += P V accumulates in TMEM across key blocks; online softmax with lazy rescaling (O is only
//   rescaled when the running max moves by more than 2^8); 256 threads = two per query row, each owning
//   half of the key columns of S/P (read from TMEM once, kept in registers) and half of O's columns
// Split-bf16 operands (NSPLIT = 2): every product is 3 MMAs, as in gemm_tc.cu.
// The CTA is deliberately simple (no warp specialisation): 96 KB smem and 256 TMEM columns let two
// CTAs share an SM, so one CTA's softmax overlaps the other's MMAs.
#include <math.h>
#include <stdlib.h>

#include "host_common.h"
#include "ptx.cuh"

namespace mtt {

constexpr int kAttnThreads = 256;  // two threads per query row: each owns half of the key columns
constexpr int kAttnTmemCols = 256;
constexpr uint32_t kAttnTile = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16

struct AttnParams {
  int B, N, H, T;
  float scale_log2;  // scale * log2(e)
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* prompt_logits;
};

// row maximum of one 32-column chunk (FULL: every column is a valid key)
template <bool FULL>
__device__ __forceinline__ float chunk_max(const uint32_t (&r)[32], int col0, int kn, float mx) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float s = __uint_as_float(r[i]);
    if (FULL || col0 + i < kn) mx = fmaxf(mx, s);
  }
  return mx;
}
// P = exp2(S*c - m*c) for one chunk, packed as bf16 hi / lo pairs; returns the chunk's row sum
template <bool FULL>
__device__ __forceinline__ float chunk_exp_pack(const uint32_t (&r)[32], int col0, int kn, float sl2, float mb,
                                                uint32_t (&ph)[16], uint32_t (&pl)[16]) {
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), sl2, -mb));
    float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), sl2, -mb));
    if (!FULL) {
      if (col0 + i >= kn) p0 = 0.f;
      if (col0 + i + 1 >= kn) p1 = 0.f;
    }
    sum += p0 + p1;
    split_pack2(p0, p1, ph[i >> 1], pl[i >> 1]);
  }
  return sum;
}

template <int NSPLIT>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                 const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                          // [NSPLIT][16 KB]
  uint8_t* sK = sQ + NSPLIT * kAttnTile;
  uint8_t* sV = sK + NSPLIT * kAttnTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NSPLIT * kAttnTile);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* xch = reinterpret_cast<float*>(bars + 6);  // [2][128] row max / row sum exchange between halves

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int half = warp >> 2;              // which 64 key columns / 32 output columns this thread owns
  const int row = (warp & 3) * 32 + (tid & 31);
  const int C = p.H * 64;
  const int nq = (p.N + 127) / 128;        // query tiles per (b, h)
  const int nkv = nq;                      // key blocks per (b, h)
  const int total = nq * p.H * p.B;        // work items; PERSISTENT: item = blockIdx.x, + gridDim.x, ...
  // (measured with %globaltimer: ~7 us of every 28 us non-persistent CTA was prologue / epilogue -- barrier
  //  init, TMEM allocation, first-load latency -- so CTAs now stay resident and walk the item list)

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_hi);
    if (NSPLIT == 2) tma_prefetch_desc(&tm_lo);
    mbar_init(bar_q, 1);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc<kAttnTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;        // S / P: columns [0, 128)
  const uint32_t tO = tmem_base + 128;  // O:      columns [128, 192)
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;

  // ---- issuer-side helpers (executed by ONE elected lane of warp 0; see gemm_tc.cu on why elected) ----
  auto item_coords = [&](int item, int& qt, int& h, int& b) {
    qt = item % nq;
    h = (item / nq) % p.H;
    b = item / (nq * p.H);
  };
  auto load_q = [&](int item) {
    int qt, h, b;
    item_coords(item, qt, h, b);
    mbar_arrive_expect_tx(bar_q, NSPLIT * kAttnTile);
    tma_load_3d(sQ, &tm_hi, bar_q, h * 64, qt * 128, b);
    if (NSPLIT == 2) tma_load_3d(sQ + kAttnTile, &tm_lo, bar_q, h * 64, qt * 128, b);
  };
  auto load_k = [&](int item, int j) {
    int qt, h, b;
    item_coords(item, qt, h, b);
    mbar_arrive_expect_tx(bar_k, NSPLIT * kAttnTile);
    tma_load_3d(sK, &tm_hi, bar_k, C + h * 64, j * 128, b);
    if (NSPLIT == 2) tma_load_3d(sK + kAttnTile, &tm_lo, bar_k, C + h * 64, j * 128, b);
  };
  auto load_v = [&](int item, int j) {
    int qt, h, b;
    item_coords(item, qt, h, b);
    mbar_arrive_expect_tx(bar_v, NSPLIT * kAttnTile);
    tma_load_3d(sV, &tm_hi, bar_v, 2 * C + h * 64, j * 128, b);
    if (NSPLIT == 2) tma_load_3d(sV + kAttnTile, &tm_lo, bar_v, 2 * C + h * 64, j * 128, b);
  };
  // S = Q K_j^T into TMEM; g = running key-block count of this CTA (barrier phase)
  auto issue_s = [&](int j, uint32_t g) {
    mbar_wait(bar_k, g & 1);
    tc_fence_after();
    const int knj = min(128, p.N - j * 128);
    const uint32_t idesc_s = umma_idesc_bf16(128, (knj + 15) & ~15, 0);
    const uint32_t qh = smem_u32(sQ), kh = smem_u32(sK);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t qdh = umma_desc_sw128(qh + ks * 32);
      const uint64_t kdh = umma_desc_sw128(kh + ks * 32);
      umma_ss(tS, qdh, kdh, idesc_s, ks > 0);
      if (NSPLIT == 2) {
        const uint64_t qdl = umma_desc_sw128(qh + kAttnTile + ks * 32);
        const uint64_t kdl = umma_desc_sw128(kh + kAttnTile + ks * 32);
        umma_ss(tS, qdh, kdl, idesc_s, 1);
        umma_ss(tS, qdl, kdh, idesc_s, 1);
      }
    }
    umma_commit(bar_s);
  };

  const int first = blockIdx.x;
  if (first < total && warp == 0 && elect_one()) {
    load_q(first);
    load_k(first, 0);
    load_v(first, 0);
    mbar_wait(bar_q, 0);
    issue_s(0, 0);
  }
  __syncwarp();

  constexpr float kLazyLog2 = 8.0f;  // lazy rescaling threshold (see below)
  uint32_t g = 0;                    // key blocks processed by this CTA: phase of bar_k / bar_v / bar_s / bar_o
  uint32_t qn = 0;                   // items processed by this CTA: phase of bar_q
  for (int item = first; item < total; item += gridDim.x, ++qn) {
    int qt, h, b;
    item_coords(item, qt, h, b);
    const int next = item + gridDim.x;
    const int q_row = qt * 128 + row;
    const bool export_row = (p.prompt_logits != nullptr) && (q_row < p.T);
    float* export_ptr =
        export_row ? p.prompt_logits + (((long long)b * p.H + h) * p.T + q_row) * p.N : nullptr;
    // Online softmax with O accumulated in TMEM across key blocks and LAZY rescaling: m_run only moves (and
    // O / l are only rescaled) when a block's maximum exceeds it by more than 2^kLazyLog2 after scaling.
    float m_run = -INFINITY, l_run = 0.f;

    for (int j = 0; j < nkv; ++j, ++g) {
      const uint32_t ph = g & 1;
      const int kn = min(128, p.N - j * 128);  // valid keys in this block
      const int kn16 = (kn + 15) & ~15;        // MMA N (S) / K extent (PV)
      const bool last = j + 1 == nkv;
      mbar_wait(bar_s, ph);                    // S_j was queued behind the previous PV by the issuer
      tc_fence_after();
      if (warp == 0 && elect_one()) {          // S_j retired: its K block (and, after the last block, Q) is free
        if (!last) {
          load_k(item, j + 1);
        } else if (next < total) {
          load_q(next);
          load_k(next, 0);
        }
      }
      __syncwarp();

      // ---- S (this thread's 64 columns) is read from TMEM ONCE and stays in registers
      const int nchunk = (kn16 + 31) >> 5;
      const bool full = kn == 128;
      const int c0 = half * 2;
      uint32_t s0[32], s1[32];
      const bool have0 = c0 < nchunk, have1 = c0 + 1 < nchunk;
      if (have0) tmem_ld32(tS + lane_addr + c0 * 32, s0);
      if (have1) tmem_ld32(tS + lane_addr + (c0 + 1) * 32, s1);
      tmem_ld_wait();
      float mx = -INFINITY;
      if (have0) mx = full ? chunk_max<true>(s0, c0 * 32, kn, mx) : chunk_max<false>(s0, c0 * 32, kn, mx);
      if (have1) mx = full ? chunk_max<true>(s1, c0 * 32 + 32, kn, mx) : chunk_max<false>(s1, c0 * 32 + 32, kn, mx);
      if (export_row) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (have0 && c0 * 32 + i < kn) export_ptr[j * 128 + c0 * 32 + i] = __uint_as_float(s0[i]);
          if (have1 && c0 * 32 + 32 + i < kn) export_ptr[j * 128 + c0 * 32 + 32 + i] = __uint_as_float(s1[i]);
        }
      }
      xch[half * 128 + row] = mx;
      __syncthreads();
      mx = fmaxf(xch[row], xch[128 + row]);
      // ---- lazy rescale of the TMEM accumulator (both threads of a row take the same decision)
      const bool need = (mx - m_run) * p.scale_log2 > kLazyLog2;
      if (j == 0) {
        m_run = mx;
      } else if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2_approx((m_run - mx) * p.scale_log2) : 1.0f;
        uint32_t o[32];
        tmem_ld32(tO + lane_addr + half * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st32(tO + lane_addr + half * 32, o);
        l_run *= alpha;
        if (need) m_run = mx;
      }
      const float mb = m_run * p.scale_log2;

      // ---- P = exp2(S*c - m*c), written in place as packed bf16 hi | lo (16 + 16 columns per chunk)
      if (have0) {
        uint32_t ph_[16], pl_[16];
        l_run += full ? chunk_exp_pack<true>(s0, c0 * 32, kn, p.scale_log2, mb, ph_, pl_)
                      : chunk_exp_pack<false>(s0, c0 * 32, kn, p.scale_log2, mb, ph_, pl_);
        tmem_st16(tS + lane_addr + c0 * 32, ph_);
        if (NSPLIT == 2) tmem_st16(tS + lane_addr + c0 * 32 + 16, pl_);
      }
      if (have1) {
        uint32_t ph_[16], pl_[16];
        l_run += full ? chunk_exp_pack<true>(s1, c0 * 32 + 32, kn, p.scale_log2, mb, ph_, pl_)
                      : chunk_exp_pack<false>(s1, c0 * 32 + 32, kn, p.scale_log2, mb, ph_, pl_);
        tmem_st16(tS + lane_addr + c0 * 32 + 32, ph_);
        if (NSPLIT == 2) tmem_st16(tS + lane_addr + c0 * 32 + 48, pl_);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncthreads();

      if (warp == 0 && elect_one()) {
        tc_fence_after();
        mbar_wait(bar_v, ph);
        tc_fence_after();
        constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
        const uint32_t vh = smem_u32(sV);
        const int ksteps = kn16 >> 4;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t a_hi = tS + (ks >> 1) * 32 + (ks & 1) * 8;
          const uint64_t vdh = umma_desc_sw128(vh + ks * 2048);
          umma_ts(tO, a_hi, vdh, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          if (NSPLIT == 2) {
            const uint64_t vdl = umma_desc_sw128(vh + kAttnTile + ks * 2048);
            umma_ts(tO, a_hi, vdl, idesc_o, 1);
            umma_ts(tO, a_hi + 16, vdh, idesc_o, 1);
          }
        }
        umma_commit(bar_o);
        // The next S (of this item or of the CTA's next item) is queued right behind PV_j: the tensor pipe
        // executes in issue order, so it cannot overwrite P_j before PV_j has read it.
        if (!last) {
          issue_s(j + 1, g + 1);
          mbar_wait(bar_o, ph);  // PV_j retired: the V buffer is free
          load_v(item, j + 1);
        } else if (next < total) {
          mbar_wait(bar_q, (qn + 1) & 1);  // next item's Q (requested when S_last retired)
          issue_s(0, g + 1);
          mbar_wait(bar_o, ph);
          load_v(next, 0);
        }
      }
      __syncwarp();
    }

    // ---- item epilogue: O / l, where l is the sum of the two halves' partial row sums (same m_run).
    // The next item's PV_0 (which overwrites O) is only issued after every thread has passed this point.
    mbar_wait(bar_o, (g - 1) & 1);
    tc_fence_after();
    uint32_t o[32];
    tmem_ld32(tO + lane_addr + half * 32, o);
    tmem_ld_wait();
    __syncthreads();
    xch[half * 128 + row] = l_run;
    __syncthreads();
    if (q_row < p.N) {
      const float inv = 1.0f / (xch[row] + xch[128 + row]);
      const long long off = ((long long)b * p.N + q_row) * C + h * 64 + half * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 hv, lv;
        split_pack2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv, hv.x, lv.x);
        split_pack2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv, hv.y, lv.y);
        split_pack2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv, hv.z, lv.z);
        split_pack2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv, hv.w, lv.w);
        *reinterpret_cast<uint4*>(p.out_hi + off + i) = hv;
        if (NSPLIT == 2) *reinterpret_cast<uint4*>(p.out_lo + off + i) = lv;
      }
    }
    __syncthreads();  // xch is rewritten by the next item's first block: every thread must have read its row sums
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<kAttnTmemCols>(tmem_base);
  }
}

template <int NSPLIT>
static int launch_attn(const CUtensorMap& mh, const CUtensorMap& ml, const AttnParams& p,
                       cudaStream_t stream) {
  constexpr uint32_t smem = 3 * NSPLIT * kAttnTile + 1024 + 64 + 2 * 128 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<NSPLIT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess)
      return set_error(MTT_ERR_LAUNCH, "attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int total = ((p.N + 127) / 128) * p.H * p.B;
  const int slots = 2 * sm_count();  // two CTAs per SM (96 KB smem, 256 TMEM columns each)
  attention_kernel<NSPLIT><<<total < slots ? total : slots, kAttnThreads, smem, stream>>>(mh, ml, p);
  return check_launch("mtt_attention");
}

}  // namespace mtt

namespace mtt {
int launch_attention3(const mtt_attn_desc* d, cudaStream_t stream);  // attention3_tc.cu
static int g_attn_variant = -1;  // -1: read MTT_ATTN_VARIANT once; 0 = default = 3 (warp-specialised kernel,
                                 // attention3_tc.cu), 1 = the single-role persistent kernel in this file
}  // namespace mtt

extern "C" void mtt_set_attention_variant(int v) { mtt::g_attn_variant = v; }

extern "C" int mtt_attention(const mtt_attn_desc* d, mtt_stream_t stream_) {
  using namespace mtt;
  if (!d) return set_error(MTT_ERR_BAD_SHAPE, "mtt_attention: null descriptor");
  if (d->B <= 0 || d->N <= 0 || d->H <= 0 || d->T < 0 || d->T > d->N)
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_attention: B=%d N=%d H=%d T=%d", d->B, d->N, d->H, d->T);
  if (d->nsplit != 1 && d->nsplit != 2)
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_attention: nsplit=%d", d->nsplit);
  if (!d->qkv_hi || !d->out_hi || (d->nsplit == 2 && (!d->qkv_lo || !d->out_lo)))
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_attention: missing plane");
  if (d->T > 128) return set_error(MTT_ERR_BAD_SHAPE, "mtt_attention: T=%d > 128 unsupported", d->T);
  if ((reinterpret_cast<uintptr_t>(d->out_hi) & 15) || (d->out_lo && (reinterpret_cast<uintptr_t>(d->out_lo) & 15)))
    return set_error(MTT_ERR_MISALIGNED, "mtt_attention: output not 16-byte aligned");
  if (g_attn_variant < 0) {
    const char* e = getenv("MTT_ATTN_VARIANT");
    g_attn_variant = e ? atoi(e) : 0;
  }
  if (g_attn_variant != 1) return launch_attention3(d, static_cast<cudaStream_t>(stream_));
  const int C = d->H * 64;
  CUtensorMap mh, ml;
  const uint64_t dims[3] = {(uint64_t)3 * C, (uint64_t)d->N, (uint64_t)d->B};
  const uint64_t str[2] = {(uint64_t)3 * C * 2, (uint64_t)d->N * 3 * C * 2};
  const uint32_t box[3] = {64, 128, 1};
  int rc;
  if ((rc = make_tmap_bf16(&mh, d->qkv_hi, 3, dims, str, box))) return rc;
  if (d->nsplit == 2) {
    if ((rc = make_tmap_bf16(&ml, d->qkv_lo, 3, dims, str, box))) return rc;
  } else {
    ml = mh;
  }
  AttnParams p;
  p.B = d->B;
  p.N = d->N;
  p.H = d->H;
  p.T = d->prompt_logits ? d->T : 0;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.out_hi = static_cast<__nv_bfloat16*>(d->out_hi);
  p.out_lo = static_cast<__nv_bfloat16*>(d->out_lo);
  p.prompt_logits = d->prompt_logits;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  return d->nsplit == 2 ? launch_attn<2>(mh, ml, p, stream) : launch_attn<1>(mh, ml, p, stream);
}
