// Fused multi-head attention, software-pipelined variant (variant 2; attention_tc.cu is the default --
// after the issue-path fixes the simpler persistent kernel measured faster: 76 vs 86 us at cfg4).
//
// Same function as attention_tc.cu (TP taskprompter.py:204-210 + prompt-row logit export, IP vit.py:189-193)
// but restructured so that the tensor pipe never waits for the softmax of the SAME CTA:
//   * keys/values stream in blocks of 64, S is double-buffered in TMEM (2 x 64 columns) and S_{j+1} = Q K_{j+1}^T
//     is issued before softmax_j starts, so softmax phases run back to back while QK^T / PV execute;
//   * a dedicated issuer warp owns TMA and tcgen05.mma; 16 softmax warps (FOUR threads per query row, 16 key
//     columns each: the softmax is bound by per-warp instruction latency, ~9 cycles per issued instruction
//     in ncu, so more, shorter warps win) only wait on "S ready" and signal "P ready" through mbarriers;
//   * O accumulates in TMEM with lazy rescaling, P overwrites S in place as packed bf16 hi|lo.
// TMEM: S0 [0,64) S1 [64,128) O [128,192) -> 256 columns; smem 96 KB -> two CTAs per SM.
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace mtt {

constexpr int kA2Parts = 4;                            // threads per query row
constexpr int kA2SoftmaxThreads = 128 * kA2Parts;       // 16 softmax warps
constexpr int kA2Cols = 64 / kA2Parts;                  // S / O columns owned by one thread (16)
constexpr int kA2Threads = kA2SoftmaxThreads + 32;
constexpr int kA2TmemCols = 256;
constexpr uint32_t kQTile = 128 * 64 * 2;  // 16 KB: 128 query rows x 64 bf16
constexpr uint32_t kKVTile = 64 * 64 * 2;  // 8 KB: 64 keys x 64 bf16
constexpr int kKB = 64;                    // keys per block

struct Attn2Params {
  int B, N, H, T;
  float scale_log2;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* prompt_logits;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int NSPLIT>
__global__ void __launch_bounds__(kA2Threads, 2)
attention2_kernel(const __grid_constant__ CUtensorMap tmq_hi, const __grid_constant__ CUtensorMap tmq_lo,
                  const __grid_constant__ CUtensorMap tmkv_hi, const __grid_constant__ CUtensorMap tmkv_lo,
                  const Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                              // [NSPLIT][16 KB]
  uint8_t* sK = sQ + NSPLIT * kQTile;              // [2 stages][NSPLIT][8 KB]
  uint8_t* sV = sK + 2 * NSPLIT * kKVTile;         // [2 stages][NSPLIT][8 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * NSPLIT * kKVTile);
  uint64_t* bar_q = bars;
  uint64_t* bar_k = bars + 1;   // [2]
  uint64_t* bar_v = bars + 3;   // [2]
  uint64_t* bar_s = bars + 5;   // [2] S_j complete            (tcgen05.commit)
  uint64_t* bar_p = bars + 7;   // [2] P_j stored               (8 warp arrivals)
  uint64_t* bar_o = bars + 9;   // [2] PV_j complete            (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  float* xch = reinterpret_cast<float*>(bars + 12);  // [2 buffers][kA2Parts][128 rows]

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int C = p.H * 64;
  const int q0 = qt * 128;
  const int nkv = (p.N + kKB - 1) / kKB;

  if (tid == 0) {
    mbar_init(bar_q, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_k[i], 1);
      mbar_init(&bar_v[i], 1);
      mbar_init(&bar_s[i], 1);
      mbar_init(&bar_p[i], kA2SoftmaxThreads / 32);
      mbar_init(&bar_o[i], 1);
    }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == kA2SoftmaxThreads / 32) tmem_alloc<kA2TmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tO = tmem_base + 128;

  if (warp == kA2SoftmaxThreads / 32) {
    // ============================================================ issuer warp: TMA + tcgen05.mma
    // one ELECTED lane (not `lane == 0`): ptxas then keeps the MMA / TMA operands in uniform registers instead
    // of wrapping every UTCHMMA in an ELECT / R2UR / BRA.U.ANY loop (~100 cycles per instruction)
    if (elect_one()) {
      tma_prefetch_desc(&tmq_hi);
      tma_prefetch_desc(&tmkv_hi);
      auto load_k = [&](int j) {
        uint8_t* dst = sK + (j & 1) * NSPLIT * kKVTile;
        mbar_arrive_expect_tx(&bar_k[j & 1], NSPLIT * kKVTile);
        tma_load_3d(dst, &tmkv_hi, &bar_k[j & 1], C + h * 64, j * kKB, b);
        if (NSPLIT == 2) tma_load_3d(dst + kKVTile, &tmkv_lo, &bar_k[j & 1], C + h * 64, j * kKB, b);
      };
      auto load_v = [&](int j) {
        uint8_t* dst = sV + (j & 1) * NSPLIT * kKVTile;
        mbar_arrive_expect_tx(&bar_v[j & 1], NSPLIT * kKVTile);
        tma_load_3d(dst, &tmkv_hi, &bar_v[j & 1], 2 * C + h * 64, j * kKB, b);
        if (NSPLIT == 2) tma_load_3d(dst + kKVTile, &tmkv_lo, &bar_v[j & 1], 2 * C + h * 64, j * kKB, b);
      };
      mbar_arrive_expect_tx(bar_q, NSPLIT * kQTile);
      tma_load_3d(sQ, &tmq_hi, bar_q, h * 64, q0, b);
      if (NSPLIT == 2) tma_load_3d(sQ + kQTile, &tmq_lo, bar_q, h * 64, q0, b);
      load_k(0);
      load_v(0);
      if (nkv > 1) {
        load_k(1);
        load_v(1);
      }
      mbar_wait(bar_q, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
      const uint32_t qh = smem_u32(sQ);
      for (int i = 0; i <= nkv; ++i) {
        if (i < nkv) {
          // ---- A: S_i = Q K_i^T into S buffer i&1. The tensor pipe runs in issue order, so this cannot
          //         overtake PV_{i-2}, the last reader of that buffer (as P_{i-2}).
          const int kn = min(kKB, p.N - i * kKB);
          const uint32_t idesc_s = umma_idesc_bf16(128, (kn + 15) & ~15, 0);
          mbar_wait(&bar_k[i & 1], (i >> 1) & 1);
          tc_fence_after();
          const uint32_t kh = smem_u32(sK + (i & 1) * NSPLIT * kKVTile);
          const uint32_t tS = tmem_base + (i & 1) * 64;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t qdh = umma_desc_sw128(qh + ks * 32);
            const uint64_t kdh = umma_desc_sw128(kh + ks * 32);
            umma_ss(tS, qdh, kdh, idesc_s, ks > 0);
            if (NSPLIT == 2) {
              const uint64_t qdl = umma_desc_sw128(qh + kQTile + ks * 32);
              const uint64_t kdl = umma_desc_sw128(kh + kKVTile + ks * 32);
              umma_ss(tS, qdh, kdl, idesc_s, 1);
              umma_ss(tS, qdl, kdh, idesc_s, 1);
            }
          }
          umma_commit(&bar_s[i & 1]);
          // ---- D: once S_i has retired its K buffer is free: prefetch K_{i+2}
          if (i + 2 < nkv) {
            mbar_wait(&bar_s[i & 1], (i >> 1) & 1);
            load_k(i + 2);
          }
        }
        if (i >= 1) {
          // ---- C: O += P_{i-1} V_{i-1} as soon as the softmax warps have stored P_{i-1}
          const int j = i - 1;
          const int kn = min(kKB, p.N - j * kKB);
          const int ksteps = ((kn + 15) & ~15) >> 4;
          mbar_wait(&bar_p[j & 1], (j >> 1) & 1);
          mbar_wait(&bar_v[j & 1], (j >> 1) & 1);
          tc_fence_after();
          const uint32_t vh = smem_u32(sV + (j & 1) * NSPLIT * kKVTile);
          const uint32_t tP = tmem_base + (j & 1) * 64;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t vdh = umma_desc_sw128(vh + ks * 2048);
            umma_ts(tO, tP + ks * 8, vdh, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
            if (NSPLIT == 2) {
              const uint64_t vdl = umma_desc_sw128(vh + kKVTile + ks * 2048);
              umma_ts(tO, tP + ks * 8, vdl, idesc_o, 1);
              umma_ts(tO, tP + 32 + ks * 8, vdh, idesc_o, 1);
            }
          }
          umma_commit(&bar_o[j & 1]);
          // ---- E: once PV_{i-1} has retired its V buffer is free: prefetch V_{i+1}
          if (i + 1 < nkv) {
            mbar_wait(&bar_o[j & 1], (j >> 1) & 1);
            load_v(i + 1);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ============================================================ softmax warps (kA2Parts threads per query row)
    const int part = warp >> 2;                 // which 16 of the block's 64 key columns / of O's 64 columns
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    constexpr float kLazyLog2 = 8.0f;           // rescale O only when the running max grows by > 2^8
    float m_run = -INFINITY, l_run = 0.f;
    const int q_row = q0 + row;
    const bool export_row = (p.prompt_logits != nullptr) && (q_row < p.T);
    float* export_ptr =
        export_row ? p.prompt_logits + (((long long)b * p.H + h) * p.T + q_row) * p.N : nullptr;

    for (int j = 0; j < nkv; ++j) {
      const int bb = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int kn = min(kKB, p.N - j * kKB);
      const int c0 = part * kA2Cols;            // first key column of this thread
      const bool mine = c0 < ((kn + 15) & ~15); // the MMA computed at least one of this thread's columns
      const bool full = c0 + kA2Cols <= kn;
      const uint32_t tS = tmem_base + bb * 64;
      mbar_wait(&bar_s[bb], ph);
      tc_fence_after();
      uint32_t s[kA2Cols];
      if (mine) tmem_ld16(tS + lane_addr + c0, s);
      tmem_ld_wait();
      float mx = -INFINITY;
      if (mine) {
#pragma unroll
        for (int i = 0; i < kA2Cols; ++i)
          if (full || c0 + i < kn) mx = fmaxf(mx, __uint_as_float(s[i]));
        if (export_row) {
#pragma unroll
          for (int i = 0; i < kA2Cols; ++i)
            if (c0 + i < kn) export_ptr[j * kKB + c0 + i] = __uint_as_float(s[i]);
        }
      }
      float* x = xch + bb * (kA2Parts * 128);
      x[part * 128 + row] = mx;
      named_bar_sync(1, kA2SoftmaxThreads);     // also orders every thread's S read before any P write
      mx = fmaxf(fmaxf(x[row], x[128 + row]), fmaxf(x[256 + row], x[384 + row]));
      const bool need = (mx - m_run) * p.scale_log2 > kLazyLog2;
      if (j == 0) {
        m_run = mx;
      } else if (__any_sync(0xffffffffu, need)) {
        // PV_{j-1} must have retired before O is touched; PV_j cannot start before P_j below is stored
        mbar_wait(&bar_o[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        const float alpha = need ? ex2_approx((m_run - mx) * p.scale_log2) : 1.0f;
        uint32_t o[kA2Cols];
        tmem_ld16(tO + lane_addr + c0, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < kA2Cols; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st16(tO + lane_addr + c0, o);
        l_run *= alpha;
        if (need) m_run = mx;
      }
      if (mine) {
        const float mb = m_run * p.scale_log2;
        uint32_t ph_[kA2Cols / 2], pl_[kA2Cols / 2];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < kA2Cols; i += 2) {
          float p0 = ex2_approx(fmaf(__uint_as_float(s[i]), p.scale_log2, -mb));
          float p1 = ex2_approx(fmaf(__uint_as_float(s[i + 1]), p.scale_log2, -mb));
          if (!full) {
            if (c0 + i >= kn) p0 = 0.f;
            if (c0 + i + 1 >= kn) p1 = 0.f;
          }
          sum += p0 + p1;
          split_pack2(p0, p1, ph_[i >> 1], pl_[i >> 1]);
        }
        l_run += sum;
        tmem_st8(tS + lane_addr + part * (kA2Cols / 2), ph_);
        if (NSPLIT == 2) tmem_st8(tS + lane_addr + 32 + part * (kA2Cols / 2), pl_);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p[bb]);
    }

    // ---- epilogue: O / l (l = sum of the row's partial sums, all taken against the same running max)
    mbar_wait(&bar_o[(nkv - 1) & 1], ((nkv - 1) >> 1) & 1);
    tc_fence_after();
    uint32_t o[kA2Cols];
    tmem_ld16(tO + lane_addr + part * kA2Cols, o);
    tmem_ld_wait();
    float* x = xch + (nkv & 1) * (kA2Parts * 128);
    x[part * 128 + row] = l_run;
    named_bar_sync(1, kA2SoftmaxThreads);
    if (q_row < p.N) {
      const float inv = 1.0f / ((x[row] + x[128 + row]) + (x[256 + row] + x[384 + row]));
      const long long off = ((long long)b * p.N + q_row) * C + h * 64 + part * kA2Cols;
#pragma unroll
      for (int i = 0; i < kA2Cols; i += 8) {
        uint4 hv, lv;
        split_pack2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv, hv.x, lv.x);
        split_pack2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv, hv.y, lv.y);
        split_pack2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv, hv.z, lv.z);
        split_pack2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv, hv.w, lv.w);
        *reinterpret_cast<uint4*>(p.out_hi + off + i) = hv;
        if (NSPLIT == 2) *reinterpret_cast<uint4*>(p.out_lo + off + i) = lv;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kA2SoftmaxThreads / 32) {
    tc_fence_after();
    tmem_dealloc<kA2TmemCols>(tmem_base);
  }
}

template <int NSPLIT>
static int launch_attn2(const CUtensorMap* maps, const Attn2Params& p, cudaStream_t stream) {
  constexpr uint32_t smem = NSPLIT * kQTile + 4 * NSPLIT * kKVTile + 1024 + 128 + 2 * kA2Parts * 128 * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention2_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem);
    if (e != cudaSuccess)
      return set_error(MTT_ERR_LAUNCH, "attention2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((p.N + 127) / 128, p.H, p.B);
  attention2_kernel<NSPLIT><<<grid, kA2Threads, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], p);
  return check_launch("mtt_attention(v2)");
}

int launch_attention2(const mtt_attn_desc* d, cudaStream_t stream) {
  const int C = d->H * 64;
  CUtensorMap maps[4];
  const uint64_t dims[3] = {(uint64_t)3 * C, (uint64_t)d->N, (uint64_t)d->B};
  const uint64_t str[2] = {(uint64_t)3 * C * 2, (uint64_t)d->N * 3 * C * 2};
  const uint32_t boxq[3] = {64, 128, 1};
  const uint32_t boxkv[3] = {64, kKB, 1};
  int rc;
  if ((rc = make_tmap_bf16(&maps[0], d->qkv_hi, 3, dims, str, boxq))) return rc;
  if ((rc = make_tmap_bf16(&maps[2], d->qkv_hi, 3, dims, str, boxkv))) return rc;
  if (d->nsplit == 2) {
    if ((rc = make_tmap_bf16(&maps[1], d->qkv_lo, 3, dims, str, boxq))) return rc;
    if ((rc = make_tmap_bf16(&maps[3], d->qkv_lo, 3, dims, str, boxkv))) return rc;
  } else {
    maps[1] = maps[0];
    maps[3] = maps[2];
  }
  Attn2Params p;
  p.B = d->B;
  p.N = d->N;
  p.H = d->H;
  p.T = d->prompt_logits ? d->T : 0;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.out_hi = static_cast<__nv_bfloat16*>(d->out_hi);
  p.out_lo = static_cast<__nv_bfloat16*>(d->out_lo);
  p.prompt_logits = d->prompt_logits;
  return d->nsplit == 2 ? launch_attn2<2>(maps, p, stream) : launch_attn2<1>(maps, p, stream);
}

}  // namespace mtt
