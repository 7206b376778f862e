// Fused multi-head attention, warp-specialised variant (variant 3 of mtt_attention; same contract as
// attention_tc.cu: TP/models/transformers/taskprompter.py:204-210, prompt-row raw logits :436-437,:482).
//
// What the profile of attention_tc.cu said (profiles/README.md): tensor pipe 36 % active, issue slots 40 %, XU 24 %,
// stalls dominated by long-scoreboard waits -- every CTA walks ONE serial chain per key block
//   S = Q K^T  ->  TMEM read  ->  row max  ->  (sync)  ->  exp / split  ->  TMEM write  ->  (sync)  ->  O += P V
// and two co-resident CTAs only hide part of it.  This variant breaks the chain instead:
//   * key blocks of 64 with TWO S buffers in TMEM: the MMA warp keeps S_{j+1} (and S_{j+2}) in flight while the
//     softmax warps work on S_j, and issues O += P_j V_j the moment P_j is published;
//   * Q lives in TMEM (copied once per item), so S = Q K^T is a TS-form MMA that reads only the 64-key K block
//     from shared memory (64 B/clk instead of 192 B/clk for an SS-form N = 64 MMA: the SS form is shared-memory
//     bound: 49 cycles per MMA against 33, scripts/mma_probe.cu -- which is what made an earlier SS-form variant slower);
//   * one thread per query row (4 softmax warps): no cross-thread max / sum exchange, no __syncthreads;
//   * the row maximum is OPTIMISTIC: P is computed against the running maximum while the block maximum is
//     tracked alongside, and only when that maximum moved by more than 2^8 (rare after the first block) is the
//     block redone with the new maximum and O / l rescaled -- same lazy-rescaling arithmetic as variant 1;
//   * dedicated TMA warp (K ring, V ring, next item's Q) and MMA warp; 192 threads, 96 KB smem, 256 TMEM
//     columns: two CTAs per SM.
// TMEM columns: S0 [0,64) | S1 [64,128) | O [128,192) | Q hi [192,224) | Q lo [224,256); P_j overwrites S_j in
// place as packed bf16 (hi in the first 32 columns, lo in the next 32).
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace mtt {

constexpr int kA3Threads = 192;            // warps 0-3: softmax (thread = query row), 4: MMA issue, 5: TMA
constexpr uint32_t kA3QTile = 128 * 64 * 2;  // one plane of the query tile (16 KB)
constexpr uint32_t kA3KVTile = 64 * 64 * 2;  // one plane of a 64-key K or V block (8 KB)
constexpr int kA3KStages = 2;
constexpr int kA3VStages = 2;
constexpr float kA3LazyLog2 = 8.0f;

struct Attn3Params {
  int B, N, H, T;
  float scale_log2;  // scale * log2(e)
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* prompt_logits;
  unsigned int* trace;  // TRACE instantiation only: [2 CTAs][2 roles][1024] clock stamps (scripts/attn_trace.py)
};

// One pass over this thread's row of S_j (64 columns in TMEM): tracks the raw block maximum, and -- against the
// scaled running maximum mb -- produces P = exp2(S c - mb) as packed bf16 hi / lo and its row sum.
template <bool FULL, int NSPLIT>
__device__ __forceinline__ float softmax_block(uint32_t taddr, int kn, float sl2, float mb, float& bmax,
                                               uint32_t (&ph)[32], uint32_t (&pl)[32], float* export_ptr) {
  float sum = 0.f, mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t s[32];
    if (!FULL && c * 32 >= ((kn + 15) & ~15)) {  // columns the (narrowed) MMA never wrote
#pragma unroll
      for (int i = 0; i < 16; ++i) ph[c * 16 + i] = pl[c * 16 + i] = 0u;
      continue;
    }
    tmem_ld32(taddr + c * 32, s);
    tmem_ld_wait();
    if (export_ptr) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (FULL || c * 32 + i < kn) export_ptr[c * 32 + i] = __uint_as_float(s[i]);
    }
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float s0 = __uint_as_float(s[i]), s1 = __uint_as_float(s[i + 1]);
      float p0 = ex2_approx(fmaf(s0, sl2, -mb));
      float p1 = ex2_approx(fmaf(s1, sl2, -mb));
      if (FULL) {
        mx = fmaxf(mx, fmaxf(s0, s1));
      } else {
        if (c * 32 + i < kn) mx = fmaxf(mx, s0); else p0 = 0.f;
        if (c * 32 + i + 1 < kn) mx = fmaxf(mx, s1); else p1 = 0.f;
      }
      sum += p0 + p1;
      uint32_t h, l;
      split_pack2(p0, p1, h, l);
      ph[c * 16 + (i >> 1)] = h;
      if (NSPLIT == 2) pl[c * 16 + (i >> 1)] = l;
    }
  }
  bmax = mx;
  return sum;
}

// raw maximum of this thread's row of S_j (first block of an item: the running maximum does not exist yet)
template <bool FULL>
__device__ __forceinline__ float block_max(uint32_t taddr, int kn) {
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (!FULL && c * 32 >= ((kn + 15) & ~15)) continue;
    uint32_t s[32];
    tmem_ld32(taddr + c * 32, s);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (FULL || c * 32 + i < kn) mx = fmaxf(mx, __uint_as_float(s[i]));
  }
  return mx;
}

template <int NSPLIT, bool TRACE>
__global__ void __launch_bounds__(kA3Threads, 2)
attention3_kernel(const __grid_constant__ CUtensorMap tmq_hi, const __grid_constant__ CUtensorMap tmq_lo,
                  const __grid_constant__ CUtensorMap tmk_hi, const __grid_constant__ CUtensorMap tmk_lo,
                  const Attn3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                  // [NSPLIT][16 KB]
  uint8_t* sK = sQ + NSPLIT * kA3QTile;                // [kA3KStages][NSPLIT][8 KB]
  uint8_t* sV = sK + kA3KStages * NSPLIT * kA3KVTile;  // [kA3VStages][NSPLIT][8 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kA3VStages * NSPLIT * kA3KVTile);
  uint64_t* q_full = bars + 0;    // TMA: next query tile landed in sQ
  uint64_t* q_ready = bars + 1;   // softmax warps: query tile copied to TMEM (sQ free again), count 4
  uint64_t* k_full = bars + 2;    // [kA3KStages]
  uint64_t* k_empty = k_full + kA3KStages;
  uint64_t* v_full = k_empty + kA3KStages;  // [kA3VStages]
  uint64_t* v_empty = v_full + kA3VStages;
  uint64_t* s_ready = v_empty + kA3VStages;  // [2] S_j complete in TMEM
  uint64_t* p_ready = s_ready + 2;           // [2] P_j published by the 4 softmax warps
  uint64_t* pv_done = p_ready + 2;           // [2] O += P_j V_j retired (two barriers: a softmax warp may be two
                                             //     PVs behind, which one parity bit cannot tell apart)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int C = p.H * 64;
  const int nq = (p.N + 127) / 128;  // query tiles per (b, h)
  const int nkv = (p.N + 63) / 64;   // key blocks per (b, h)
  const int total = nq * p.H * p.B;  // work items; persistent: item = blockIdx.x, + gridDim.x, ...

  if (warp == 5 && elect_one()) {
    tma_prefetch_desc(&tmq_hi);
    tma_prefetch_desc(&tmk_hi);
    if (NSPLIT == 2) {
      tma_prefetch_desc(&tmq_lo);
      tma_prefetch_desc(&tmk_lo);
    }
    mbar_init(q_full, 1);
    mbar_init(q_ready, 4);
    for (int s = 0; s < kA3KStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < kA3VStages; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_ready[s], 1);
      mbar_init(&p_ready[s], 4);
      mbar_init(&pv_done[s], 1);
    }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 4) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tO = tmem_base + 128;
  const uint32_t tQ = tmem_base + 192;
  // debug trace: CTAs 0 and gridDim.x / 2 (co-resident on SM 0 when the grid is 2 x #SM), MMA warp and softmax warp 0
  const int tr_cta = (blockIdx.x == 0) ? 0 : ((blockIdx.x == gridDim.x / 2) ? 1 : -1);
  unsigned int tr_n = 1;
  auto stamp = [&](int role) {
    if (TRACE && tr_cta >= 0 && tid == role * 128 && tr_n < 1024) p.trace[(tr_cta * 2 + role) * 1024 + tr_n++] = clock32();
  };
  if (TRACE && tr_cta >= 0 && (tid == 0 || tid == 128)) p.trace[(tr_cta * 2 + (tid == 128)) * 1024] = smid();

  auto item_coords = [&](int item, int& qt, int& h, int& b) {
    qt = item % nq;
    h = (item / nq) % p.H;
    b = item / (nq * p.H);
  };

  if (warp == 5) {
    // ------------------------------------------------------------------ TMA warp (elected lane issues)
    uint32_t g = 0, qi = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x, ++qi) {
      int qt, h, b;
      item_coords(item, qt, h, b);
      if (qi == 0 && elect_one()) {
        mbar_arrive_expect_tx(q_full, NSPLIT * kA3QTile);
        tma_load_3d(sQ, &tmq_hi, q_full, h * 64, qt * 128, b);
        if (NSPLIT == 2) tma_load_3d(sQ + kA3QTile, &tmq_lo, q_full, h * 64, qt * 128, b);
      }
      __syncwarp();
      for (int j = 0; j < nkv; ++j, ++g) {
        const int ks = g % kA3KStages, vs = g % kA3VStages;
        mbar_wait(&k_empty[ks], ((g / kA3KStages) & 1) ^ 1);
        if (elect_one()) {
          uint8_t* dk = sK + ks * NSPLIT * kA3KVTile;
          mbar_arrive_expect_tx(&k_full[ks], NSPLIT * kA3KVTile);
          tma_load_3d(dk, &tmk_hi, &k_full[ks], C + h * 64, j * 64, b);
          if (NSPLIT == 2) tma_load_3d(dk + kA3KVTile, &tmk_lo, &k_full[ks], C + h * 64, j * 64, b);
        }
        __syncwarp();
        mbar_wait(&v_empty[vs], ((g / kA3VStages) & 1) ^ 1);
        if (elect_one()) {
          uint8_t* dv = sV + vs * NSPLIT * kA3KVTile;
          mbar_arrive_expect_tx(&v_full[vs], NSPLIT * kA3KVTile);
          tma_load_3d(dv, &tmk_hi, &v_full[vs], 2 * C + h * 64, j * 64, b);
          if (NSPLIT == 2) tma_load_3d(dv + kA3KVTile, &tmk_lo, &v_full[vs], 2 * C + h * 64, j * 64, b);
        }
        __syncwarp();
        if (j == 0 && item + (int)gridDim.x < total) {
          // this item's Q sits in TMEM by now (or soon): sQ can take the next item's query tile
          mbar_wait(q_ready, qi & 1);
          if (elect_one()) {
            int qt2, h2, b2;
            item_coords(item + gridDim.x, qt2, h2, b2);
            mbar_arrive_expect_tx(q_full, NSPLIT * kA3QTile);
            tma_load_3d(sQ, &tmq_hi, q_full, h2 * 64, qt2 * 128, b2);
            if (NSPLIT == 2) tma_load_3d(sQ + kA3QTile, &tmq_lo, q_full, h2 * 64, qt2 * 128, b2);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA warp (elected lane issues)
    // Issue order: S_0 S_1 | PV_0 S_2 | PV_1 S_3 | ...  (global block index across the CTA's items).  S_{g+2}
    // targets the buffer PV_g reads; the tensor pipe executes in issue order, so it cannot overwrite P_g early.
    const int my_items = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t nblocks = (uint32_t)my_items * (uint32_t)nkv;
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 1);
    uint32_t gs = 0;          // next S to issue
    uint32_t js = 0, qis = 0; // its block index inside the item / item ordinal
    auto issue_s = [&]() {
      if (js == 0) {
        mbar_wait(q_ready, qis & 1);  // Q of this item is in TMEM
        tc_fence_after();
      }
      const int ks_ = gs % kA3KStages;
      mbar_wait(&k_full[ks_], (gs / kA3KStages) & 1);
      tc_fence_after();
      const int kn = min(64, p.N - (int)js * 64);
      const uint32_t idesc_s = umma_idesc_bf16(128, (kn + 15) & ~15, 0);
      const uint32_t tS = tmem_base + (gs & 1) * 64;
      const uint32_t kh = smem_u32(sK + ks_ * NSPLIT * kA3KVTile);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t kdh = umma_desc_sw128(kh + ks * 32);
          umma_ts(tS, tQ + ks * 8, kdh, idesc_s, ks > 0);
          if (NSPLIT == 2) {
            const uint64_t kdl = umma_desc_sw128(kh + kA3KVTile + ks * 32);
            umma_ts(tS, tQ + ks * 8, kdl, idesc_s, 1);
            umma_ts(tS, tQ + 32 + ks * 8, kdh, idesc_s, 1);
          }
        }
        umma_commit(&s_ready[gs & 1]);
        umma_commit(&k_empty[ks_]);
      }
      __syncwarp();
      ++gs;
      if (++js == (uint32_t)nkv) {
        js = 0;
        ++qis;
      }
    };
    if (nblocks > 0) issue_s();
    if (nblocks > 1) issue_s();
    uint32_t jp = 0;
    for (uint32_t gp = 0; gp < nblocks; ++gp) {
      const int vs = gp % kA3VStages;
      mbar_wait(&p_ready[gp & 1], (gp >> 1) & 1);
      stamp(1);
      mbar_wait(&v_full[vs], (gp / kA3VStages) & 1);
      tc_fence_after();
      const int kn = min(64, p.N - (int)jp * 64);
      const int ksteps = (kn + 15) >> 4;
      const uint32_t tP = tmem_base + (gp & 1) * 64;
      const uint32_t vh = smem_u32(sV + vs * NSPLIT * kA3KVTile);
      if (elect_one()) {
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t vdh = umma_desc_sw128(vh + ks * 2048);
          umma_ts(tO, tP + ks * 8, vdh, idesc_o, (jp > 0 || ks > 0) ? 1u : 0u);
          if (NSPLIT == 2) {
            const uint64_t vdl = umma_desc_sw128(vh + kA3KVTile + ks * 2048);
            umma_ts(tO, tP + ks * 8, vdl, idesc_o, 1);
            umma_ts(tO, tP + 32 + ks * 8, vdh, idesc_o, 1);
          }
        }
        umma_commit(&v_empty[vs]);
        umma_commit(&pv_done[gp & 1]);
      }
      __syncwarp();
      stamp(1);
      if (++jp == (uint32_t)nkv) jp = 0;
      if (gs < nblocks) issue_s();
      stamp(1);
    }
  } else {
    // ------------------------------------------------------------------ softmax warps: thread = query row
    const int row = tid;  // 0..127
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    auto copy_q = [&](uint32_t qi) {  // sQ (128-byte swizzled rows) -> TMEM, this thread's row
      mbar_wait(q_full, qi & 1);
#pragma unroll
      for (int pl_ = 0; pl_ < NSPLIT; ++pl_) {
        uint32_t r[32];
        const uint8_t* base = sQ + pl_ * kA3QTile + row * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = *reinterpret_cast<const uint4*>(base + ((c ^ (row & 7)) << 4));
          r[c * 4 + 0] = v.x;
          r[c * 4 + 1] = v.y;
          r[c * 4 + 2] = v.z;
          r[c * 4 + 3] = v.w;
        }
        tmem_st32(tQ + lane_addr + pl_ * 32, r);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (elect_one()) mbar_arrive(q_ready);
      __syncwarp();
    };

    uint32_t g = 0, qi = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x, ++qi) {
      int qt, h, b;
      item_coords(item, qt, h, b);
      const bool has_next = item + (int)gridDim.x < total;
      const int q_row = qt * 128 + row;
      const bool export_row = (p.prompt_logits != nullptr) && (q_row < p.T);
      float* export_base =
          export_row ? p.prompt_logits + (((long long)b * p.H + h) * p.T + q_row) * p.N : nullptr;
      if (qi == 0) copy_q(0);
      float m_run = -INFINITY, l_run = 0.f;

      for (int j = 0; j < nkv; ++j, ++g) {
        const int kn = min(64, p.N - j * 64);
        const bool full = kn == 64;
        const uint32_t tS = tmem_base + (g & 1) * 64 + lane_addr;
        stamp(0);
        mbar_wait(&s_ready[g & 1], (g >> 1) & 1);
        stamp(0);
        tc_fence_after();
        // the item's last S has retired: nothing reads this item's Q any more -> stage the next item's Q now,
        // so the MMA warp can run ahead into the next item while this block's softmax is still in flight
        if (j == nkv - 1 && has_next) copy_q(qi + 1);

        float* ex = export_row ? export_base + j * 64 : nullptr;
        uint32_t ph[32], pl[32];
        float bmax, sum;
        if (j == 0) m_run = full ? block_max<true>(tS, kn) : block_max<false>(tS, kn);
        sum = full ? softmax_block<true, NSPLIT>(tS, kn, p.scale_log2, m_run * p.scale_log2, bmax, ph, pl, ex)
                   : softmax_block<false, NSPLIT>(tS, kn, p.scale_log2, m_run * p.scale_log2, bmax, ph, pl, ex);
        const bool need = (bmax - m_run) * p.scale_log2 > kA3LazyLog2;
        if (__any_sync(0xffffffffu, need)) {
          // rare: the block maximum ran away from the running maximum.  O / l are rescaled (O is quiescent: PV_{g-1}
          // has retired and PV_g needs this warp's P_g) and the block is redone against the new maximum.
          mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);  // S_g retired => PV_{g-2} retired: no aliasing
          tc_fence_after();
          const float alpha = need ? ex2_approx((m_run - bmax) * p.scale_log2) : 1.0f;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + lane_addr + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tO + lane_addr + c * 32, o);
          }
          l_run *= alpha;
          if (need) m_run = bmax;
          sum = full ? softmax_block<true, NSPLIT>(tS, kn, p.scale_log2, m_run * p.scale_log2, bmax, ph, pl, nullptr)
                     : softmax_block<false, NSPLIT>(tS, kn, p.scale_log2, m_run * p.scale_log2, bmax, ph, pl,
                                                    nullptr);
        }
        l_run += sum;
        stamp(0);
        tmem_st32(tS, ph);
        if (NSPLIT == 2) tmem_st32(tS + 32, pl);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (elect_one()) mbar_arrive(&p_ready[g & 1]);
        __syncwarp();
        stamp(0);
      }

      // ---- item epilogue: O / l.  PV_0 of the next item (which overwrites O) needs this warp's next P.
      mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / l_run;
      const long long off = ((long long)b * p.N + q_row) * C + h * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(tO + lane_addr + c * 32, o);
        tmem_ld_wait();
        if (q_row < p.N) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 hv, lv;
            split_pack2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv, hv.x, lv.x);
            split_pack2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv, hv.y, lv.y);
            split_pack2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv, hv.z, lv.z);
            split_pack2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv, hv.w, lv.w);
            *reinterpret_cast<uint4*>(p.out_hi + off + c * 32 + i) = hv;
            if (NSPLIT == 2) *reinterpret_cast<uint4*>(p.out_lo + off + c * 32 + i) = lv;
          }
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

template <int NSPLIT, bool TRACE>
static int launch_attn3(const CUtensorMap* maps, const Attn3Params& p, cudaStream_t stream) {
  constexpr uint32_t smem =
      NSPLIT * kA3QTile + (kA3KStages + kA3VStages) * NSPLIT * kA3KVTile + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention3_kernel<NSPLIT, TRACE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess)
      return set_error(MTT_ERR_LAUNCH, "attention3: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int total = ((p.N + 127) / 128) * p.H * p.B;
  const int slots = 2 * sm_count();
  attention3_kernel<NSPLIT, TRACE><<<total < slots ? total : slots, kA3Threads, smem, stream>>>(maps[0], maps[1], maps[2],
                                                                                       maps[3], p);
  return check_launch("mtt_attention(variant 3)");
}

unsigned int* g_attn_trace = nullptr;  // shared with attention5_tc.cu
}  // namespace mtt
// debug: a device buffer of 4096 uint32 receives clock stamps of the next parity-mode launches (NULL = off)
extern "C" void mtt_set_attention_trace(void* buf) { mtt::g_attn_trace = static_cast<unsigned int*>(buf); }
namespace mtt {

int launch_attention3(const mtt_attn_desc* d, cudaStream_t stream) {
  const int C = d->H * 64;
  CUtensorMap maps[4];
  const uint64_t dims[3] = {(uint64_t)3 * C, (uint64_t)d->N, (uint64_t)d->B};
  const uint64_t str[2] = {(uint64_t)3 * C * 2, (uint64_t)d->N * 3 * C * 2};
  const uint32_t qbox[3] = {64, 128, 1};
  const uint32_t kbox[3] = {64, 64, 1};
  int rc;
  if ((rc = make_tmap_bf16(&maps[0], d->qkv_hi, 3, dims, str, qbox))) return rc;
  if ((rc = make_tmap_bf16(&maps[2], d->qkv_hi, 3, dims, str, kbox))) return rc;
  if (d->nsplit == 2) {
    if ((rc = make_tmap_bf16(&maps[1], d->qkv_lo, 3, dims, str, qbox))) return rc;
    if ((rc = make_tmap_bf16(&maps[3], d->qkv_lo, 3, dims, str, kbox))) return rc;
  } else {
    maps[1] = maps[0];
    maps[3] = maps[2];
  }
  Attn3Params p;
  p.B = d->B;
  p.N = d->N;
  p.H = d->H;
  p.T = d->prompt_logits ? d->T : 0;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.out_hi = static_cast<__nv_bfloat16*>(d->out_hi);
  p.out_lo = static_cast<__nv_bfloat16*>(d->out_lo);
  p.prompt_logits = d->prompt_logits;
  p.trace = g_attn_trace;
  if (g_attn_trace && d->nsplit == 2) return launch_attn3<2, true>(maps, p, stream);
  return d->nsplit == 2 ? launch_attn3<2, false>(maps, p, stream) : launch_attn3<1, false>(maps, p, stream);
}

}  // namespace mtt
