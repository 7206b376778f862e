// Fused multi-head attention, the default kernel of mtt_attention (variant 5; contract in attention_tc.cu:
// TP/models/transformers/taskprompter.py:204-210, prompt-row raw logits :436-437,:482; IP vit.py:189-193).
//
// Warp-specialised, persistent: each CTA (192 threads, 99 KB smem, 256 TMEM columns -> two CTAs per SM) walks
// (batch, head, 128-query tile) items:
//   * warp 5 = TMA (K ring, V ring, the next item's Q), warp 4 = MMA issue (one elected lane), warps 0-3 = softmax,
//     ONE thread per query row: no cross-thread max / sum exchange, no __syncthreads in the loop;
//   * key blocks of 64 with TWO S buffers in TMEM: S_{j+1} is in flight while the softmax warps work on S_j, and
//     O += P_j V_j is issued the moment P_j is published (issue order S_0 S_1 | PV_0 S_2 | PV_1 S_3 ...; the tensor pipe
//     executes in issue order, so S_{j+2} cannot overwrite P_j early);
//   * Q is copied to TMEM once per item, so S = Q K^T is a TS-form MMA that reads only the 64-key K block from shared
//     memory: 33 cycles per 128x64x16 MMA against 49 for the SS form, which is shared-memory bound (scripts/mma_probe.cu);
//   * P = exp2(S c - m) is written back IN PLACE over S as packed bf16 hi / lo; the hi plane is rounded on the integer
//     pipe (add 0x8000, PRMT), only the lo plane goes through F2FP;
//   * the row maximum is OPTIMISTIC: P is computed against the running maximum while the block maximum is tracked
//     alongside, and only when it moved by more than 2^8 (rare after the first block) is the block redone and O / l
//     rescaled (lazy rescaling);
//   * two tcgen05.commit per block: the TMA warp reuses s_ready / pv_done as "slot free" signals (the K slot of block g
//     is free when S_{g-2} is complete, the V slot when PV_{g-2} has retired; two stages = two S buffers);
//   * the item epilogue (O / l, split, store) writes whole 32-byte sectors per lane (st.global.v8).
// Measured and what bounds it: profiles/r1k_attention_analysis.md (a block costs a CTA ~2150 cycles for 768 cycles of
// tensor work; the per-block softmax hand-off, not the tensor pipe, sets the pace). attention3_tc.cu is the predecessor
// (four commits per block, cvt-based hi plane, 16-byte epilogue stores), kept as variant 3.
// TMEM columns: S0 [0,64) | S1 [64,128) | O [128,192) | Q hi [192,224) | Q lo [224,256); P_j overwrites S_j in
// place as packed bf16 (hi in the first 32 columns, lo in the next 32).
#include <math.h>

#include "host_common.h"
#include "ptx.cuh"

namespace mtt {

constexpr int kA5Threads = 192;            // warps 0-3: softmax (thread = query row), 4: MMA issue, 5: TMA
constexpr uint32_t kA5QTile = 128 * 64 * 2;  // one plane of the query tile (16 KB)
constexpr uint32_t kA5KVTile = 64 * 64 * 2;  // one plane of a 64-key K or V block (8 KB)
constexpr int kA5KStages = 2;
constexpr int kA5VStages = 2;
constexpr float kA5LazyLog2 = 8.0f;
static_assert(kA5KStages == 2 && kA5VStages == 2, "the TMA warp reuses s_ready / pv_done (two S buffers) as slot-free signals");

struct Attn5Params {
  int B, N, H, T;
  float scale_log2;  // scale * log2(e)
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  float* prompt_logits;
  int wide_store;       // out planes 32-byte aligned: st.global.v8
  unsigned int* trace;  // TRACE instantiation only: [2 CTAs][2 roles][1024] clock stamps (scripts/attn_trace.py)
};

// One pass over this thread's row of S_j (64 columns in TMEM): tracks the raw block maximum, and -- against the
// scaled running maximum mb -- produces P = exp2(S c - mb) as packed bf16 hi / lo and its row sum.
template <bool FULL, int NSPLIT>
__device__ __forceinline__ float softmax_block5(uint32_t taddr, int kn, float sl2, float mb, float& bmax,
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
      // hi = bf16(p) by round-half-up on the integer pipe (p is finite and >= 0), lo = bf16(p - hi) by F2FP
      const uint32_t u0 = __float_as_uint(p0) + 0x8000u, u1 = __float_as_uint(p1) + 0x8000u;
      ph[c * 16 + (i >> 1)] = __byte_perm(u0, u1, 0x7632);
      if (NSPLIT == 2)
        pl[c * 16 + (i >> 1)] = pack_bf16x2(p0 - __uint_as_float(u0 & 0xFFFF0000u), p1 - __uint_as_float(u1 & 0xFFFF0000u));
    }
  }
  bmax = mx;
  return sum;
}

// The same pass with the instruction stream cut to ~4.5 issue slots per element (the pass above: ~8 plus an F2FP on
// the exp unit's pipe): packed FFMA2 / FADD2, truncation split through PRMT, both 32-column TMEM loads in flight
// before the first use, and NO per-element maximum -- an exponent that ran away from the running maximum shows up
// in the row sum (any p > 2^kA5LazyLog2 makes sum exceed it), which is all the caller needs to trigger its redo path.
// RN_LO: round the lo plane of P to nearest instead of truncating it (two more ALU instructions per pair; the default
// truncates and removes the mean truncation loss in the item epilogue). Measured and dropped: taking a quarter of the
// exponentials from a degree-5 FMA-pipe polynomial instead of MUFU.EX2 (57-59 us against 55: the softmax warps are
// bound by issue slots and dependent-issue latency, not by the exp unit; profiles/r2_attention.md).
template <bool FULL, int NSPLIT, bool RN_LO>
__device__ __forceinline__ float softmax_block6(uint32_t taddr, int kn, float sl2, float mb, uint32_t (&ph)[32],
                                               uint32_t (&pl)[32], float* export_ptr) {
  const int ncols = FULL ? 64 : ((kn + 15) & ~15);
  uint32_t s0[32], s1[32];
  tmem_ld32(taddr, s0);
  if (FULL || ncols > 32) tmem_ld32(taddr + 32, s1);
  tmem_ld_wait();
  if (export_ptr) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (FULL || i < kn) export_ptr[i] = __uint_as_float(s0[i]);
      if (FULL || 32 + i < kn) export_ptr[32 + i] = __uint_as_float(s1[i]);
    }
  }
  const float2 c2 = make_float2(sl2, sl2), m2 = make_float2(-mb, -mb);
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const uint32_t(&s)[32] = c ? s1 : s0;
    if (!FULL && c * 32 >= ncols) {  // columns the (narrowed) MMA never wrote
#pragma unroll
      for (int i = 0; i < 16; ++i) ph[c * 16 + i] = pl[c * 16 + i] = 0u;
      continue;
    }
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float2 t = ffma2(make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])), c2, m2);
      float2 p = make_float2(ex2_approx(t.x), ex2_approx(t.y));
      if (!FULL) {
        if (c * 32 + i >= kn) p.x = 0.f;
        if (c * 32 + i + 1 >= kn) p.y = 0.f;
      }
      acc = fadd2(acc, p);
      if (NSPLIT == 2) {
        if (RN_LO) split_trunc_rn2(p, ph[c * 16 + (i >> 1)], pl[c * 16 + (i >> 1)]);
        else split_trunc2(p, ph[c * 16 + (i >> 1)], pl[c * 16 + (i >> 1)]);
      } else {
        ph[c * 16 + (i >> 1)] = pack_bf16x2(p.x, p.y);
      }
    }
  }
  return acc.x + acc.y;
}

// raw maximum of this thread's row of S_j (first block of an item: the running maximum does not exist yet)
template <bool FULL>
__device__ __forceinline__ float block_max5(uint32_t taddr, int kn) {
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

// MODE 0: the round-1 softmax pass; 1: packed-math pass, truncated P planes with the mean loss folded into the
// normalisation (default); 2: packed-math pass with the lo plane of P rounded to nearest
template <int NSPLIT, bool TRACE, int MODE>
__global__ void __launch_bounds__(kA5Threads, 2)
attention5_kernel(const __grid_constant__ CUtensorMap tmq_hi, const __grid_constant__ CUtensorMap tmq_lo,
                  const __grid_constant__ CUtensorMap tmk_hi, const __grid_constant__ CUtensorMap tmk_lo,
                  const __grid_constant__ CUtensorMap tmo_hi, const __grid_constant__ CUtensorMap tmo_lo,
                  const Attn5Params p) {
  constexpr bool FAST = MODE > 0;
  constexpr bool RN_LO = MODE == 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                  // [NSPLIT][16 KB]
  uint8_t* sK = sQ + NSPLIT * kA5QTile;                // [kA5KStages][NSPLIT][8 KB]
  uint8_t* sV = sK + kA5KStages * NSPLIT * kA5KVTile;  // [kA5VStages][NSPLIT][8 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kA5VStages * NSPLIT * kA5KVTile);
  uint64_t* q_full = bars + 0;    // TMA: next query tile landed in sQ
  uint64_t* q_ready = bars + 1;   // softmax warps: query tile copied to TMEM (sQ free again), count 4
  uint64_t* k_full = bars + 2;    // [kA5KStages]
  uint64_t* k_empty = k_full + kA5KStages;
  uint64_t* v_full = k_empty + kA5KStages;  // [kA5VStages]
  uint64_t* v_empty = v_full + kA5VStages;
  uint64_t* s_ready = v_empty + kA5VStages;  // [2] S_j complete in TMEM
  uint64_t* p_ready = s_ready + 2;           // [2] P_j published by the 4 softmax warps
  uint64_t* pv_done = p_ready + 2;           // [2] O += P_j V_j retired (two barriers: a softmax warp may be two
                                             //     PVs behind, which one parity bit cannot tell apart)
  uint64_t* o_staged = pv_done + 2;          // the item's normalised output tile sits in sQ (MODE > 0), count 4
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_staged + 1);

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
    for (int s = 0; s < kA5KStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < kA5VStages; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_ready[s], 1);
      mbar_init(&p_ready[s], 4);
      mbar_init(&pv_done[s], 1);
    }
    mbar_init(o_staged, 4);
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
    if (TRACE && tr_cta >= 0 && tid == role * 128 && tr_n < 960) p.trace[(tr_cta * 2 + role) * 1024 + tr_n++] = clock32();
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
        mbar_arrive_expect_tx(q_full, NSPLIT * kA5QTile);
        tma_load_3d(sQ, &tmq_hi, q_full, h * 64, qt * 128, b);
        if (NSPLIT == 2) tma_load_3d(sQ + kA5QTile, &tmq_lo, q_full, h * 64, qt * 128, b);
      }
      __syncwarp();
      for (int j = 0; j < nkv; ++j, ++g) {
        const int ks = g % kA5KStages, vs = g % kA5VStages;
        if (g >= 2) mbar_wait(&s_ready[g & 1], ((g - 2) >> 1) & 1);  // S_{g-2} complete: its K slot is free
        if (elect_one()) {
          uint8_t* dk = sK + ks * NSPLIT * kA5KVTile;
          mbar_arrive_expect_tx(&k_full[ks], NSPLIT * kA5KVTile);
          tma_load_3d(dk, &tmk_hi, &k_full[ks], C + h * 64, j * 64, b);
          if (NSPLIT == 2) tma_load_3d(dk + kA5KVTile, &tmk_lo, &k_full[ks], C + h * 64, j * 64, b);
        }
        __syncwarp();
        if (g >= 2) mbar_wait(&pv_done[g & 1], ((g - 2) >> 1) & 1);  // PV_{g-2} retired: its V slot is free
        if (elect_one()) {
          uint8_t* dv = sV + vs * NSPLIT * kA5KVTile;
          mbar_arrive_expect_tx(&v_full[vs], NSPLIT * kA5KVTile);
          tma_load_3d(dv, &tmk_hi, &v_full[vs], 2 * C + h * 64, j * 64, b);
          if (NSPLIT == 2) tma_load_3d(dv + kA5KVTile, &tmk_lo, &v_full[vs], 2 * C + h * 64, j * 64, b);
        }
        __syncwarp();
        // MODE > 0: the softmax warps stage an item's normalised O tile in sQ (free between the copy of the next Q into
        // TMEM and the next Q prefetch) and this warp writes it out with a TMA store, so the global-memory burst of the
        // item boundary (all CTAs finish an item together: 9.5 MB) drains behind the next item's blocks instead of in
        // front of them. The store of item qi - 1 is issued here, at block jq of item qi; only then may the Q tile of
        // item qi + 1 be prefetched into sQ (MODE 0 prefetches at block 0).
        const int jq = (MODE > 0) ? (nkv - 1 < 2 ? nkv - 1 : 2) : 0;
        if (j == jq) {
          if (MODE > 0 && qi >= 1) {
            mbar_wait(o_staged, (qi - 1) & 1);
            if (elect_one()) {
              int qt0, h0, b0;
              item_coords(item - (int)gridDim.x, qt0, h0, b0);
              tma_store_3d(&tmo_hi, sQ, h0 * 64, qt0 * 128, b0);
              if (NSPLIT == 2) tma_store_3d(&tmo_lo, sQ + kA5QTile, h0 * 64, qt0 * 128, b0);
              tma_store_commit();
              tma_store_wait_read();
            }
            __syncwarp();
          }
          if (item + (int)gridDim.x < total) {
            // this item's Q sits in TMEM by now (or soon): sQ can take the next item's query tile
            mbar_wait(q_ready, qi & 1);
            if (elect_one()) {
              int qt2, h2, b2;
              item_coords(item + gridDim.x, qt2, h2, b2);
              mbar_arrive_expect_tx(q_full, NSPLIT * kA5QTile);
              tma_load_3d(sQ, &tmq_hi, q_full, h2 * 64, qt2 * 128, b2);
              if (NSPLIT == 2) tma_load_3d(sQ + kA5QTile, &tmq_lo, q_full, h2 * 64, qt2 * 128, b2);
            }
            __syncwarp();
          }
        }
      }
    }
    if (MODE > 0 && qi >= 1) {   // the last item's output tile
      mbar_wait(o_staged, (qi - 1) & 1);
      if (elect_one()) {
        int qt0, h0, b0;
        item_coords((int)blockIdx.x + (int)(qi - 1) * (int)gridDim.x, qt0, h0, b0);
        tma_store_3d(&tmo_hi, sQ, h0 * 64, qt0 * 128, b0);
        if (NSPLIT == 2) tma_store_3d(&tmo_lo, sQ + kA5QTile, h0 * 64, qt0 * 128, b0);
        tma_store_commit();
        tma_store_wait_all();
      }
      __syncwarp();
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
      const int ks_ = gs % kA5KStages;
      mbar_wait(&k_full[ks_], (gs / kA5KStages) & 1);
      tc_fence_after();
      const int kn = min(64, p.N - (int)js * 64);
      const uint32_t idesc_s = umma_idesc_bf16(128, (kn + 15) & ~15, 0);
      const uint32_t tS = tmem_base + (gs & 1) * 64;
      const uint32_t kh = smem_u32(sK + ks_ * NSPLIT * kA5KVTile);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t kdh = umma_desc_sw128(kh + ks * 32);
          umma_ts(tS, tQ + ks * 8, kdh, idesc_s, ks > 0);
          if (NSPLIT == 2) {
            const uint64_t kdl = umma_desc_sw128(kh + kA5KVTile + ks * 32);
            umma_ts(tS, tQ + ks * 8, kdl, idesc_s, 1);
            umma_ts(tS, tQ + 32 + ks * 8, kdh, idesc_s, 1);
          }
        }
        umma_commit(&s_ready[gs & 1]);
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
      const int vs = gp % kA5VStages;
      mbar_wait(&p_ready[gp & 1], (gp >> 1) & 1);
      stamp(1);
      mbar_wait(&v_full[vs], (gp / kA5VStages) & 1);
      tc_fence_after();
      const int kn = min(64, p.N - (int)jp * 64);
      const int ksteps = (kn + 15) >> 4;
      const uint32_t tP = tmem_base + (gp & 1) * 64;
      const uint32_t vh = smem_u32(sV + vs * NSPLIT * kA5KVTile);
      if (elect_one()) {
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t vdh = umma_desc_sw128(vh + ks * 2048);
          umma_ts(tO, tP + ks * 8, vdh, idesc_o, (jp > 0 || ks > 0) ? 1u : 0u);
          if (NSPLIT == 2) {
            const uint64_t vdl = umma_desc_sw128(vh + kA5KVTile + ks * 2048);
            umma_ts(tO, tP + ks * 8, vdl, idesc_o, 1);
            umma_ts(tO, tP + 32 + ks * 8, vdh, idesc_o, 1);
          }
        }
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
        const uint8_t* base = sQ + pl_ * kA5QTile + row * 128;
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
        if (j == 0) m_run = full ? block_max5<true>(tS, kn) : block_max5<false>(tS, kn);
        bool need;
        if (FAST) {
          sum = full ? softmax_block6<true, NSPLIT, RN_LO>(tS, kn, p.scale_log2, m_run * p.scale_log2, ph, pl, ex)
                     : softmax_block6<false, NSPLIT, RN_LO>(tS, kn, p.scale_log2, m_run * p.scale_log2, ph, pl, ex);
          // some p above 2^kA5LazyLog2 => the sum is above it too (the converse may fire early: a harmless redo);
          // an overflowed (inf) or NaN sum takes the redo path as well
          need = !(sum <= exp2f(kA5LazyLog2));
        } else {
          sum = full ? softmax_block5<true, NSPLIT>(tS, kn, p.scale_log2, m_run * p.scale_log2, bmax, ph, pl, ex)
                     : softmax_block5<false, NSPLIT>(tS, kn, p.scale_log2, m_run * p.scale_log2, bmax, ph, pl, ex);
          need = (bmax - m_run) * p.scale_log2 > kA5LazyLog2;
        }
        if (__any_sync(0xffffffffu, need)) {
          // rare: the block maximum ran away from the running maximum.  O / l are rescaled (O is quiescent: PV_{g-1}
          // has retired and PV_g needs this warp's P_g) and the block is redone against the new maximum.
          mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);  // S_g retired => PV_{g-2} retired: no aliasing
          tc_fence_after();
          if (FAST) {
            bmax = full ? block_max5<true>(tS, kn) : block_max5<false>(tS, kn);
            need = bmax > m_run;  // any upward move is taken now that the block is redone anyway
          }
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
          sum = full ? softmax_block5<true, NSPLIT>(tS, kn, p.scale_log2, m_run * p.scale_log2, bmax, ph, pl, nullptr)
                     : softmax_block5<false, NSPLIT>(tS, kn, p.scale_log2, m_run * p.scale_log2, bmax, ph, pl,
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
      auto stamp_e = [&](int k) {   // TRACE: epilogue timeline of the first 20 items at trace[.. + 960 + 3 * item + k]
        if (TRACE && tr_cta >= 0 && tid == 0 && qi < 20) p.trace[(tr_cta * 2) * 1024 + 960 + 3 * qi + k] = clock32();
      };
      stamp_e(0);
      mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
      tc_fence_after();
      stamp_e(1);
      // MODE 1 truncates both planes of P: their mean loss (kSplitTruncBias per element) is removed here, where the
      // probabilities are normalised by the sum l of the un-truncated p
      const float inv = (MODE == 1 && NSPLIT == 2) ? 1.0f / (l_run * (1.0f - kSplitTruncBias)) : 1.0f / l_run;
      const long long off = ((long long)b * p.N + q_row) * C + h * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(tO + lane_addr + c * 32, o);
        tmem_ld_wait();
        if (MODE > 0) {
          // stage this thread's row of the normalised tile in sQ (this thread copied ITS row of the next Q to TMEM
          // before it got here, nobody else touches the row) in the 128-byte-swizzled layout of the output tensor map:
          // 16-byte chunk k of row r lives at chunk k ^ (r & 7); the TMA warp stores the tile (rows >= N are clipped)
#pragma unroll
          for (int i = 0; i < 32; i += 16) {
            U32x8 hv, lv;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              split_pack2(__uint_as_float(o[i + 2 * k]) * inv, __uint_as_float(o[i + 2 * k + 1]) * inv, hv.v[k], lv.v[k]);
            const int ch = (c * 32 + i) >> 3;   // first of the two 16-byte chunks of these 16 columns
            uint8_t* rowp = sQ + row * 128;
            *reinterpret_cast<uint4*>(rowp + (((ch) ^ (row & 7)) << 4)) = make_uint4(hv.v[0], hv.v[1], hv.v[2], hv.v[3]);
            *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (row & 7)) << 4)) = make_uint4(hv.v[4], hv.v[5], hv.v[6], hv.v[7]);
            if (NSPLIT == 2) {
              *reinterpret_cast<uint4*>(rowp + kA5QTile + (((ch) ^ (row & 7)) << 4)) = make_uint4(lv.v[0], lv.v[1], lv.v[2], lv.v[3]);
              *reinterpret_cast<uint4*>(rowp + kA5QTile + (((ch + 1) ^ (row & 7)) << 4)) =
                  make_uint4(lv.v[4], lv.v[5], lv.v[6], lv.v[7]);
            }
          }
        } else if (q_row < p.N) {
#pragma unroll
          for (int i = 0; i < 32; i += 16) {
            U32x8 hv, lv;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              split_pack2(__uint_as_float(o[i + 2 * k]) * inv, __uint_as_float(o[i + 2 * k + 1]) * inv, hv.v[k], lv.v[k]);
            if (p.wide_store) {  // whole 32-byte sectors per lane
              st_global_v8(p.out_hi + off + c * 32 + i, hv);
              if (NSPLIT == 2) st_global_v8(p.out_lo + off + c * 32 + i, lv);
            } else {
              *reinterpret_cast<uint4*>(p.out_hi + off + c * 32 + i) = make_uint4(hv.v[0], hv.v[1], hv.v[2], hv.v[3]);
              *reinterpret_cast<uint4*>(p.out_hi + off + c * 32 + i + 8) = make_uint4(hv.v[4], hv.v[5], hv.v[6], hv.v[7]);
              if (NSPLIT == 2) {
                *reinterpret_cast<uint4*>(p.out_lo + off + c * 32 + i) = make_uint4(lv.v[0], lv.v[1], lv.v[2], lv.v[3]);
                *reinterpret_cast<uint4*>(p.out_lo + off + c * 32 + i + 8) = make_uint4(lv.v[4], lv.v[5], lv.v[6], lv.v[7]);
              }
            }
          }
        }
      }
      if (MODE > 0) {
        fence_proxy_async();   // the generic-proxy writes above must be visible to the TMA (async proxy) store
        __syncwarp();
        if (elect_one()) mbar_arrive(o_staged);
        __syncwarp();
      }
      tc_fence_before();
      stamp_e(2);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

template <int NSPLIT, bool TRACE, int MODE>
static int launch_attn5(const CUtensorMap* maps, const Attn5Params& p, cudaStream_t stream) {
  constexpr uint32_t smem =
      NSPLIT * kA5QTile + (kA5KStages + kA5VStages) * NSPLIT * kA5KVTile + 1024 + 256;
  static bool attr_set[kMaxDevices] = {};  // the opt-in is per device (and per kernel instantiation)
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    cudaError_t e = cudaFuncSetAttribute(attention5_kernel<NSPLIT, TRACE, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess)
      return set_error(MTT_ERR_LAUNCH, "attention5: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[dev_] = true;
  }
  const int total = ((p.N + 127) / 128) * p.H * p.B;
  const int slots = 2 * sm_count();
  attention5_kernel<NSPLIT, TRACE, MODE><<<total < slots ? total : slots, kA5Threads, smem, stream>>>(
      maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
  return check_launch("mtt_attention(variant 5)");
}

extern unsigned int* g_attn_trace;  // attention_tc.cu (mtt_set_attention_trace)

int launch_attention5(const mtt_attn_desc* d, int mode, cudaStream_t stream) {
  const int C = d->H * 64;
  CUtensorMap maps[6];   // q hi/lo (128-row box), k|v hi/lo (64-row box), out hi/lo (128-row box, TMA store)
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
  {
    const uint64_t odims[3] = {(uint64_t)C, (uint64_t)d->N, (uint64_t)d->B};
    const uint64_t ostr[2] = {(uint64_t)C * 2, (uint64_t)d->N * C * 2};
    if ((rc = make_tmap_bf16(&maps[4], d->out_hi, 3, odims, ostr, qbox))) return rc;
    if (d->nsplit == 2) {
      if ((rc = make_tmap_bf16(&maps[5], d->out_lo, 3, odims, ostr, qbox))) return rc;
    } else {
      maps[5] = maps[4];
    }
  }
  Attn5Params p;
  p.B = d->B;
  p.N = d->N;
  p.H = d->H;
  p.T = d->prompt_logits ? d->T : 0;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.out_hi = static_cast<__nv_bfloat16*>(d->out_hi);
  p.out_lo = static_cast<__nv_bfloat16*>(d->out_lo);
  p.prompt_logits = d->prompt_logits;
  p.wide_store = ((reinterpret_cast<uintptr_t>(d->out_hi) | reinterpret_cast<uintptr_t>(d->out_lo)) & 31) == 0;
  p.trace = g_attn_trace;
  if (g_attn_trace && d->nsplit == 2)
    return mode == 2 ? launch_attn5<2, true, 2>(maps, p, stream)
                     : (mode == 1 ? launch_attn5<2, true, 1>(maps, p, stream) : launch_attn5<2, true, 0>(maps, p, stream));
  if (d->nsplit == 2)
    return mode == 2 ? launch_attn5<2, false, 2>(maps, p, stream)
                     : (mode == 1 ? launch_attn5<2, false, 1>(maps, p, stream) : launch_attn5<2, false, 0>(maps, p, stream));
  return mode == 0 ? launch_attn5<1, false, 0>(maps, p, stream) : launch_attn5<1, false, 1>(maps, p, stream);
}

}  // namespace mtt
