// CTA-pair (tcgen05 cta_group::2) variant of the GEMM / implicit-GEMM convolution in gemm_tc.cu.
//
// Two CTAs of one cluster (same TPC) share every MMA: the pair computes a 256 x BN2 output tile,
// CTA r holds rows [128 r, 128 r + 128) of A and rows [BN2/2 r, ...) of the weight tile in its own
// shared memory and owns the 128 x BN2 fp32 accumulator slice in its own TMEM.  Per MMA each SM
// reads only 128 x 16 of A and BN2/2 x 16 of B from shared memory (half the B traffic of the 1-CTA
// kernel), which is what lifts the shared-memory-bandwidth cap of the 128x128 single-CTA tile
// (profiles/r1a_gemm_tc_metrics.csv: tensor pipe 60-67 % active).
//
// Protocol (leader = cluster rank 0):
//   * both CTAs' producer threads TMA their halves into their own smem; all complete_tx land on the
//     LEADER's full barrier (cta_group::2 loads), which the leader's producer arms with 2x the bytes;
//   * the leader's MMA thread issues tcgen05.mma.cta_group::2 and commits with a cluster multicast so
//     the smem-empty and accumulator-full barriers fire in BOTH CTAs;
//   * each CTA's 8 epilogue warps drain their own TMEM slice and arrive on the LEADER's
//     accumulator-empty barrier (16 arrivals).
#include <stdlib.h>

#include "gemm_common.cuh"

namespace mtt {

template <int NSPLIT>
struct Gemm2Cfg {
  static constexpr int kStages = (NSPLIT == 2) ? 3 : 6;
  static constexpr uint32_t kStageBytes = NSPLIT * 2 * kTileBytes;  // A 128x64 + B (<=128)x64 per plane
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

// Ragged last N tile: the pair issues a NARROWER MMA (N rounded up to 16) instead of multiplying zero-filled weight
// rows. With N = n the two CTAs contribute n / 2 weight rows each, so CTA r loads the rows starting at
// tile_start + r * n / 2 (not r * BN2 / 2): accumulator column c is then output column tile_start + c for c < n.
// (350 -> 350 convs on 256-wide tiles: the second tile is 96 wide instead of 256 = 31 % fewer MMAs.)
__device__ __forceinline__ int pair_tile_n(const GemmParams& p, int nt, int bn2) {
  const int left = p.N - nt * bn2;
  return left >= bn2 ? bn2 : ((left + 15) & ~15);
}

// ---- stream-K tail (SK = true, single-problem kernel with 256-wide tiles) ------------------------------------------------
// With T tiles on P pairs the plain persistent schedule runs ceil(T / P) rounds: 204 qkv tiles on 74 pairs are 3 rounds
// for 2.76 rounds of work (fc1: 4 for 3.68, fc2: 1 for 0.92). Here the R = T mod P tiles that would form the ragged last
// round (tile indices [0, R)) are split along K instead: their R * k_iters k-blocks are dealt out evenly, pair p taking
// the contiguous range [p W / P, (p + 1) W / P) of W = R * k_iters, which touches at most two tiles. The remaining
// T - R tiles run one pair per tile as before (tile R + p + j P). Every pair does its stream-K range FIRST.
//   * a range piece that starts inside a tile (k0 > 0) is a CONTRIBUTION: the epilogue warps dump the raw fp32
//     accumulator to this pair's slot of p.sk_part and publish it per warp (release store to p.sk_flags);
//   * the piece that holds a tile's first k-block OWNS the tile: its epilogue warps wait for the warps of the following
//     pairs whose ranges end the tile (a contribution is always a pair's first piece, so it never waits on anything),
//     add their partials in pair order (a fixed order: results are reproducible run to run) and run the fused epilogue.
//     The owner resets each flag it consumed, so the flag array is all-zero again when the launch retires.
// All pairs are co-resident (grid <= number of SM pairs, one CTA per SM), so the owner's spin cannot starve a contributor.
struct SkSched {
  int n_sk;                  // 0..2 stream-K pieces of this pair
  int tile0, k0_0, k1_0;     // first piece (may start inside its tile)
  int k1_1;                  // second piece: tile0 + 1, k-blocks [0, k1_1)
  int dp_first, dp_step;     // whole tiles: dp_first + j * dp_step < num_tiles
  int n_seg;
  long long W;               // total stream-K k-blocks
};
template <bool SK>
__host__ __device__ __forceinline__ SkSched sk_schedule(const GemmParams& p, int pair, int num_pairs, int num_tiles, int k_iters) {
  SkSched s;
  s.n_sk = 0;
  s.tile0 = s.k0_0 = s.k1_0 = s.k1_1 = 0;
  s.W = 0;
  s.dp_first = pair;
  s.dp_step = num_pairs;
  if (SK && p.sk_tiles > 0) {
    s.W = (long long)p.sk_tiles * k_iters;
    const long long b = (long long)pair * s.W / num_pairs, e = (long long)(pair + 1) * s.W / num_pairs;
    if (e > b) {
      const int t0 = (int)(b / k_iters);
      const long long t0_end = (long long)(t0 + 1) * k_iters;
      s.tile0 = t0;
      s.k0_0 = (int)(b - (long long)t0 * k_iters);
      s.k1_0 = (int)((e < t0_end ? e : t0_end) - (long long)t0 * k_iters);
      s.n_sk = 1;
      if (e > t0_end) {
        s.k1_1 = (int)(e - t0_end);
        s.n_sk = 2;
      }
    }
    s.dp_first = p.sk_tiles + pair;
  }
  const int left = num_tiles - s.dp_first;
  s.n_seg = s.n_sk + (left > 0 ? (left + s.dp_step - 1) / s.dp_step : 0);
  return s;
}
// piece `si` of a pair's schedule: tile index and k-block range [k0, k1)
__host__ __device__ __forceinline__ void sk_piece(const SkSched& s, int si, int k_iters, int& tile, int& k0, int& k1) {
  if (si < s.n_sk) {
    tile = s.tile0 + si;
    k0 = si == 0 ? s.k0_0 : 0;
    k1 = si == 0 ? s.k1_0 : s.k1_1;
  } else {
    tile = s.dp_first + (si - s.n_sk) * s.dp_step;
    k0 = 0;
    k1 = k_iters;
  }
}
__device__ __forceinline__ void sk_flag_publish(unsigned int* f) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(1u) : "memory");
}
__device__ __forceinline__ unsigned int sk_flag_peek(const unsigned int* f) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
  return v;
}

// The kernel body, shared by the single-problem kernel (GROUPED = false) and the grouped one (tile -> (problem, tile)).
template <int NSPLIT, int BN2, bool GROUPED, bool SK>
__device__ __forceinline__ void gemm2_tc_body(const CUtensorMap (*maps)[4], const GemmParams& p, const GemmGroup* grp) {
  static_assert(!SK || (BN2 == 256 && !GROUPED), "stream-K: single-problem kernel with 256-wide tiles only");
  using Cfg = Gemm2Cfg<NSPLIT>;
  constexpr int ST = Cfg::kStages;
  constexpr int BNH = BN2 / 2;                       // weight rows held by each CTA
  constexpr uint32_t kBTile = BNH * BK * 2;          // bytes of one B plane per CTA
  constexpr int kTmemCols = 2 * BN2;                 // two accumulator buffers
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + ST * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + ST;
  uint64_t* tfull_bar = empty_bar + ST;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int pairs_m = (p.tiles_m + 1) >> 1;
  const int tpp = pairs_m * p.tiles_n;                        // pair tiles per problem
  const int num_tiles = GROUPED ? tpp * grp->count : tpp;
  const int k_iters = p.taps * p.num_kb;
  const SkSched sched = sk_schedule<SK>(p, pair, num_pairs, num_tiles, k_iters);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps[0][0]);
    tma_prefetch_desc(&maps[0][2]);
    if (NSPLIT == 2) {
      tma_prefetch_desc(&maps[0][1]);
      tma_prefetch_desc(&maps[0][3]);
    }
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * kEpiWarps);
    }
    fence_barrier_init();
  }
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / TMA signal
  if (warp == 1) tmem_alloc_cg2<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    // warp-uniform loop, one ELECTED lane issues (see gemm_tc.cu: `if (lane == 0)` costs ~100 cycles/instr)
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = 2 * NSPLIT * (p.a_box_bytes + kBTile);  // both CTAs' bytes
      for (int si = 0; si < sched.n_seg; ++si) {
        int tile, kbeg, kend;
        sk_piece(sched, si, k_iters, tile, kbeg, kend);
        const int g = GROUPED ? tile / tpp : 0;
        const int tl = GROUPED ? tile - g * tpp : tile;
        const int ms = (tl % pairs_m) * 2 + (int)rank;
        const int nt = tl / pairs_m;
        const int nrow = nt * BN2 + (int)rank * (pair_tile_n(p, nt, BN2) >> 1);
        const CUtensorMap* tmA_hi = &maps[g][0];
        const CUtensorMap* tmA_lo = &maps[g][1];
        const CUtensorMap* tmB_hi = &maps[g][2];
        const CUtensorMap* tmB_lo = &maps[g][3];
        for (int ki = kbeg; ki < kend; ++ki) {
          const int tap = ki / p.num_kb, kb = ki - tap * p.num_kb;
          const int dy = (tap / p.ksize - p.ksize / 2) * p.dil;
          const int dx = (tap % p.ksize - p.ksize / 2) * p.dil;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + NSPLIT * kTileBytes;
          if (elect_one()) {
            if (p.debug & 1) {  // profiling aid: no loads, the MMAs run on whatever is in shared memory
              if (leader) mbar_arrive(&full_bar[stage]);
            } else {
              if (leader) mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
              load_a_tile<NSPLIT, 1>(p, tmA_hi, tmA_lo, sa, &full_bar[stage], ms, kb, dy, dx);
              const int kcoord = tap * p.cin_pad + kb * BK;
              tma_load_2d_cg2(sb, tmB_hi, &full_bar[stage], kcoord, nrow);
              if (NSPLIT == 2) tma_load_2d_cg2(sb + kBTile, tmB_lo, &full_bar[stage], kcoord, nrow);
            }
          }
          __syncwarp();
          if (++stage == ST) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only; elected lane)
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < sched.n_seg; ++it) {
        int tile, kbeg, kend;
        sk_piece(sched, it, k_iters, tile, kbeg, kend);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN2;
        const uint32_t idesc = umma_idesc_bf16(256, pair_tile_n(p, (GROUPED ? tile % tpp : tile) / pairs_m, BN2), 0);
        uint32_t accum = 0;
        int kb = kbeg % p.num_kb;
        for (int ki = kbeg; ki < kend; ++ki) {
          const int nks = (++kb == p.num_kb) ? p.k_last_steps : BK / 16;  // zero-padded tail of K: no MMAs
          if (kb == p.num_kb) kb = 0;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_hi = a_hi + NSPLIT * kTileBytes;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks) {
              if (ks >= nks) break;
              const uint64_t adh = umma_desc_sw128(a_hi + ks * 32);
              const uint64_t bdh = umma_desc_sw128(b_hi + ks * 32);
              umma_ss_cg2(tacc, adh, bdh, idesc, (ks > 0) ? 1u : accum);
              if (NSPLIT == 2) {
                const uint64_t adl = umma_desc_sw128(a_hi + kTileBytes + ks * 32);
                const uint64_t bdl = umma_desc_sw128(b_hi + kBTile + ks * 32);
                umma_ss_cg2(tacc, adh, bdl, idesc, 1);
                umma_ss_cg2(tacc, adl, bdh, idesc, 1);
              }
            }
            umma_commit_cg2(&empty_bar[stage]);  // frees the smem slot in both CTAs
          }
          __syncwarp();
          accum = 1;
          if (++stage == ST) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit_cg2(&tfull_bar[as]);  // accumulator complete, both CTAs
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    constexpr int kColsPerWarp = BN2 / 2;  // 128 or 64 columns per epilogue warp
    for (int it = 0; it < sched.n_seg; ++it) {
      int tile, kbeg, kend;
      sk_piece(sched, it, k_iters, tile, kbeg, kend);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int g = GROUPED ? tile / tpp : 0;
      const int tl = GROUPED ? tile - g * tpp : tile;
      const int ms = (tl % pairs_m) * 2 + (int)rank;
      const int nt = tl / pairs_m;
      const RowInfo ri = row_info(p, ms, row);
      GemmParams pq = p;                      // grouped: this problem's pointers over the shared geometry
      if (GROUPED) {
        const GroupProblem& gp = grp->prob[g];
        pq.bias = gp.bias;
        pq.residual = gp.residual;
        pq.out_f32 = gp.out_f32;
        pq.out_hi = gp.out_hi;
        pq.out_lo = gp.out_lo;
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN2 + half * kColsPerWarp;
      const int nbase = nt * BN2 + half * kColsPerWarp;
      if (kColsPerWarp == 64) {
        uint32_t r0[32], r1[32];
        tmem_ld32(taddr, r0);
        tmem_ld32(taddr + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
        if (ri.ok && !(p.debug & 2)) {
          epilogue_store32(pq, r0, nbase, ri);
          epilogue_store32(pq, r1, nbase + 32, ri);
        }
      } else {
        // stream-K roles of this piece: a contribution starts inside its tile; the owner holds the tile's first
        // k-block and, unless it holds all of them, adds the partials of the pairs that follow it
        const bool sk_contrib = SK && kbeg > 0;
        const bool sk_owner = SK && kbeg == 0 && kend < k_iters;
        const long long tile_kend = (long long)(tile + 1) * k_iters;
        // 128 columns per warp: two 64-column halves to keep the register footprint at 64 accumulators
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r0[32], r1[32];
          tmem_ld32(taddr + hh * 64, r0);
          tmem_ld32(taddr + hh * 64 + 32, r1);
          tmem_ld_wait();
          if (hh == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
          }
          if (SK && sk_contrib) {
            if (ri.ok) {
              float* dst = p.sk_part + ((size_t)(pair * 2 + (int)rank) * BM + row) * 256 + half * kColsPerWarp + hh * 64;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                U32x8 a, b;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  a.v[j] = r0[c * 8 + j];
                  b.v[j] = r1[c * 8 + j];
                }
                st_global_v8(dst + c * 8, a);
                st_global_v8(dst + 32 + c * 8, b);
              }
            }
          } else {
            if (SK && sk_owner) {
              for (int pp = pair + 1; pp < num_pairs && (long long)pp * sched.W / num_pairs < tile_kend; ++pp) {
                unsigned int* flag = p.sk_flags + (pp * 2 + (int)rank) * kEpiWarps + (warp - 2);
                if (hh == 0) {
                  if (lane == 0) {
                    while (sk_flag_peek(flag) == 0) {
                    }
                  }
                  __syncwarp();
                  __threadfence();
                }
                if (ri.ok) {
                  const float4* src = reinterpret_cast<const float4*>(
                      p.sk_part + ((size_t)(pp * 2 + (int)rank) * BM + row) * 256 + half * kColsPerWarp + hh * 64);
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    const float4 a = __ldcg(src + c), b = __ldcg(src + 8 + c);
                    r0[c * 4 + 0] = __float_as_uint(__uint_as_float(r0[c * 4 + 0]) + a.x);
                    r0[c * 4 + 1] = __float_as_uint(__uint_as_float(r0[c * 4 + 1]) + a.y);
                    r0[c * 4 + 2] = __float_as_uint(__uint_as_float(r0[c * 4 + 2]) + a.z);
                    r0[c * 4 + 3] = __float_as_uint(__uint_as_float(r0[c * 4 + 3]) + a.w);
                    r1[c * 4 + 0] = __float_as_uint(__uint_as_float(r1[c * 4 + 0]) + b.x);
                    r1[c * 4 + 1] = __float_as_uint(__uint_as_float(r1[c * 4 + 1]) + b.y);
                    r1[c * 4 + 2] = __float_as_uint(__uint_as_float(r1[c * 4 + 2]) + b.z);
                    r1[c * 4 + 3] = __float_as_uint(__uint_as_float(r1[c * 4 + 3]) + b.w);
                  }
                }
                if (hh == 1) {   // both halves read: hand the flag back as zero for the next launch
                  __syncwarp();
                  if (lane == 0) *reinterpret_cast<volatile unsigned int*>(flag) = 0u;
                }
              }
            }
            if (ri.ok && !(p.debug & 2)) {
              epilogue_store32(pq, r0, nbase + hh * 64, ri);
              epilogue_store32(pq, r1, nbase + hh * 64 + 32, ri);
            }
          }
          __syncwarp();
        }
        if (SK && sk_contrib) {   // this warp's 32 rows x 128 columns of the partial are written: publish them
          __threadfence();
          __syncwarp();
          if (lane == 0) sk_flag_publish(p.sk_flags + (pair * 2 + (int)rank) * kEpiWarps + (warp - 2));
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still touch this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2<kTmemCols>(tmem_base);
  }
}

struct Gemm2Maps1 {
  CUtensorMap m[1][4];
};

template <int NSPLIT, int BN2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_tc_kernel(const __grid_constant__ Gemm2Maps1 maps, const GemmParams p) {
  gemm2_tc_body<NSPLIT, BN2, false, false>(maps.m, p, nullptr);
}

// the same kernel with the stream-K tail (separate instantiation: the plain kernel keeps its register budget)
template <int NSPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_tc_streamk_kernel(const __grid_constant__ Gemm2Maps1 maps, const GemmParams p) {
  gemm2_tc_body<NSPLIT, 256, false, true>(maps.m, p, nullptr);
}

template <int NSPLIT, int BN2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_tc_grouped_kernel(const __grid_constant__ GemmGroupMaps maps, const __grid_constant__ GemmGroup grp,
                        const GemmParams p) {
  gemm2_tc_body<NSPLIT, BN2, true, false>(maps.m, p, &grp);
}

template <int NSPLIT, int BN2>
static int launch_gemm2(const CUtensorMap* maps, const GemmParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<NSPLIT>;
  static bool attr_set[kMaxDevices] = {};  // the opt-in is per device (and per kernel instantiation)
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_tc_kernel<NSPLIT, BN2>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return set_error(MTT_ERR_LAUNCH, "gemm2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[dev_] = true;
  }
  const int pairs_m = (p.tiles_m + 1) / 2;
  const int tiles = pairs_m * p.tiles_n;
  const int max_pairs = sm_count() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  Gemm2Maps1 gm;
  for (int i = 0; i < 4; ++i) gm.m[0][i] = maps[i];
  gemm2_tc_kernel<NSPLIT, BN2><<<pairs * 2, kGemmThreads, Cfg::kSmemBytes, stream>>>(gm, p);
  return check_launch("mtt_gemm(cta pair)");
}

// Stream-K launch: all SM pairs, the ragged last round (p.sk_tiles tiles) split along K (see sk_schedule above).
template <int NSPLIT>
static int launch_gemm2_streamk(const CUtensorMap* maps, const GemmParams& p, int pairs, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<NSPLIT>;
  static bool attr_set[kMaxDevices] = {};
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_tc_streamk_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return set_error(MTT_ERR_LAUNCH, "gemm2(stream-K): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[dev_] = true;
  }
  Gemm2Maps1 gm;
  for (int i = 0; i < 4; ++i) gm.m[0][i] = maps[i];
  gemm2_tc_streamk_kernel<NSPLIT><<<pairs * 2, kGemmThreads, Cfg::kSmemBytes, stream>>>(gm, p);
  return check_launch("mtt_gemm(cta pair, stream-K)");
}

static int g_streamk = -1;  // -1: read MTT_GEMM_STREAMK once. 0 = off, 1 = automatic (default), 2 = whenever legal
void set_gemm_streamk(int mode) { g_streamk = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }

// How many tiles of a `tiles`-tile, k_iters-deep problem go to the stream-K schedule on `pairs` CTA pairs (0 = none).
// Legal: a ragged last round whose split leaves every pair >= 4 k-blocks (the owner's wait assumes no pair is empty).
// Automatic: only a SINGLE partial round (tiles < pairs) that leaves >= 1/2 of the pairs idle and is >= 32 k-blocks
// deep. Measured per shape (profiles/r3_streamk.md, split / plain launch time): fc2 at batch 1 (20 tiles x 64 k-blocks)
// 0.91, dW of proj (16 x 65) 0.91 -- but dW of qkv (48 x 65) 1.05, fc2 at batch 4 (68 x 64) 1.21, qkv (204 x 16) 1.17:
// with all 148 SMs streaming operands the k-block rate drops (the pair kernel already sits at the L2 -> SM operand
// rate) and every split tile adds a 256 KB partial written and read back plus two epilogue passes; cfg4 bs 4 forward
// 328 -> 308 images/s with the split forced on, bs 1 forward 159.4 -> 161.9 images/s with this policy.
int streamk_tiles(int tiles, int k_iters, int pairs) {
  if (g_streamk < 0) {
    const char* e = getenv("MTT_GEMM_STREAMK");
    set_gemm_streamk(e ? atoi(e) : 1);
  }
  if (!g_streamk || pairs < 2 || pairs * 2 * kEpiWarps * 4 > (int)kSkFlagBytes) return 0;
  const int r = tiles % pairs;
  if (r == 0 || (pairs - r) * 16 < pairs || (long long)r * k_iters / pairs < 4) return 0;
  if (g_streamk == 1 && (tiles >= pairs || (pairs - r) * 2 < pairs || k_iters < 32)) return 0;
  return r;
}

// Test hook (mtt_debug_streamk_schedule): the pieces pair `pair` of `pairs` runs, from the same code the kernel uses.
int streamk_schedule_host(int tiles, int k_iters, int pairs, int pair, int* out, int max_pieces) {
  GemmParams p = {};
  p.sk_tiles = streamk_tiles(tiles, k_iters, pairs);
  const int num_pairs = p.sk_tiles > 0 ? pairs : (tiles < pairs ? tiles : pairs);
  if (pair >= num_pairs) return 0;
  const SkSched s = sk_schedule<true>(p, pair, num_pairs, tiles, k_iters);
  int n = 0;
  for (int si = 0; si < s.n_seg && n < max_pieces; ++si, ++n) sk_piece(s, si, k_iters, out[3 * n], out[3 * n + 1], out[3 * n + 2]);
  return s.n_seg;
}

template <int NSPLIT>
static int launch_gemm2_grouped(const GemmGroupMaps& gm, const GemmGroup& grp, const GemmParams& p, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<NSPLIT>;
  static bool attr_set[kMaxDevices] = {};
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_tc_grouped_kernel<NSPLIT, 256>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return set_error(MTT_ERR_LAUNCH, "gemm2(grouped): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[dev_] = true;
  }
  const int tiles = grp.tiles_per_problem * grp.count;
  const int max_pairs = sm_count() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  gemm2_tc_grouped_kernel<NSPLIT, 256><<<pairs * 2, kGemmThreads, Cfg::kSmemBytes, stream>>>(gm, grp, p);
  return check_launch("mtt_gemm_grouped(cta pair)");
}

// count problems of identical geometry on the CTA-pair kernel (256 x 256 tiles, ragged last N tile narrowed)
int launch_gemm_2cta_grouped(const mtt_gemm_desc* d, int count, cudaStream_t stream) {
  GemmParams p;
  GemmGroupMaps gm;
  GemmGroup grp;
  grp.count = count;
  for (int g = 0; g < count; ++g) {
    GemmParams pg;
    int rc = gemm_prepare(&d[g], 128, pg, gm.m[g]);
    if (rc) return rc;
    pg.tiles_n = (d[g].N + 255) / 256;
    if (g == 0) {
      p = pg;
    } else {
      p.vec_ok = p.vec_ok && pg.vec_ok;
      p.vec32_ok = p.vec32_ok && pg.vec32_ok;
    }
    grp.prob[g] = GroupProblem{pg.bias, pg.residual, pg.out_f32, pg.out_hi, pg.out_lo};
  }
  grp.tiles_per_problem = ((p.tiles_m + 1) / 2) * p.tiles_n;
  return d[0].nsplit == 2 ? launch_gemm2_grouped<2>(gm, grp, p, stream) : launch_gemm2_grouped<1>(gm, grp, p, stream);
}

int launch_gemm_2cta(const mtt_gemm_desc* d, int bn2, cudaStream_t stream) {
  GemmParams p;
  CUtensorMap maps[4];
  int rc = gemm_prepare(d, bn2 / 2, p, maps);
  if (rc) return rc;
  p.tiles_n = (d->N + bn2 - 1) / bn2;
  if (bn2 == 256 && d->sk_ws) {
    // the caller lent a stream-K workspace (mtt_gemm_streamk_bytes): split the ragged last round of tiles along K
    const int pairs = sm_count() / 2;
    const int tiles = ((p.tiles_m + 1) / 2) * p.tiles_n;
    const int r = streamk_tiles(tiles, p.taps * p.num_kb, pairs);
    if (r > 0) {
      if ((size_t)d->sk_ws_bytes < sk_workspace_bytes(pairs) || (reinterpret_cast<uintptr_t>(d->sk_ws) & 255))
        return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm: sk_ws of %lld bytes (256-byte aligned) < required %zu "
                         "(mtt_gemm_streamk_bytes)", (long long)d->sk_ws_bytes, sk_workspace_bytes(pairs));
      p.sk_tiles = r;
      p.sk_flags = static_cast<unsigned int*>(d->sk_ws);
      p.sk_part = reinterpret_cast<float*>(static_cast<uint8_t*>(d->sk_ws) + kSkFlagBytes);
      return d->nsplit == 2 ? launch_gemm2_streamk<2>(maps, p, pairs, stream) : launch_gemm2_streamk<1>(maps, p, pairs, stream);
    }
  }
  if (bn2 == 256)
    return d->nsplit == 2 ? launch_gemm2<2, 256>(maps, p, stream) : launch_gemm2<1, 256>(maps, p, stream);
  return d->nsplit == 2 ? launch_gemm2<2, 128>(maps, p, stream) : launch_gemm2<1, 128>(maps, p, stream);
}

}  // namespace mtt
