// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   D[m, n] = act( sum_k A[m, k] * Bw[n, k] + bias[n] ) + residual[m, n]
//
// Operands are split-bf16 planes (hi, lo).  With NSPLIT = 2 every 128x128x16 product is issued as
// three tcgen05.mma (hi*hi + hi*lo + lo*hi) accumulating in fp32 TMEM, which carries ~16 mantissa
// bits per operand -- the reference computes these contractions in fp32 (cuBLAS / cuDNN eager,
// TP/models/transformers/taskprompter.py:201,212,274,362,691) and single-pass bf16/tf32 does not
// meet its 1e-3 parity bar.  NSPLIT = 1 is plain bf16.
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..9 = epilogue (TMEM -> registers -> bias/act/residual -> global).  Three barrier rings:
// smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, two accumulator buffers so the
// epilogue of tile i overlaps the main loop of tile i+1), and a static persistent tile schedule.
//
// Convolution (mode 1) is the same kernel: the A tile of 128 output pixels is a TH x TW patch of one
// NHWC image and each filter tap is a shifted rank-4 TMA box; TMA's out-of-bounds zero fill is the
// zero padding.  The K loop runs over taps x channel blocks.
#include "gemm_common.cuh"

namespace mtt {

constexpr int BN = 128;
constexpr int kTmemCols = 2 * BN;

template <int NSPLIT>
struct GemmCfg {
  static constexpr int kStages = (NSPLIT == 2) ? 3 : 6;
  static constexpr uint32_t kStageBytes = NSPLIT * 2 * kTileBytes;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

// The kernel body, shared by the single-problem kernel (GROUPED = false: `maps` holds one set of four tensor maps)
// and the grouped one (GROUPED = true: one set per problem, tile -> (problem, tile) through grp).
template <int NSPLIT, bool GROUPED>
__device__ __forceinline__ void gemm_tc_body(const CUtensorMap (*maps)[4], const GemmParams& p, const GemmGroup* grp) {
  using Cfg = GemmCfg<NSPLIT>;
  constexpr int ST = Cfg::kStages;
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
  const int tpp = p.tiles_m * p.tiles_n;                       // tiles per problem
  const int num_tiles = GROUPED ? tpp * grp->count : tpp;
  const int k_iters = p.taps * p.num_kb;

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
      mbar_init(&tempty_bar[s], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp runs the (warp-uniform) loop and ONE ELECTED lane issues: with `if (lane == 0)` ptxas
    // cannot prove the operands uniform and wraps every UTMALDG / UTCHMMA in an ELECT / R2UR / BRA.U.ANY
    // uniformisation loop (~100 cycles per instruction, measured with clock64 in the attention kernel).
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = NSPLIT * (p.a_box_bytes + kTileBytes);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int g = GROUPED ? tile / tpp : 0;
        const int tl = GROUPED ? tile - g * tpp : tile;
        const int mt = tl % p.tiles_m;
        const int nt = tl / p.tiles_m;
        const CUtensorMap* tmA_hi = &maps[g][0];
        const CUtensorMap* tmA_lo = &maps[g][1];
        const CUtensorMap* tmB_hi = &maps[g][2];
        const CUtensorMap* tmB_lo = &maps[g][3];
        for (int ki = 0; ki < k_iters; ++ki) {
          const int tap = ki / p.num_kb, kb = ki - tap * p.num_kb;
          const int dy = (tap / p.ksize - p.ksize / 2) * p.dil;
          const int dx = (tap % p.ksize - p.ksize / 2) * p.dil;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + NSPLIT * kTileBytes;
          if (elect_one()) {
            if (p.debug & 1) {  // profiling aid: no loads
              mbar_arrive(&full_bar[stage]);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
              load_a_tile<NSPLIT, 0>(p, tmA_hi, tmA_lo, sa, &full_bar[stage], mt, kb, dy, dx);
              const int kcoord = tap * p.cin_pad + kb * BK;
              tma_load_2d(sb, tmB_hi, &full_bar[stage], kcoord, nt * BN);
              if (NSPLIT == 2) tma_load_2d(sb + kTileBytes, tmB_lo, &full_bar[stage], kcoord, nt * BN);
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
    // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        // ragged last N tile: a narrower MMA (N rounded up to 16) instead of multiplying zero-filled weight rows
        const int n_left = p.N - (((GROUPED ? tile % tpp : tile)) / p.tiles_m) * BN;
        const uint32_t idesc = umma_idesc_bf16(BM, n_left >= BN ? BN : ((n_left + 15) & ~15), 0);
        uint32_t accum = 0;
        int kb = 0;
        for (int ki = 0; ki < k_iters; ++ki) {
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
              umma_ss(tacc, adh, bdh, idesc, (ks > 0) ? 1u : accum);
              if (NSPLIT == 2) {
                const uint64_t adl = umma_desc_sw128(a_hi + kTileBytes + ks * 32);
                const uint64_t bdl = umma_desc_sw128(b_hi + kTileBytes + ks * 32);
                umma_ss(tacc, adh, bdl, idesc, 1);
                umma_ss(tacc, adl, bdh, idesc, 1);
              }
            }
            umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          }
          __syncwarp();
          accum = 1;
          if (++stage == ST) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tfull_bar[as]);  // accumulator complete
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;    // which pair of 32-column chunks
    const int row = q * 32 + lane;       // row inside the tile
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int g = GROUPED ? tile / tpp : 0;
      const int tl = GROUPED ? tile - g * tpp : tile;
      const int mt = tl % p.tiles_m;
      const int nt = tl / p.tiles_m;
      const RowInfo ri = row_info(p, mt, row);

      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + half * 64;
      tmem_ld32(taddr, r0);
      tmem_ld32(taddr + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);  // accumulator buffer may be overwritten

      if (ri.ok && !(p.debug & 2)) {
        const int n0 = nt * BN + half * 64;
        if (GROUPED) {
          GemmParams q = p;                       // this problem's pointers over the shared geometry
          const GroupProblem& gp = grp->prob[g];
          q.bias = gp.bias;
          q.residual = gp.residual;
          q.out_f32 = gp.out_f32;
          q.out_hi = gp.out_hi;
          q.out_lo = gp.out_lo;
          epilogue_store32(q, r0, n0, ri);
          epilogue_store32(q, r1, n0 + 32, ri);
        } else {
          epilogue_store32(p, r0, n0, ri);
          epilogue_store32(p, r1, n0 + 32, ri);
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

struct GemmMaps1 {
  CUtensorMap m[1][4];
};

template <int NSPLIT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ GemmMaps1 maps, const GemmParams p) {
  gemm_tc_body<NSPLIT, false>(maps.m, p, nullptr);
}

template <int NSPLIT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_grouped_kernel(const __grid_constant__ GemmGroupMaps maps, const __grid_constant__ GemmGroup grp,
                       const GemmParams p) {
  gemm_tc_body<NSPLIT, true>(maps.m, p, &grp);
}

template <int NSPLIT>
static int launch_gemm(const CUtensorMap* maps, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<NSPLIT>;
  static bool attr_set[kMaxDevices] = {};  // the opt-in is per device (and per kernel instantiation)
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<NSPLIT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return set_error(MTT_ERR_LAUNCH, "gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[dev_] = true;
  }
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  GemmMaps1 gm;
  for (int i = 0; i < 4; ++i) gm.m[0][i] = maps[i];
  gemm_tc_kernel<NSPLIT><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(gm, p);
  return check_launch("mtt_gemm");
}

template <int NSPLIT>
static int launch_gemm_grouped(const GemmGroupMaps& gm, const GemmGroup& grp, const GemmParams& p,
                               cudaStream_t stream) {
  using Cfg = GemmCfg<NSPLIT>;
  static bool attr_set[kMaxDevices] = {};
  const int dev_ = current_device();
  if (!attr_set[dev_]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_grouped_kernel<NSPLIT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return set_error(MTT_ERR_LAUNCH, "gemm(grouped): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[dev_] = true;
  }
  const int tiles = grp.tiles_per_problem * grp.count;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_tc_grouped_kernel<NSPLIT><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(gm, grp, p);
  return check_launch("mtt_gemm_grouped");
}

int launch_gemm_1cta_grouped(const mtt_gemm_desc* d, int count, cudaStream_t stream) {
  GemmParams p;
  GemmGroupMaps gm;
  GemmGroup grp;
  grp.count = count;
  for (int g = 0; g < count; ++g) {
    GemmParams pg;
    int rc = gemm_prepare(&d[g], BN, pg, gm.m[g]);
    if (rc) return rc;
    pg.tiles_n = (d[g].N + BN - 1) / BN;
    if (g == 0) {
      p = pg;
    } else if (pg.vec_ok != p.vec_ok || pg.vec32_ok != p.vec32_ok) {  // the epilogue's vector width is shared
      p.vec_ok = p.vec_ok && pg.vec_ok;
      p.vec32_ok = p.vec32_ok && pg.vec32_ok;
    }
    grp.prob[g] = GroupProblem{pg.bias, pg.residual, pg.out_f32, pg.out_hi, pg.out_lo};
  }
  grp.tiles_per_problem = p.tiles_m * p.tiles_n;
  return d[0].nsplit == 2 ? launch_gemm_grouped<2>(gm, grp, p, stream) : launch_gemm_grouped<1>(gm, grp, p, stream);
}


int launch_gemm_1cta(const mtt_gemm_desc* d, cudaStream_t stream) {
  GemmParams p;
  CUtensorMap maps[4];
  int rc = gemm_prepare(d, BN, p, maps);
  if (rc) return rc;
  p.tiles_n = (d->N + BN - 1) / BN;
  return d->nsplit == 2 ? launch_gemm<2>(maps, p, stream) : launch_gemm<1>(maps, p, stream);
}

}  // namespace mtt
