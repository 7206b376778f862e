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
#include "host_common.h"
#include "ptx.cuh"

namespace mtt {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;
constexpr uint32_t kTileBytes = BM * BK * 2;  // one bf16 operand tile: 16 KB
constexpr int kTmemCols = 2 * BN;

struct GemmParams {
  int M, N;
  int num_kb, taps, ksize, dil, mode;
  int H, W, TW, TH, tiles_x, tiles_y;
  int tiles_m, tiles_n;
  int cin_pad;
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
  int vec_ok;
  int a_groups_per_tile;  // >0: A rows are gathered in groups through a rank-3 tensor map
};

template <int NSPLIT>
struct GemmCfg {
  static constexpr int kStages = (NSPLIT == 2) ? 3 : 6;
  static constexpr uint32_t kStageBytes = NSPLIT * 2 * kTileBytes;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int NSPLIT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const GemmParams p) {
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
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int k_iters = p.taps * p.num_kb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    if (NSPLIT == 2) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmB_lo);
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
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = NSPLIT * (p.a_box_bytes + kTileBytes);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile % p.tiles_m;
        const int nt = tile / p.tiles_m;
        int cb = 0, cy0 = 0, cx0 = 0;
        if (p.mode == 1) {
          const int per_img = p.tiles_x * p.tiles_y;
          cb = mt / per_img;
          const int r = mt - cb * per_img;
          cy0 = (r / p.tiles_x) * p.TH;
          cx0 = (r % p.tiles_x) * p.TW;
        }
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = (tap / p.ksize - p.ksize / 2) * p.dil;
          const int dx = (tap % p.ksize - p.ksize / 2) * p.dil;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + NSPLIT * kTileBytes;
            mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
            if (p.mode == 0 && p.a_groups_per_tile > 0) {
              tma_load_3d(sa, &tmA_hi, &full_bar[stage], kb * BK, 0, mt * p.a_groups_per_tile);
              if (NSPLIT == 2)
                tma_load_3d(sa + kTileBytes, &tmA_lo, &full_bar[stage], kb * BK, 0, mt * p.a_groups_per_tile);
            } else if (p.mode == 0) {
              tma_load_2d(sa, &tmA_hi, &full_bar[stage], kb * BK, mt * BM);
              if (NSPLIT == 2) tma_load_2d(sa + kTileBytes, &tmA_lo, &full_bar[stage], kb * BK, mt * BM);
            } else {
              tma_load_4d(sa, &tmA_hi, &full_bar[stage], kb * BK, cx0 + dx, cy0 + dy, cb);
              if (NSPLIT == 2)
                tma_load_4d(sa + kTileBytes, &tmA_lo, &full_bar[stage], kb * BK, cx0 + dx, cy0 + dy, cb);
            }
            const int kcoord = tap * p.cin_pad + kb * BK;
            tma_load_2d(sb, &tmB_hi, &full_bar[stage], kcoord, nt * BN);
            if (NSPLIT == 2) tma_load_2d(sb + kTileBytes, &tmB_lo, &full_bar[stage], kcoord, nt * BN);
            if (++stage == ST) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        uint32_t accum = 0;
        for (int ki = 0; ki < k_iters; ++ki) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_hi = a_hi + NSPLIT * kTileBytes;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t adh = umma_desc_sw128(a_hi + ks * 32);
            const uint64_t bdh = umma_desc_sw128(b_hi + ks * 32);
            umma_ss(tacc, adh, bdh, idesc, accum);
            accum = 1;
            if (NSPLIT == 2) {
              const uint64_t adl = umma_desc_sw128(a_hi + kTileBytes + ks * 32);
              const uint64_t bdl = umma_desc_sw128(b_hi + kTileBytes + ks * 32);
              umma_ss(tacc, adh, bdl, idesc, 1);
              umma_ss(tacc, adl, bdh, idesc, 1);
            }
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == ST) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete
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
      const int mt = tile % p.tiles_m;
      const int nt = tile / p.tiles_m;
      // ---- where does this thread's row go?
      long long m;  // logical output row (before regrouping)
      bool row_ok;
      if (p.mode == 0) {
        m = (long long)mt * BM + row;
        row_ok = m < p.M;
      } else {
        const int per_img = p.tiles_x * p.tiles_y;
        const int cb = mt / per_img;
        const int r = mt - cb * per_img;
        const int y = (r / p.tiles_x) * p.TH + row / p.TW;
        const int x = (r % p.tiles_x) * p.TW + row % p.TW;
        row_ok = (row < p.TW * p.TH) && (y < p.H) && (x < p.W);
        m = ((long long)cb * p.H + y) * p.W + x;
      }
      long long mo = m;
      if (p.in_group > 0) mo = (m / p.in_group) * p.out_group + p.out_offset + (m % p.in_group);
      const long long mr = (p.res_row_mod > 0) ? (m % p.res_row_mod) : mo;

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

      if (row_ok) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t* rr = c ? r1 : r0;
          const int n0 = nt * BN + half * 64 + c * 32;
          if (n0 >= p.N) break;
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
                const float4* rp = reinterpret_cast<const float4*>(p.residual + mr * p.ldr + n);
                const float4 a0 = rp[0], a1 = rp[1];
                v[0] += a0.x; v[1] += a0.y; v[2] += a0.z; v[3] += a0.w;
                v[4] += a1.x; v[5] += a1.y; v[6] += a1.z; v[7] += a1.w;
              }
              if (p.out_f32) {
                float4* op = reinterpret_cast<float4*>(p.out_f32 + mo * p.ldo_f32 + n);
                op[0] = make_float4(v[0], v[1], v[2], v[3]);
                op[1] = make_float4(v[4], v[5], v[6], v[7]);
              }
              if (p.out_hi) {
                uint4 h, l;
                split_pack2(v[0], v[1], h.x, l.x);
                split_pack2(v[2], v[3], h.y, l.y);
                split_pack2(v[4], v[5], h.z, l.z);
                split_pack2(v[6], v[7], h.w, l.w);
                *reinterpret_cast<uint4*>(p.out_hi + mo * p.ldo_bf + n) = h;
                if (p.out_lo) *reinterpret_cast<uint4*>(p.out_lo + mo * p.ldo_bf + n) = l;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (n + j >= p.N) break;
                float x = v[j];
                if (p.bias) x += __ldg(p.bias + n + j);
                if (p.act == MTT_ACT_GELU) x = gelu_erf(x);
                else if (p.act == MTT_ACT_RELU) x = fmaxf(x, 0.f);
                if (p.residual) x += p.residual[mr * p.ldr + n + j];
                if (p.out_f32) p.out_f32[mo * p.ldo_f32 + n + j] = x;
                if (p.out_hi) {
                  __nv_bfloat16 h, l;
                  split_bf16(x, h, l);
                  p.out_hi[mo * p.ldo_bf + n + j] = h;
                  if (p.out_lo) p.out_lo[mo * p.ldo_bf + n + j] = l;
                }
              }
            }
          }
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

// Pick the TH x TW (<= 128 pixel) output patch that wastes the fewest MMA rows.
static void pick_conv_tile(int H, int W, int* TW, int* TH) {
  double best = -1;
  int btw = 1, bth = 1;
  for (int tw = 1; tw <= 128 && tw <= W; ++tw) {
    int th = 128 / tw;
    if (th > H) th = H;
    if (th < 1) continue;
    const long long tiles = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th);
    const double eff = (double)H * W / ((double)tiles * 128.0);
    if (eff > best + 1e-9 || (eff > best - 1e-9 && tw > btw)) {
      best = eff;
      btw = tw;
      bth = th;
    }
  }
  *TW = btw;
  *TH = bth;
}

template <int NSPLIT>
static int launch_gemm(const CUtensorMap* maps, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<NSPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<NSPLIT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess)
      return set_error(MTT_ERR_LAUNCH, "gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_tc_kernel<NSPLIT><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(maps[0], maps[1], maps[2],
                                                                         maps[3], p);
  return check_launch("mtt_gemm");
}

}  // namespace mtt

extern "C" int mtt_gemm(const mtt_gemm_desc* d, mtt_stream_t stream_) {
  using namespace mtt;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d) return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm: null descriptor");
  if (d->M <= 0 || d->N <= 0 || d->K <= 0)
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm: M=%d N=%d K=%d must be positive", d->M, d->N, d->K);
  if (d->nsplit != 1 && d->nsplit != 2)
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm: nsplit=%d (1 or 2)", d->nsplit);
  if (!d->a_hi || !d->b_hi || (d->nsplit == 2 && (!d->a_lo || !d->b_lo)))
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm: missing operand plane");
  if (!d->out_f32 && !d->out_hi) return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm: no output");
  if (d->lda % 8 || d->ldb % 8)
    return set_error(MTT_ERR_MISALIGNED, "mtt_gemm: lda=%lld ldb=%lld must be multiples of 8",
                     (long long)d->lda, (long long)d->ldb);

  GemmParams p{};
  p.M = d->M;
  p.N = d->N;
  p.mode = d->mode;
  p.num_kb = (d->K + BK - 1) / BK;
  p.tiles_n = (d->N + BN - 1) / BN;
  p.bias = d->bias;
  p.act = d->act;
  p.residual = d->residual;
  p.ldr = d->ldr;
  p.res_row_mod = d->res_row_mod;
  p.out_f32 = d->out_f32;
  p.ldo_f32 = d->ldo_f32;
  p.out_hi = static_cast<__nv_bfloat16*>(d->out_hi);
  p.out_lo = d->nsplit == 2 ? static_cast<__nv_bfloat16*>(d->out_lo) : nullptr;
  p.ldo_bf = d->ldo_bf;
  p.in_group = d->in_group;
  p.out_group = d->out_group;
  p.out_offset = d->out_offset;
  if (p.out_hi && d->nsplit == 2 && !p.out_lo)
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm: out_lo missing for nsplit=2");

  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  p.vec_ok = 1;
  if (d->bias && !al16(d->bias)) p.vec_ok = 0;
  if (d->residual && (!al16(d->residual) || d->ldr % 4)) p.vec_ok = 0;
  if (d->out_f32 && (!al16(d->out_f32) || d->ldo_f32 % 4)) p.vec_ok = 0;
  if (d->out_hi && (!al16(d->out_hi) || d->ldo_bf % 8 || (p.out_lo && !al16(p.out_lo)))) p.vec_ok = 0;

  CUtensorMap maps[4];
  int rc;
  const int ksq = (d->mode == 1) ? d->ksize * d->ksize : 1;
  if (d->mode == 0 && d->a_group_rows > 0) {
    // gathered A: logical row r = (g, i) lives at physical row g * a_group_stride + i, i < a_group_rows
    const int g = d->a_group_rows;
    if (g > BM || d->M % g || d->a_group_stride < g)
      return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm: a_group_rows=%d must divide M=%d and be <= %d", g, d->M, BM);
    const int ngroups = d->M / g;
    const int gpt = BM / g;  // groups per tile
    if (ngroups > gpt && BM % g)
      return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm: %d groups of %d rows need %d %% %d == 0", ngroups, g, BM, g);
    p.taps = 1;
    p.ksize = 1;
    p.dil = 1;
    p.cin_pad = 0;
    p.a_groups_per_tile = ngroups < gpt ? ngroups : gpt;
    p.tiles_m = (ngroups + p.a_groups_per_tile - 1) / p.a_groups_per_tile;
    if (p.tiles_m > 1 && p.a_groups_per_tile * g != BM)
      return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm: gathered A tiles must be full");
    p.a_box_bytes = (uint32_t)(p.a_groups_per_tile * g) * BK * 2;
    const uint64_t dims[3] = {(uint64_t)d->K, (uint64_t)g, (uint64_t)ngroups};
    const uint64_t str[2] = {(uint64_t)d->lda * 2, (uint64_t)d->a_group_stride * d->lda * 2};
    const uint32_t box[3] = {BK, (uint32_t)g, (uint32_t)p.a_groups_per_tile};
    if ((rc = make_tmap_bf16(&maps[0], d->a_hi, 3, dims, str, box))) return rc;
    if (d->nsplit == 2) {
      if ((rc = make_tmap_bf16(&maps[1], d->a_lo, 3, dims, str, box))) return rc;
    } else {
      maps[1] = maps[0];
    }
  } else if (d->mode == 0) {
    p.taps = 1;
    p.ksize = 1;
    p.dil = 1;
    p.cin_pad = 0;
    p.tiles_m = (d->M + BM - 1) / BM;
    p.a_box_bytes = kTileBytes;
    const uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->M};
    const uint64_t str[1] = {(uint64_t)d->lda * 2};
    const uint32_t box[2] = {BK, BM};
    if ((rc = make_tmap_bf16(&maps[0], d->a_hi, 2, dims, str, box))) return rc;
    if (d->nsplit == 2) {
      if ((rc = make_tmap_bf16(&maps[1], d->a_lo, 2, dims, str, box))) return rc;
    } else {
      maps[1] = maps[0];
    }
  } else if (d->mode == 1) {
    if (d->B <= 0 || d->H <= 0 || d->W <= 0 || (long long)d->B * d->H * d->W != d->M)
      return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm(conv): B*H*W = %d*%d*%d != M = %d", d->B, d->H,
                       d->W, d->M);
    if (d->ksize != 1 && d->ksize != 3)
      return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm(conv): ksize=%d (1 or 3)", d->ksize);
    if (d->dil < 1) return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm(conv): dil=%d", d->dil);
    p.taps = ksq;
    p.ksize = d->ksize;
    p.dil = d->dil;
    p.cin_pad = p.num_kb * BK;
    p.H = d->H;
    p.W = d->W;
    pick_conv_tile(d->H, d->W, &p.TW, &p.TH);
    p.tiles_x = (d->W + p.TW - 1) / p.TW;
    p.tiles_y = (d->H + p.TH - 1) / p.TH;
    p.tiles_m = d->B * p.tiles_x * p.tiles_y;
    p.a_box_bytes = (uint32_t)(p.TW * p.TH) * BK * 2;
    const uint64_t dims[4] = {(uint64_t)d->K, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t str[3] = {(uint64_t)d->lda * 2, (uint64_t)d->W * d->lda * 2,
                             (uint64_t)d->H * d->W * d->lda * 2};
    const uint32_t box[4] = {BK, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    if ((rc = make_tmap_bf16(&maps[0], d->a_hi, 4, dims, str, box))) return rc;
    if (d->nsplit == 2) {
      if ((rc = make_tmap_bf16(&maps[1], d->a_lo, 4, dims, str, box))) return rc;
    } else {
      maps[1] = maps[0];
    }
  } else {
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm: mode=%d", d->mode);
  }
  {
    const uint64_t ktot = (d->mode == 1) ? (uint64_t)ksq * p.cin_pad : (uint64_t)d->K;
    if ((uint64_t)d->ldb < ktot)
      return set_error(MTT_ERR_BAD_SHAPE, "mtt_gemm: ldb=%lld < packed K=%llu", (long long)d->ldb,
                       (unsigned long long)ktot);
    const uint64_t dims[2] = {ktot, (uint64_t)d->N};
    const uint64_t str[1] = {(uint64_t)d->ldb * 2};
    const uint32_t box[2] = {BK, BN};
    if ((rc = make_tmap_bf16(&maps[2], d->b_hi, 2, dims, str, box))) return rc;
    if (d->nsplit == 2) {
      if ((rc = make_tmap_bf16(&maps[3], d->b_lo, 2, dims, str, box))) return rc;
    } else {
      maps[3] = maps[2];
    }
  }
  return d->nsplit == 2 ? launch_gemm<2>(maps, p, stream) : launch_gemm<1>(maps, p, stream);
}
