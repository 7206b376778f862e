// Micro-probe: how many SM cycles one tcgen05.mma (M = 128, K = 16, 16-bit operands) takes as a function of N, of
// where A comes from (shared memory / TMEM), of the B layout (K-major / MN-major), of whether consecutive MMAs
// accumulate into the SAME TMEM tile or alternate between two tiles, and of a second CTA on the same SM doing the
// same.  Also checks that kind::f16 accepts A = fp16 with B = bf16 (numerically).  Development aid for the
// attention kernel (its MMAs are N = 64, i.e. 32 cycles of math each at full rate); not part of the product.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/mma_probe scripts/mma_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>

#include "../multi-task-transformer_b200/csrc/ptx.cuh"

using namespace mtt;

struct Cfg {
  int n;       // MMA N
  int a_tmem;  // 1: A from TMEM (TS form)
  int b_mn;    // 1: B is MN-major (like V in P V)
  int ntiles;  // number of D tiles the sequence alternates over (1, 2)
  int run;     // consecutive MMAs on one tile before switching (1 = strict alternation, 12 = like attention)
  int count;   // MMAs per measurement
};

__global__ void __launch_bounds__(128, 2) probe(Cfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x < 32) tmem_alloc<256>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x < 32) {
    const uint32_t idesc = umma_idesc_bf16(128, c.n, c.b_mn);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
        // descriptors and tile addresses are loop-invariant: the loop body is 12 bare MMAs, as in the kernels
        const uint64_t b0 = umma_desc_sw128(sb), b1 = umma_desc_sw128(sb + (c.b_mn ? 2048 : 32));
        const uint64_t a0 = umma_desc_sw128(sa), a1 = umma_desc_sw128(sa + 32);
        const uint32_t ta0 = tb + 224, ta1 = tb + 232;
        const uint32_t d0 = tb, d1 = tb + (c.ntiles > 1 ? c.n : 0);
        const bool alt1 = c.ntiles > 1 && c.run == 1;   // strict alternation
        const bool alt12 = c.ntiles > 1 && c.run != 1;  // 12 on one tile, 12 on the other
        for (int i = 0; i < c.count; i += 12) {
          const uint32_t dA = (alt12 && ((i / 12) & 1)) ? d1 : d0;
          const uint32_t dB = alt1 ? d1 : dA;
          if (c.a_tmem) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
              umma_ts(dA, ta0, b0, idesc, 1);
              umma_ts(dB, ta1, b1, idesc, 1);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
              umma_ss(dA, a0, b0, idesc, 1);
              umma_ss(dB, a1, b1, idesc, 1);
            }
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<256>(tb);
  }
}

// D = A * B^T with A = fp16 (all 1.5) and B = bf16 (all 2.0), K = 16: 48 if the formats are honoured.
__global__ void __launch_bounds__(128, 1) mixed_check(float* out, int a_fmt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint16_t av = a_fmt == 0 ? 0x3E00 /* fp16 1.5 */ : 0x3FC0 /* bf16 1.5 */;
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) {
    reinterpret_cast<uint16_t*>(smem)[i] = av;                // A tile
    reinterpret_cast<uint16_t*>(smem + 32768)[i] = 0x4000;    // B tile: bf16 2.0
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x < 32) tmem_alloc<64>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x < 32) {
    // idesc: D f32 (1<<4), A format (0 = f16, 1 = bf16) << 7, B format bf16 (1 << 10), N = 64, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)a_fmt << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    if (elect_one()) {
      umma_ss(tb, umma_desc_sw128(smem_u32(smem)), umma_desc_sw128(smem_u32(smem + 32768)), idesc, 0);
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t r[16];
    tmem_ld16(tb, r);
    tmem_ld_wait();
    if (threadIdx.x == 0) out[0] = __uint_as_float(r[0]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<64>(tb);
  }
}


// Interference between the tensor pipe and tcgen05.ld / tcgen05.st issued by other warps of the same CTA:
// warp 0 issues `count` TS-form N = 64 MMAs (or none); warps 1..W loop `iters` times over
//   tcgen05.ld 32x32b.x32 + wait  [+ tcgen05.st 32x32b.x32 + wait]   on columns the MMAs do not touch.
// out[0] = cycles of the MMA stream, out[1] = cycles of warp 1's load/store loop.
__global__ void __launch_bounds__(288, 1) probe_ldst(int count, int iters, int do_st, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x < 32) tmem_alloc<256>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0);
    const uint32_t sb = smem_u32(smem + 32768);
    __syncwarp();
    const long long t0 = clock64();
    if (count > 0) {
      if (elect_one()) {
        const uint64_t b0 = umma_desc_sw128(sb), b1 = umma_desc_sw128(sb + 32);
        for (int i = 0; i < count; i += 12) {
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            umma_ts(tb, tb + 224, b0, idesc, 1);
            umma_ts(tb + 64, tb + 232, b1, idesc, 1);
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, 0);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  } else {
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col = tb + 128 + (warp > 4 ? 32 : 0) + lane_addr;
    uint32_t r[32];
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      tmem_ld32(col, r);
      tmem_ld_wait();
      if (do_st) {
#pragma unroll
        for (int k = 0; k < 32; ++k) r[k] += 1;
        tmem_st32(col, r);
        tmem_st_wait();
      }
    }
    const long long t1 = clock64();
    if (warp == 1 && (threadIdx.x & 31) == 0 && blockIdx.x == 0) out[1] = (iters > 0) ? (t1 - t0) + (r[0] & 0) : 0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<256>(tb);
  }
}

static void run_ldst(int warps, int count, int iters, int do_st, long long* d) {
  cudaMemset(d, 0, 16);
  probe_ldst<<<148, 32 * (1 + warps), 66 * 1024>>>(count, iters, do_st, d);
  long long h[2] = {0, 0};
  cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    printf("error %s\n", cudaGetErrorString(e));
    exit(1);
  }
  printf("ld/st warps %d  MMAs %4d  ld%s iters %4d | MMA stream %7lld cycles (%.1f / MMA)   ld loop %7lld cycles (%.1f / iter)\n",
         warps, count, do_st ? "+st" : "   ", iters, h[0], count ? (double)h[0] / count : 0.0, h[1],
         iters ? (double)h[1] / iters : 0.0);
}

static double run(const Cfg& c, int ctas, long long* d) {
  probe<<<ctas, 128, 66 * 1024>>>(c, d);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    printf("error %s\n", cudaGetErrorString(e));
    exit(1);
  }
  return (double)h / c.count;
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  cudaFuncSetAttribute(mixed_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  printf("  N a_src b_layout tiles run  | cycles per MMA: 1 CTA, 148 CTAs (1/SM), 296 CTAs (2/SM, per CTA)   ideal N/2\n");
  struct Row { int n, a, b, nt, run; };
  const Row rows[] = {
      {64, 0, 0, 1, 1},  {64, 0, 0, 2, 1},  {64, 0, 0, 2, 12}, {64, 1, 0, 1, 1},  {64, 1, 0, 2, 1}, {64, 1, 0, 2, 12},
      {64, 0, 1, 1, 1},  {64, 1, 1, 1, 1},  {64, 1, 1, 2, 1},  {64, 1, 1, 2, 12}, {128, 0, 0, 1, 1}, {128, 0, 0, 2, 1},
      {128, 1, 0, 1, 1}, {256, 0, 0, 1, 1}, {32, 1, 0, 1, 1},  {32, 1, 0, 2, 1},  {16, 1, 0, 1, 1},
  };
  for (const Row& r : rows) {
    Cfg c{r.n, r.a, r.b, r.nt, r.run, 240};
    printf("%3d %s %s %d %2d  | %6.1f %6.1f %6.1f   %d\n", r.n, r.a ? "tmem" : "smem", r.b ? "mn" : "k ", r.nt, r.run,
           run(c, 1, d), run(c, 148, d), run(c, 296, d), r.n / 2);
  }
  cudaFuncSetAttribute(probe_ldst, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  run_ldst(4, 960, 0, 0, d);      // MMAs alone
  run_ldst(1, 0, 200, 0, d);      // one warp loading alone
  run_ldst(4, 0, 200, 0, d);      // four warps (all lane quarters) loading
  run_ldst(8, 0, 200, 0, d);
  run_ldst(4, 0, 200, 1, d);      // load + store
  run_ldst(4, 960, 200, 0, d);    // MMAs under load traffic
  run_ldst(8, 960, 200, 0, d);
  run_ldst(4, 960, 200, 1, d);    // MMAs under load + store traffic
  run_ldst(8, 960, 200, 1, d);
  run_ldst(4, 960, 40, 1, d);     // a shorter burst of traffic
  float* f;
  cudaMalloc(&f, 4);
  for (int a_fmt = 1; a_fmt >= 0; --a_fmt) {
    mixed_check<<<1, 128, 66 * 1024>>>(f, a_fmt);
    float h = 0;
    cudaError_t e = cudaMemcpy(&h, f, 4, cudaMemcpyDeviceToHost);
    printf("A %s x B bf16, K = 16, all 1.5 * 2.0: D = %g (expected 48)  [%s]\n", a_fmt ? "bf16" : "fp16", h,
           cudaGetErrorString(e));
  }
  return 0;
}
