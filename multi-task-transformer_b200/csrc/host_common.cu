#include <string.h>

#include "host_common.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>

#include <vector>

namespace mtt {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch() { ++g_launches; }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return set_error(MTT_ERR_LAUNCH, "%s: launch failed: %s", what, cudaGetErrorString(e));
  count_launch();
  return MTT_OK;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box) {
  auto enc = get_encode();
  if (!enc) return set_error(MTT_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not found");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return set_error(MTT_ERR_MISALIGNED, "TMA base pointer %p is not 16-byte aligned", base);
  for (int i = 0; i + 1 < rank; ++i)
    if (strides_bytes[i] % 16 != 0)
      return set_error(MTT_ERR_MISALIGNED, "TMA stride %d = %llu bytes is not a multiple of 16", i,
                       (unsigned long long)strides_bytes[i]);
  // The descriptor is a pure function of (base, rank, dims, strides, box): a small per-thread direct-mapped cache keeps
  // the eager (non-graph) launch path and batch-1 latency from paying four driver encodes per GEMM / attention call
  // (inside a CUDA graph the maps are baked into the kernel parameters at capture and none of this runs on replay).
  struct Key {
    const void* base;
    int rank;
    uint64_t dims[5], strides[4];
    uint32_t box[5];
  };
  struct Slot {
    bool used;
    Key key;
    CUtensorMap map;
  };
  constexpr int kSlots = 512;
  static thread_local Slot* cache = nullptr;   // heap, not a 120 KB TLS block in a dlopen'ed library
  if (!cache) cache = new Slot[kSlots]();
  Key k;
  memset(&k, 0, sizeof(k));
  k.base = base;
  k.rank = rank;
  uint64_t h = reinterpret_cast<uintptr_t>(base) * 0x9E3779B97F4A7C15ull + (uint64_t)rank;
  for (int i = 0; i < rank; ++i) {
    k.dims[i] = dims[i];
    k.box[i] = box[i];
    if (i + 1 < rank) k.strides[i] = strides_bytes[i];
    h = (h ^ dims[i]) * 0x9E3779B97F4A7C15ull;
    h = (h ^ box[i]) * 0x9E3779B97F4A7C15ull;
    if (i + 1 < rank) h = (h ^ strides_bytes[i]) * 0x9E3779B97F4A7C15ull;
  }
  Slot& slot = cache[(h >> 32) % kSlots];
  if (slot.used && memcmp(&slot.key, &k, sizeof(k)) == 0) {
    *out = slot.map;
    return MTT_OK;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                   gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(MTT_ERR_DRIVER,
                     "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu,%llu] box "
                     "[%u,%u,%u,%u,%u]",
                     (int)r, rank, (unsigned long long)dims[0],
                     (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0),
                     (unsigned long long)(rank > 3 ? dims[3] : 0),
                     (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0,
                     rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
  }
  slot.used = true;
  slot.key = k;
  slot.map = *out;
  return MTT_OK;
}

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

int sm_count() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (!n[dev]) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
}

// ---------------------------------------------------------------- per-launch timing (bench.py's roofline block)
struct ProfileEntry {
  cudaEvent_t start, stop;
  int kind, M, N, K;
  double flops;
};
static bool g_prof_on = false;
static std::vector<ProfileEntry> g_prof;

ProfileScope::ProfileScope(cudaStream_t stream, int kind, double flops, int M, int N, int K) : stream_(stream), slot_(-1) {
  if (!g_prof_on) return;
  ProfileEntry e;
  e.kind = kind;
  e.M = M;
  e.N = N;
  e.K = K;
  e.flops = flops;
  if (cudaEventCreate(&e.start) != cudaSuccess || cudaEventCreate(&e.stop) != cudaSuccess) return;
  cudaEventRecord(e.start, stream);
  g_prof.push_back(e);
  slot_ = (int)g_prof.size() - 1;
}
ProfileScope::~ProfileScope() {
  if (slot_ >= 0) cudaEventRecord(g_prof[slot_].stop, stream_);
}

}  // namespace mtt

extern "C" {

int mtt_profile_begin(void) {
  for (auto& e : mtt::g_prof) {
    cudaEventDestroy(e.start);
    cudaEventDestroy(e.stop);
  }
  mtt::g_prof.clear();
  mtt::g_prof_on = true;
  return MTT_OK;
}

int mtt_profile_end(mtt_profile_rec* out, int32_t max_recs, int32_t* n_recs) {
  mtt::g_prof_on = false;
  if (cudaDeviceSynchronize() != cudaSuccess) return mtt::set_error(MTT_ERR_LAUNCH, "mtt_profile_end: device sync failed");
  int n = 0;
  for (auto& e : mtt::g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e.start, e.stop);
    if (out && n < max_recs) {
      out[n].kind = e.kind;
      out[n].M = e.M;
      out[n].N = e.N;
      out[n].K = e.K;
      out[n].ms = ms;
      out[n].flops = e.flops;
    }
    ++n;
    cudaEventDestroy(e.start);
    cudaEventDestroy(e.stop);
  }
  mtt::g_prof.clear();
  if (n_recs) *n_recs = n;
  return MTT_OK;
}

int mtt_version(void) { return 200; }

const char* mtt_last_error(void) { return mtt::g_err; }

int mtt_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess)
    return mtt::set_error(MTT_ERR_UNSUPPORTED_ARCH, "no CUDA device");
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10)
    return mtt::set_error(MTT_ERR_UNSUPPORTED_ARCH,
                          "device compute capability %d.%d; this library is sm_100a only", major,
                          minor);
  return MTT_OK;
}

int64_t mtt_launch_count(void) { return mtt::g_launches; }
void mtt_launch_count_reset(void) { mtt::g_launches = 0; }
}
