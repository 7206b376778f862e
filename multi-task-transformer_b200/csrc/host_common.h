// Host-side helpers shared by the C-ABI entry points: error slot, launch counter, TMA descriptors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mtt_b200.h"

namespace mtt {

int set_error(int code, const char* fmt, ...);
void count_launch();

// Checks the last launch; returns 0 or MTT_ERR_LAUNCH with the CUDA error text recorded.
int check_launch(const char* what);

// rank-N (N<=5) bf16 tensor map, 128-byte swizzle, zero OOB fill.
// dims/strides innermost first; strides_bytes[i] is the byte stride of dim i+1 (rank-1 entries).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box);

// ordinal of the calling thread's current device (< kMaxDevices), and its SM count (cached per device)
constexpr int kMaxDevices = 64;
int current_device();
int sm_count();

// Profiling aid (mtt_profile_begin / mtt_profile_end): while active, every tensor-core launch of the library is
// bracketed by CUDA events on its own stream and recorded with its algorithmic FLOPs. Not usable during graph capture.
struct ProfileScope {
  ProfileScope(cudaStream_t stream, int kind, double flops, int M, int N, int K);
  ~ProfileScope();
  cudaStream_t stream_;
  int slot_;
};

}  // namespace mtt
