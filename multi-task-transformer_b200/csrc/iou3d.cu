// Bird's-eye-view IoU of rotated boxes and rotated / axis-aligned NMS for the 3D-detection post-processing of the
// Cityscapes-3D branch. Replaces the reference's only native code:
//   TP/detection_toolbox/iou3d/src/iou3d_kernel.cu:124-250 box_overlap / iou_bev (edge-intersection + contained
//     corners, angular sort, shoelace), :253-283 the pairwise kernels, :285-330 nms_kernel, :332-385 nms_normal_kernel;
//   TP/detection_toolbox/iou3d/src/iou3d.cpp:96-202 nms_gpu / nms_normal_gpu (mask matrix copied to the HOST with a
//     blocking cudaMemcpy, greedy sweep on the CPU, cudaMalloc / cudaFree per call).
// Here: box format and results contract are the reference's ([x1, y1, x2, y2, ry], boxes sorted by score; keep =
// indices of the survivors in order), but
//   * the overlap polygon is obtained by clipping box B against the four half-planes of box A in A's own frame
//     (Sutherland-Hodgman): no trigonometric sort, at most 8 vertices, branch-light;
//   * the greedy sweep runs ON THE DEVICE (one thread block, the removal bit-set in shared memory), so a call is
//     enqueue-only: no host round trip, no allocation (caller-provided workspace), the count lands in device memory.
#include <math.h>

#include "host_common.h"

namespace mtt {

constexpr float kIouEps = 1e-8f;  // iou3d_kernel.cu:18

struct P2 {
  float x, y;
};

// Corners of box [x1,y1,x2,y2,ry] rotated about its centre the way the reference rotates them
// (rotate_around_center, iou3d_kernel.cu:104-113: x' = dx cos + dy sin, y' = -dx sin + dy cos).
__device__ __forceinline__ void box_corners(const float* b, P2 (&c)[4]) {
  const float cx = 0.5f * (b[0] + b[2]), cy = 0.5f * (b[1] + b[3]);
  const float cs = cosf(b[4]), sn = sinf(b[4]);
  const float xs[4] = {b[0], b[2], b[2], b[0]}, ys[4] = {b[1], b[1], b[3], b[3]};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float dx = xs[k] - cx, dy = ys[k] - cy;
    c[k].x = dx * cs + dy * sn + cx;
    c[k].y = -dx * sn + dy * cs + cy;
  }
}

// Clip polygon `in` (n vertices) against the half-plane a * x + b * y <= c.
__device__ __forceinline__ int clip_halfplane(const P2* in, int n, float a, float b, float c, P2* out) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const P2 p = in[i], q = in[(i + 1 == n) ? 0 : i + 1];
    const float dp = a * p.x + b * p.y - c, dq = a * q.x + b * q.y - c;
    if (dp <= 0.f) out[m++] = p;
    if ((dp < 0.f && dq > 0.f) || (dp > 0.f && dq < 0.f)) {
      const float t = dp / (dp - dq);
      out[m].x = p.x + t * (q.x - p.x);
      out[m].y = p.y + t * (q.y - p.y);
      ++m;
    }
  }
  return m;
}

// Area of the intersection of two rotated boxes (iou3d_kernel.cu:124-241 computes the same polygon another way).
__device__ float box_overlap_bev(const float* box_a, const float* box_b) {
  P2 cb[4];
  box_corners(box_b, cb);
  // B's corners in A's frame: undo A's rotation about A's centre, then A is the axis-aligned rectangle it is stored as
  const float cx = 0.5f * (box_a[0] + box_a[2]), cy = 0.5f * (box_a[1] + box_a[3]);
  const float cs = cosf(box_a[4]), sn = sinf(box_a[4]);
  P2 poly[8], tmp[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float dx = cb[k].x - cx, dy = cb[k].y - cy;
    poly[k].x = dx * cs - dy * sn + cx;      // inverse of x' = dx cos + dy sin, y' = -dx sin + dy cos
    poly[k].y = dx * sn + dy * cs + cy;
  }
  int n = 4;
  n = clip_halfplane(poly, n, -1.f, 0.f, -box_a[0], tmp);   // x >= x1
  if (n < 3) return 0.f;
  n = clip_halfplane(tmp, n, 1.f, 0.f, box_a[2], poly);     // x <= x2
  if (n < 3) return 0.f;
  n = clip_halfplane(poly, n, 0.f, -1.f, -box_a[1], tmp);   // y >= y1
  if (n < 3) return 0.f;
  n = clip_halfplane(tmp, n, 0.f, 1.f, box_a[3], poly);     // y <= y2
  if (n < 3) return 0.f;
  float area = 0.f;
  for (int k = 1; k + 1 < n; ++k)
    area += (poly[k].x - poly[0].x) * (poly[k + 1].y - poly[0].y) - (poly[k].y - poly[0].y) * (poly[k + 1].x - poly[0].x);
  return 0.5f * fabsf(area);
}

__device__ __forceinline__ float iou_bev(const float* a, const float* b) {   // iou3d_kernel.cu:243-251
  const float sa = (a[2] - a[0]) * (a[3] - a[1]), sb = (b[2] - b[0]) * (b[3] - b[1]);
  const float so = box_overlap_bev(a, b);
  return so / fmaxf(sa + sb - so, kIouEps);
}

__device__ __forceinline__ float iou_normal(const float* a, const float* b) {   // iou3d_kernel.cu:332-340
  const float w = fmaxf(fminf(a[2], b[2]) - fmaxf(a[0], b[0]), 0.f);
  const float h = fmaxf(fminf(a[3], b[3]) - fmaxf(a[1], b[1]), 0.f);
  const float inter = w * h;
  const float sa = (a[2] - a[0]) * (a[3] - a[1]), sb = (b[2] - b[0]) * (b[3] - b[1]);
  return inter / fmaxf(sa + sb - inter, kIouEps);
}

// mode 0: overlap area, 1: rotated IoU (boxes_overlap_kernel / boxes_iou_bev_kernel, iou3d_kernel.cu:253-283)
__global__ void __launch_bounds__(256)
pairwise_bev_kernel(const float* __restrict__ a, int na, const float* __restrict__ b, int nb, int mode,
                    float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)na * nb) return;
  const float* pa = a + (i / nb) * 5;
  const float* pb = b + (i % nb) * 5;
  out[i] = mode == 0 ? box_overlap_bev(pa, pb) : iou_bev(pa, pb);
}

// Suppression bit matrix: mask[i][w] bit j set when box 64 w + j (> i) overlaps box i above the threshold
// (nms_kernel / nms_normal_kernel, iou3d_kernel.cu:285-330, :342-385). One block = 64 rows x one 64-box column block.
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float* __restrict__ boxes, int n, float thresh, int rotated, unsigned long long* __restrict__ mask) {
  const int row0 = blockIdx.y * 64, col0 = blockIdx.x * 64;
  const int words = (n + 63) / 64;
  __shared__ float cb[64 * 5];
  const int ncol = min(64, n - col0);
  if ((int)threadIdx.x < ncol)
    for (int k = 0; k < 5; ++k) cb[threadIdx.x * 5 + k] = boxes[(long long)(col0 + threadIdx.x) * 5 + k];
  __syncthreads();
  const int i = row0 + threadIdx.x;
  if (i >= n) return;
  unsigned long long bits = 0;
  if (blockIdx.x >= blockIdx.y) {          // only boxes after i can be suppressed by i
    float me[5];
    for (int k = 0; k < 5; ++k) me[k] = boxes[(long long)i * 5 + k];
    const int start = (row0 == col0) ? threadIdx.x + 1 : 0;
    for (int j = start; j < ncol; ++j) {
      const float v = rotated ? iou_bev(me, cb + j * 5) : iou_normal(me, cb + j * 5);
      if (v > thresh) bits |= 1ULL << j;
    }
  }
  mask[(long long)i * words + blockIdx.x] = bits;
}

// Greedy sweep in score order on the device (iou3d.cpp:131-143 does this on the host after a blocking copy).
__global__ void __launch_bounds__(1024)
nms_sweep_kernel(const unsigned long long* __restrict__ mask, int n, long long* __restrict__ keep, int* __restrict__ num_keep) {
  extern __shared__ unsigned long long removed[];   // words
  const int words = (n + 63) / 64;
  for (int w = threadIdx.x; w < words; w += blockDim.x) removed[w] = 0;
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  for (int i = 0; i < n; ++i) {
    const bool alive = !(removed[i >> 6] & (1ULL << (i & 63)));   // every thread reads the same word: uniform branch
    __syncthreads();
    if (alive) {
      if (threadIdx.x == 0) keep[cnt++] = i;
      for (int w = (i >> 6) + threadIdx.x; w < words; w += blockDim.x) removed[w] |= mask[(long long)i * words + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_keep = cnt;
}

}  // namespace mtt

using namespace mtt;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" {

int mtt_boxes_bev_pairwise(const float* boxes_a, int32_t num_a, const float* boxes_b, int32_t num_b, int32_t mode,
                           float* out, mtt_stream_t stream) {
  if (!boxes_a || !boxes_b || !out || num_a < 0 || num_b < 0 || mode < 0 || mode > 1)
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_boxes_bev_pairwise: bad arguments");
  const long long total = (long long)num_a * num_b;
  if (total == 0) return MTT_OK;
  pairwise_bev_kernel<<<(unsigned)((total + 255) / 256), 256, 0, STREAM>>>(boxes_a, num_a, boxes_b, num_b, mode, out);
  return check_launch("mtt_boxes_bev_pairwise");
}

size_t mtt_nms_workspace_bytes(int32_t n) {
  const size_t words = ((size_t)(n > 0 ? n : 0) + 63) / 64;
  return (size_t)(n > 0 ? n : 0) * words * sizeof(unsigned long long) + 16;
}

int mtt_nms_bev(const float* boxes, int32_t n, float thresh, int32_t rotated, int64_t* keep, int32_t* num_keep,
                void* workspace, size_t ws_bytes, mtt_stream_t stream) {
  if (!num_keep || n < 0 || (n > 0 && (!boxes || !keep)))
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_nms_bev: bad arguments (n=%d)", n);
  if (n == 0) {
    cudaMemsetAsync(num_keep, 0, sizeof(int32_t), STREAM);
    return MTT_OK;
  }
  if (!workspace || ws_bytes < mtt_nms_workspace_bytes(n) || (reinterpret_cast<uintptr_t>(workspace) & 7))
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_nms_bev: workspace %zu bytes < %zu (mtt_nms_workspace_bytes) or misaligned",
                     ws_bytes, mtt_nms_workspace_bytes(n));
  const int words = (n + 63) / 64;
  if ((size_t)words * 8 > 200 * 1024) return set_error(MTT_ERR_BAD_SHAPE, "mtt_nms_bev: n=%d too large", n);
  unsigned long long* mask = static_cast<unsigned long long*>(workspace);
  nms_mask_kernel<<<dim3(words, words), 64, 0, STREAM>>>(boxes, n, thresh, rotated, mask);
  int rc = check_launch("mtt_nms_bev(mask)");
  if (rc) return rc;
  static bool attr[kMaxDevices] = {};
  const int dev_ = current_device();
  if (!attr[dev_] && (size_t)words * 8 > 48 * 1024) {
    cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr[dev_] = true;
  }
  const int threads = words >= 1024 ? 1024 : (words >= 256 ? 256 : 64);
  nms_sweep_kernel<<<1, threads, (size_t)words * 8, STREAM>>>(mask, n, reinterpret_cast<long long*>(keep), num_keep);
  return check_launch("mtt_nms_bev(sweep)");
}

}  // extern "C"
