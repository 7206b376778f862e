// Inference-time image pre-processing of the reference, one kernel (the step in front of
// the hot path): TaskPrompter/inference.py:127-133 (cv2.imread -> float32 -> BGR2RGB), :93-115 get_infer_transforms =
// Normalize (data/transforms.py:236-251: x / 255, - mean, / std) -> DirectResize (inference.py:66-81, cv2.resize
// INTER_LINEAR) -> ToTensor (transforms.py:265-273, HWC -> CHW).  HBM-bound and tiny (0.8 MB in, 3 MB out at 512^2):
// one thread per output pixel, coalesced NCHW stores; the point is that the image reaches the patch-embed im2col
// without a host round trip.
//
// Arithmetic follows the reference's order in fp32 (IEEE division, no FMA contraction) so the result matches the
// CPU pipeline to the last bit or two: normalise the four neighbours, interpolate horizontally, then vertically.
// Source coordinates are computed in double like cv2 (resize.cpp: fx = (dx + 0.5) * scale - 0.5).
#include "host_common.h"

namespace mtt {

struct Norm3 {
  float mean[3];
  float std[3];
};

__device__ __forceinline__ void src_coord(int d, double scale, int n_in, int& i0, int& i1, float& l1) {
  double s = ((double)d + 0.5) * scale - 0.5;
  if (s < 0.0) s = 0.0;
  i0 = (int)floor(s);
  if (i0 > n_in - 1) i0 = n_in - 1;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  l1 = (float)(s - (double)i0);
}

__global__ void __launch_bounds__(256)
preprocess_kernel(const uint8_t* __restrict__ img, int B, int h, int w, int bgr, Norm3 nm, float* __restrict__ out,
                  int H, int W, double sy, double sx) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W;
  if (idx >= total) return;
  const int x = (int)(idx % W);
  const int y = (int)((idx / W) % H);
  const int b = (int)(idx / ((long long)W * H));
  int y0, y1, x0, x1;
  float ly, lx;
  src_coord(y, sy, h, y0, y1, ly);
  src_coord(x, sx, w, x0, x1, lx);
  const uint8_t* base = img + (long long)b * h * w * 3;
  const uint8_t* p00 = base + ((long long)y0 * w + x0) * 3;
  const uint8_t* p01 = base + ((long long)y0 * w + x1) * 3;
  const uint8_t* p10 = base + ((long long)y1 * w + x0) * 3;
  const uint8_t* p11 = base + ((long long)y1 * w + x1) * 3;
  const float wx0 = __fsub_rn(1.0f, lx), wy0 = __fsub_rn(1.0f, ly);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int sc = bgr ? 2 - c : c;  // output channel c is R, G, B
    const float m = nm.mean[c], sd = nm.std[c];
    const float v00 = __fdiv_rn(__fsub_rn(__fdiv_rn((float)p00[sc], 255.0f), m), sd);
    const float v01 = __fdiv_rn(__fsub_rn(__fdiv_rn((float)p01[sc], 255.0f), m), sd);
    const float v10 = __fdiv_rn(__fsub_rn(__fdiv_rn((float)p10[sc], 255.0f), m), sd);
    const float v11 = __fdiv_rn(__fsub_rn(__fdiv_rn((float)p11[sc], 255.0f), m), sd);
    const float top = __fadd_rn(__fmul_rn(v00, wx0), __fmul_rn(v01, lx));
    const float bot = __fadd_rn(__fmul_rn(v10, wx0), __fmul_rn(v11, lx));
    out[(((long long)b * 3 + c) * H + y) * W + x] = __fadd_rn(__fmul_rn(top, wy0), __fmul_rn(bot, ly));
  }
}

}  // namespace mtt

extern "C" int mtt_preprocess_image(const uint8_t* img, int32_t B, int32_t h, int32_t w, int32_t bgr,
                                    const float* mean3, const float* std3, float* out, int32_t H, int32_t W,
                                    mtt_stream_t stream) {
  using namespace mtt;
  if (!img || !out || !mean3 || !std3 || B <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0)
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_preprocess_image: bad arguments (B=%d h=%d w=%d H=%d W=%d)", B, h, w, H, W);
  Norm3 nm;
  for (int c = 0; c < 3; ++c) {
    nm.mean[c] = mean3[c];
    nm.std[c] = std3[c];
    if (!(std3[c] > 0.f)) return set_error(MTT_ERR_BAD_SHAPE, "mtt_preprocess_image: std[%d] = %g", c, std3[c]);
  }
  const long long total = (long long)B * H * W;
  preprocess_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      img, B, h, w, bgr, nm, out, H, W, (double)h / (double)H, (double)w / (double)W);
  return check_launch("mtt_preprocess_image");
}
