// mtt_attention: fused multi-head attention for the joint [prompts; patches] token sequence (head_dim 64).
//
// Replaces, per (batch, head), the eager sequence of TP/models/transformers/taskprompter.py:204-210
//   raw = q @ k^T ; attn = softmax(raw * scale) ; x = attn @ v
// (and IP/models/transformers/vit.py:189-193) without materialising the [B,H,N,N] maps, and emits
// the only part of `raw` the reference ever consumes: the un-scaled logits of the first T (prompt)
// query rows (taskprompter.py:436-437 spatial gates, :482 cross-task reweighting).
//
// This file is the C-ABI entry: descriptor validation and kernel selection.  The kernels:
//   attention5_tc.cu  (default)  warp-specialised: TMA warp, MMA warp, four softmax warps, 64-key blocks with two
//                                S buffers in TMEM, Q in TMEM (TS-form S = Q K^T), P written back in place over S,
//                                O accumulated in TMEM with lazy rescaling; two CTAs per SM
// Measured and dropped (profiles/README.md): a single-role kernel with 128-key blocks (tensor pipe 40 %), an
// SS-form variant (shared-memory bound at N = 64: 49 cycles per MMA against 33 for TS, scripts/mma_probe.cu) and a
// variant with eight softmax warps / two threads per row (20 % slower: pair barriers, 96-register cap).
#include <stdlib.h>

#include "host_common.h"

namespace mtt {
int launch_attention5(const mtt_attn_desc* d, int mode, cudaStream_t stream);  // attention5_tc.cu
static int g_attn_variant = -1;  // -1: read MTT_ATTN_VARIANT once; 0 = default
unsigned int* g_attn_trace = nullptr;  // mtt_set_attention_trace: clock-stamp buffer of the TRACE instantiation
}  // namespace mtt

extern "C" void mtt_set_attention_variant(int v) { mtt::g_attn_variant = v; }
extern "C" void mtt_set_attention_trace(void* buf) { mtt::g_attn_trace = static_cast<unsigned int*>(buf); }

extern "C" int mtt_attention(const mtt_attn_desc* d, mtt_stream_t stream_) {
  using namespace mtt;
  if (!d) return set_error(MTT_ERR_BAD_SHAPE, "mtt_attention: null descriptor");
  if (d->B <= 0 || d->N <= 0 || d->H <= 0 || d->T < 0 || d->T > d->N)
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_attention: B=%d N=%d H=%d T=%d", d->B, d->N, d->H, d->T);
  if (d->nsplit != 1 && d->nsplit != 2)
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_attention: nsplit=%d", d->nsplit);
  if (!d->qkv_hi || !d->out_hi || (d->nsplit == 2 && (!d->qkv_lo || !d->out_lo)))
    return set_error(MTT_ERR_BAD_SHAPE, "mtt_attention: missing plane");
  if (d->T > 128) return set_error(MTT_ERR_BAD_SHAPE, "mtt_attention: T=%d > 128 unsupported", d->T);
  if ((reinterpret_cast<uintptr_t>(d->out_hi) & 15) || (d->out_lo && (reinterpret_cast<uintptr_t>(d->out_lo) & 15)))
    return set_error(MTT_ERR_MISALIGNED, "mtt_attention: output not 16-byte aligned");
  if (g_attn_variant < 0) {
    const char* e = getenv("MTT_ATTN_VARIANT");
    g_attn_variant = e ? atoi(e) : 0;
  }
  ProfileScope prof(static_cast<cudaStream_t>(stream_), 1, 4.0 * d->B * d->H * (double)d->N * d->N * 64, d->B * d->N, d->N,
                    d->H * 64);
  // variant 5 = the round-1 softmax pass (rounded split, per-element maximum); 0 / 6 = the packed-math pass with
  // truncated P planes (default); 7 = packed-math pass with the lo plane of P rounded to nearest
  const int mode = g_attn_variant == 5 ? 0 : (g_attn_variant == 7 ? 2 : 1);
  return launch_attention5(d, mode, static_cast<cudaStream_t>(stream_));
}
