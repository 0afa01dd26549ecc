// Training-side losses of the hot path's outputs (BASELINE config 5).
//
//  ce_rebalance : the token-level term of AnchorColorProbLoss (reference models/loss.py:59-76): mean cross-entropy over the
//                 B*S super-pixel tokens of ONE 313-way logit map (pal_logit or ref_logit) against the hard labels
//                 (nearest gamut bin of the pooled ground-truth colour, models/basic.py:177-194 + train_colorizer.py:143),
//                 and its gradient with respect to the logits as autograd delivers it through basic.RebalanceLoss
//                 (models/basic.py:120-134): (softmax - onehot) / n_valid, multiplied by the token's weight (the class
//                 weight of its label, ColorLabel.get_classweights, models/basic.py:173-175).  label -1 = ignore_index.
//  encode_soft  : ColorLabel.encode_ab2ind (models/basic.py:177-194): 5 nearest gamut bins, Gaussian (sigma 5) weights,
//                 normalised -> soft 313-way code per token (the trainer takes its arg-max, train_colorizer.py:143).
//  spixel_recon : the two terms of SPixelLoss (models/loss.py:17-30) given the reconstruction upfeat(poolfeat(target)):
//                 mean over pixels of ||recon - target||_2 over the feature channels and over the last two (position)
//                 channels.
// Logits are [B, 313, S] exactly as the forward returns them (token index fastest): one thread per token walks the 313
// classes with coalesced loads; three passes (max, sum of exponentials, gradient), L2-resident.
#include "common.cuh"
#include <cfloat>

namespace {

constexpr int NV = 313;

__global__ void ce_token_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, const float* __restrict__ tw,
                                int B, int S, float* __restrict__ token_loss, float* __restrict__ dlogits) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * S) return;
  const int n = idx / S, t = idx - n * S;
  const float* col = logits + (size_t)n * NV * S + t;
  const int lab = labels[idx];
  float m = -FLT_MAX;
  for (int c = 0; c < NV; ++c) m = fmaxf(m, col[(size_t)c * S]);
  float s = 0.f;
  for (int c = 0; c < NV; ++c) s += expf(col[(size_t)c * S] - m);
  const bool valid = lab >= 0 && lab < NV;
  token_loss[idx] = valid ? (m + logf(s)) - col[(size_t)lab * S] : 0.f;
  if (dlogits) {
    float* g = dlogits + (size_t)n * NV * S + t;
    const float w = valid ? tw[idx] : 0.f, inv = 1.0f / s;
    for (int c = 0; c < NV; ++c) {
      const float p = expf(col[(size_t)c * S] - m) * inv;
      g[(size_t)c * S] = (p - (c == lab ? 1.f : 0.f)) * w;       // scaled by 1 / n_valid in ce_finish_kernel
    }
  }
}

// one block: deterministic tree reduction of the token losses and the valid count; then scales the gradient
__global__ void ce_reduce_kernel(const float* __restrict__ token_loss, const int32_t* __restrict__ labels, int M,
                                 float* __restrict__ out /* [0] mean loss, [1] n_valid */) {
  __shared__ float ssum[1024];
  __shared__ int scnt[1024];
  float s = 0.f;
  int c = 0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    s += token_loss[i];
    c += (labels[i] >= 0 && labels[i] < NV) ? 1 : 0;
  }
  ssum[threadIdx.x] = s;
  scnt[threadIdx.x] = c;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { ssum[threadIdx.x] += ssum[threadIdx.x + o]; scnt[threadIdx.x] += scnt[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[1] = (float)scnt[0];
    out[0] = scnt[0] > 0 ? ssum[0] / (float)scnt[0] : 0.f;
  }
}

__global__ void ce_scale_kernel(float* __restrict__ dlogits, size_t n, const float* __restrict__ out) {
  const float inv = out[1] > 0.f ? 1.0f / out[1] : 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dlogits[i] *= inv;
}

// ab: NCHW fp32 [B,2,S] normalised by 110; q: [B,313,S]
__global__ void encode_soft_kernel(const float* __restrict__ ab, const float* __restrict__ table, int B, int S, float* __restrict__ q) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * S) return;
  const int n = idx / S, t = idx - n * S;
  const float a0 = ab[((size_t)n * 2 + 0) * S + t] * 110.0f, a1 = ab[((size_t)n * 2 + 1) * S + t] * 110.0f;
  float bd[5];
  int bi[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) { bd[i] = FLT_MAX; bi[i] = 0; }
  for (int c = 0; c < NV; ++c) {
    const float d0 = table[2 * c] - a0, d1 = table[2 * c + 1] - a1;
    float d = d0 * d0 + d1 * d1;          // ordering by squared distance == ordering by distance
    int ci = c;
    if (d < bd[4]) {
#pragma unroll
      for (int i = 0; i < 5; ++i)
        if (d < bd[i]) { const float fd = bd[i]; const int fi = bi[i]; bd[i] = d; bi[i] = ci; d = fd; ci = fi; }
    }
  }
  float w[5], s = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i) { w[i] = expf(-bd[i] / 50.0f) / (2.0f * 3.14159265358979f * 5.0f); s += w[i]; }
  float* out = q + (size_t)n * NV * S + t;
  for (int c = 0; c < NV; ++c) out[(size_t)c * S] = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i) out[(size_t)bi[i] * S] = w[i] / s;
}

// recon, target: NCHW fp32 [B, C, H, W]; per pixel L2 norm over channels [0, C-2) and over [C-2, C); block partial sums
__global__ void spixel_recon_kernel(const float* __restrict__ recon, const float* __restrict__ target, int B, int C, size_t plane,
                                    float* __restrict__ partial /* [gridDim.x][2] */) {
  __shared__ float sf[256], sp[256];
  float f = 0.f, p = 0.f;
  const size_t total = (size_t)B * plane;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / plane, px = i - n * plane;
    float a = 0.f, b = 0.f;
    for (int c = 0; c < C; ++c) {
      const float d = recon[(n * C + c) * plane + px] - target[(n * C + c) * plane + px];
      if (c < C - 2) a = fmaf(d, d, a); else b = fmaf(d, d, b);
    }
    f += sqrtf(a);
    p += sqrtf(b);
  }
  sf[threadIdx.x] = f;
  sp[threadIdx.x] = p;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { sf[threadIdx.x] += sf[threadIdx.x + o]; sp[threadIdx.x] += sp[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = sf[0]; partial[2 * blockIdx.x + 1] = sp[0]; }
}

__global__ void spixel_finish_kernel(const float* __restrict__ partial, int nblocks, float n_pixels, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float f = 0.f, p = 0.f;
    for (int i = 0; i < nblocks; ++i) { f += partial[2 * i]; p += partial[2 * i + 1]; }
    out[0] = f / n_pixels;
    out[1] = p / n_pixels;
  }
}

}  // namespace

extern "C" int disco_encode_ab2ind(disco_handle* h, const float* ab, const float* q_to_ab, int batch, int S, float* q, void* stream) {
  DISCO_CHECK_ARG(h && ab && q_to_ab && q && batch > 0 && S > 0, "encode_ab2ind: bad argument");
  DiscoDeviceGuard guard(h);
  encode_soft_kernel<<<(batch * S + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ab, q_to_ab, batch, S, q);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}

extern "C" int disco_ce_rebalance(disco_handle* h, const float* logits, const int32_t* labels, const float* token_weights, int batch,
                                  int S, float* token_loss, float* loss_out, float* dlogits, void* stream) {
  DISCO_CHECK_ARG(h && logits && labels && token_weights && token_loss && loss_out, "ce_rebalance: null pointer");
  DISCO_CHECK_ARG(batch > 0 && S > 0, "ce_rebalance: bad shape");
  DiscoDeviceGuard guard(h);
  cudaStream_t st = (cudaStream_t)stream;
  const int M = batch * S;
  ce_token_kernel<<<(M + 127) / 128, 128, 0, st>>>(logits, labels, token_weights, batch, S, token_loss, dlogits);
  DISCO_LAUNCH_CHECK(h);
  ce_reduce_kernel<<<1, 1024, 0, st>>>(token_loss, labels, M, loss_out);
  DISCO_LAUNCH_CHECK(h);
  if (dlogits) {
    ce_scale_kernel<<<h->sm_count * 4, 256, 0, st>>>(dlogits, (size_t)M * NV, loss_out);
    DISCO_LAUNCH_CHECK(h);
  }
  return DISCO_OK;
}

extern "C" int disco_spixel_recon_loss(disco_handle* h, const float* recon, const float* target, int batch, int C, int H, int W,
                                       float* partial, int n_partial, float* loss_out, void* stream) {
  DISCO_CHECK_ARG(h && recon && target && partial && loss_out, "spixel_recon_loss: null pointer");
  DISCO_CHECK_ARG(batch > 0 && C > 2 && H > 0 && W > 0 && n_partial >= 1, "spixel_recon_loss: bad shape (needs C > 2)");
  DiscoDeviceGuard guard(h);
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = n_partial < h->sm_count * 4 ? n_partial : h->sm_count * 4;
  spixel_recon_kernel<<<nb, 256, 0, st>>>(recon, target, batch, C, (size_t)H * W, partial);
  DISCO_LAUNCH_CHECK(h);
  spixel_finish_kernel<<<1, 32, 0, st>>>(partial, nb, (float)((double)batch * H * W), loss_out);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}
