// Perceptual-loss side ops: everything AnchorColorProbLoss._perceptual_loss needs around the
// VGG19 convolutions (which run through disco_conv):
//   * disco_lab2rgb_norm  -- basic.lab2rgb (reference models/basic.py:431-475: lab2xyz -> xyz2rgb, Richard Zhang's
//                            formulas) fused with VGG19Loss.normalize (models/loss.py:198-203), written as NHWC activations
//                            padded to `cpad` channels for the first VGG layer, and/or as the plain fp32 NCHW RGB image
//   * disco_maxpool2      -- nn.MaxPool2d(2, 2) of torchvision's vgg19.features on NHWC activations
//   * disco_l1_mean       -- nn.L1Loss() (mean |x - y|) between two activation tensors, deterministic two-stage reduction
#include "common.cuh"

namespace {

// reference basic.lab2xyz + basic.xyz2rgb, fp32, one pixel
__device__ __forceinline__ void lab_to_rgb(float l_rs, float a_rs, float b_rs, float& r, float& g, float& b) {
  const float L = l_rs * 50.0f + 50.0f, A = a_rs * 110.0f, Bc = b_rs * 110.0f;
  const float y_int = (L + 16.0f) / 116.0f;
  const float x_int = A / 500.0f + y_int;
  const float z_int = fmaxf(0.0f, y_int - Bc / 200.0f);
  float xyz[3] = {x_int, y_int, z_int};
  const float sc[3] = {0.95047f, 1.0f, 1.08883f};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float v = xyz[i];
    xyz[i] = (v > 0.2068966f ? v * v * v : (v - 16.0f / 116.0f) / 7.787f) * sc[i];
  }
  float rgb[3];
  rgb[0] = 3.24048134f * xyz[0] - 1.53715152f * xyz[1] - 0.49853633f * xyz[2];
  rgb[1] = -0.96925495f * xyz[0] + 1.87599f * xyz[1] + 0.04155593f * xyz[2];
  rgb[2] = 0.05564664f * xyz[0] - 0.20404134f * xyz[1] + 1.05731107f * xyz[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float v = fmaxf(rgb[i], 0.0f);
    rgb[i] = v > 0.0031308f ? 1.055f * powf(v, 1.0f / 2.4f) - 0.055f : 12.92f * v;
  }
  r = rgb[0]; g = rgb[1]; b = rgb[2];
}

// ab == nullptr: `gray` is an RGB image (B,3,H,W) and only the normalisation / packing runs
template <typename T>
__global__ void lab2rgb_norm_kernel(const float* __restrict__ gray, const float* __restrict__ ab, int B, int H, int W,
                                    float* __restrict__ rgb_nchw, T* __restrict__ norm_nhwc, int cpad, float m0, float m1, float m2,
                                    float s0, float s1, float s2) {
  const size_t plane = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * plane) return;
  const size_t n = idx / plane, p = idx - n * plane;
  float r, g, b;
  if (ab) {
    lab_to_rgb(gray[idx], ab[(n * 2) * plane + p], ab[(n * 2 + 1) * plane + p], r, g, b);
  } else {
    r = gray[(n * 3) * plane + p]; g = gray[(n * 3 + 1) * plane + p]; b = gray[(n * 3 + 2) * plane + p];
  }
  if (rgb_nchw) {
    rgb_nchw[(n * 3) * plane + p] = r;
    rgb_nchw[(n * 3 + 1) * plane + p] = g;
    rgb_nchw[(n * 3 + 2) * plane + p] = b;
  }
  if (norm_nhwc) {
    T* o = norm_nhwc + idx * cpad;
    o[0] = from_f32<T>((r - m0) / s0);
    o[1] = from_f32<T>((g - m1) / s1);
    o[2] = from_f32<T>((b - m2) / s2);
    for (int c = 3; c < cpad; ++c) o[c] = from_f32<T>(0.0f);
  }
}

// 8 channels (bf16: one 16-byte vector) or 4 channels (fp32) per thread
template <typename T, int V>
__global__ void maxpool2_kernel(const T* __restrict__ x, int B, int H, int W, int C, T* __restrict__ y) {
  const int Ho = H >> 1, Wo = W >> 1, CV = C / V;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * Ho * Wo * CV) return;
  const int cv = (int)(idx % CV);
  size_t r = idx / CV;
  const int ox = (int)(r % Wo);
  r /= Wo;
  const int oy = (int)(r % Ho), n = (int)(r / Ho);
  const T* base = x + (((size_t)n * H + 2 * oy) * W + 2 * ox) * C + cv * V;
  float m[V];
#pragma unroll
  for (int v = 0; v < V; ++v) m[v] = -3.4e38f;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      T tmp[V];
      *reinterpret_cast<uint4*>(tmp) = *reinterpret_cast<const uint4*>(base + ((size_t)dy * W + dx) * C);
#pragma unroll
      for (int v = 0; v < V; ++v) m[v] = fmaxf(m[v], to_f32(tmp[v]));
    }
  T out[V];
#pragma unroll
  for (int v = 0; v < V; ++v) out[v] = from_f32<T>(m[v]);
  *reinterpret_cast<uint4*>(y + (((size_t)n * Ho + oy) * Wo + ox) * C + cv * V) = *reinterpret_cast<const uint4*>(out);
}

constexpr int L1_BLOCK = 256;

template <typename T>
__global__ void l1_partial_kernel(const T* __restrict__ x, const T* __restrict__ y, size_t n, float* __restrict__ partial) {
  __shared__ float red[L1_BLOCK];
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * L1_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * L1_BLOCK)
    s += fabsf(to_f32(x[i]) - to_f32(y[i]));
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = L1_BLOCK / 2; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void l1_final_kernel(const float* __restrict__ partial, int nb, double inv_n, float weight, float* __restrict__ out, int accumulate) {
  __shared__ double red[L1_BLOCK];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += L1_BLOCK) s += (double)partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = L1_BLOCK / 2; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float v = weight * (float)(red[0] * inv_n);
    out[0] = accumulate ? out[0] + v : v;
  }
}

// AnchorColorProbLoss._laplace_gradient (models/loss.py:51-57): depthwise 3x3 Laplacian [[1,1,1],[1,-8,1],[1,1,1]] without
// padding of prediction and target, L1 mean of the difference.  The Laplacian is linear, so one pass over d = target - pred.
__device__ __forceinline__ float lap_at(const float* __restrict__ t, const float* __restrict__ p, int W, size_t c) {
  float s = 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const size_t i = c + (ptrdiff_t)dy * W + dx;
      s += (dy == 0 && dx == 0 ? -8.0f : 1.0f) * (t[i] - p[i]);
    }
  return s;
}

__global__ void laplace_l1_kernel(const float* __restrict__ pred, const float* __restrict__ target, int planes, int H, int W,
                                  float* __restrict__ sign_map, float* __restrict__ partial) {
  __shared__ float red[L1_BLOCK];
  const int Ho = H - 2, Wo = W - 2;
  const size_t n = (size_t)planes * Ho * Wo;
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * L1_BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * L1_BLOCK) {
    const int ox = (int)(i % Wo);
    const size_t r = i / Wo;
    const int oy = (int)(r % Ho);
    const size_t pl = r / Ho;
    const float v = lap_at(target, pred, W, (pl * H + oy + 1) * W + ox + 1);
    s += fabsf(v);
    if (sign_map) sign_map[i] = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = L1_BLOCK / 2; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// d loss / d pred[y][x] = -(1/n) * sum over the outputs (oy, ox) whose window holds (y, x) of k[y-oy][x-ox] * sign(oy, ox)
__global__ void laplace_l1_grad_kernel(const float* __restrict__ sign_map, int planes, int H, int W, float inv_n, float* __restrict__ grad) {
  const int Ho = H - 2, Wo = W - 2;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)planes * H * W) return;
  const int x = (int)(idx % W);
  const size_t r = idx / W;
  const int y = (int)(r % H);
  const size_t pl = r / H;
  float s = 0.f;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const int oy = y - 1 - dy, ox = x - 1 - dx;          // output whose tap (dy, dx) reads (y, x)
      if (oy >= 0 && oy < Ho && ox >= 0 && ox < Wo) s += (dy == 0 && dx == 0 ? -8.0f : 1.0f) * sign_map[(pl * Ho + oy) * Wo + ox];
    }
  grad[idx] = -s * inv_n;
}

}  // namespace

extern "C" int disco_laplace_l1(disco_handle* h, const float* pred, const float* target, int batch, int C, int H, int W,
                                float* sign_scratch, float* partial, int n_partial, float* out, float* grad_pred, void* stream) {
  DISCO_CHECK_ARG(h && pred && target && partial && out, "laplace_l1: null pointer");
  DISCO_CHECK_ARG(batch > 0 && C > 0 && H >= 3 && W >= 3 && n_partial > 0, "laplace_l1: needs H, W >= 3 (got %dx%d)", H, W);
  DISCO_CHECK_ARG(!grad_pred || sign_scratch, "laplace_l1: the gradient needs the sign scratch buffer");
  DiscoDeviceGuard guard(h);
  const long long n = (long long)batch * C * (H - 2) * (W - 2);
  const long long want = (n + L1_BLOCK - 1) / L1_BLOCK;
  const int nb = (int)(want < n_partial ? want : n_partial);
  laplace_l1_kernel<<<nb, L1_BLOCK, 0, (cudaStream_t)stream>>>(pred, target, batch * C, H, W, sign_scratch, partial);
  DISCO_LAUNCH_CHECK(h);
  l1_final_kernel<<<1, L1_BLOCK, 0, (cudaStream_t)stream>>>(partial, nb, 1.0 / (double)n, 1.0f, out, 0);
  DISCO_LAUNCH_CHECK(h);
  if (grad_pred) {
    const size_t m = (size_t)batch * C * H * W;
    laplace_l1_grad_kernel<<<(unsigned)((m + 255) / 256), 256, 0, (cudaStream_t)stream>>>(sign_scratch, batch * C, H, W, (float)(1.0 / (double)n), grad_pred);
    DISCO_LAUNCH_CHECK(h);
  }
  return DISCO_OK;
}

static int lab2rgb_norm_impl(disco_handle* h, const float* gray, const float* ab, int batch, int H, int W, float* rgb_nchw,
                             void* norm_nhwc, int dtype, int cpad, const float* mean3, const float* std3, void* stream);

extern "C" int disco_lab2rgb_norm(disco_handle* h, const float* gray, const float* ab, int batch, int H, int W, float* rgb_nchw,
                                  void* norm_nhwc, int dtype, int cpad, const float* mean3, const float* std3, void* stream) {
  DISCO_CHECK_ARG(ab != nullptr, "lab2rgb_norm: null pointer");
  return lab2rgb_norm_impl(h, gray, ab, batch, H, W, rgb_nchw, norm_nhwc, dtype, cpad, mean3, std3, stream);
}

extern "C" int disco_rgb_norm(disco_handle* h, const float* rgb, int batch, int H, int W, void* norm_nhwc, int dtype, int cpad,
                              const float* mean3, const float* std3, void* stream) {
  DISCO_CHECK_ARG(norm_nhwc != nullptr, "rgb_norm: null pointer");
  return lab2rgb_norm_impl(h, rgb, nullptr, batch, H, W, nullptr, norm_nhwc, dtype, cpad, mean3, std3, stream);
}

static int lab2rgb_norm_impl(disco_handle* h, const float* gray, const float* ab, int batch, int H, int W, float* rgb_nchw,
                             void* norm_nhwc, int dtype, int cpad, const float* mean3, const float* std3, void* stream) {
  DISCO_CHECK_ARG(h && gray && (rgb_nchw || norm_nhwc), "lab2rgb_norm: null pointer");
  DISCO_CHECK_ARG(batch > 0 && H > 0 && W > 0, "lab2rgb_norm: bad shape");
  DISCO_CHECK_ARG(!norm_nhwc || (cpad >= 3 && mean3 && std3 && (dtype == DISCO_F32 || dtype == DISCO_BF16)), "lab2rgb_norm: bad packing arguments");
  DiscoDeviceGuard guard(h);
  const size_t n = (size_t)batch * H * W;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  const float m0 = mean3 ? mean3[0] : 0.f, m1 = mean3 ? mean3[1] : 0.f, m2 = mean3 ? mean3[2] : 0.f;
  const float s0 = std3 ? std3[0] : 1.f, s1 = std3 ? std3[1] : 1.f, s2 = std3 ? std3[2] : 1.f;
  if (dtype == DISCO_BF16)
    lab2rgb_norm_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(gray, ab, batch, H, W, rgb_nchw,
                                                                                reinterpret_cast<__nv_bfloat16*>(norm_nhwc), cpad, m0, m1, m2, s0, s1, s2);
  else
    lab2rgb_norm_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(gray, ab, batch, H, W, rgb_nchw, reinterpret_cast<float*>(norm_nhwc),
                                                                        cpad, m0, m1, m2, s0, s1, s2);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}

extern "C" int disco_maxpool2(disco_handle* h, int dtype, const void* x, int batch, int H, int W, int C, void* y, void* stream) {
  DISCO_CHECK_ARG(h && x && y, "maxpool2: null pointer");
  DISCO_CHECK_ARG(batch > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "maxpool2: H, W must be even (got %dx%d)", H, W);
  DISCO_CHECK_ARG(dtype == DISCO_F32 || dtype == DISCO_BF16, "maxpool2: dtype");
  const int V = dtype == DISCO_BF16 ? 8 : 4;
  DISCO_CHECK_ARG(C > 0 && C % V == 0, "maxpool2: C must be a multiple of %d (got %d)", V, C);
  DiscoDeviceGuard guard(h);
  const size_t n = (size_t)batch * (H / 2) * (W / 2) * (C / V);
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (dtype == DISCO_BF16)
    maxpool2_kernel<__nv_bfloat16, 8><<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), batch, H, W, C,
                                                                               reinterpret_cast<__nv_bfloat16*>(y));
  else
    maxpool2_kernel<float, 4><<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(x), batch, H, W, C,
                                                                       reinterpret_cast<float*>(y));
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}

extern "C" int disco_l1_mean(disco_handle* h, int dtype, const void* x, const void* y, long long n, float weight, int accumulate,
                             float* partial, int n_partial, float* out, void* stream) {
  DISCO_CHECK_ARG(h && x && y && partial && out, "l1_mean: null pointer");
  DISCO_CHECK_ARG(n > 0 && n_partial > 0, "l1_mean: empty input");
  DISCO_CHECK_ARG(dtype == DISCO_F32 || dtype == DISCO_BF16, "l1_mean: dtype");
  DiscoDeviceGuard guard(h);
  const long long want = (n + L1_BLOCK - 1) / L1_BLOCK;
  const int nb = (int)(want < n_partial ? want : n_partial);
  if (dtype == DISCO_BF16)
    l1_partial_kernel<__nv_bfloat16><<<nb, L1_BLOCK, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                               reinterpret_cast<const __nv_bfloat16*>(y), (size_t)n, partial);
  else
    l1_partial_kernel<float><<<nb, L1_BLOCK, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(y),
                                                                       (size_t)n, partial);
  DISCO_LAUNCH_CHECK(h);
  l1_final_kernel<<<1, L1_BLOCK, 0, (cudaStream_t)stream>>>(partial, nb, 1.0 / (double)n, weight, out, accumulate);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}
