// Lab -> sRGB uint8 on the device: the last step of the inference CLI (reference main/colorizer/inference.py:119-127 ->
// utils/util.py:91-106: de-normalise L = (l + 1) * 50, ab = ab * 110, cv2.cvtColor(COLOR_LAB2RGB) on float32, * 255,
// astype(uint8)).  The formula below is OpenCV's float Lab2RGB (D65 white point, sRGB transfer curve, clip to [0, 1] before
// the curve, truncation to uint8), the same colour science as the reference's own torch lab2rgb (models/basic.py:395-475).
// Doing it here turns the per-image D2H copy from 12 B/pixel of fp32 Lab into 3 B/pixel of finished RGB and takes the
// conversion off the single Python thread.  HBM-bound: 12 B read + 3 B written per pixel.
#include "common.cuh"

namespace {

__device__ __forceinline__ float lab_finv(float f) {          // inverse of the CIE f() companding
  return f <= 0.20689655f /* 7.787 * 0.008856 + 16/116 */ ? (f - 16.0f / 116.0f) / 7.787f : f * f * f;
}
__device__ __forceinline__ float srgb_curve(float v) {
  v = fminf(fmaxf(v, 0.f), 1.f);
  if (v >= 1.f) return 1.f;          // 1.055f - 0.055f rounds to 0.99999994f, which would truncate to 254
  return v <= 0.0031308f ? v * 12.92f : 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
}

// gray [B,1,H,W], ab [B,2,H,W] (normalised as the forward returns them) -> rgb uint8 [B, Hc, Wc, 3] (top-left crop)
__global__ void lab2rgb_u8_kernel(const float* __restrict__ gray, const float* __restrict__ ab, int B, int H, int W, int Hc, int Wc,
                                  uint8_t* __restrict__ rgb) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * Hc * Wc;
  if (idx >= total) return;
  const int x = (int)(idx % Wc), y = (int)((idx / Wc) % Hc), n = (int)(idx / ((size_t)Wc * Hc));
  const size_t plane = (size_t)H * W, pix = (size_t)y * W + x;
  const float L = gray[(size_t)n * plane + pix] * 50.0f + 50.0f;
  const float a = ab[((size_t)n * 2 + 0) * plane + pix] * 110.0f;
  const float b = ab[((size_t)n * 2 + 1) * plane + pix] * 110.0f;
  float fy, Y;
  if (L <= 7.9996248f /* 0.008856 * 903.3 */) { Y = L / 903.3f; fy = 7.787f * Y + 16.0f / 116.0f; }
  else { fy = (L + 16.0f) / 116.0f; Y = fy * fy * fy; }
  const float X = lab_finv(a / 500.0f + fy) * 0.950456f;
  const float Z = lab_finv(fy - b / 200.0f) * 1.088754f;
  const float r = 3.240479f * X - 1.53715f * Y - 0.498535f * Z;
  const float g = -0.969256f * X + 1.875991f * Y + 0.041556f * Z;
  const float bl = 0.055648f * X - 0.204043f * Y + 1.057311f * Z;
  uint8_t* o = rgb + idx * 3;
  o[0] = (uint8_t)(srgb_curve(r) * 255.0f);
  o[1] = (uint8_t)(srgb_curve(g) * 255.0f);
  o[2] = (uint8_t)(srgb_curve(bl) * 255.0f);
}

}  // namespace

extern "C" int disco_lab2rgb_u8(disco_handle* h, const float* gray, const float* ab, int batch, int H, int W, int crop_h, int crop_w,
                                uint8_t* rgb, void* stream) {
  DISCO_CHECK_ARG(h && gray && ab && rgb, "lab2rgb_u8: null pointer");
  DISCO_CHECK_ARG(batch > 0 && H > 0 && W > 0 && crop_h > 0 && crop_h <= H && crop_w > 0 && crop_w <= W, "lab2rgb_u8: bad shape");
  DiscoDeviceGuard guard(h);
  const size_t total = (size_t)batch * crop_h * crop_w;
  lab2rgb_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gray, ab, batch, H, W, crop_h, crop_w, rgb);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}
