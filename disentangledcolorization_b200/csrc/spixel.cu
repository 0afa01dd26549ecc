// Super-pixel soft pooling / un-pooling (HBM-bound glue of the forward).
//
//   poolfeat_partial : one CTA per 16x16 cell; one pass over the cell's feature pixels produces, for
//                      each of the 9 neighbour directions k, the mass the cell emits:
//                      sum_pix prob_k * [feat(64), ab(2), 1]  and the hard-assignment count.
//   poolfeat_gather  : each cell collects the 9 masses addressed to it (k ascending, like the
//                      reference's running sum), normalises, and writes tokens / colours / sizes.
//   upfeat           : pixel = sum_k prob_k * token(neighbour k), tokens of the 3x3 neighbourhood in smem.
// Reference: models/basic.py:274-376.  All sums are fp32.
#include "common.cuh"

namespace {

constexpr int SP = 16;            // super-pixel cell size (reference --psize default, all BASELINE configs)
constexpr int PART = 68;          // 64 feats + 2 ab + prob mass + hard mass

template <typename T>
__device__ __forceinline__ void ld4(const T* p, float o[4]);
template <>
__device__ __forceinline__ void ld4<float>(const float* p, float o[4]) {
  float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
__device__ __forceinline__ void ld4<__nv_bfloat16>(const __nv_bfloat16* p, float o[4]) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}
template <typename T>
__device__ __forceinline__ void st4(T* p, const float o[4]);
template <>
__device__ __forceinline__ void st4<float>(float* p, const float o[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, const float o[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(o[2], o[3]);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}

// grid (w, h, B), 288 threads: warps 0..7 = (pixel column x, 4-channel group) over the cell's 16 rows, warp 8 = the four
// scalar masses (ab0, ab1, soft mass, hard mass).  feats NHWC C=64.
// r2: the affinity of the cell is staged pixel-major ([pixel][12]) so that a thread fetches all nine directions of a
// pixel with three LDS.128 (one wavefront each) instead of nine scalar LDS -- the r1 kernel ran at the shared-memory
// pipe's limit (1 LDS per 4 FMA) -- and the 36 scalar masses, which cost 180 warp shuffles per thread, now ride on a
// ninth warp with the same row loop.
constexpr int PKP = 12;           // floats per pixel in the staged affinity (9 used)
template <typename T>
__global__ void __launch_bounds__(288) poolfeat_partial_kernel(const T* __restrict__ feats, const float* __restrict__ ab,
                                                               const float* __restrict__ aff, int H, int W,
                                                               float* __restrict__ partial) {
  // phase 1-2: Pk[256][12]; phase 3 (after a barrier): Red[16][9*64] + RedX[32][36] over the same bytes
  __shared__ __align__(16) float smem[16 * 9 * 64 + 32 * 36];
  float* Pk = smem;
  float* Red = smem;
  float* RedX = smem + 16 * 9 * 64;
  const int cx = blockIdx.x, cy = blockIdx.y, n = blockIdx.z;
  const int tid = threadIdx.x;
  const size_t plane = (size_t)H * W;
  const int h = H / SP, w = W / SP;

  if (tid < 256) {                 // thread = pixel (row-major in the cell)
    const int py = tid >> 4, px = tid & 15;
    const size_t pix = (size_t)(cy * SP + py) * W + cx * SP + px;
    float p[PKP];
#pragma unroll
    for (int k = 0; k < 9; ++k) p[k] = aff[((size_t)n * 9 + k) * plane + pix];
    p[9] = p[10] = p[11] = 0.f;
#pragma unroll
    for (int q = 0; q < 3; ++q)
      *reinterpret_cast<float4*>(Pk + tid * PKP + 4 * q) = make_float4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
  }
  __syncthreads();

  float acc[9][4];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
  if (tid < 256) {
    const int x = tid >> 4, cg = tid & 15;
    const T* base = feats + (((size_t)n * H + cy * SP) * W + cx * SP + x) * 64 + cg * 4;
#pragma unroll 4
    for (int y = 0; y < SP; ++y) {
      float f[4];
      ld4<T>(base + (size_t)y * W * 64, f);
      const float4 p0 = *reinterpret_cast<const float4*>(Pk + (y * SP + x) * PKP);
      const float4 p1 = *reinterpret_cast<const float4*>(Pk + (y * SP + x) * PKP + 4);
      const float p8 = Pk[(y * SP + x) * PKP + 8];
      const float p[9] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p8};
#pragma unroll
      for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[k][j] = fmaf(p[k], f[j], acc[k][j]);
    }
  } else {
    // scalar masses: lane = (column x, half of the rows); acc[k] = (sum p_k ab0, sum p_k ab1, sum p_k, #pixels whose max is p_k)
    const int lane = tid - 256, x = lane & 15, yh = lane >> 4;
#pragma unroll 2
    for (int yy = 0; yy < SP / 2; ++yy) {
      const int y = yh * (SP / 2) + yy;
      const size_t pix = (size_t)(cy * SP + y) * W + cx * SP + x;
      const float a0 = ab ? ab[((size_t)n * 2 + 0) * plane + pix] : 0.f;
      const float a1 = ab ? ab[((size_t)n * 2 + 1) * plane + pix] : 0.f;
      const float4 p0 = *reinterpret_cast<const float4*>(Pk + (y * SP + x) * PKP);
      const float4 p1 = *reinterpret_cast<const float4*>(Pk + (y * SP + x) * PKP + 4);
      const float p8 = Pk[(y * SP + x) * PKP + 8];
      const float p[9] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p8};
      float m = p[0];
#pragma unroll
      for (int k = 1; k < 9; ++k) m = fmaxf(m, p[k]);
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        acc[k][0] = fmaf(p[k], a0, acc[k][0]);
        acc[k][1] = fmaf(p[k], a1, acc[k][1]);
        acc[k][2] += p[k];
        acc[k][3] += (p[k] == m) ? 1.f : 0.f;
      }
    }
  }
  __syncthreads();                 // every thread is done with Pk: the reduction buffers reuse its bytes
  if (tid < 256) {
    const int x = tid >> 4, cg = tid & 15;
#pragma unroll
    for (int k = 0; k < 9; ++k)
      *reinterpret_cast<float4*>(Red + x * (9 * 64) + k * 64 + cg * 4) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
  } else {
    const int lane = tid - 256;
#pragma unroll
    for (int k = 0; k < 9; ++k)
      *reinterpret_cast<float4*>(RedX + lane * 36 + k * 4) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
  }
  __syncthreads();
  float* out = partial + (((size_t)n * h + cy) * w + cx) * 9 * PART;
  const float inv = 1.0f / (SP * SP);   // avg_pool2d
  for (int e = tid; e < 9 * 64; e += 288) {
    float s = 0.f;
#pragma unroll
    for (int x = 0; x < 16; ++x) s += Red[x * (9 * 64) + e];
    out[(e >> 6) * PART + (e & 63)] = s * inv;
  }
  if (tid < 36) {
    // same order of additions as a sum over the cell's pixels grouped by (half, column): deterministic
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < 32; ++l) s += RedX[l * 36 + tid];
    out[(tid >> 2) * PART + 64 + (tid & 3)] = s * inv;
  }
}

// grid (B*S), 64 threads: thread c = feature channel; threads 0..3 also do (ab0, ab1, mass, hard)
__global__ void poolfeat_gather_kernel(const float* __restrict__ partial, int h, int w, float* __restrict__ tokens,
                                       float* __restrict__ spix_ab, float* __restrict__ conf,
                                       float* __restrict__ sizes) {
  const int S = h * w;
  const int n = blockIdx.x / S, cell = blockIdx.x % S;
  const int cy = cell / w, cx = cell % w;
  const int c = threadIdx.x;
  float fsum = 0.f, mass = 0.f, extra = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int ky = k / 3 - 1, kx = k % 3 - 1;
    const int ey = cy - ky, ex = cx - kx;      // emitting cell
    if (ey < 0 || ey >= h || ex < 0 || ex >= w) continue;
    const float* p = partial + ((((size_t)n * h + ey) * w + ex) * 9 + k) * PART;
    fsum += p[c];
    mass += p[66];
    if (c < 4) extra += p[64 + c];
  }
  tokens[((size_t)n * S + cell) * 64 + c] = fsum / (mass + 1e-8f);
  if (c < 2) spix_ab[((size_t)n * 2 + c) * S + cell] = extra / (mass + 1e-8f);
  if (c == 2) conf[(size_t)n * S + cell] = extra;
  if (c == 3) sizes[(size_t)n * S + cell] = extra;
}

// grid (w, h, B), 256 threads
template <typename T>
__global__ void __launch_bounds__(256) upfeat_kernel(const float* __restrict__ tokens, const float* __restrict__ aff,
                                                     int H, int W, T* __restrict__ out) {
  __shared__ __align__(16) float Pk[SP * SP * PKP];   // pixel-major: three LDS.128 fetch a pixel's nine directions
  __shared__ float Tk[9][64];
  const int cx = blockIdx.x, cy = blockIdx.y, n = blockIdx.z;
  const int tid = threadIdx.x;
  const int h = H / SP, w = W / SP;
  const size_t plane = (size_t)H * W;
  {
    const int py = tid >> 4, px = tid & 15;
    const size_t pix = (size_t)(cy * SP + py) * W + cx * SP + px;
    float p[PKP];
#pragma unroll
    for (int k = 0; k < 9; ++k) p[k] = aff[((size_t)n * 9 + k) * plane + pix];
    p[9] = p[10] = p[11] = 0.f;
#pragma unroll
    for (int q = 0; q < 3; ++q)
      *reinterpret_cast<float4*>(Pk + tid * PKP + 4 * q) = make_float4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
  }
  for (int e = tid; e < 9 * 64; e += 256) {
    const int k = e >> 6, c = e & 63;
    const int ny = cy + k / 3 - 1, nx = cx + k % 3 - 1;
    Tk[k][c] = (ny >= 0 && ny < h && nx >= 0 && nx < w) ? tokens[(((size_t)n * h + ny) * w + nx) * 64 + c] : 0.f;
  }
  __syncthreads();
  const int x = tid >> 4, cg = tid & 15;
  f32x2 t2[9][2];                    // the 3x3 neighbourhood's tokens, this thread's 4 channels, as packed fp32 pairs (FFMA2)
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    t2[k][0] = pack2(Tk[k][cg * 4 + 0], Tk[k][cg * 4 + 1]);
    t2[k][1] = pack2(Tk[k][cg * 4 + 2], Tk[k][cg * 4 + 3]);
  }
  T* base = out + (((size_t)n * H + cy * SP) * W + cx * SP + x) * 64 + cg * 4;
#pragma unroll 4
  for (int y = 0; y < SP; ++y) {
    f32x2 o01 = pack2(0.f, 0.f), o23 = pack2(0.f, 0.f);
    const float4 p0 = *reinterpret_cast<const float4*>(Pk + (y * SP + x) * PKP);
    const float4 p1 = *reinterpret_cast<const float4*>(Pk + (y * SP + x) * PKP + 4);
    const float p8 = Pk[(y * SP + x) * PKP + 8];
    const float pv[9] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p8};
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const f32x2 pp = pack2(pv[k], pv[k]);
      o01 = ffma2(t2[k][0], pp, o01);
      o23 = ffma2(t2[k][1], pp, o23);
    }
    float o[4];
    unpack2(o01, o[0], o[1]);
    unpack2(o23, o[2], o[3]);
    st4<T>(base + (size_t)y * W * 64, o);
  }
}

}  // namespace

extern "C" int disco_poolfeat(disco_handle* h, int dtype, const void* feats, const float* ab, const float* affinity,
                              int batch, int H, int W, int C, float* partial, float* tokens, float* spix_ab,
                              float* conf, float* sizes, void* stream) {
  DISCO_CHECK_ARG(h && feats && affinity && partial && tokens && spix_ab && conf && sizes, "poolfeat: null pointer");
  DISCO_CHECK_ARG(C == 64, "poolfeat: C must be 64 (got %d)", C);
  DISCO_CHECK_ARG(H > 0 && W > 0 && H % SP == 0 && W % SP == 0, "poolfeat: H, W must be multiples of 16 (got %dx%d)", H, W);
  DiscoDeviceGuard guard(h);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(W / SP, H / SP, batch);
  if (dtype == DISCO_F32)
    poolfeat_partial_kernel<float><<<grid, 288, 0, st>>>((const float*)feats, ab, affinity, H, W, partial);
  else
    poolfeat_partial_kernel<__nv_bfloat16><<<grid, 288, 0, st>>>((const __nv_bfloat16*)feats, ab, affinity, H, W, partial);
  DISCO_LAUNCH_CHECK(h);
  poolfeat_gather_kernel<<<batch * (H / SP) * (W / SP), 64, 0, st>>>(partial, H / SP, W / SP, tokens, spix_ab, conf, sizes);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}

extern "C" int disco_upfeat(disco_handle* h, int dtype, const float* tokens, const float* affinity, int batch, int H,
                            int W, int C, void* out, void* stream) {
  DISCO_CHECK_ARG(h && tokens && affinity && out, "upfeat: null pointer");
  DISCO_CHECK_ARG(C == 64, "upfeat: C must be 64 (got %d)", C);
  DISCO_CHECK_ARG(H > 0 && W > 0 && H % SP == 0 && W % SP == 0, "upfeat: H, W must be multiples of 16 (got %dx%d)", H, W);
  DiscoDeviceGuard guard(h);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(W / SP, H / SP, batch);
  if (dtype == DISCO_F32)
    upfeat_kernel<float><<<grid, 256, 0, st>>>(tokens, affinity, H, W, (float*)out);
  else
    upfeat_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(tokens, affinity, H, W, (__nv_bfloat16*)out);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}
