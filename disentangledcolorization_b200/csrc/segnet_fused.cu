// Fused narrow layers of SpixelNet (reference models/network.py:264-267,293-295: conv0a -> conv0b -> conv1a, each
// conv(no bias) + BatchNorm + LeakyReLU(0.1) with the BatchNorm folded into weights and bias).
//
// Why.  At full resolution the 16-channel layers move 32 B per pixel per tensor and do 2.3 kMAC per pixel: launched one
// by one they are bound by memory traffic and by the TMA box-row rate of the tcgen05 path (r1: 0.07 + 0.16 + 0.20 ms at
// batch 64 against an HBM floor of 0.09 ms for the three).  Here one CTA takes a 32 x 16 pixel tile through all three
// layers: the L-channel tile (+3 pixel halo) is read once, conv0a runs on CUDA cores in fp32 (Cin = 1), its 16-channel
// output stays in shared memory (bf16, +1 halo), conv0b and the stride-2 conv1a run on tensor cores (mma.sync m16n8k16 bf16,
// fp32 accumulation) with A fragments gathered by ldmatrix at the tap's pixel offset (implicit GEMM, no im2col), and only
// the two tensors later layers read -- out1 (skip connection of conv0_1) and conv1a's output -- go to HBM, as coalesced
// 16-byte copies out of shared memory.  HBM traffic per pixel: 4 B in, 32 + 16 B out.
//
// Shared-memory tiles are [pixel][16 channels] bf16 = 32 B per pixel; the two 16-byte halves of pixel q are stored at
// half ^ ((q >> 2) & 1), so that the eight 16-byte rows an ldmatrix reads from eight consecutive pixels land in eight
// different bank groups.  An M tile of the implicit GEMM is 16 consecutive pixels of the flattened halo region (it may wrap
// around row ends: every ldmatrix lane computes its own pixel's address).
#include "common.cuh"
#include <cstring>

namespace {

constexpr int SG_THREADS = 256;
constexpr int TWF = 32, THF = 16;            // owned full-resolution tile
constexpr int R1W = TWF + 1, R1H = THF + 1;  // out1 region conv1a needs: rows/cols -1 .. +15/+31   (33 x 17)
constexpr int R0W = R1W + 2, R0H = R1H + 2;  // conv0a region conv0b needs                          (35 x 19)
constexpr int RGW = R0W + 2, RGH = R0H + 2;  // L-channel region                                    (37 x 21)
constexpr int P16 = 32;                      // pixel pitch of a 16-channel tile in bytes
// byte offset of 16-byte half `c` of pixel `q` in a swizzled 16-channel tile
__device__ __forceinline__ int sw16(int q, int c) { return q * P16 + (((c ^ (q >> 2)) & 1) << 4); }
constexpr int N1 = R1W * R1H;                // 561 out1 pixels
constexpr int N0 = R0W * R0H;                // 665 conv0a pixels
constexpr int MT1 = (N1 + 15) / 16;          // 36 M tiles of conv0b
constexpr int A1W = TWF / 2, A1H = THF / 2;  // conv1a tile at half resolution (16 x 8 = 8 M tiles)

struct SgParams {
  const float* gray;        // [B, H, W]
  const float* w0a;         // [9][16] fp32 (BatchNorm folded)
  const float* b0a;         // [16]
  const uint16_t* w0b;      // [9][16 co][16 ci] bf16
  const float* b0b;         // [16]
  const uint16_t* w1a;      // [9][32 co][16 ci] bf16
  const float* b1a;         // [32]
  __nv_bfloat16* out1;      // [B, H, W, 16]
  __nv_bfloat16* a1;        // [B, H/2, W/2, 32]
  int B, H, W;
  float slope;
};

struct SgSmem {
  __align__(16) float gray[(RGH * RGW + 3) / 4 * 4];
  __align__(16) float w0a[9 * 16 + 16];      // read as float4
  __align__(16) float bias[16 + 32];         // read as float2
  __align__(16) uint16_t w0b[9 * 16 * 16];
  __align__(16) uint16_t w1a[9 * 32 * 16];
  __align__(16) uint8_t t0a[N0 * P16];        // conv0a output (+1 halo), later reused to stage conv1a's output tile
  __align__(16) uint8_t t1[(MT1 * 16) * P16]; // out1 region
};

__device__ __forceinline__ uint32_t sg_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void sg_ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void sg_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float lrelu(float v, float s) { return fmaxf(v, v * s); }   // 0 <= s < 1

__global__ void __launch_bounds__(SG_THREADS, 3) segnet_head_kernel(const SgParams P) {
  extern __shared__ __align__(16) uint8_t sg_raw[];
  SgSmem& S = *reinterpret_cast<SgSmem*>(sg_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int H = P.H, W = P.W;
  const int x0 = blockIdx.x * TWF, y0 = blockIdx.y * THF, n = blockIdx.z;
  const float slope = P.slope;

  // ---- stage the L-channel tile (zero outside the image = the convolution's zero padding) and the parameters
  {
    const float* gimg = P.gray + (size_t)n * H * W;
    for (int i = tid; i < RGH * RGW; i += SG_THREADS) {
      const int ry = i / RGW, rx = i - ry * RGW;
      const int y = y0 - 3 + ry, x = x0 - 3 + rx;
      S.gray[i] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(gimg + (size_t)y * W + x) : 0.f;
    }
    for (int i = tid; i < 9 * 16; i += SG_THREADS) S.w0a[i] = P.w0a[i];
    if (tid < 16) { S.w0a[144 + tid] = P.b0a[tid]; S.bias[tid] = P.b0b[tid]; }
    if (tid < 32) S.bias[16 + tid] = P.b1a[tid];
    for (int i = tid; i < 9 * 16 * 16 / 8; i += SG_THREADS)
      reinterpret_cast<uint4*>(S.w0b)[i] = __ldg(reinterpret_cast<const uint4*>(P.w0b) + i);
    for (int i = tid; i < 9 * 32 * 16 / 8; i += SG_THREADS)
      reinterpret_cast<uint4*>(S.w1a)[i] = __ldg(reinterpret_cast<const uint4*>(P.w1a) + i);
  }
  __syncthreads();

  // ---- conv0a (1 -> 16, fp32 CUDA cores): one thread per pixel of the 35 x 19 region; bias first, taps in order, exactly
  //      like conv_c1_kernel; pixels outside the image are conv0b's zero padding
  for (int p = tid; p < N0; p += SG_THREADS) {
    const int ry = p / R0W, rx = p - ry * R0W;
    const int y = y0 - 2 + ry, x = x0 - 2 + rx;
    uint32_t o[8];
    if (y >= 0 && y < H && x >= 0 && x < W) {
      float v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = S.w0a[144 + c];
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        const float gv = S.gray[(ry + tp / 3) * RGW + rx + tp % 3];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w = *reinterpret_cast<const float4*>(&S.w0a[tp * 16 + 4 * q]);
          v[4 * q] = fmaf(gv, w.x, v[4 * q]); v[4 * q + 1] = fmaf(gv, w.y, v[4 * q + 1]);
          v[4 * q + 2] = fmaf(gv, w.z, v[4 * q + 2]); v[4 * q + 3] = fmaf(gv, w.w, v[4 * q + 3]);
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = pack_bf16(lrelu(v[2 * c], slope), lrelu(v[2 * c + 1], slope));
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = 0u;
    }
    *reinterpret_cast<uint4*>(S.t0a + sw16(p, 0)) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4*>(S.t0a + sw16(p, 1)) = make_uint4(o[4], o[5], o[6], o[7]);
  }
  __syncthreads();

  // ---- conv0b (16 -> 16) on tensor cores over the 33 x 17 region: weights as B fragments in registers (9 taps x 2 n-tiles)
  {
    uint32_t bw[9][2][2];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint16_t* wr = S.w0b + (tp * 16 + 8 * j + g) * 16;
        bw[tp][j][0] = *reinterpret_cast<const uint32_t*>(wr + 2 * t);
        bw[tp][j][1] = *reinterpret_cast<const uint32_t*>(wr + 8 + 2 * t);
      }
    const float2 bia0 = *reinterpret_cast<const float2*>(&S.bias[2 * t]);
    const float2 bia1 = *reinterpret_cast<const float2*>(&S.bias[8 + 2 * t]);
    const uint32_t t0a = sg_smem_u32(S.t0a);
    const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1), khalf = lane >> 4;
    for (int mt = warp; mt < MT1; mt += SG_THREADS / 32) {
      int p = mt * 16 + lrow;
      p = p < N1 ? p : N1 - 1;
      const int ry = p / R1W, rx = p - ry * R1W;                       // out1 region pixel of this lane's ldmatrix row
      const int qbase = ry * R0W + rx;                                  // tap (dy, dx) = conv0a pixel (ry + dy + 1, rx + dx + 1)
      float acc[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        uint32_t a[4];
        sg_ldsm_x4(a, t0a + (uint32_t)sw16(qbase + (tp / 3) * R0W + tp % 3, khalf));
        sg_mma(acc[0], a, bw[tp][0][0], bw[tp][0][1]);
        sg_mma(acc[1], a, bw[tp][1][0], bw[tp][1][1]);
      }
      // rows g and g+8 of the tile = region pixels mt*16+g, +8; zero outside the image (conv1a's padding)
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int q = mt * 16 + g + 8 * hrow;
        const int qy = q / R1W, qx = q - qy * R1W;
        const int y = y0 - 1 + qy, x = x0 - 1 + qx;
        const bool in = q < N1 && y >= 0 && y < H && x >= 0 && x < W;
        const uint32_t v0 = in ? pack_bf16(lrelu(acc[0][2 * hrow] + bia0.x, slope), lrelu(acc[0][2 * hrow + 1] + bia0.y, slope)) : 0u;
        const uint32_t v1 = in ? pack_bf16(lrelu(acc[1][2 * hrow] + bia1.x, slope), lrelu(acc[1][2 * hrow + 1] + bia1.y, slope)) : 0u;
        *reinterpret_cast<uint32_t*>(S.t1 + sw16(q, 0) + 4 * t) = v0;
        *reinterpret_cast<uint32_t*>(S.t1 + sw16(q, 1) + 4 * t) = v1;
      }
    }
  }
  __syncthreads();

  // ---- out1 -> HBM: the owned 32 x 16 pixels, 2 x 16 B per pixel, consecutive threads on consecutive bytes of a row
  for (int i = tid; i < TWF * THF * 2; i += SG_THREADS) {
    const int half = i & 1, px = (i >> 1) & (TWF - 1), py = i >> 6;
    const int y = y0 + py, x = x0 + px;
    if (y < H && x < W) {
      const uint4 v = *reinterpret_cast<const uint4*>(S.t1 + sw16((py + 1) * R1W + px + 1, half));
      *reinterpret_cast<uint4*>(P.out1 + (((size_t)n * H + y) * W + x) * 16 + half * 8) = v;
    }
  }

  // ---- conv1a (16 -> 32, stride 2): one M tile (16 half-resolution pixels of one row) per warp; weights from shared memory
  {
    const uint32_t t1 = sg_smem_u32(S.t1);
    const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1), khalf = lane >> 4;
    const int Y = warp;                                              // half-resolution row of the tile (A1H == 8 warps)
    // output pixel (Y, X = lrow) reads out1 region pixel (2Y + dy + 1, 2X + dx + 1), dy, dx in {-1, 0, 1}
    const int qbase = (2 * Y) * R1W + 2 * lrow;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      uint32_t a[4];
      sg_ldsm_x4(a, t1 + (uint32_t)sw16(qbase + (tp / 3) * R1W + tp % 3, khalf));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint16_t* wr = S.w1a + (tp * 32 + 8 * j + g) * 16;
        sg_mma(acc[j], a, *reinterpret_cast<const uint32_t*>(wr + 2 * t), *reinterpret_cast<const uint32_t*>(wr + 8 + 2 * t));
      }
    }
    // stage the 16 x 8 x 32-channel tile in the (dead) conv0a buffer: [pixel][64 B]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 bb = *reinterpret_cast<const float2*>(&S.bias[16 + 8 * j + 2 * t]);
      *reinterpret_cast<uint32_t*>(S.t0a + (Y * A1W + g) * 64 + 16 * j + 4 * t) =
          pack_bf16(lrelu(acc[j][0] + bb.x, slope), lrelu(acc[j][1] + bb.y, slope));
      *reinterpret_cast<uint32_t*>(S.t0a + (Y * A1W + g + 8) * 64 + 16 * j + 4 * t) =
          pack_bf16(lrelu(acc[j][2] + bb.x, slope), lrelu(acc[j][3] + bb.y, slope));
    }
  }
  __syncthreads();
  {
    const int H2 = H >> 1, W2 = W >> 1;
    for (int i = tid; i < A1W * A1H * 4; i += SG_THREADS) {
      const int piece = i & 3, px = (i >> 2) & (A1W - 1), py = i >> 6;
      const int y = (y0 >> 1) + py, x = (x0 >> 1) + px;
      if (y < H2 && x < W2) {
        const uint4 v = *reinterpret_cast<const uint4*>(S.t0a + (py * A1W + px) * 64 + piece * 16);
        *reinterpret_cast<uint4*>(P.a1 + (((size_t)n * H2 + y) * W2 + x) * 32 + piece * 8) = v;
      }
    }
  }
}

}  // namespace

extern "C" int disco_segnet_head(disco_handle* h, const float* gray, const float* w0a, const float* b0a, const uint16_t* w0b,
                                 const float* b0b, const uint16_t* w1a, const float* b1a, float slope, int batch, int H, int W,
                                 void* out1, void* a1, void* stream) {
  DISCO_CHECK_ARG(h && gray && w0a && b0a && w0b && b0b && w1a && b1a && out1 && a1, "segnet_head: null pointer");
  DISCO_CHECK_ARG(batch > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "segnet_head: H, W must be positive and even (got %dx%d)", H, W);
  DISCO_CHECK_ARG(slope >= 0.f && slope < 1.f, "segnet_head: LeakyReLU slope must be in [0, 1)");
  DISCO_CHECK_ARG(batch <= 65535 && (H + THF - 1) / THF <= 65535, "segnet_head: grid too large");
  DiscoDeviceGuard guard(h);
  SgParams P;
  P.gray = gray; P.w0a = w0a; P.b0a = b0a; P.w0b = w0b; P.b0b = b0b; P.w1a = w1a; P.b1a = b1a;
  P.out1 = reinterpret_cast<__nv_bfloat16*>(out1); P.a1 = reinterpret_cast<__nv_bfloat16*>(a1);
  P.B = batch; P.H = H; P.W = W; P.slope = slope;
  if (int rc = disco_ensure_smem(h, (const void*)segnet_head_kernel, (int)sizeof(SgSmem))) return rc;
  dim3 grid((W + TWF - 1) / TWF, (H + THF - 1) / THF, batch);
  segnet_head_kernel<<<grid, SG_THREADS, sizeof(SgSmem), (cudaStream_t)stream>>>(P);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}
