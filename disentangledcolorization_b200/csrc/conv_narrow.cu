// The 16-channel full-resolution layers of SpixelNet's decoder tail on mma.sync (reference models/network.py:278-282,
// 307-313: deconv0 -> conv0_1 -> pred_mask0 + Softmax(1)).
//
// Why not the tcgen05 kernels of conv_tc.cu.  With 16 channels a pixel is a 32-byte row: the resident-weight tcgen05
// kernel is paced by the TMA box-row rate on such rows (r2 per-op table: deconv0 0.127, conv0_1 0.185, pred_mask0 0.138 ms at
// batch 64), and an N = 16 MMA is capped by the shared-memory read of its A operand on either tensor path (1024
// MAC/clk/SM).  Here a persistent CTA stages the 18 x 34 pixel halo tile of each source with 16-byte cp.async
// copies (no TMA), runs the implicit GEMM with mma.sync.m16n8k16 (bf16 in, fp32 accumulate; A fragments by ldmatrix at the
// tap's pixel offset, B fragments = weights pre-packed by the host in fragment order), applies bias + activation (or the
// 9-way softmax) in registers and writes the tile through shared memory as full 16-byte / 128-byte rows.
// Measured (batch 64, 256 x 256): deconv0 0.088, conv0_1 0.158, pred_mask0 0.127-0.133 ms; ncu: tensor pipe 44 % / 29 %
// active, issue slots 46-52 %, shared-memory wavefronts halved by the row-sliding accumulation with no change in time --
// what is left is instruction issue and fixed-latency dependencies at 16 warps per SM (125 registers: the weights live in
// registers).  The 32-channel half-resolution layers (conv1b, conv1_1) were tried on this path and are faster on the
// tcgen05 resident kernel (0.058 / 0.082 against 0.067 / 0.12 ms): they are not claimed.
//
// Routed from disco_conv like the tcgen05 kernels: conv_narrow_match() claims a descriptor, disco_conv_tc_weight_elems /
// _pack_weights produce this file's packing for it, conv_narrow_launch() runs it.  DISCO_NARROW=0 disables the route.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int NC_THREADS = 256;
constexpr int NC_TW = 32;                    // output tile width (two 16-pixel M tiles per row)
constexpr int NC_PLANE = NC_TW * 16 + 4;     // fp32 plane pitch of the softmax staging (banks: 2t planes apart by 8)

struct NcParams {
  const __nv_bfloat16* src0;
  const __nv_bfloat16* src1;
  const uint32_t* wfrag;    // [tap][k-step][n-tile][lane][2] bf16 pairs (deconv: [phase][tap][k-step][n-tile][lane][2])
  const float* bias;
  void* out;
  int B, H, W;              // OUTPUT dims (deconv: the sources are H/2 x W/2)
  int act;
  float slope;
  int tiles_x, tiles_y, tiles;   // persistent CTAs walk tiles blockIdx.x, + gridDim.x, ... (x fastest, then y, then image)
};

__device__ __forceinline__ uint32_t nc_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void nc_ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void nc_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void nc_cp16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g));
}
__device__ __forceinline__ void nc_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint32_t nc_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float nc_act(float v, int act, float slope) { return act == DISCO_ACT_NONE ? v : fmaxf(v, v * slope); }

// Source tiles: [pixel][C channels] bf16 with the pixel pitch padded by 16 B (an odd number of 16-byte units), so that the
// eight rows of an ldmatrix (eight consecutive pixels, same chunk) hit eight bank groups AND a tap is a constant byte
// offset from the lane's base address (an XOR swizzle would cost an address computation per tap and M tile).
template <int C>
__device__ __forceinline__ constexpr int nc_pitch() { return C * 2 + 16; }
template <int C>
__device__ __forceinline__ int nc_off(int q, int c) { return q * nc_pitch<C>() + (c << 4); }
// Output staging: dense [pixel][C] with the 16-byte chunk index XOR-ed with bits of the pixel index (conflict-free
// fragment writes, 16-byte reads by the copy-out)
template <int C>
__device__ __forceinline__ int nc_sw(int q, int c) {
  constexpr int NCH = C / 8;                                  // 2, 4 or 8 chunks per pixel
  constexpr int SH = NCH == 2 ? 2 : (NCH == 4 ? 1 : 0);
  return q * (C * 2) + (((c ^ (q >> SH)) & (NCH - 1)) << 4);
}

// halo tile of one source: RH x RW pixels from (yb, xb), zero outside the Hs x Ws image.  Warp w copies rows w, w + 8, ...:
// the row test and the row base addresses are warp-uniform, a lane's chunk k of the row is k * 16 bytes into the row in
// global memory (the first version's flat chunk loop spent ~30 instructions per 16 bytes: half of the kernel's dynamic
// instruction count).
template <int C, int RH, int RW>
__device__ __forceinline__ void nc_load_tile(const __nv_bfloat16* src, uint8_t* tile, int n, int yb, int xb, int Hs, int Ws, int tid) {
  constexpr int NCH = C / 8, CPR = RW * NCH;                  // 16-byte chunks per pixel / per tile row
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t tb = nc_smem_u32(tile);
  const __nv_bfloat16* img = src + (size_t)n * Hs * Ws * C;
#pragma unroll 1
  for (int ry = warp; ry < RH; ry += NC_THREADS / 32) {
    const int y = yb + ry;
    const bool rowok = y >= 0 && y < Hs;
    const __nv_bfloat16* grow = img + ((long long)y * Ws + xb) * C;   // dereferenced only where (y, x) is inside the image
    const int rbase = ry * RW * nc_pitch<C>();
#pragma unroll
    for (int k0 = 0; k0 < CPR; k0 += 32) {
      const int k = k0 + lane;
      if (k0 + 32 <= CPR || k < CPR) {
        const int rx = k / NCH, c = k % NCH, x = xb + rx;
        const int off = rbase + rx * nc_pitch<C>() + (c << 4);
        if (rowok && x >= 0 && x < Ws) nc_cp16(tb + off, grow + k * 8);
        else *reinterpret_cast<uint4*>(tile + off) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
}

// ---- 3 x 3, stride 1 -----------------------------------------------------------------------------------------------------
// C0 (+ C1): channels of the one or two concatenated sources; NT: 8-channel output tiles; HEAD: Cout = 9 + Softmax(1),
// fp32 NCHW.  Tile 32 x 16; warp w owns the 16-pixel column strip w & 1 and the four output rows 4 (w >> 1) .. + 3.
//
// Row-sliding accumulation.  With 16 output channels an A fragment (16 pixels x 16 channels, 512 B of shared memory) feeds
// only two MMAs, and at 128 B/clk the ldmatrix traffic alone equals the MMA time (first version: 42 % of the mma.sync
// peak, shared-memory pipe 63 % busy).  The A fragment of input row r shifted by dx is the operand of tap (dy, dx) for
// output row r - dy, for all three dy: the warp keeps the accumulators of its four output rows in registers, walks the six
// input rows once, and issues up to 3 x NT MMAs per ldmatrix.  The weights (9 x K-steps x NT B fragments, packed by the host
// in fragment order) stay in registers for the whole persistent CTA.
constexpr int NC_TH = 16, NC_ROWS = 4;        // tile height, output rows per warp
template <int C0, int C1, int NT, bool HEAD>
struct NcCfg {
  static constexpr int RW = NC_TW + 2, RH = NC_TH + 2, NPIX = RW * RH;
  static constexpr int KS0 = C0 / 16, KS1 = C1 / 16, KS = KS0 + KS1;
  static constexpr int T0 = NPIX * (C0 * 2 + 16), T1 = C1 > 0 ? NPIX * (C1 * 2 + 16) : 0;
  static constexpr int OUTB = HEAD ? 9 * NC_PLANE * 4 : NC_TW * NC_TH * NT * 16;
  static constexpr int SMEM = T0 + T1 + OUTB;
};

template <int C0, int C1, int NT, bool HEAD, int MINB>
__global__ void __launch_bounds__(NC_THREADS, MINB) narrow_conv3_kernel(const NcParams P) {
  using Cf = NcCfg<C0, C1, NT, HEAD>;
  constexpr int RW = Cf::RW, KS0 = Cf::KS0, KS = Cf::KS, TH = NC_TH;
  static_assert(!HEAD || NT == 2, "softmax head: 9 channels in two n-tiles");
  static_assert(NC_THREADS / 32 == 2 * (NC_TH / NC_ROWS), "8 warps = 2 column strips x 4 row groups");
  extern __shared__ __align__(16) uint8_t nc_smem[];
  uint8_t* so = nc_smem + Cf::T0 + Cf::T1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int H = P.H, W = P.W;

  auto load = [&](int tile) {
    const int tx = tile % P.tiles_x, r = tile / P.tiles_x, ty = r % P.tiles_y, n = r / P.tiles_y;
    nc_load_tile<C0, Cf::RH, RW>(P.src0, nc_smem, n, ty * TH - 1, tx * NC_TW - 1, H, W, tid);
    if constexpr (C1 > 0) nc_load_tile<C1, Cf::RH, RW>(P.src1, nc_smem + Cf::T0, n, ty * TH - 1, tx * NC_TW - 1, H, W, tid);
  };
  int tile = blockIdx.x;
  if (tile < P.tiles) load(tile);
  uint32_t bw[9 * KS][NT][2];
#pragma unroll
  for (int k = 0; k < 9 * KS; ++k)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(P.wfrag) + (k * NT + j) * 32 + lane);
      bw[k][j][0] = v.x;
      bw[k][j][1] = v.y;
    }
  float bias_r[NT][2];        // this lane's output channels 8j + 2t, + 1 (the softmax head pads channel 9 with channel 8)
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    bias_r[j][0] = __ldg(P.bias + (HEAD && j ? 8 : 8 * j + 2 * t));
    bias_r[j][1] = __ldg(P.bias + (HEAD && j ? 8 : 8 * j + 2 * t + 1));
  }
  const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1), khalf = lane >> 4;
  const int xs = (warp & 1) * 16, R0 = (warp >> 1) * NC_ROWS;
  // this lane's ldmatrix row address for halo pixel (R0, xs + lrow), chunk khalf, per source
  const uint32_t a0 = nc_smem_u32(nc_smem) + (uint32_t)nc_off<C0>(R0 * RW + xs + lrow, khalf);
  const uint32_t a1 = nc_smem_u32(nc_smem + Cf::T0) + (uint32_t)nc_off<(C1 > 0 ? C1 : 16)>(R0 * RW + xs + lrow, khalf);
#pragma unroll 1
  for (; tile < P.tiles; tile += gridDim.x) {
    nc_cp_wait();
    __syncthreads();            // the tile has landed; every warp is past the previous tile's copy-out (staging is free)
    const int x0 = (tile % P.tiles_x) * NC_TW, y0 = ((tile / P.tiles_x) % P.tiles_y) * TH, n = tile / (P.tiles_x * P.tiles_y);
    float acc[NC_ROWS][NT][4];
#pragma unroll
    for (int o = 0; o < NC_ROWS; ++o)
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[o][j][0] = acc[o][j][1] = acc[o][j][2] = acc[o][j][3] = 0.f;
#pragma unroll
    for (int r = 0; r < NC_ROWS + 2; ++r) {                  // input halo row R0 + r
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t a[4];
          if (ks < KS0) nc_ldsm_x4(a, a0 + (uint32_t)nc_off<C0>(r * RW + dx, ks * 2));
          else nc_ldsm_x4(a, a1 + (uint32_t)nc_off<(C1 > 0 ? C1 : 16)>(r * RW + dx, (ks - KS0) * 2));
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int o = r - dy;                             // output row fed by tap (dy, dx) of this input row
            if (o >= 0 && o < NC_ROWS) {
#pragma unroll
              for (int j = 0; j < NT; ++j) nc_mma(acc[o][j], a, bw[(dy * 3 + dx) * KS + ks][j][0], bw[(dy * 3 + dx) * KS + ks][j][1]);
            }
          }
        }
      }
    }
    // ---- epilogue into the staging buffer
#pragma unroll
    for (int o = 0; o < NC_ROWS; ++o) {
      const int pbase = (R0 + o) * NC_TW + xs;               // tile-local pixel of fragment row 0
      if constexpr (HEAD) {
        float* sf = reinterpret_cast<float*>(so);
        const float b0 = bias_r[0][0], b1 = bias_r[0][1], b8 = bias_r[1][0];
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          const float v0 = acc[o][0][2 * hr] + b0, v1 = acc[o][0][2 * hr + 1] + b1;
          const float v8 = __shfl_sync(0xffffffffu, acc[o][1][2 * hr] + b8, lane & ~3);   // channel 8 sits in the quad's lane 0
          float mx = fmaxf(v0, v1);
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          mx = fmaxf(mx, v8);
          const float e0 = expf(v0 - mx), e1 = expf(v1 - mx), e8 = expf(v8 - mx);
          float s = e0 + e1;
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += e8;
          const float inv = 1.0f / s;
          const int px = pbase + g + 8 * hr;
          sf[(2 * t) * NC_PLANE + px] = e0 * inv;
          sf[(2 * t + 1) * NC_PLANE + px] = e1 * inv;
          if (t == 0) sf[8 * NC_PLANE + px] = e8 * inv;
        }
      } else {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float b0 = bias_r[j][0], b1 = bias_r[j][1];
          *reinterpret_cast<uint32_t*>(so + nc_sw<NT * 8>(pbase + g, j) + 4 * t) =
              nc_pack(nc_act(acc[o][j][0] + b0, P.act, P.slope), nc_act(acc[o][j][1] + b1, P.act, P.slope));
          *reinterpret_cast<uint32_t*>(so + nc_sw<NT * 8>(pbase + g + 8, j) + 4 * t) =
              nc_pack(nc_act(acc[o][j][2] + b0, P.act, P.slope), nc_act(acc[o][j][3] + b1, P.act, P.slope));
        }
      }
    }
    __syncthreads();            // staging complete, source tile dead: the next tile streams in under the copy-out
    if (tile + (int)gridDim.x < P.tiles) load(tile + gridDim.x);

    if constexpr (HEAD) {
      // fp32 NCHW: 9 planes, rows of 32 pixels = 128 B
      const float* sf = reinterpret_cast<const float*>(so);
      float* outp = reinterpret_cast<float*>(P.out);
      for (int i = tid; i < 9 * TH * 8; i += NC_THREADS) {
        const int j = i / (TH * 8), rem = i - j * (TH * 8), r = rem >> 3, c4 = rem & 7;
        const int y = y0 + r, x = x0 + 4 * c4;
        if (y < H && x < W)
          *reinterpret_cast<float4*>(outp + (((size_t)n * 9 + j) * H + y) * W + x) =
              *reinterpret_cast<const float4*>(sf + j * NC_PLANE + r * NC_TW + 4 * c4);
      }
    } else {
      constexpr int PCS = NT;                                 // 16-byte pieces per pixel
      __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(P.out);
      for (int i = tid; i < NC_TW * TH * PCS; i += NC_THREADS) {
        const int piece = i % PCS, px = i / PCS;
        const int y = y0 + px / NC_TW, x = x0 + px % NC_TW;
        if (y < H && x < W)
          *reinterpret_cast<uint4*>(outp + (((size_t)n * H + y) * W + x) * (NT * 8) + piece * 8) =
              *reinterpret_cast<const uint4*>(so + nc_sw<NT * 8>(px, piece));
      }
    }
  }
}

// ---- ConvTranspose2d(4, 2, 1): four 2 x 2 phase convolutions ------------------------------------------------------------------
// Output tile 32 x 16; the source region is 18 x 10 half-resolution pixels.  Warp w takes phase w & 3 (py = phase >> 1,
// px = phase & 1) and four of the eight half-resolution rows: their M tiles (16 half-resolution pixels of a row) share
// the phase's B fragments.  Output pixel (2Y + py, 2X + px) sums taps (a, b) in {0,1}^2 with kernel index
// ky = 1 - py + 2a, kx = 1 - px + 2b at source pixel (Y + py - a, X + px - b)   [iy = (oy + 1 - ky) / 2].
template <int C, int NT>
struct NdCfg {
  static constexpr int RW = NC_TW / 2 + 2, RH = 16 / 2 + 2, NPIX = RW * RH;
  static constexpr int KS = C / 16;
  static constexpr int T0 = NPIX * (C * 2 + 16);
  static constexpr int WF = 4 * 4 * KS * NT * 256;
  static constexpr int PLANE = 8 * 16 * NT * 16 + 32;       // one phase of the staged output tile
  static constexpr int OUTB = 4 * PLANE;
  static constexpr int SMEM = T0 + WF + OUTB;
};

template <int C, int NT, int MINB>
__global__ void __launch_bounds__(NC_THREADS, MINB) narrow_deconv4_kernel(const NcParams P) {
  using Cf = NdCfg<C, NT>;
  constexpr int RW = Cf::RW, KS = Cf::KS, ND_PLANE = Cf::PLANE;
  extern __shared__ __align__(16) uint8_t nc_smem[];
  uint32_t* wf = reinterpret_cast<uint32_t*>(nc_smem + Cf::T0);
  uint8_t* so = nc_smem + Cf::T0 + Cf::WF;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int H = P.H, W = P.W;

  auto load = [&](int tile) {
    const int tx = tile % P.tiles_x, r = tile / P.tiles_x, ty = r % P.tiles_y, n = r / P.tiles_y;
    nc_load_tile<C, Cf::RH, RW>(P.src0, nc_smem, n, ty * 8 - 1, tx * (NC_TW / 2) - 1, H >> 1, W >> 1, tid);
  };
  int tile = blockIdx.x;
  if (tile < P.tiles) load(tile);
  {
    const uint32_t wb = nc_smem_u32(wf);
    for (int i = tid; i < Cf::WF / 16; i += NC_THREADS) nc_cp16(wb + 16 * i, reinterpret_cast<const uint4*>(P.wfrag) + i);
  }
  const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1), khalf = lane >> 4;
#pragma unroll 1
  for (; tile < P.tiles; tile += gridDim.x) {
  nc_cp_wait();
  __syncthreads();
  const int x0 = (tile % P.tiles_x) * NC_TW, y0 = ((tile / P.tiles_x) % P.tiles_y) * 16, n = tile / (P.tiles_x * P.tiles_y);
  const uint32_t s0 = nc_smem_u32(nc_smem);
  const int phase = warp & 3, py = phase >> 1, px = phase & 1, Y0 = (warp >> 2) * 4;
  float acc[4][NT][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[m][j][0] = acc[m][j][1] = acc[m][j][2] = acc[m][j][3] = 0.f;
#pragma unroll
  for (int tp = 0; tp < 4; ++tp) {
    const int a_ = tp >> 1, b_ = tp & 1;
    const int qt = (1 + py - a_) * RW + (1 + px - b_) + lrow;          // + Y * RW per M tile
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t b[NT][2];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const uint2 v = *reinterpret_cast<const uint2*>(wf + ((((phase * 4 + tp) * KS + ks) * NT + j) * 32 + lane) * 2);
        b[j][0] = v.x;
        b[j][1] = v.y;
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        uint32_t a[4];
        nc_ldsm_x4(a, s0 + (uint32_t)nc_off<C>(qt + (Y0 + m) * RW, ks * 2 + khalf));
#pragma unroll
        for (int j = 0; j < NT; ++j) nc_mma(acc[m][j], a, b[j][0], b[j][1]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    // staged per phase ([phase][Y][X] pixels, planes 32 B apart in the bank map) so that a quad's eight rows hit eight
    // bank groups; the copy-out below interleaves the phases back into rows
    uint8_t* sp = so + phase * ND_PLANE + (Y0 + m) * 16 * (NT * 16);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const float b0 = __ldg(P.bias + 8 * j + 2 * t), b1 = __ldg(P.bias + 8 * j + 2 * t + 1);
      *reinterpret_cast<uint32_t*>(sp + nc_sw<NT * 8>(g, j) + 4 * t) =
          nc_pack(nc_act(acc[m][j][0] + b0, P.act, P.slope), nc_act(acc[m][j][1] + b1, P.act, P.slope));
      *reinterpret_cast<uint32_t*>(sp + nc_sw<NT * 8>(g + 8, j) + 4 * t) =
          nc_pack(nc_act(acc[m][j][2] + b0, P.act, P.slope), nc_act(acc[m][j][3] + b1, P.act, P.slope));
    }
  }
  __syncthreads();
  if (tile + (int)gridDim.x < P.tiles) load(tile + gridDim.x);
  {
    constexpr int PCS = NT;
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(P.out);
    for (int i = tid; i < NC_TW * 16 * PCS; i += NC_THREADS) {
      const int piece = i % PCS, pxl = i / PCS;
      const int y = y0 + pxl / NC_TW, x = x0 + pxl % NC_TW;
      if (y < H && x < W)
        *reinterpret_cast<uint4*>(outp + (((size_t)n * H + y) * W + x) * (NT * 8) + piece * 8) =
            *reinterpret_cast<const uint4*>(so + (((pxl / NC_TW) & 1) * 2 + (pxl & 1)) * ND_PLANE + ((pxl / NC_TW) >> 1) * 16 * (NT * 16) +
                                            nc_sw<NT * 8>((pxl % NC_TW) >> 1, piece));
    }
  }
  }  // tile loop
}

// ---- host side -------------------------------------------------------------------------------------------------------------
enum NcKind { NC_NONE = 0, NC_C16X2_16, NC_C16_SOFTMAX9, NC_DC32_16 };

bool nc_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DISCO_NARROW");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

NcKind nc_classify(const disco_conv_desc* d) {
  if (!d || !nc_enabled() || d->dtype != DISCO_BF16 || d->residual || d->post_scale || d->post_shift) return NC_NONE;
  if (d->n_src < 1 || d->n_src > 2 || d->batch <= 0 || d->batch > 65535) return NC_NONE;
  if (d->act != DISCO_ACT_NONE && !(d->slope >= 0.f && d->slope < 1.f) && d->act != DISCO_ACT_RELU) return NC_NONE;
  for (int s = 0; s < d->n_src; ++s)
    if (d->src[s].is_f32 || d->src[s].up2 || !d->src[s].ptr) return NC_NONE;
  const int c0 = d->src[0].C, c1 = d->n_src == 2 ? d->src[1].C : 0;
  if (d->kind == DISCO_CONV3) {
    if (d->stride != 1) return NC_NONE;
    for (int s = 0; s < d->n_src; ++s)
      if (d->src[s].H != d->Ho || d->src[s].W != d->Wo) return NC_NONE;
    if (d->head == DISCO_HEAD_SOFTMAX9)
      return (c0 == 16 && c1 == 0 && d->Cout == 9 && d->Wo % 4 == 0) ? NC_C16_SOFTMAX9 : NC_NONE;
    if (d->head != DISCO_HEAD_NONE) return NC_NONE;
    if (c0 == 16 && c1 == 16 && d->Cout == 16) return NC_C16X2_16;
    return NC_NONE;                                           // 32-channel layers: the tcgen05 resident kernel is faster (measured)
  }
  if (d->kind == DISCO_DECONV4) {
    if (d->head != DISCO_HEAD_NONE || d->n_src != 1 || d->Ho % 2 || d->Wo % 2) return NC_NONE;
    if (d->src[0].H * 2 != d->Ho || d->src[0].W * 2 != d->Wo) return NC_NONE;
    if (c0 == 32 && d->Cout == 16) return NC_DC32_16;
  }
  return NC_NONE;
}

uint16_t nc_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (u & 0x7fffffffu) > 0x7f800000u ? (uint16_t)0x7fff : (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

template <typename K, typename Cfg>
int nc_launch(disco_handle* h, K kernel, NcParams& P, int TH, cudaStream_t st) {
  if (int rc = disco_ensure_smem(h, (const void*)kernel, Cfg::SMEM)) return rc;
  P.tiles_x = (P.W + NC_TW - 1) / NC_TW;
  P.tiles_y = (P.H + TH - 1) / TH;
  const long long tiles = (long long)P.tiles_x * P.tiles_y * P.B;
  DISCO_CHECK_ARG(tiles < (1ll << 30), "narrow_launch: too many tiles");
  P.tiles = (int)tiles;
  int per_sm = 0;
  DISCO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NC_THREADS, Cfg::SMEM));
  DISCO_CHECK_ARG(per_sm > 0, "narrow_launch: kernel does not fit on an SM");
  const int grid = (int)std::min<long long>(tiles, (long long)per_sm * h->sm_count);
  kernel<<<grid, NC_THREADS, Cfg::SMEM, st>>>(P);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}

}  // namespace

bool conv_narrow_match(const disco_conv_desc* d) { return nc_classify(d) != NC_NONE; }

// bf16 elements of the packing (fragment order: [tap][k-step][n-tile][lane][b0.lo b0.hi b1.lo b1.hi])
int64_t conv_narrow_weight_elems(const disco_conv_desc* d) {
  const NcKind k = nc_classify(d);
  if (k == NC_NONE) return 0;
  const int ks = (d->src[0].C + (d->n_src == 2 ? d->src[1].C : 0)) / 16;
  const int nt = (d->Cout + 7) / 8;
  const int taps = d->kind == DISCO_DECONV4 ? 16 : 9;
  return (int64_t)taps * ks * nt * 32 * 4;
}

int conv_narrow_pack(const disco_conv_desc* d, const float* w32, uint16_t* out) {
  const NcKind k = nc_classify(d);
  DISCO_CHECK_ARG(k != NC_NONE, "narrow_pack: descriptor not claimed by the narrow-layer kernels");
  const int nt = (d->Cout + 7) / 8, Cout = d->Cout;
  const bool dc = d->kind == DISCO_DECONV4;
  size_t o = 0;
  for (int slot = 0; slot < (dc ? 16 : 9); ++slot) {
    int tap = slot;
    if (dc) {                                                 // slot = phase * 4 + (a * 2 + b)
      const int phase = slot >> 2, a = (slot >> 1) & 1, b = slot & 1;
      const int ky = 1 - (phase >> 1) + 2 * a, kx = 1 - (phase & 1) + 2 * b;
      tap = ky * 4 + kx;
    }
    for (int s = 0; s < d->n_src; ++s) {
      const int Cs = d->src[s].C;
      const float* wb = w32 + d->src[s].w_off + (size_t)tap * Cs * Cout;
      for (int ks = 0; ks < Cs / 16; ++ks)
        for (int j = 0; j < nt; ++j)
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3, co = 8 * j + g;
            const int ci[4] = {16 * ks + 2 * t, 16 * ks + 2 * t + 1, 16 * ks + 8 + 2 * t, 16 * ks + 9 + 2 * t};
            for (int e = 0; e < 4; ++e) out[o++] = co < Cout ? nc_bf16(wb[(size_t)ci[e] * Cout + co]) : (uint16_t)0;
          }
    }
  }
  return DISCO_OK;
}

int conv_narrow_launch(disco_handle* h, const disco_conv_desc* d, cudaStream_t st) {
  const NcKind k = nc_classify(d);
  DISCO_CHECK_ARG(k != NC_NONE, "narrow_launch: descriptor not claimed by the narrow-layer kernels");
  DISCO_CHECK_ARG(d->weights && d->bias && d->out, "narrow_launch: null pointer");
  NcParams P;
  P.src0 = reinterpret_cast<const __nv_bfloat16*>(d->src[0].ptr);
  P.src1 = d->n_src == 2 ? reinterpret_cast<const __nv_bfloat16*>(d->src[1].ptr) : nullptr;
  P.wfrag = reinterpret_cast<const uint32_t*>(d->weights);
  P.bias = d->bias;
  P.out = d->out;
  P.B = d->batch; P.H = d->Ho; P.W = d->Wo;
  P.act = d->act;
  P.slope = d->act == DISCO_ACT_RELU ? 0.f : d->slope;
  switch (k) {
    case NC_C16X2_16:
      return nc_launch<decltype(&narrow_conv3_kernel<16, 16, 2, false, 2>), NcCfg<16, 16, 2, false>>(
          h, narrow_conv3_kernel<16, 16, 2, false, 2>, P, NC_TH, st);
    case NC_C16_SOFTMAX9:
      return nc_launch<decltype(&narrow_conv3_kernel<16, 0, 2, true, 2>), NcCfg<16, 0, 2, true>>(
          h, narrow_conv3_kernel<16, 0, 2, true, 2>, P, NC_TH, st);
    case NC_DC32_16:
      return nc_launch<decltype(&narrow_deconv4_kernel<32, 2, 2>), NdCfg<32, 2>>(h, narrow_deconv4_kernel<32, 2, 2>, P, 16, st);
    default:
      break;
  }
  return DISCO_ERR_UNSUPPORTED;
}
