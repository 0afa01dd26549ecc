// Generic fused convolution on CUDA cores (fp32 accumulate, fp32 weights).
//
// This is the *exact* path: with DISCO_F32 activations it reproduces the reference's fp32
// convolution stack to ~1e-6 (parity tests, |d ab| <= 1e-3 gate) and serves as the on-device
// cross-check for the tensor-core kernel (conv_tc.cu).  It also covers the shapes the tcgen05
// kernel does not take (Cin == 1 first layers).  Implicit GEMM: tile = 64 output pixels (8x8)
// x 64 output channels per CTA, K stepped by (source, tap, 16 input channels) through shared memory,
// 4x4 register micro-tiles.
#include "common.cuh"
#include <cstdlib>

namespace {

constexpr int TP = 64;   // pixels per tile (8 x 8)
constexpr int TC = 64;   // output channels per tile
constexpr int KC = 16;   // input channels per K step

struct SimtParams {
  disco_conv_desc d;
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, bool ok, int c, int C, float out[4]);

template <>
__device__ __forceinline__ void load4<float>(const float* p, bool ok, int c, int C, float out[4]) {
  if (!ok) { out[0] = out[1] = out[2] = out[3] = 0.f; return; }
  if (c + 4 <= C && (C & 3) == 0) {
    float4 v = *reinterpret_cast<const float4*>(p + c);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = (c + j < C) ? p[c + j] : 0.f;
  }
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, bool ok, int c, int C, float out[4]) {
  if (!ok) { out[0] = out[1] = out[2] = out[3] = 0.f; return; }
  if (c + 4 <= C && (C & 3) == 0) {
    uint2 raw = *reinterpret_cast<const uint2*>(p + c);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
    out[0] = __low2float(a); out[1] = __high2float(a); out[2] = __low2float(b); out[3] = __high2float(b);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = (c + j < C) ? __bfloat162float(p[c + j]) : 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtParams P) {
  const disco_conv_desc& d = P.d;
  __shared__ float As[KC][TP + 4];
  __shared__ float Bs[KC][TC + 4];
  __shared__ float Cs[TP][TC + 1];

  const int tid = threadIdx.x;
  const int tiles_x = (d.Wo + 7) >> 3;
  const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x % tiles_x;
  const int co0 = blockIdx.y * TC;
  const int n = blockIdx.z;

  const int tx = tid & 15;   // channel group: co0 + tx*4 .. +3
  const int ty = tid >> 4;   // pixel group:   ty*4 .. +3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // A-load role: pixel lp = tid/4, channels (tid%4)*4..+3 of the current chunk
  const int lp = tid >> 2, lc = (tid & 3) * 4;
  const int l_oy = tile_y * 8 + (lp >> 3), l_ox = tile_x * 8 + (lp & 7);
  // B-load role: k = tid/16, channels (tid%16)*4..+3
  const int bk = tid >> 4, bc = (tid & 15) * 4;

  const int ntaps = d.kind == DISCO_DECONV4 ? 16 : 9;
  for (int s = 0; s < d.n_src; ++s) {
    const disco_conv_src& src = d.src[s];
    const int Cs_ = src.C;
    const float* wbase = reinterpret_cast<const float*>(d.weights) + src.w_off;
    for (int tap = 0; tap < ntaps; ++tap) {
      int iy, ix;
      bool ok;
      if (d.kind == DISCO_DECONV4) {
        const int ky = tap >> 2, kx = tap & 3;
        const int vy = l_oy + 1 - ky, vx = l_ox + 1 - kx;
        ok = vy >= 0 && vx >= 0 && !(vy & 1) && !(vx & 1);
        iy = vy >> 1; ix = vx >> 1;
        ok = ok && iy < src.H && ix < src.W;
      } else {
        const int ky = tap / 3, kx = tap % 3;
        const int vy = l_oy * d.stride + ky - 1, vx = l_ox * d.stride + kx - 1;
        ok = vy >= 0 && vx >= 0 && vy < (src.H << src.up2) && vx < (src.W << src.up2);
        iy = vy >> src.up2; ix = vx >> src.up2;
      }
      ok = ok && l_oy < d.Ho && l_ox < d.Wo;
      const size_t pix = ((size_t)n * src.H + (ok ? iy : 0)) * src.W + (ok ? ix : 0);
      for (int c0 = 0; c0 < Cs_; c0 += KC) {
        float a[4];
        if (src.is_f32 || sizeof(T) == 4)
          load4<float>(reinterpret_cast<const float*>(src.ptr) + pix * Cs_, ok, c0 + lc, Cs_, a);
        else
          load4<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(src.ptr) + pix * Cs_, ok, c0 + lc, Cs_, a);
#pragma unroll
        for (int j = 0; j < 4; ++j) As[lc + j][lp] = a[j];
        {
          const int ci = c0 + bk;
          const float* wrow = wbase + ((size_t)tap * Cs_ + ci) * d.Cout;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int co = co0 + bc + j;
            Bs[bk][bc + j] = (ci < Cs_ && co < d.Cout) ? wrow[co] : 0.f;
          }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
          const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
          const float aa[4] = {av.x, av.y, av.z, av.w};
          const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }

  // epilogue: bias -> (+residual) -> activation -> post affine, staged through smem
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = ty * 4 + i;
    const int oy = tile_y * 8 + (p >> 3), ox = tile_x * 8 + (p & 7);
    const bool pok = oy < d.Ho && ox < d.Wo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      float v = acc[i][j];
      if (co < d.Cout) {
        v += d.bias[co];
        if (d.residual && pok) {
          const size_t o = (((size_t)n * d.Ho + oy) * d.Wo + ox) * d.Cout + co;
          v += to_f32(reinterpret_cast<const T*>(d.residual)[o]);
        }
        v = apply_act(v, d.act, d.slope);
        if (d.post_scale) v = v * d.post_scale[co] + d.post_shift[co];
      }
      Cs[p][tx * 4 + j] = v;
    }
  }
  __syncthreads();
  if (d.head == DISCO_HEAD_NONE) {
    T* out = reinterpret_cast<T*>(d.out);
    for (int e = tid; e < TP * TC; e += 256) {
      const int p = e >> 6, c = e & 63;
      const int oy = tile_y * 8 + (p >> 3), ox = tile_x * 8 + (p & 7), co = co0 + c;
      if (oy < d.Ho && ox < d.Wo && co < d.Cout)
        out[(((size_t)n * d.Ho + oy) * d.Wo + ox) * d.Cout + co] = from_f32<T>(Cs[p][c]);
    }
  } else if (tid < TP) {
    const int p = tid;
    const int oy = tile_y * 8 + (p >> 3), ox = tile_x * 8 + (p & 7);
    if (oy < d.Ho && ox < d.Wo) {
      float* out = reinterpret_cast<float*>(d.out);
      const size_t plane = (size_t)d.Ho * d.Wo;
      const size_t base = (size_t)n * d.Cout * plane + (size_t)oy * d.Wo + ox;
      if (d.head == DISCO_HEAD_SOFTMAX9) {
        float m = Cs[p][0];
        for (int c = 1; c < 9; ++c) m = fmaxf(m, Cs[p][c]);
        float e[9], s = 0.f;
        for (int c = 0; c < 9; ++c) { e[c] = expf(Cs[p][c] - m); s += e[c]; }
        for (int c = 0; c < 9; ++c) out[base + c * plane] = e[c] / s;
      } else {
        for (int c = 0; c < 2; ++c) out[base + c * plane] = d.head == DISCO_HEAD_TANH2 ? tanhf(Cs[p][c]) : Cs[p][c];
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Cin == 1 first layers (repnet.conv1_2.0: 1->64, segnet.conv0a: 1->16): HBM-bound (4 B in, 2*Cout B out
// per pixel).  One thread = 4 consecutive pixels of a row x 4 output channels: the 3x6 gray window slides across the
// 4 pixels and down a 16-row strip, the 36 weights stay in registers (packed fp32 pairs), and every store instruction
// of a warp writes whole 128-byte lines of the NHWC output.
// grid = (ceil(W / pixels_per_block), batch * ceil(H / 16)), blockDim = 256.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256, 2) conv_c1_kernel(const float* __restrict__ gray, const float* __restrict__ w,
                                                      const float* __restrict__ bias, const float* __restrict__ ps,
                                                      const float* __restrict__ pb, int rows, int H, int W, int Cout, int act,
                                                      float slope, T* __restrict__ out) {
  constexpr int PX = 4;                               // consecutive pixels of a row per thread
  constexpr int CPT = 4;                              // output channels per thread (two packed fp32 pairs)
  constexpr int RB = 16;                              // rows per blockIdx.y
  const int tpp = Cout / CPT;                         // threads per pixel group
  const int cgi = threadIdx.x % tpp, pg = threadIdx.x / tpp;
  const int c0 = cgi * CPT;
  const int x0 = (blockIdx.x * (256 / tpp) + pg) * PX;
  // weights / bias as packed fp32 pairs (FFMA2): 18 instead of 36 FMA instructions per pixel and thread
  f32x2 wr2[9][CPT / 2], br2[CPT / 2];
  float sr[CPT], hr[CPT];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < CPT / 2; ++j) wr2[t][j] = pack2(w[t * Cout + c0 + 2 * j], w[t * Cout + c0 + 2 * j + 1]);
#pragma unroll
  for (int j = 0; j < CPT / 2; ++j) br2[j] = pack2(bias[c0 + 2 * j], bias[c0 + 2 * j + 1]);
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    sr[j] = ps ? ps[c0 + j] : 1.f;
    hr[j] = pb ? pb[c0 + j] : 0.f;
  }
  if (x0 >= W) return;
  // blockIdx.y owns RB consecutive rows of one image: the 3-row gray window slides DOWN the strip, so every new output
  // row needs one new window row (6 loads instead of 18), and that row is requested before the arithmetic of the
  // current row starts -- its latency hides behind the row's FFMA2 work instead of being paid once per row.
  const int blocks_per_img = (H + RB - 1) / RB;
  const int n = blockIdx.y / blocks_per_img, y_begin = (blockIdx.y % blocks_per_img) * RB;
  const int y_end = y_begin + RB < H ? y_begin + RB : H;
  const float* g = gray + (size_t)n * H * W;               // image base
  auto load_row = [&](int yy, float (&dst)[PX + 2]) {
    const bool yok = yy >= 0 && yy < H;
#pragma unroll
    for (int dx = 0; dx < PX + 2; ++dx) {
      const int xx = x0 + dx - 1;
      dst[dx] = (yok && xx >= 0 && xx < W) ? __ldg(g + (size_t)yy * W + xx) : 0.f;
    }
  };
  float win[3][PX + 2], nxt[PX + 2];
  load_row(y_begin - 1, win[0]);
  load_row(y_begin, win[1]);
  load_row(y_begin + 1, win[2]);
  for (int y = y_begin; y < y_end; ++y) {
    const size_t r = (size_t)n * H + y;
    load_row(y + 2, nxt);                                  // consumed only after this row's arithmetic
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      if (x0 + p >= W) break;
      f32x2 v2[CPT / 2];
#pragma unroll
      for (int j = 0; j < CPT / 2; ++j) v2[j] = br2[j];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float gv = win[t / 3][p + t % 3];
        const f32x2 G = pack2(gv, gv);
#pragma unroll
        for (int j = 0; j < CPT / 2; ++j) v2[j] = ffma2(G, wr2[t][j], v2[j]);
      }
      float v[CPT];
#pragma unroll
      for (int j = 0; j < CPT / 2; ++j) unpack2(v2[j], v[2 * j], v[2 * j + 1]);
#pragma unroll
      for (int j = 0; j < CPT; ++j) v[j] = apply_act(v[j], act, slope) * sr[j] + hr[j];
      T* o = out + (r * W + x0 + p) * Cout + c0;
      if (sizeof(T) == 2) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(o) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
      } else {
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
#pragma unroll
    for (int dx = 0; dx < PX + 2; ++dx) { win[0][dx] = win[1][dx]; win[1][dx] = win[2][dx]; win[2][dx] = nxt[dx]; }
  }
}

// ------------------------------------------------------------------------------------------------
// Cin == 1 -> 64 channels on tensor cores (repnet.conv1_2.0, reference models/network.py:152).  The CUDA-core kernel above
// is issue-bound (r2a capture: 63 % of the issue slots, 36 % of the HBM peak): 9 x 64 FMAs + the epilogue per pixel.
// As a GEMM the layer is [pixels x 32] x [32 x 64] with the K axis holding, per pixel, the nine L-channel taps split into
// bf16 hi + lo against the weights split the same way (g_hi w_hi + g_lo w_hi + g_hi w_lo: fp32-grade products; the
// L channel is fp32 and must not be rounded to bf16) and two constant-one columns that carry the bias (hi, lo):
//   k = 8 s + j;  s = 0: g_hi[tap j] x w_hi[j];  s = 1: g_lo[tap j] x w_hi[j];  s = 2: g_hi[tap j] x w_lo[j]   (taps 0..7)
//   s = 3: j = 0: g_hi[8] x w_hi[8];  1: g_lo[8] x w_hi[8];  2: g_hi[8] x w_lo[8];  3: 1 x b_hi;  4: 1 x b_lo;  5..7: 0
// so that in the m16n8k16 A fragment thread t needs exactly the taps 2t, 2t+1 (and 8): 16 MMAs per 16 pixels x 64 channels
// instead of 144 FFMA2 warp instructions.  The B operand (32 x 64, 32 registers) is built once per thread.
// One warp = 16 consecutive pixels of a row; the bf16 tile is staged in a private shared-memory patch and written out as
// whole 128-byte pixel rows.  Persistent CTAs (2 per SM) loop over 64 x 16 pixel tiles so that the B-operand set-up (about
// 60 dependent loads per thread) is paid once per CTA, not once per tile.
// ------------------------------------------------------------------------------------------------
constexpr int C1M_TW = 64, C1M_TH = 16;
__device__ __forceinline__ void c1m_split(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffff0000u);        // truncation: hi is exactly a bf16, lo = x - hi is exact in fp32
  lo = x - hi;
}
__device__ __forceinline__ uint32_t c1m_pack(float a, float b) {   // bf16x2, element a in the low half (round to nearest)
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void c1m_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 2) conv_c1_mma_kernel(const float* __restrict__ gray, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const float* __restrict__ ps,
                                                             const float* __restrict__ pb, int B, int H, int W, int act, float slope,
                                                             __nv_bfloat16* __restrict__ out) {
  __shared__ float sg[(C1M_TH + 2) * (C1M_TW + 2)];
  __shared__ __align__(16) uint8_t patch[8][16 * 144];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int tiles_x = (W + C1M_TW - 1) / C1M_TW, tiles_y = (H + C1M_TH - 1) / C1M_TH;
  const int n_tiles = tiles_x * tiles_y * B;
  // ---- B fragments: b[ks][j][0] = B[k = 16 ks + 2t, +1][n = 8j + g], b[ks][j][1] = B[k = 16 ks + 8 + 2t, +1][n]
  uint32_t bfr[2][8][2];
  {
    auto whi = [&](int tap, int co) { float hi, lo; c1m_split(w[tap * 64 + co], hi, lo); return hi; };
    auto wlo = [&](int tap, int co) { float hi, lo; c1m_split(w[tap * 64 + co], hi, lo); return lo; };
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = 8 * j + g;
      float bh, bl;
      c1m_split(bias[co], bh, bl);
      // slot s = k / 8: ks 0 holds slots 0 (reg 0) and 1 (reg 1); ks 1 holds slots 2 (reg 0) and 3 (reg 1)
      bfr[0][j][0] = c1m_pack(whi(2 * t, co), whi(2 * t + 1, co));
      bfr[0][j][1] = bfr[0][j][0];
      bfr[1][j][0] = c1m_pack(wlo(2 * t, co), wlo(2 * t + 1, co));
      const float s3a = t == 0 ? whi(8, co) : (t == 1 ? wlo(8, co) : (t == 2 ? bl : 0.f));
      const float s3b = t == 0 ? whi(8, co) : (t == 1 ? bh : 0.f);
      bfr[1][j][1] = c1m_pack(s3a, s3b);
    }
  }
  float sc[8][2], sh[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      sc[j][e] = ps ? ps[8 * j + 2 * t + e] : 1.f;
      sh[j][e] = pb ? pb[8 * j + 2 * t + e] : 0.f;
    }
  // taps 2t, 2t+1 and 8 as offsets into the staged tile
  const int o0 = ((2 * t) / 3) * (C1M_TW + 2) + (2 * t) % 3, o1 = ((2 * t + 1) / 3) * (C1M_TW + 2) + (2 * t + 1) % 3;
  const int o8 = 2 * (C1M_TW + 2) + 2;
  uint8_t* mp = patch[warp];
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  const int n = tile / (tiles_x * tiles_y), tr = tile - n * (tiles_x * tiles_y);
  const int y0 = (tr / tiles_x) * C1M_TH, x0 = (tr % tiles_x) * C1M_TW;
  __syncthreads();                                          // every warp is done with the previous tile's L-channel patch
  {
    const float* gi = gray + (size_t)n * H * W;
    for (int i = tid; i < (C1M_TH + 2) * (C1M_TW + 2); i += 256) {
      const int ry = i / (C1M_TW + 2), rx = i - ry * (C1M_TW + 2);
      const int y = y0 - 1 + ry, x = x0 - 1 + rx;
      sg[i] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(gi + (size_t)y * W + x) : 0.f;
    }
  }
  __syncthreads();
  for (int mt = warp; mt < (C1M_TW / 16) * C1M_TH; mt += 8) {
    const int ry = mt >> 2, rx = (mt & 3) * 16;             // tile row, first pixel of the 16-pixel run
    uint32_t a[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {                          // fragment rows g and g + 8
      const float* p = sg + ry * (C1M_TW + 2) + rx + g + 8 * r;
      float h0, l0, h1, l1, h8, l8;
      c1m_split(p[o0], h0, l0);
      c1m_split(p[o1], h1, l1);
      c1m_split(p[o8], h8, l8);
      const uint32_t s0 = c1m_pack(h0, h1);                 // exact: both are bf16 values
      const uint32_t s1 = c1m_pack(l0, l1);
      const uint32_t s3 = t == 0 ? c1m_pack(h8, l8) : (t == 1 ? c1m_pack(h8, 1.0f) : (t == 2 ? c1m_pack(1.0f, 0.f) : 0u));
      a[0][r] = s0; a[0][2 + r] = s1;                       // k-step 0: slots 0 | 1
      a[1][r] = s0; a[1][2 + r] = s3;                       // k-step 1: slots 2 | 3
    }
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
      c1m_mma(acc[j], a[0][0], a[0][1], a[0][2], a[0][3], bfr[0][j][0], bfr[0][j][1]);
      c1m_mma(acc[j], a[1][0], a[1][1], a[1][2], a[1][3], bfr[1][j][0], bfr[1][j][1]);
    }
    __syncwarp();                                           // the previous tile's copy-out has left the patch
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = apply_act(acc[j][e], act, slope) * sc[j][e & 1] + sh[j][e & 1];
      *reinterpret_cast<uint32_t*>(mp + g * 144 + 16 * j + 4 * t) = c1m_pack(v[0], v[1]);
      *reinterpret_cast<uint32_t*>(mp + (g + 8) * 144 + 16 * j + 4 * t) = c1m_pack(v[2], v[3]);
    }
    __syncwarp();
    const int y = y0 + ry;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = lane + 32 * r, px = i >> 3, pc = i & 7, x = x0 + rx + px;
      if (y < H && x < W)
        *reinterpret_cast<uint4*>(out + (((size_t)n * H + y) * W + x) * 64 + pc * 8) = *reinterpret_cast<const uint4*>(mp + px * 144 + pc * 16);
    }
  }
  }
}

}  // namespace

int conv_simt_launch(disco_handle* h, const disco_conv_desc* d, cudaStream_t st) {
  DISCO_CHECK_ARG(d->n_src >= 1 && d->n_src <= 2, "conv: n_src must be 1 or 2");
  DISCO_CHECK_ARG(d->head == DISCO_HEAD_NONE || d->Cout <= TC, "conv: head needs Cout <= 64");
  // dedicated kernel for the single-channel fp32 input layers (bf16 storage path only: the fp32 exact path keeps
  // the reference summation structure of the generic kernel)
  if (d->dtype == DISCO_BF16 && d->kind == DISCO_CONV3 && d->n_src == 1 && d->src[0].C == 1 && d->src[0].is_f32 &&
      d->stride == 1 && !d->src[0].up2 && d->head == DISCO_HEAD_NONE && !d->residual && d->Cout % 4 == 0 &&
      256 % (d->Cout / 4) == 0 && d->batch * ((d->Ho + 15) / 16) <= 65535 && (long long)d->Ho * d->Wo * (d->Cout / 4) < (1ll << 31)) {
    if (d->Cout == 64 && d->batch <= 65535 && (getenv("DISCO_C1_MMA") == nullptr || getenv("DISCO_C1_MMA")[0] != '0')) {
      const long long tiles = (long long)((d->Wo + C1M_TW - 1) / C1M_TW) * ((d->Ho + C1M_TH - 1) / C1M_TH) * d->batch;
      const int grid = (int)(tiles < 2 * h->sm_count ? tiles : 2 * h->sm_count);
      conv_c1_mma_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(d->src[0].ptr),
                                               reinterpret_cast<const float*>(d->weights) + d->src[0].w_off, d->bias, d->post_scale,
                                               d->post_shift, d->batch, d->Ho, d->Wo, d->act, d->slope,
                                               reinterpret_cast<__nv_bfloat16*>(d->out));
      DISCO_LAUNCH_CHECK(h);
      return DISCO_OK;
    }
    const int tpp = d->Cout / 4, px_per_block = (256 / tpp) * 4;
    const int rows = d->batch * d->Ho;
    const int gx = (d->Wo + px_per_block - 1) / px_per_block;
    const int gy = d->batch * ((d->Ho + 15) / 16);     // 16-row strips (RB in the kernel)
    conv_c1_kernel<__nv_bfloat16><<<dim3(gx, gy), 256, 0, st>>>(
        reinterpret_cast<const float*>(d->src[0].ptr), reinterpret_cast<const float*>(d->weights) + d->src[0].w_off, d->bias,
        d->post_scale, d->post_shift, rows, d->Ho, d->Wo, d->Cout, d->act, d->slope,
        reinterpret_cast<__nv_bfloat16*>(d->out));
    DISCO_LAUNCH_CHECK(h);
    return DISCO_OK;
  }
  SimtParams P;
  P.d = *d;
  dim3 grid(((d->Wo + 7) / 8) * ((d->Ho + 7) / 8), (d->Cout + TC - 1) / TC, d->batch);
  if (d->dtype == DISCO_F32)
    conv_simt_kernel<float><<<grid, 256, 0, st>>>(P);
  else
    conv_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(P);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}
