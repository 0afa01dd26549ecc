// Generic fused convolution on CUDA cores (fp32 accumulate, fp32 weights).
//
// This is the *exact* path: with DISCO_F32 activations it reproduces the reference's fp32
// convolution stack to ~1e-6 (parity tests, |d ab| <= 1e-3 gate) and serves as the on-device
// cross-check for the tensor-core kernel (conv_tc.cu).  It also covers the shapes the tcgen05
// kernel does not take (Cin == 1 first layers).  Implicit GEMM: tile = 64 output pixels (8x8)
// x 64 output channels per CTA, K stepped by (source, tap, 16 input channels) through shared memory,
// 4x4 register micro-tiles.
#include "common.cuh"

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

}  // namespace

int conv_simt_launch(disco_handle* h, const disco_conv_desc* d, cudaStream_t st) {
  DISCO_CHECK_ARG(d->n_src >= 1 && d->n_src <= 2, "conv: n_src must be 1 or 2");
  DISCO_CHECK_ARG(d->head == DISCO_HEAD_NONE || d->Cout <= TC, "conv: head needs Cout <= 64");
  // dedicated kernel for the single-channel fp32 input layers (bf16 storage path only: the fp32 exact path keeps
  // the reference summation structure of the generic kernel)
  if (d->dtype == DISCO_BF16 && d->kind == DISCO_CONV3 && d->n_src == 1 && d->src[0].C == 1 && d->src[0].is_f32 &&
      d->stride == 1 && !d->src[0].up2 && d->head == DISCO_HEAD_NONE && !d->residual && d->Cout % 4 == 0 &&
      256 % (d->Cout / 4) == 0 && d->batch * ((d->Ho + 15) / 16) <= 65535 && (long long)d->Ho * d->Wo * (d->Cout / 4) < (1ll << 31)) {
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
