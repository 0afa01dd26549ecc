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
#include <cstdlib>

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

// ------------------------------------------------------------------------------------------------
// bf16 path: the same per-cell masses on tensor cores.  Per cell the 9 x 64 feature masses are a small GEMM
//   M[dir][ch] = sum_px  P[px][dir] * F[px][ch]      (16 x 256) x (256 x 64), K = the cell's 256 pixels,
// so the kernel stages the cell's feature tile once (cp.async, 16-byte pieces, whole 2 KB rows), writes the affinity
// transposed and split into bf16 hi + lo (the features ARE bf16 on this path, so hi*f + lo*f is the fp32-grade product),
// and four warps run mma.sync m16n8k16 with A = affinity^T by ldmatrix and B = features by ldmatrix.trans: 64 MMAs per warp
// instead of ~1150 FFMA + LDS instructions per warp in the CUDA-core kernel, which was issue-bound at 42 % of the HBM
// peak.  A fifth warp adds up the four scalar masses (ab0, ab1, soft mass, hard mass) like the CUDA-core kernel does.
// grid (w, h, B), 160 threads.
// ------------------------------------------------------------------------------------------------
constexpr int PF_FP = 144;                 // feature tile pixel pitch in bytes (128 + 16: conflict-free ldmatrix)
constexpr int PF_AP = (256 + 8) * 2;       // affinity^T row pitch in bytes (one row per direction)
constexpr int PF_STAGES = 2;               // feature tile stages per CTA (persistent CTAs, two per SM: one computes while the other loads)
struct PfSmem {
  __align__(16) uint8_t F[PF_STAGES][256 * PF_FP];
  __align__(16) uint8_t Ahi[16 * PF_AP];
  __align__(16) uint8_t Alo[16 * PF_AP];
  __align__(16) uint8_t Hot[16 * PF_AP];   // indicator of the pixel's maximal direction(s), transposed like the affinity
  __align__(16) uint8_t Ehi[256 * 16];     // per pixel 8 "extra channels" (ab0, ab1, 1, 0...) as bf16 hi | lo
  __align__(16) uint8_t Elo[256 * 16];
};

__device__ __forceinline__ uint32_t pf_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void pf_ldsm(uint32_t (&r)[4], uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void pf_ldsm_t(uint32_t (&r)[4], uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void pf_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void pf_split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<uint32_t*>(&h);
  __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xffff0000u));
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// Persistent CTAs (grid = 2 x #SMs): cell i of a CTA is computed while the feature tile of cell i+1 is in flight
// (cp.async, two stages) and the affinity / ab values of cell i+1 sit in registers; the second CTA of the SM fills the
// phases in which the first one waits -- an HBM-streaming kernel has to keep
// tens of KB per SM in flight, which one short-lived CTA per cell did not (r2j capture: 27 % of DRAM peak, long-scoreboard
// stalls 4.7 cycles per issue).
__global__ void __launch_bounds__(160, 2) poolfeat_partial_mma_kernel(const __nv_bfloat16* __restrict__ feats, const float* __restrict__ ab,
                                                                      const float* __restrict__ aff, int B, int H, int W,
                                                                      float* __restrict__ partial) {
  extern __shared__ __align__(16) uint8_t pf_raw[];
  PfSmem& S = *reinterpret_cast<PfSmem*>(pf_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t plane = (size_t)H * W;
  const int h = H / SP, w = W / SP;
  const int n_cells = B * h * w;
  const int step = gridDim.x;
  const float inv = 1.0f / (SP * SP);   // avg_pool2d

  auto cell_origin = [&](int c, int& n, size_t& pix0) {      // image index and pixel offset of the cell's top-left pixel
    n = c / (h * w);
    const int r = c - n * (h * w), cy = r / w, cx = r - cy * w;
    pix0 = (size_t)(cy * SP) * W + cx * SP;
  };
  auto issue_features = [&](int c, int stage) {
    if (c < n_cells) {
      int n; size_t pix0;
      cell_origin(c, n, pix0);
      const __nv_bfloat16* fb = feats + ((size_t)n * plane + pix0) * 64;
      const uint32_t fs = pf_smem_u32(S.F[stage]);
      for (int i = tid; i < 256 * 8; i += 160) {
        const int px = i >> 3, pc = i & 7;
        const __nv_bfloat16* src = fb + ((size_t)(px >> 4) * W + (px & 15)) * 64 + pc * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(fs + (uint32_t)(px * PF_FP + pc * 16)), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");       // (possibly empty) group: uniform counting
  };
  // registers that hold the NEXT cell's affinity (threads 0..127: one pixel pair, 9 directions) and ab (warp 4: 8 pixels)
  float2 paff[9], pab[2];
  auto prefetch_scalars = [&](int c) {
    if (c >= n_cells || tid >= 128) return;
    int n; size_t pix0;
    cell_origin(c, n, pix0);
    const int py = tid >> 3, px = (tid & 7) * 2;
    const size_t off = pix0 + (size_t)py * W + px;
    const float* a = aff + (size_t)n * 9 * plane + off;
#pragma unroll
    for (int k = 0; k < 9; ++k) paff[k] = *reinterpret_cast<const float2*>(a + (size_t)k * plane);
    pab[0] = ab ? *reinterpret_cast<const float2*>(ab + ((size_t)n * 2 + 0) * plane + off) : make_float2(0.f, 0.f);
    pab[1] = ab ? *reinterpret_cast<const float2*>(ab + ((size_t)n * 2 + 1) * plane + off) : make_float2(0.f, 0.f);
  };

  // rows 9..15 of the transposed affinity / indicator stay zero for the whole kernel
  for (int i = tid; i < 7 * (PF_AP / 16); i += 160) {
    const int r = 9 + i / (PF_AP / 16), c = i % (PF_AP / 16);
    *reinterpret_cast<uint4*>(S.Ahi + r * PF_AP + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(S.Alo + r * PF_AP + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(S.Hot + r * PF_AP + c * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  const int c0 = blockIdx.x;
  issue_features(c0, 0);
  prefetch_scalars(c0);

  int it = 0;
  for (int c = c0; c < n_cells; c += step, ++it) {
    const int stage = it % PF_STAGES;
    // ---- this cell's affinity transposed as bf16 hi | lo, the indicator of each pixel's maximal direction(s) (hard
    //      assignment, ties counted like the reference's ==), and the per-pixel "extra channels" (ab0, ab1, 1)
    if (tid < 128) {
      const int py = tid >> 3, px = (tid & 7) * 2, q = py * SP + px;
      float m0 = paff[0].x, m1 = paff[0].y;
#pragma unroll
      for (int k = 1; k < 9; ++k) { m0 = fmaxf(m0, paff[k].x); m1 = fmaxf(m1, paff[k].y); }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        uint32_t hi, lo;
        pf_split2(paff[k].x, paff[k].y, hi, lo);
        *reinterpret_cast<uint32_t*>(S.Ahi + k * PF_AP + q * 2) = hi;
        *reinterpret_cast<uint32_t*>(S.Alo + k * PF_AP + q * 2) = lo;
        *reinterpret_cast<uint32_t*>(S.Hot + k * PF_AP + q * 2) =
            (paff[k].x == m0 ? 0x3f80u : 0u) | (paff[k].y == m1 ? 0x3f800000u : 0u);       // bf16 1.0 = 0x3f80
      }
      uint32_t h0, l0, h1, l1;
      pf_split2(pab[0].x, pab[1].x, h0, l0);           // pixel q:     (ab0, ab1)
      pf_split2(pab[0].y, pab[1].y, h1, l1);           // pixel q + 1
      *reinterpret_cast<uint4*>(S.Ehi + q * 16) = make_uint4(h0, 0x3f80u, 0u, 0u);
      *reinterpret_cast<uint4*>(S.Elo + q * 16) = make_uint4(l0, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(S.Ehi + (q + 1) * 16) = make_uint4(h1, 0x3f80u, 0u, 0u);
      *reinterpret_cast<uint4*>(S.Elo + (q + 1) * 16) = make_uint4(l1, 0u, 0u, 0u);
    }
    prefetch_scalars(c + step);
    issue_features(c + step, (it + 1) % PF_STAGES);
    asm volatile("cp.async.wait_group 1;" ::: "memory");        // this cell's feature tile has landed (the next one may be pending)
    __syncthreads();

    int n; size_t pix0;
    cell_origin(c, n, pix0);
    float* out = partial + (size_t)c * 9 * PART;
    if (warp < 4) {
      // channels 16*warp .. 16*warp+15 (two n-tiles), all 16 k-steps of 16 pixels; hi and lo products in separate accumulators
      const int g = lane >> 2, t = lane & 3;
      const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1), lhalf = lane >> 4;
      const uint32_t a_hi = pf_smem_u32(S.Ahi) + (uint32_t)(lrow * PF_AP + lhalf * 16);
      const uint32_t a_lo = pf_smem_u32(S.Alo) + (uint32_t)(lrow * PF_AP + lhalf * 16);
      const uint32_t f_b = pf_smem_u32(S.F[stage]) + (uint32_t)(lrow * PF_FP + (16 * warp + 8 * lhalf) * 2);
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll 4
      for (int ks = 0; ks < 16; ++ks) {
        uint32_t ah[4], al[4], b[4];
        pf_ldsm(ah, a_hi + ks * 32);
        pf_ldsm(al, a_lo + ks * 32);
        pf_ldsm_t(b, f_b + ks * 16 * PF_FP);
        pf_mma(acc[0], ah, b[0], b[1]);
        pf_mma(acc[1], al, b[0], b[1]);
        pf_mma(acc[2], ah, b[2], b[3]);
        pf_mma(acc[3], al, b[2], b[3]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int ch = 16 * warp + 8 * j + 2 * t;
        *reinterpret_cast<float2*>(out + g * PART + ch) =
            make_float2((acc[2 * j][0] + acc[2 * j + 1][0]) * inv, (acc[2 * j][1] + acc[2 * j + 1][1]) * inv);
        if (g == 0)
          *reinterpret_cast<float2*>(out + 8 * PART + ch) =
              make_float2((acc[2 * j][2] + acc[2 * j + 1][2]) * inv, (acc[2 * j][3] + acc[2 * j + 1][3]) * inv);
      }
    } else {
      // fifth warp: the four scalar masses as one more n-tile.  columns (ab0, ab1, 1): sum_px p_k * (ab0, ab1, 1) with both
      // operands split (p_hi e_hi + p_lo e_hi + p_hi e_lo); hard mass = indicator^T x ones (exact)
      const int g = lane >> 2, t = lane & 3;
      const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1), lhalf = lane >> 4;
      const uint32_t a_hi = pf_smem_u32(S.Ahi) + (uint32_t)(lrow * PF_AP + lhalf * 16);
      const uint32_t a_lo = pf_smem_u32(S.Alo) + (uint32_t)(lrow * PF_AP + lhalf * 16);
      const uint32_t a_ho = pf_smem_u32(S.Hot) + (uint32_t)(lrow * PF_AP + lhalf * 16);
      // lanes 0..15: rows (pixels) of e_hi, lanes 16..31: rows of e_lo -> {b0, b1} of e_hi, {b0, b1} of e_lo
      const uint32_t e_b = (lane < 16 ? pf_smem_u32(S.Ehi) : pf_smem_u32(S.Elo)) + (uint32_t)((lane & 15) * 16);
      float acc[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll 4
      for (int ks = 0; ks < 16; ++ks) {
        uint32_t ah[4], al[4], ho[4], e[4];
        pf_ldsm(ah, a_hi + ks * 32);
        pf_ldsm(al, a_lo + ks * 32);
        pf_ldsm(ho, a_ho + ks * 32);
        pf_ldsm_t(e, e_b + ks * 16 * 16);
        pf_mma(acc[0], ah, e[0], e[1]);
        pf_mma(acc[1], al, e[0], e[1]);
        pf_mma(acc[2], ah, e[2], e[3]);
        pf_mma(acc[3], ho, e[0], e[1]);
      }
      // accumulator columns 2t, 2t+1: t == 0 -> (ab0, ab1), t == 1 -> (mass | hard count, unused)
      if (t == 0) {
        *reinterpret_cast<float2*>(out + g * PART + 64) =
            make_float2((acc[0][0] + acc[1][0] + acc[2][0]) * inv, (acc[0][1] + acc[1][1] + acc[2][1]) * inv);
        if (g == 0)
          *reinterpret_cast<float2*>(out + 8 * PART + 64) =
              make_float2((acc[0][2] + acc[1][2] + acc[2][2]) * inv, (acc[0][3] + acc[1][3] + acc[2][3]) * inv);
      } else if (t == 1) {
        *reinterpret_cast<float2*>(out + g * PART + 66) = make_float2((acc[0][0] + acc[1][0] + acc[2][0]) * inv, acc[3][0] * inv);
        if (g == 0)
          *reinterpret_cast<float2*>(out + 8 * PART + 66) = make_float2((acc[0][2] + acc[1][2] + acc[2][2]) * inv, acc[3][2] * inv);
      }
    }
    __syncthreads();             // the affinity buffers and this feature stage are free for the next iterations
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
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
  else if (getenv("DISCO_POOL_MMA") == nullptr || getenv("DISCO_POOL_MMA")[0] != '0') {
    if (int rc = disco_ensure_smem(h, (const void*)poolfeat_partial_mma_kernel, (int)sizeof(PfSmem))) return rc;
    const int cells = batch * (H / SP) * (W / SP);
    poolfeat_partial_mma_kernel<<<cells < 2 * h->sm_count ? cells : 2 * h->sm_count, 160, sizeof(PfSmem), st>>>(
        (const __nv_bfloat16*)feats, ab, affinity, batch, H, W, partial);
  } else
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

// ---------------------------------------------------------------------------------------------
// split_spixels of the SpixelSeg inference script (reference main/spixelseg/inference.py:67-75): winner-take-all super-pixel
// id per pixel.  The 9-channel neighbour-id map of basic.init_spixel_grid (models/basic.py:221-251: cell ids with edge
// replication) is evaluated on the fly; like the reference, ids of ALL channels that equal the maximum are summed.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void spixel_ids_kernel(const float* __restrict__ prob, int B, int H, int W, int sp, int32_t* __restrict__ ids) {
  const size_t plane = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * plane) return;
  const size_t n = idx / plane, p = idx - n * plane;
  const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
  const int nh = H / sp, nw = W / sp, cy = y / sp, cx = x / sp;
  const float* pr = prob + n * 9 * plane + p;
  float v[9], mx = -3.4e38f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { v[k] = pr[(size_t)k * plane]; mx = fmaxf(mx, v[k]); }
  int id = 0;
#pragma unroll
  for (int k = 0; k < 9; ++k)
    if (v[k] == mx) id += min(max(cy + k / 3 - 1, 0), nh - 1) * nw + min(max(cx + k % 3 - 1, 0), nw - 1);
  ids[idx] = id;
}
}  // namespace

extern "C" int disco_spixel_ids(disco_handle* h, const float* prob, int batch, int H, int W, int sp, int32_t* ids, void* stream) {
  DISCO_CHECK_ARG(h && prob && ids, "spixel_ids: null pointer");
  DISCO_CHECK_ARG(batch > 0 && sp > 0 && H >= sp && W >= sp && H % sp == 0 && W % sp == 0,
                  "spixel_ids: H, W must be positive multiples of the super-pixel size (got %dx%d, %d)", H, W, sp);
  DiscoDeviceGuard guard(h);
  const size_t n = (size_t)batch * H * W;
  spixel_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(prob, batch, H, W, sp, ids);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}
