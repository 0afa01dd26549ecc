// Fused token transformer: ONE kernel runs a whole stack of post-norm encoder layers
// (TransformerEncoder / EncoderLayer, reference models/transformer2d.py:17-28,52-60: MHA(q = k = x + pos, v = x; 8 heads of
// 8 channels) -> +residual -> LayerNorm -> FFN 64->256->64 (ReLU) -> +residual -> LayerNorm, eval mode, no mask).
//
// Work split.  One CTA = 128 consecutive tokens of one image, one warp = 16 tokens.  A warp keeps its 16 x 64 token state in
// registers for the whole stack and chains every product through the register fragments of mma.sync (the C fragment of
// one m16n8 tile is the A fragment of the next product): QKV projection, Q K^T, softmax (online, in registers), P V, output
// projection, both FFN GEMMs and both LayerNorms never leave the register file.  Shared memory only stages the operands all
// warps share -- weight chunks (64 x 64) and key/value blocks (32 keys) -- through one 4-slot cp.async ring.
//
// Images with more than 128 tokens span a thread-block cluster (S = 256 -> 2 CTAs, S = 1024 -> 8): every CTA writes the
// keys/values of its own tokens to a ping-pong scratch buffer in global memory (L2-resident), the cluster meets at one
// barrier.cluster per layer (arrive right after the K/V stores, wait after the Q projection, so the skew is hidden), then
// streams all key/value blocks of the image back through the ring.
//
// Precision.  The token path feeds discrete decisions (k-means anchors, arg-max labels), so it keeps fp32-grade accuracy on
// bf16 tensor cores: every operand is split x = hi + lo (two bf16) and each product is hi*hi + lo*hi + hi*lo with fp32
// accumulation -- relative error ~2^-16 per product instead of 2^-9.  Softmax statistics, LayerNorm and residuals are fp32.
// The one product whose left operand is bounded -- P V, with the probabilities relative to a running reference in
// (0, 2^10] -- runs on the f16 tensor path instead: P as a single fp16 (11 significant bits) against V split into fp16
// hi + lo, two MMAs per k-step and one conversion per pair instead of three MMAs and a six-instruction split.
#include "common.cuh"
#include <cuda_fp16.h>
#include <cstring>

namespace {

#ifndef ES_PV_F16
#define ES_PV_F16 1          // 1: P V on the f16 path (P single fp16, V fp16 hi + lo); 0: bf16 path, P and V both split (3 MMAs)
#endif
constexpr int ES_THREADS = 256;
constexpr int ES_D = 4;                    // ring slots
constexpr int ES_ROWB = 144;               // shared-memory row pitch: 64 bf16 + 8 padding -> conflict-free ldmatrix
constexpr int ES_SLOT = 128 * ES_ROWB;     // 18432 B = weight chunk (hi|lo x 64 rows) = K/V block (Khi|Klo|Vhi|Vlo x 32 keys)
constexpr int ES_VEC = 832;                // per-layer vectors: bq' bk bv bo (4x64) | b1 (256) | b2 | g1 be1 g2 be2 (5x64)
constexpr int ES_CHUNKS = 12;              // Wk Wv Wq' Wo (W1_i W2_i) x 4, 64 x 64 each, in consumption order
constexpr int ES_SMEM = ES_D * ES_SLOT + 2 * ES_VEC * 4;

struct EsParams {
  const float* x_in;       // [B, S, 64]
  const float* pos;        // [S, 64]
  const uint16_t* w;       // [L][12][2 (hi, lo)][64][64] bf16
  const float* vec;        // [L][832]
  uint16_t* kv;            // [2][B][4 (Khi, Klo, Vhi, Vlo)][csize * 128][64] bf16
  float* y;                // [B, S, 64]
  int B, S, n_layers, csize, n_kv, srows;
};

__device__ __forceinline__ uint32_t es_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(int n) {   // at most n groups still in flight
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
// D(16x8, fp32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// same shape, fp16 operands
__device__ __forceinline__ void mma16816h(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float x0, float x1) {      // fp16 pair, element 0 in the low half-word
  __half2 h = __floats2half2_rn(x0, x1);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void split2h(float x0, float x1, uint32_t& hi, uint32_t& lo) {   // x ~= hi + lo, both fp16
  __half2 h = __floats2half2_rn(x0, x1);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float2 hf = __half22float2(h);
  lo = pack_h2(x0 - hf.x, x1 - hf.y);
}
__device__ __forceinline__ float es_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void es_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void es_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t es_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// (x0, x1) -> packed bf16 pairs hi and lo with x ~= hi + lo (element 0 in the low half-word)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(x0 - h0, x1 - h1);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// 16 x 64 tile in accumulator layout (c[j] = columns 8j..8j+7: [0],[1] row g, [2],[3] row g+8, columns 2t, 2t+1)
// -> A fragments of the four k16 steps of the next product
__device__ __forceinline__ void make_frags(const float (&c)[8][4], uint32_t (&hi)[4][4], uint32_t (&lo)[4][4]) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    split2(c[2 * s][0], c[2 * s][1], hi[s][0], lo[s][0]);
    split2(c[2 * s][2], c[2 * s][3], hi[s][1], lo[s][1]);
    split2(c[2 * s + 1][0], c[2 * s + 1][1], hi[s][2], lo[s][2]);
    split2(c[2 * s + 1][2], c[2 * s + 1][3], hi[s][3], lo[s][3]);
  }
}

// acc(16 x 64) += A(16 x 64) * W^T, W = [64 outputs][64 inputs] staged at `slot` (hi rows, then lo rows at +64 rows)
__device__ __forceinline__ void gemm_chunk(float (&acc)[8][4], const uint32_t (&ahi)[4][4], const uint32_t (&alo)[4][4],
                                           uint32_t slot, int lane) {
  const uint32_t lane_off = (uint32_t)((lane & 7) * ES_ROWB + (lane >> 3) * 16);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int sp = 0; sp < 2; ++sp) {
      uint32_t bh[4], bl[4];
      const uint32_t a = slot + (uint32_t)(j * 8 * ES_ROWB + sp * 64) + lane_off;
      ldsm_x4(bh, a);
      ldsm_x4(bl, a + 64 * ES_ROWB);
      mma16816(acc[j], ahi[2 * sp], bh[0], bh[1]);
      mma16816(acc[j], alo[2 * sp], bh[0], bh[1]);
      mma16816(acc[j], ahi[2 * sp], bl[0], bl[1]);
      mma16816(acc[j], ahi[2 * sp + 1], bh[2], bh[3]);
      mma16816(acc[j], alo[2 * sp + 1], bh[2], bh[3]);
      mma16816(acc[j], ahi[2 * sp + 1], bl[2], bl[3]);
    }
  }
}

__device__ __forceinline__ void zero_tile(float (&c)[8][4]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f;
}

// LayerNorm over the 64 columns of both rows a thread shares with its 3 quad neighbours (eps 1e-5, biased variance)
__device__ __forceinline__ void layer_norm(float (&x)[8][4], const float* gamma, const float* beta, int t) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { s0 += x[j][0] + x[j][1]; s1 += x[j][2] + x[j][3]; }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
  s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float m0 = s0 * (1.f / 64.f), m1 = s1 * (1.f / 64.f);
  float v0 = 0.f, v1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float d;
    d = x[j][0] - m0; v0 = fmaf(d, d, v0); d = x[j][1] - m0; v0 = fmaf(d, d, v0);
    d = x[j][2] - m1; v1 = fmaf(d, d, v1); d = x[j][3] - m1; v1 = fmaf(d, d, v1);
  }
  v0 += __shfl_xor_sync(0xffffffffu, v0, 1); v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
  v0 += __shfl_xor_sync(0xffffffffu, v0, 2); v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
  const float r0 = 1.0f / sqrtf(v0 * (1.f / 64.f) + 1e-5f), r1 = 1.0f / sqrtf(v1 * (1.f / 64.f) + 1e-5f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 g = *reinterpret_cast<const float2*>(gamma + 8 * j + 2 * t);
    const float2 b = *reinterpret_cast<const float2*>(beta + 8 * j + 2 * t);
    x[j][0] = (x[j][0] - m0) * r0 * g.x + b.x; x[j][1] = (x[j][1] - m0) * r0 * g.y + b.y;
    x[j][2] = (x[j][2] - m1) * r1 * g.x + b.x; x[j][3] = (x[j][3] - m1) * r1 * g.y + b.y;
  }
}

// MINB = CTAs per SM the register allocation is capped for.  1: 255 registers, no spills -- fastest per CTA, used when the
// launch has no more CTAs than SMs (batch 64 at 256x256: 128 CTAs).  2: 128 registers (~0.5 KB of spills per thread) so that
// two CTAs share an SM -- 16 warps per SM hide the dependency latencies of the register-chained MMAs, and a launch with more
// CTAs than SMs (batch 32 at 512x512: 256 CTAs in 8-CTA clusters) runs in one wave: 1.83 -> 1.11 ms per stack.
template <int MINB>
__global__ void __launch_bounds__(ES_THREADS, MINB) encoder_stack_kernel(const EsParams P) {
  extern __shared__ __align__(128) uint8_t es_smem[];
  const uint32_t ring = es_smem_u32(es_smem);
  float* sv = reinterpret_cast<float*>(es_smem + ES_D * ES_SLOT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int crank = (int)es_cluster_ctarank();
  const int b = blockIdx.x / P.csize;
  const int S = P.S, n_kv = P.n_kv;
  const int row0 = crank * 128 + warp * 16 + g, row1 = row0 + 8;     // this thread's two tokens
  const bool ok0 = row0 < S, ok1 = row1 < S;
  const int ipl = ES_CHUNKS + n_kv;                                  // ring items per layer
  const int total = ipl * P.n_layers;

  // ---- operand ring: item gi = (layer, j); j = 0,1,2: Wk Wv Wq | 3..3+n_kv-1: K/V blocks | then Wo, (W1_i, W2_i) x 4
  int issued = 0, synced = -1, kv_ready = -1;
  auto issue = [&](int gi) {
    const int l = gi / ipl, j = gi - l * ipl;
    const uint32_t slot = ring + (uint32_t)(gi % ES_D) * ES_SLOT;
    if (j < 3 || j >= 3 + n_kv) {
      const int c = j < 3 ? j : j - n_kv;
      const uint16_t* src = P.w + (size_t)(l * ES_CHUNKS + c) * 2 * 4096;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int p = tid + ES_THREADS * q, arr = p >> 9, r = (p >> 3) & 63, pc = p & 7;
        cp_async16(slot + (uint32_t)(arr * 64 * ES_ROWB + r * ES_ROWB + pc * 16), src + arr * 4096 + r * 64 + pc * 8);
      }
    } else {
      const int kb = j - 3;
      const uint16_t* base = P.kv + ((size_t)((l & 1) * P.B + b) * 4) * P.srows * 64;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int p = tid + ES_THREADS * q, arr = p >> 8, r = (p >> 3) & 31, pc = p & 7;
        cp_async16(slot + (uint32_t)(arr * 32 * ES_ROWB + r * ES_ROWB + pc * 16),
                   base + ((size_t)arr * P.srows + kb * 32 + r) * 64 + pc * 8);
      }
    }
    cp_async_commit();
  };
  // every warp has reached item `synced`, so the slots of all earlier items are free: items up to synced + D - 1 may be
  // in flight; key/value blocks of a layer only after that layer's cluster barrier
  auto try_issue = [&]() {
    while (issued < total && issued < synced + ES_D) {
      const int l = issued / ipl, j = issued - l * ipl;
      if (j >= 3 && j < 3 + n_kv && kv_ready < l) break;
      issue(issued);
      ++issued;
    }
  };
  // call before computing item i: its operands have landed for every thread, and the slot of item i-1 is free again
  auto advance = [&](int i) -> uint32_t {
    cp_async_wait_dyn(issued - i - 1);
    __syncthreads();
    synced = i;
    try_issue();
    return ring + (uint32_t)(i % ES_D) * ES_SLOT;
  };

  // ---- token state (accumulator layout)
  float x[8][4];
  {
    const float* xr0 = P.x_in + ((size_t)b * S + (ok0 ? row0 : 0)) * 64 + 2 * t;
    const float* xr1 = P.x_in + ((size_t)b * S + (ok1 ? row1 : 0)) * 64 + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 a = ok0 ? *reinterpret_cast<const float2*>(xr0 + 8 * j) : make_float2(0.f, 0.f);
      const float2 c = ok1 ? *reinterpret_cast<const float2*>(xr1 + 8 * j) : make_float2(0.f, 0.f);
      x[j][0] = a.x; x[j][1] = a.y; x[j][2] = c.x; x[j][3] = c.y;
    }
  }
  const float* pr0 = P.pos + (size_t)(ok0 ? row0 : 0) * 64 + 2 * t;
  const float* pr1 = P.pos + (size_t)(ok1 ? row1 : 0) * 64 + 2 * t;
  try_issue();

  for (int l = 0; l < P.n_layers; ++l) {
    float* svl = sv + (l & 1) * ES_VEC;
    for (int i = tid; i < ES_VEC; i += ES_THREADS) svl[i] = P.vec[(size_t)l * ES_VEC + i];   // visible after the next advance()
    const int it = l * ipl;
    uint16_t* kvb = P.kv + ((size_t)((l & 1) * P.B + b) * 4) * P.srows * 64;
    uint32_t qh[8][2], ql[8][2];
    {
      // ---- K = (x + pos) Wk^T + bk,  V = x Wv^T + bv  -> scratch (hi | lo), then Q' = (x + pos) Wq'^T + bq'
      uint32_t phi[4][4], plo[4][4];
      {
        float xp[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 a = *reinterpret_cast<const float2*>(pr0 + 8 * j), c = *reinterpret_cast<const float2*>(pr1 + 8 * j);
          xp[j][0] = x[j][0] + a.x; xp[j][1] = x[j][1] + a.y; xp[j][2] = x[j][2] + c.x; xp[j][3] = x[j][3] + c.y;
        }
        make_frags(xp, phi, plo);
      }
      float acc[8][4];
      auto store_kv = [&](int arr_hi, const float* bias, bool as_f16) {
        uint16_t* d0 = kvb + ((size_t)arr_hi * P.srows + row0) * 64 + 2 * t;
        uint16_t* d1 = kvb + ((size_t)arr_hi * P.srows + row1) * 64 + 2 * t;
        const size_t lo_off = (size_t)P.srows * 64;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 bb = *reinterpret_cast<const float2*>(bias + 8 * j + 2 * t);
          uint32_t h, lw;
          if (as_f16) split2h(acc[j][0] + bb.x, acc[j][1] + bb.y, h, lw); else split2(acc[j][0] + bb.x, acc[j][1] + bb.y, h, lw);
          *reinterpret_cast<uint32_t*>(d0 + 8 * j) = h;
          *reinterpret_cast<uint32_t*>(d0 + lo_off + 8 * j) = lw;
          if (as_f16) split2h(acc[j][2] + bb.x, acc[j][3] + bb.y, h, lw); else split2(acc[j][2] + bb.x, acc[j][3] + bb.y, h, lw);
          *reinterpret_cast<uint32_t*>(d1 + 8 * j) = h;
          *reinterpret_cast<uint32_t*>(d1 + lo_off + 8 * j) = lw;
        }
      };
      uint32_t slot = advance(it + 0);
      zero_tile(acc);
      gemm_chunk(acc, phi, plo, slot, lane);
      store_kv(0, svl + 64, false);      // keys: bf16 hi | lo
      {
        uint32_t xhi[4][4], xlo[4][4];
        make_frags(x, xhi, xlo);
        slot = advance(it + 1);
        zero_tile(acc);
        gemm_chunk(acc, xhi, xlo, slot, lane);
        store_kv(2, svl + 128, ES_PV_F16 != 0);     // values: fp16 hi | lo (the P V product runs on the f16 path)
      }
      es_cluster_arrive();                       // this CTA's keys/values of layer l are written
      slot = advance(it + 2);
      zero_tile(acc);
      gemm_chunk(acc, phi, plo, slot, lane);
#pragma unroll
      for (int h = 0; h < 8; ++h) {              // head h = columns 8h..8h+7 = accumulator tile h = one k8 A fragment
        const float2 bb = *reinterpret_cast<const float2*>(svl + 8 * h + 2 * t);
        split2(acc[h][0] + bb.x, acc[h][1] + bb.y, qh[h][0], ql[h][0]);
        split2(acc[h][2] + bb.x, acc[h][3] + bb.y, qh[h][1], ql[h][1]);
      }
    }
    es_cluster_wait();                           // every CTA of the image has written its keys/values
    kv_ready = l;
    try_issue();

    // ---- attention over all keys of the image, 32 per ring item; scores are in log2 units (scale folded into Wq').
    // Softmax against a LAZY reference: exp2(s - m) / sum exp2(s - m) is the same for any m, so the reference of a row is
    // re-centred (cross-lane max, rescale of O and the row sum) only for the first block and when a score exceeds it by
    // more than 2^10 -- one warp vote per (head, block) instead of a shuffle / max / exp2 / rescale chain.  The reference
    // enters through the accumulator's initial value (-m), so the MMA delivers s - m directly.
    float o[8][4], mrow[8][2], lsum[8][2];
    zero_tile(o);
#pragma unroll
    for (int h = 0; h < 8; ++h) { mrow[h][0] = mrow[h][1] = 0.f; lsum[h][0] = lsum[h][1] = 0.f; }
    for (int kb = 0; kb < n_kv; ++kb) {
      const uint32_t slot = advance(it + 3 + kb);
      const uint32_t la = slot + (uint32_t)lane * ES_ROWB;
      const bool tail = kb * 32 + 32 > S;
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        uint32_t kh[4], kl[4];
        ldsm_x4(kh, la + h * 16);
        ldsm_x4(kl, la + 32 * ES_ROWB + h * 16);
        // head dimension 8 = half a k16 step: A = [q_hi | q_lo] against B = [k_hi ; k_hi] gives q_hi k_hi + q_lo k_hi in one
        // instruction, against [k_lo ; k_lo] the remaining cross term (+ the negligible lo*lo).  Measured on B200: an
        // m16n8k8 issues at the same 2 clk/SM as an m16n8k16, so three k8 products would cost 6 clk, these two cost 4.
        const uint32_t qa[4] = {qh[h][0], qh[h][1], ql[h][0], ql[h][1]};
        float s[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          s[n][0] = s[n][1] = -mrow[h][0];
          s[n][2] = s[n][3] = -mrow[h][1];
          mma16816(s[n], qa, kh[n], kh[n]);
          mma16816(s[n], qa, kl[n], kl[n]);
        }
        if (tail) {
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const int key = kb * 32 + 8 * n + 2 * t;
            if (key >= S) { s[n][0] = -INFINITY; s[n][2] = -INFINITY; }
            if (key + 1 >= S) { s[n][1] = -INFINITY; s[n][3] = -INFINITY; }
          }
        }
        float mx0 = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]));
        float mx1 = fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3]));
        mx0 = fmaxf(mx0, fmaxf(fmaxf(s[2][0], s[2][1]), fmaxf(s[3][0], s[3][1])));
        mx1 = fmaxf(mx1, fmaxf(fmaxf(s[2][2], s[2][3]), fmaxf(s[3][2], s[3][3])));
        if (kb == 0 || __any_sync(0xffffffffu, fmaxf(mx0, mx1) > 10.0f)) {
          // re-centre: row maxima over the quad; block 0 always (the initial reference 0 may be far off either way),
          // later blocks only upwards
          mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
          mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
          const float d0 = (kb == 0 || mx0 > 0.f) ? mx0 : 0.f, d1 = (kb == 0 || mx1 > 0.f) ? mx1 : 0.f;   // finite: a valid key exists
          const float c0 = kb == 0 ? 1.f : es_ex2(-d0), c1 = kb == 0 ? 1.f : es_ex2(-d1);   // block 0: O and the sums are still zero
          mrow[h][0] += d0; mrow[h][1] += d1;
          lsum[h][0] *= c0; lsum[h][1] *= c1;
          o[h][0] *= c0; o[h][1] *= c0; o[h][2] *= c1; o[h][3] *= c1;
#pragma unroll
          for (int n = 0; n < 4; ++n) { s[n][0] -= d0; s[n][1] -= d0; s[n][2] -= d1; s[n][3] -= d1; }
        }
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          s[n][0] = es_ex2(s[n][0]); s[n][1] = es_ex2(s[n][1]);
          s[n][2] = es_ex2(s[n][2]); s[n][3] = es_ex2(s[n][3]);
          rs0 += s[n][0] + s[n][1]; rs1 += s[n][2] + s[n][3];
        }
        lsum[h][0] += rs0; lsum[h][1] += rs1;      // per-thread partial row sums
        uint32_t vh[4], vl[4];
        ldsm_x4_t(vh, la + 64 * ES_ROWB + h * 16);
        ldsm_x4_t(vl, la + 96 * ES_ROWB + h * 16);
#pragma unroll
        for (int u = 0; u < 2; ++u) {            // keys 16u..16u+15 of the block = score tiles 2u, 2u+1
#if ES_PV_F16
          const uint32_t pf[4] = {pack_h2(s[2 * u][0], s[2 * u][1]), pack_h2(s[2 * u][2], s[2 * u][3]),
                                  pack_h2(s[2 * u + 1][0], s[2 * u + 1][1]), pack_h2(s[2 * u + 1][2], s[2 * u + 1][3])};
          mma16816h(o[h], pf, vh[2 * u], vh[2 * u + 1]);
          mma16816h(o[h], pf, vl[2 * u], vl[2 * u + 1]);
#else
          uint32_t ph[4], pl[4];
          split2(s[2 * u][0], s[2 * u][1], ph[0], pl[0]);
          split2(s[2 * u][2], s[2 * u][3], ph[1], pl[1]);
          split2(s[2 * u + 1][0], s[2 * u + 1][1], ph[2], pl[2]);
          split2(s[2 * u + 1][2], s[2 * u + 1][3], ph[3], pl[3]);
          mma16816(o[h], ph, vh[2 * u], vh[2 * u + 1]);
          mma16816(o[h], pl, vh[2 * u], vh[2 * u + 1]);
          mma16816(o[h], ph, vl[2 * u], vl[2 * u + 1]);
#endif
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      float l0 = lsum[h][0], l1 = lsum[h][1];
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float i0 = 1.0f / l0, i1 = 1.0f / l1;
      o[h][0] *= i0; o[h][1] *= i0; o[h][2] *= i1; o[h][3] *= i1;
    }

    // ---- x1 = LayerNorm1(x + o Wo^T + bo)
    {
      uint32_t ohi[4][4], olo[4][4];
      make_frags(o, ohi, olo);
      const uint32_t slot = advance(it + 3 + n_kv);
      float acc[8][4];
      zero_tile(acc);
      gemm_chunk(acc, ohi, olo, slot, lane);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 bb = *reinterpret_cast<const float2*>(svl + 192 + 8 * j + 2 * t);
        x[j][0] += acc[j][0] + bb.x; x[j][1] += acc[j][1] + bb.y; x[j][2] += acc[j][2] + bb.x; x[j][3] += acc[j][3] + bb.y;
      }
      layer_norm(x, svl + 576, svl + 640, t);
    }

    // ---- x = LayerNorm2(x1 + relu(x1 W1^T + b1) W2^T + b2), hidden units in four chunks of 64
    {
      uint32_t xhi[4][4], xlo[4][4];
      make_frags(x, xhi, xlo);
      float yacc[8][4];
      zero_tile(yacc);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t slot = advance(it + 4 + n_kv + 2 * c);
        float hacc[8][4];
        zero_tile(hacc);
        gemm_chunk(hacc, xhi, xlo, slot, lane);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 bb = *reinterpret_cast<const float2*>(svl + 256 + 64 * c + 8 * j + 2 * t);
          hacc[j][0] = fmaxf(hacc[j][0] + bb.x, 0.f); hacc[j][1] = fmaxf(hacc[j][1] + bb.y, 0.f);
          hacc[j][2] = fmaxf(hacc[j][2] + bb.x, 0.f); hacc[j][3] = fmaxf(hacc[j][3] + bb.y, 0.f);
        }
        uint32_t hhi[4][4], hlo[4][4];
        make_frags(hacc, hhi, hlo);
        slot = advance(it + 5 + n_kv + 2 * c);
        gemm_chunk(yacc, hhi, hlo, slot, lane);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 bb = *reinterpret_cast<const float2*>(svl + 512 + 8 * j + 2 * t);
        x[j][0] += yacc[j][0] + bb.x; x[j][1] += yacc[j][1] + bb.y; x[j][2] += yacc[j][2] + bb.x; x[j][3] += yacc[j][3] + bb.y;
      }
      layer_norm(x, svl + 704, svl + 768, t);
    }
  }

  {
    float* y0 = P.y + ((size_t)b * S + (ok0 ? row0 : 0)) * 64 + 2 * t;
    float* y1 = P.y + ((size_t)b * S + (ok1 ? row1 : 0)) * 64 + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (ok0) *reinterpret_cast<float2*>(y0 + 8 * j) = make_float2(x[j][0], x[j][1]);
      if (ok1) *reinterpret_cast<float2*>(y1 + 8 * j) = make_float2(x[j][2], x[j][3]);
    }
  }
  // no CTA of the cluster may exit while a peer can still wait at a barrier it has to join
  es_cluster_arrive();
  es_cluster_wait();
}

}  // namespace

extern "C" int64_t disco_encoder_stack_scratch_elems(int batch, int S) {
  if (batch <= 0 || S <= 0) return 0;
  const int64_t csize = (S + 127) / 128;
  return 2 * (int64_t)batch * 4 * csize * 128 * 64;
}

extern "C" int disco_encoder_stack_pack(const disco_encoder_layer_weights* layers, int n_layers, uint16_t* w_out, float* vec_out) {
  DISCO_CHECK_ARG(layers && w_out && vec_out && n_layers > 0, "encoder_stack_pack: bad argument");
  const float qscale = 0.35355339059327373f * 1.4426950408889634f;     // 8^-0.5 * log2(e): softmax runs on exp2
  auto put = [&](uint16_t* dst, const float* src, int ld, float scale) {  // dst: hi[64][64] then lo[64][64]; src[n][k] with row pitch ld
    for (int n = 0; n < 64; ++n)
      for (int k = 0; k < 64; ++k) {
        const float v = src[(size_t)n * ld + k] * scale;
        __nv_bfloat16 hb = __float2bfloat16_rn(v);
        __nv_bfloat16 lb = __float2bfloat16_rn(v - __bfloat162float(hb));
        memcpy(dst + n * 64 + k, &hb, 2);
        memcpy(dst + 4096 + n * 64 + k, &lb, 2);
      }
  };
  for (int l = 0; l < n_layers; ++l) {
    const disco_encoder_layer_weights& L = layers[l];
    DISCO_CHECK_ARG(L.in_w && L.in_b && L.out_w && L.out_b && L.l1_w && L.l1_b && L.l2_w && L.l2_b && L.n1_w && L.n1_b && L.n2_w && L.n2_b,
                    "encoder_stack_pack: layer %d has a null pointer", l);
    uint16_t* w = w_out + (size_t)l * ES_CHUNKS * 8192;
    put(w + 0 * 8192, L.in_w + 64 * 64, 64, 1.f);          // Wk
    put(w + 1 * 8192, L.in_w + 128 * 64, 64, 1.f);         // Wv
    put(w + 2 * 8192, L.in_w, 64, qscale);                 // Wq'
    put(w + 3 * 8192, L.out_w, 64, 1.f);                   // Wo
    for (int c = 0; c < 4; ++c) {
      put(w + (4 + 2 * c) * 8192, L.l1_w + (size_t)c * 64 * 64, 64, 1.f);      // W1 rows 64c..64c+63
      put(w + (5 + 2 * c) * 8192, L.l2_w + c * 64, 256, 1.f);                  // W2 columns 64c..64c+63
    }
    float* v = vec_out + (size_t)l * ES_VEC;
    for (int i = 0; i < 64; ++i) {
      v[i] = L.in_b[i] * qscale; v[64 + i] = L.in_b[64 + i]; v[128 + i] = L.in_b[128 + i]; v[192 + i] = L.out_b[i];
      v[512 + i] = L.l2_b[i]; v[576 + i] = L.n1_w[i]; v[640 + i] = L.n1_b[i]; v[704 + i] = L.n2_w[i]; v[768 + i] = L.n2_b[i];
    }
    for (int i = 0; i < 256; ++i) v[256 + i] = L.l1_b[i];
  }
  return DISCO_OK;
}

extern "C" int disco_encoder_stack(disco_handle* h, const float* x_in, const float* pos, const uint16_t* w_packed, const float* vec,
                                   int n_layers, int batch, int S, uint16_t* kv_scratch, float* y, void* stream) {
  DISCO_CHECK_ARG(h && x_in && pos && w_packed && vec && kv_scratch && y, "encoder_stack: null pointer");
  DISCO_CHECK_ARG(n_layers > 0 && batch > 0 && S > 0, "encoder_stack: bad shape");
  DISCO_CHECK_ARG(S <= 1024, "encoder_stack: S = %d tokens exceeds the 8-CTA cluster an image may span (1024)", S);
  DiscoDeviceGuard guard(h);
  EsParams P;
  P.x_in = x_in; P.pos = pos; P.w = w_packed; P.vec = vec; P.kv = kv_scratch; P.y = y;
  P.B = batch; P.S = S; P.n_layers = n_layers;
  P.csize = (S + 127) / 128;
  P.n_kv = (S + 31) / 32;
  P.srows = P.csize * 128;
  const bool two_per_sm = batch * P.csize > h->sm_count;
  const void* kern = two_per_sm ? (const void*)encoder_stack_kernel<2> : (const void*)encoder_stack_kernel<1>;
  if (int rc = disco_ensure_smem(h, kern, ES_SMEM)) return rc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(batch * P.csize), 1, 1);
  cfg.blockDim = dim3(ES_THREADS, 1, 1);
  cfg.dynamicSmemBytes = ES_SMEM;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)P.csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (two_per_sm) DISCO_CUDA(cudaLaunchKernelEx(&cfg, encoder_stack_kernel<2>, P));
  else DISCO_CUDA(cudaLaunchKernelEx(&cfg, encoder_stack_kernel<1>, P));
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}
