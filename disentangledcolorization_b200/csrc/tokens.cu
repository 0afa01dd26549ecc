// Token path: linear layers with fused epilogues, attention core, batched k-means anchor selection,
// token labels.  Everything is fp32 (the token path is 0.2 % of the FLOPs and feeds discrete
// decisions -- k-means assignments, argmax labels -- so it is kept in full precision).
#include "common.cuh"
#include <cfloat>
#include <cstdlib>

namespace {

// ------------------------------------------------------------------------------------------
// linear: 64x64 output tile per CTA, K stepped by 16, 4x4 register micro-tiles
// ------------------------------------------------------------------------------------------
constexpr int LT = 64;
constexpr int LK = 16;

__global__ void __launch_bounds__(256) linear_kernel(const disco_linear_desc d) {
  __shared__ float As[LK][LT + 4];
  __shared__ float Bs[LK][LT + 4];
  __shared__ float Cs[LT][LT + 1];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * LT, n0 = blockIdx.y * LT;
  const int tx = tid & 15, ty = tid >> 4;
  const bool use_pos = d.pos != nullptr && n0 < d.pos_cols;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = tid >> 2, lc = (tid & 3) * 4;   // load role: row lr, k offset lc..lc+3
  for (int k0 = 0; k0 < d.K; k0 += LK) {
    {
      const int row = m0 + lr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < d.M) {
        v = *reinterpret_cast<const float4*>(d.X + (size_t)row * d.K + k0 + lc);
        if (use_pos) {
          const float4 p = *reinterpret_cast<const float4*>(d.pos + (size_t)(row % d.S) * d.K + k0 + lc);
          v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
        }
      }
      As[lc + 0][lr] = v.x; As[lc + 1][lr] = v.y; As[lc + 2][lr] = v.z; As[lc + 3][lr] = v.w;
      const int col = n0 + lr;
      float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col < d.N) wv = *reinterpret_cast<const float4*>(d.W + (size_t)col * d.K + k0 + lc);
      Bs[lc + 0][lr] = wv.x; Bs[lc + 1][lr] = wv.y; Bs[lc + 2][lr] = wv.z; Bs[lc + 3][lr] = wv.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LK; ++k) {
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      float v = acc[i][j];
      if (row < d.M && col < d.N) {
        if (d.b) v += d.b[col];
        if (col < d.scale_cols) v *= d.col_scale;
        if (d.hint_mask) {
          const float m = d.hint_mask[row];
          if (m != 0.f) v += m * (d.emb[(size_t)d.labels[row] * d.N + col] + d.emb[(size_t)313 * d.N + col]);
        }
        if (d.relu) v = v > 0.f ? v : 0.f;
        if (d.residual) v += d.residual[(size_t)row * d.N + col];
      }
      Cs[ty * 4 + i][tx * 4 + j] = v;
    }
  }
  __syncthreads();
  if (d.ln_gamma) {   // N == 64: the whole row lives in this tile
    if (tid < LT) {
      float mean = 0.f;
      for (int c = 0; c < 64; ++c) mean += Cs[tid][c];
      mean *= (1.f / 64.f);
      float var = 0.f;
      for (int c = 0; c < 64; ++c) { const float t = Cs[tid][c] - mean; var = fmaf(t, t, var); }
      var *= (1.f / 64.f);
      const float rstd = 1.0f / sqrtf(var + 1e-5f);
      for (int c = 0; c < 64; ++c) Cs[tid][c] = (Cs[tid][c] - mean) * rstd * d.ln_gamma[c] + d.ln_beta[c];
    }
    __syncthreads();
  }
  if (d.transpose_S > 0) {
    // Y[(row / S)][col][row % S]; consecutive threads -> consecutive rows (contiguous in memory)
    for (int e = tid; e < LT * LT; e += 256) {
      const int c = e >> 6, r = e & 63;
      const int row = m0 + r, col = n0 + c;
      if (row < d.M && col < d.N)
        d.Y[((size_t)(row / d.transpose_S) * d.N + col) * d.transpose_S + row % d.transpose_S] = Cs[r][c];
    }
  } else {
    for (int e = tid; e < LT * LT; e += 256) {
      const int r = e >> 6, c = e & 63;
      const int row = m0 + r, col = n0 + c;
      if (row < d.M && col < d.N) d.Y[(size_t)row * d.N + col] = Cs[r][c];
    }
  }
}

// ------------------------------------------------------------------------------------------
// fused encoder-layer tail: y = LN2(x1 + W2 relu(W1 x1 + b1) + b2),  x1 = LN1(x + Wo att + bo)
// (EncoderLayer.forward after the attention core, models/transformer2d.py:55-59).  One CTA = 64 tokens, 256 threads,
// 4x4 register micro-tiles; x1 and the 256-wide hidden activations never leave shared memory (stored k-major so
// they feed the next GEMM directly); weights stream through a 16-deep shared-memory tile.
// ------------------------------------------------------------------------------------------
struct EncTailArgs {
  const float* att; const float* x; float* y; int M;
  const float* wo; const float* bo; const float* g1; const float* be1;
  const float* w1; const float* b1; const float* w2; const float* b2; const float* g2; const float* be2;
};

__device__ __forceinline__ void mm_chunk(const float (*Ak)[68], int k0, const float (*Bs)[68], int tx, int ty, float (&acc)[4][4]) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float4 av = *reinterpret_cast<const float4*>(&Ak[k0 + k][ty * 4]);
    const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
    const float aa[4] = {av.x, av.y, av.z, av.w};
    const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
  }
}

// LayerNorm over the 64 columns of each of the 64 rows held in Cs; result written k-major into XT and (optionally) to global
__device__ __forceinline__ void ln_rows(float (*Cs)[65], const float* __restrict__ gamma, const float* __restrict__ beta,
                                        float (*XT)[68], float* __restrict__ gout, int m0, int M, int tid) {
  if (tid < 64) {
    float mean = 0.f;
    for (int c = 0; c < 64; ++c) mean += Cs[tid][c];
    mean *= (1.f / 64.f);
    float var = 0.f;
    for (int c = 0; c < 64; ++c) { const float t = Cs[tid][c] - mean; var = fmaf(t, t, var); }
    var *= (1.f / 64.f);
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    for (int c = 0; c < 64; ++c) Cs[tid][c] = (Cs[tid][c] - mean) * rstd * gamma[c] + beta[c];
  }
  __syncthreads();
  for (int e = tid; e < 64 * 64; e += 256) {
    const int r = e >> 6, c = e & 63;
    if (XT) XT[c][r] = Cs[r][c];
    if (gout && m0 + r < M) gout[(size_t)(m0 + r) * 64 + c] = Cs[r][c];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) encoder_tail_kernel(const EncTailArgs a) {
  extern __shared__ float et_smem[];
  float (*X1T)[68] = reinterpret_cast<float (*)[68]>(et_smem);                    // [64 k][68]  x1, k-major
  float (*HdT)[68] = reinterpret_cast<float (*)[68]>(et_smem + 64 * 68);          // [256 k][68] hidden, k-major
  float (*As)[68] = reinterpret_cast<float (*)[68]>(et_smem + (64 + 256) * 68);   // [16][68]
  float (*Bs)[68] = reinterpret_cast<float (*)[68]>(et_smem + (64 + 256 + 16) * 68);
  float (*Cs)[65] = reinterpret_cast<float (*)[65]>(et_smem + (64 + 256 + 32) * 68);   // [64][65]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 64;
  const int lr = tid >> 2, lc = (tid & 3) * 4;   // load role: row lr, k offset lc..lc+3
  float acc[4][4];

  // The three GEMMs consume 36 weight tiles of 64 rows x 16 k (4 of Wo, 16 of W1, 16 of W2).  The tile of step s+1
  // (and, in phase 1, the matching slice of the attention output) is fetched into registers BEFORE the FMAs of step s,
  // so the L2 latency of the weight stream is off the critical path.  Measured neutral (52 us per launch at M = 16384
  // either way): ncu shows the kernel bound by the shared-memory pipe -- a 4x4 register tile costs two LDS.128 = 8
  // wavefronts per 16 FFMA (62 k LSU wavefronts per SM in 103 k cycles, FMA pipe 39 %); the next step for the token
  // GEMMs is a tensor-core formulation with split-bf16 operands, not more SIMT tuning.
  constexpr int kSteps = 36;
  auto load_w = [&](int step) -> float4 {
    if (step < 4) return *reinterpret_cast<const float4*>(a.wo + (size_t)lr * 64 + step * 16 + lc);
    if (step < 20) {
      const int nb = (step - 4) >> 2, k0 = ((step - 4) & 3) * 16;
      return *reinterpret_cast<const float4*>(a.w1 + (size_t)(nb * 64 + lr) * 64 + k0 + lc);
    }
    return *reinterpret_cast<const float4*>(a.w2 + (size_t)lr * 256 + (step - 20) * 16 + lc);
  };
  auto load_att = [&](int step) -> float4 {
    if (m0 + lr < a.M) return *reinterpret_cast<const float4*>(a.att + (size_t)(m0 + lr) * 64 + step * 16 + lc);
    return make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto zero_acc = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  };
  float4 wnext = load_w(0), anext = load_att(0);
  zero_acc();
  for (int step = 0; step < kSteps; ++step) {
    Bs[lc + 0][lr] = wnext.x; Bs[lc + 1][lr] = wnext.y; Bs[lc + 2][lr] = wnext.z; Bs[lc + 3][lr] = wnext.w;
    if (step < 4) { As[lc + 0][lr] = anext.x; As[lc + 1][lr] = anext.y; As[lc + 2][lr] = anext.z; As[lc + 3][lr] = anext.w; }
    __syncthreads();
    if (step + 1 < kSteps) wnext = load_w(step + 1);
    if (step + 1 < 4) anext = load_att(step + 1);
    if (step < 4) mm_chunk(As, 0, Bs, tx, ty, acc);                          // x + att Wo^T
    else if (step < 20) mm_chunk(X1T, ((step - 4) & 3) * 16, Bs, tx, ty, acc);   // x1 W1^T, block nb
    else mm_chunk(HdT, (step - 20) * 16, Bs, tx, ty, acc);                   // hidden W2^T
    __syncthreads();
    if (step == 3) {
      // ---- end of phase 1: x1 = LN1(x + att Wo^T + bo)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = ty * 4 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = tx * 4 + j;
          float v = acc[i][j] + a.bo[c];
          if (m0 + r < a.M) v += a.x[(size_t)(m0 + r) * 64 + c];
          Cs[r][c] = v;
        }
      }
      __syncthreads();
      ln_rows(Cs, a.g1, a.be1, X1T, nullptr, m0, a.M, tid);
      zero_acc();
    } else if (step >= 4 && step < 20 && ((step - 4) & 3) == 3) {
      // ---- end of one 64-column block of phase 2: hidden = relu(x1 W1^T + b1)
      const int nb = (step - 4) >> 2;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = nb * 64 + tx * 4 + j;
          const float v = acc[i][j] + a.b1[c];
          HdT[c][ty * 4 + i] = v > 0.f ? v : 0.f;
        }
      zero_acc();
      if (step == 19) __syncthreads();
    }
  }
  // ---- end of phase 3: y = LN2(x1 + hidden W2^T + b2)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = tx * 4 + j;
      Cs[r][c] = acc[i][j] + a.b2[c] + X1T[c][r];
    }
  }
  __syncthreads();
  ln_rows(Cs, a.g2, a.be2, nullptr, a.y, m0, a.M, tid);
}

// ------------------------------------------------------------------------------------------
// attention helpers (fp32)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot8(const float (&q)[8], const float* k) {
  const float4 a = *reinterpret_cast<const float4*>(k);
  const float4 b = *reinterpret_cast<const float4*>(k + 4);
  float s = q[0] * a.x;
  s = fmaf(q[1], a.y, s); s = fmaf(q[2], a.z, s); s = fmaf(q[3], a.w, s);
  s = fmaf(q[4], b.x, s); s = fmaf(q[5], b.y, s); s = fmaf(q[6], b.z, s); s = fmaf(q[7], b.w, s);
  return s;
}
__device__ __forceinline__ void axpy8(float (&o)[8], float p, const float* v) {
  const float4 a = *reinterpret_cast<const float4*>(v);
  const float4 b = *reinterpret_cast<const float4*>(v + 4);
  o[0] = fmaf(p, a.x, o[0]); o[1] = fmaf(p, a.y, o[1]); o[2] = fmaf(p, a.z, o[2]); o[3] = fmaf(p, a.w, o[3]);
  o[4] = fmaf(p, b.x, o[4]); o[5] = fmaf(p, b.y, o[5]); o[6] = fmaf(p, b.z, o[6]); o[7] = fmaf(p, b.w, o[7]);
}

__device__ __forceinline__ void lds2x2(const float* p, f32x2& a, f32x2& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One query against all keys with a running maximum (online softmax): the exact fall-back of attention_kernel.
__device__ __noinline__ void attend_online(const float (&q)[8], const float* Ks, const float* Vs, int S, int S4, float (&o)[8],
                                           float& sum_out) {
  float mx = -FLT_MAX, sum = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) o[c] = 0.f;
  for (int j = 0; j < S4; j += 4) {
    float s0 = dot8(q, Ks + j * 8), s1 = dot8(q, Ks + j * 8 + 8), s2 = dot8(q, Ks + j * 8 + 16), s3 = dot8(q, Ks + j * 8 + 24);
    if (j + 3 >= S) {                      // tail: padded keys must not contribute
      if (j + 1 >= S) s1 = -FLT_MAX;
      if (j + 2 >= S) s2 = -FLT_MAX;
      s3 = -FLT_MAX;
    }
    const float mnew = fmaxf(fmaxf(mx, fmaxf(s0, s1)), fmaxf(s2, s3));
    const float corr = exp2f(mx - mnew);
    const float p0 = exp2f(s0 - mnew), p1 = exp2f(s1 - mnew), p2 = exp2f(s2 - mnew), p3 = exp2f(s3 - mnew);
    sum = fmaf(sum, corr, (p0 + p1) + (p2 + p3));
#pragma unroll
    for (int c = 0; c < 8; ++c) o[c] *= corr;
    axpy8(o, p0, Vs + j * 8); axpy8(o, p1, Vs + j * 8 + 8); axpy8(o, p2, Vs + j * 8 + 16); axpy8(o, p3, Vs + j * 8 + 24);
    mx = mnew;
  }
  sum_out = sum;
}

// ------------------------------------------------------------------------------------------
// attention core: grid (B*8, ceil(S / (2*blockDim))), thread = TWO queries; K/V of the (image, head) in smem.
// softmax(s) = exp(s - m) / sum exp(s - m) for ANY reference m.  The running maximum of online softmax is a serial
// max -> rescale chain per key group; here the reference is simply m = 0: scores are kept in log2 units, fp32 exp2 covers
// [-126, 128), and attention logits of LayerNorm'd tokens sit far inside that range.  The rare query whose sum leaves
// [2^-60, 2^100] (or is not finite) is recomputed with the online-softmax routine above, so the result is exact in
// every case; the argument of every exponential is the score itself, i.e. as accurate as subtracting the true maximum.
// The inner loop is straight-line packed-fp32 work, four keys per iteration: per (query, key) 4 FFMA2 (q.k) + add +
// ex2 + add + pack + 4 FFMA2 (p.v), K/V rows shared by the thread's two queries.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ qkv, int S, float* __restrict__ out) {
  extern __shared__ float kv[];   // K [S4][8] then V [S4][8], S4 = S rounded up to 4 (padding rows are zero)
  const int S4 = (S + 3) & ~3;
  float* Ks = kv;
  float* Vs = kv + (size_t)S4 * 8;
  const int n = blockIdx.x >> 3, hd = blockIdx.x & 7;
  const float* base = qkv + (size_t)n * S * 192;
  for (int e = threadIdx.x; e < S4 * 2; e += blockDim.x) {
    const int t = e >> 1, half = e & 1;
    float4 k4 = make_float4(0.f, 0.f, 0.f, 0.f), v4 = k4;
    if (t < S) {
      k4 = *reinterpret_cast<const float4*>(base + (size_t)t * 192 + 64 + hd * 8 + half * 4);
      v4 = *reinterpret_cast<const float4*>(base + (size_t)t * 192 + 128 + hd * 8 + half * 4);
    }
    *reinterpret_cast<float4*>(Ks + t * 8 + half * 4) = k4;
    *reinterpret_cast<float4*>(Vs + t * 8 + half * 4) = v4;
  }
  __syncthreads();

  const int q0i = blockIdx.y * (2 * blockDim.x) + threadIdx.x, q1i = q0i + blockDim.x;
  if (q0i >= S) return;
  const bool has1 = q1i < S;
  float qa[8], qb[8];
  {
    const float* p0 = base + (size_t)q0i * 192 + hd * 8;
    const float* p1 = base + (size_t)(has1 ? q1i : q0i) * 192 + hd * 8;
    const float4 a0 = *reinterpret_cast<const float4*>(p0), a1 = *reinterpret_cast<const float4*>(p0 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(p1), b1 = *reinterpret_cast<const float4*>(p1 + 4);
    qa[0] = a0.x; qa[1] = a0.y; qa[2] = a0.z; qa[3] = a0.w; qa[4] = a1.x; qa[5] = a1.y; qa[6] = a1.z; qa[7] = a1.w;
    qb[0] = b0.x; qb[1] = b0.y; qb[2] = b0.z; qb[3] = b0.w; qb[4] = b1.x; qb[5] = b1.y; qb[6] = b1.z; qb[7] = b1.w;
#pragma unroll
    for (int c = 0; c < 8; ++c) { qa[c] *= 1.4426950408889634f; qb[c] *= 1.4426950408889634f; }
  }
  f32x2 QA[4], QB[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { QA[c] = pack2(qa[2 * c], qa[2 * c + 1]); QB[c] = pack2(qb[2 * c], qb[2 * c + 1]); }
  const f32x2 zero2 = pack2(0.f, 0.f);
  f32x2 OA[4], OB[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { OA[c] = zero2; OB[c] = zero2; }
  float suma = 0.f, sumb = 0.f;
  for (int j = 0; j < S4; j += 4) {
    f32x2 K[4][4], V[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      lds2x2(Ks + (j + u) * 8, K[u][0], K[u][1]);
      lds2x2(Ks + (j + u) * 8 + 4, K[u][2], K[u][3]);
      lds2x2(Vs + (j + u) * 8, V[u][0], V[u][1]);
      lds2x2(Vs + (j + u) * 8 + 4, V[u][2], V[u][3]);
    }
    f32x2 sa[4], sb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { sa[u] = ffma2(QA[0], K[u][0], zero2); sb[u] = ffma2(QB[0], K[u][0], zero2); }
#pragma unroll
    for (int c = 1; c < 4; ++c)
#pragma unroll
      for (int u = 0; u < 4; ++u) { sa[u] = ffma2(QA[c], K[u][c], sa[u]); sb[u] = ffma2(QB[c], K[u][c], sb[u]); }
    float pa[4], pb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float al, ah, bl, bh;
      unpack2(sa[u], al, ah);
      unpack2(sb[u], bl, bh);
      pa[u] = ex2f(al + ah);
      pb[u] = ex2f(bl + bh);
    }
    if (j + 4 > S) {                       // tail: padded (all-zero) keys must not contribute
#pragma unroll
      for (int u = 1; u < 4; ++u)
        if (j + u >= S) { pa[u] = 0.f; pb[u] = 0.f; }
    }
    suma += (pa[0] + pa[1]) + (pa[2] + pa[3]);
    sumb += (pb[0] + pb[1]) + (pb[2] + pb[3]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const f32x2 PA = pack2(pa[u], pa[u]), PB = pack2(pb[u], pb[u]);
#pragma unroll
      for (int c = 0; c < 4; ++c) { OA[c] = ffma2(PA, V[u][c], OA[c]); OB[c] = ffma2(PB, V[u][c], OB[c]); }
    }
  }
  float oa[8], ob[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) { unpack2(OA[c], oa[2 * c], oa[2 * c + 1]); unpack2(OB[c], ob[2 * c], ob[2 * c + 1]); }
  // validity of the m = 0 evaluation: the sum stayed inside [2^-60, 2^100] and the outputs are finite; otherwise exact path
  float ca = 0.f, cb = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) { ca += fabsf(oa[c]); cb += fabsf(ob[c]); }
  if (!(suma >= 8.67e-19f && suma <= 1.2e30f && ca <= 3.0e38f)) attend_online(qa, Ks, Vs, S, S4, oa, suma);
  if (has1 && !(sumb >= 8.67e-19f && sumb <= 1.2e30f && cb <= 3.0e38f)) attend_online(qb, Ks, Vs, S, S4, ob, sumb);
  {
    const float inv = 1.0f / suma;
    float* dst = out + ((size_t)n * S + q0i) * 64 + hd * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(oa[0] * inv, oa[1] * inv, oa[2] * inv, oa[3] * inv);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(oa[4] * inv, oa[5] * inv, oa[6] * inv, oa[7] * inv);
  }
  if (has1) {
    const float inv = 1.0f / sumb;
    float* dst = out + ((size_t)n * S + q1i) * 64 + hd * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(ob[0] * inv, ob[1] * inv, ob[2] * inv, ob[3] * inv);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(ob[4] * inv, ob[5] * inv, ob[6] * inv, ob[7] * inv);
  }
}

// ------------------------------------------------------------------------------------------
// k-means (Lloyd) + anchor pick, one CTA per image.  KMAX clusters, D = 64.
// ------------------------------------------------------------------------------------------
constexpr int KMAX = 32;
constexpr int KD = 64;

struct KmArgs {
  const float* X; const int32_t* init_idx; const int32_t* draws; int n_draws;
  const float* sizes; int B, S, K, iter_limit; float tol;
  int32_t* assign; float* hint_mask; int32_t* events; int32_t* iters;
};

// Runs one image.  XT: optional shared-memory copy of the image's tokens, transposed and padded:
// XT[c * (S + 1) + t] (conflict-free for both "thread = token" and "thread = (cluster, channel)" access).
// Assignment step for K <= KT clusters: squared distances accumulated channel by channel (same order for every KT, so the
// specialisations are bit-identical), first minimum wins like torch.argmin.
template <int KT>
__device__ __forceinline__ void kmeans_assign(const float* __restrict__ X, const float* XT, const float* C, int S, int K, int SP1,
                                              int* s_assign, int32_t* __restrict__ assign, int tid, int nt) {
  for (int t = tid; t < S; t += nt) {
    float dk[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) dk[k] = 0.f;
    for (int c = 0; c < KD; ++c) {
      const float xv = XT ? XT[c * SP1 + t] : X[(size_t)t * KD + c];
#pragma unroll
      for (int k = 0; k < KT; ++k)
        if (k < K) { const float df = xv - C[k * KD + c]; dk[k] = fmaf(df, df, dk[k]); }
    }
    float best = FLT_MAX; int bi = 0;
#pragma unroll
    for (int k = 0; k < KT; ++k)
      if (k < K && dk[k] < best) { best = dk[k]; bi = k; }
    if (s_assign) s_assign[t] = bi;
    assign[t] = bi;
  }
}

__device__ void kmeans_one_image(const KmArgs& a, int n, int draw_offset, float* C, float* Cprev, int* cnt,
                                 int* ridx, float* shiftk, int* s_flag, float* XT, int* s_assign, int* members, int* moff) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int S = a.S, K = a.K, SP1 = S + 1;
  const float* X = a.X + (size_t)n * S * KD;
  int32_t* assign = a.assign + (size_t)n * S;
  if (XT) {
    for (int e = tid; e < S * KD; e += nt) XT[(e % KD) * SP1 + e / KD] = X[e];
  }
  for (int e = tid; e < K * KD; e += nt) C[e] = X[(size_t)a.init_idx[n * K + e / KD] * KD + (e % KD)];
  if (tid == 0) { s_flag[0] = 0; /* events */ s_flag[1] = 0; /* stop */ s_flag[2] = 0; /* iterations */ }
  __syncthreads();
  while (true) {
    // 1. assignment (first minimum wins, like torch.argmin); the unrolled cluster loop is sized to K
    if (K <= 8) kmeans_assign<8>(X, XT, C, S, K, SP1, s_assign, assign, tid, nt);
    else if (K <= 16) kmeans_assign<16>(X, XT, C, S, K, SP1, s_assign, assign, tid, nt);
    else kmeans_assign<KMAX>(X, XT, C, S, K, SP1, s_assign, assign, tid, nt);
    for (int e = tid; e < K * KD; e += nt) Cprev[e] = C[e];
    __syncthreads();
    // 2. member counts (one warp per cluster), then draws for empty clusters in cluster order
    //    (reference: clusterkit.py:178-184)
    {
      const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
      const int* asg = s_assign ? s_assign : assign;
      for (int k = warp; k < K; k += nw) {
        int c = 0;
        for (int t = lane; t < S; t += 32) c += (asg[t] == k);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) cnt[k] = c;
      }
    }
    __syncthreads();
    if (tid == 0) {
      int off = 0;
      for (int k = 0; k < K; ++k) { moff[k] = off; off += cnt[k]; }
      for (int k = 0; k < K; ++k) {
        ridx[k] = -1;
        if (cnt[k] == 0) {
          const int di = draw_offset + s_flag[0];
          if (di < a.n_draws) ridx[k] = a.draws[di]; else { ridx[k] = 0; a.events[a.B + 1] = 1; }
          s_flag[0]++;
        }
      }
    }
    __syncthreads();
    // 2b. member lists in token order (one warp per cluster, ballot scan): the centre update then walks cnt[k] members
    //     instead of testing all S tokens for every (cluster, channel) -- same additions in the same order
    if (members) {
      const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
      const int* asg = s_assign ? s_assign : assign;
      for (int k = warp; k < K; k += nw) {
        int base = moff[k];
        for (int t0 = 0; t0 < S; t0 += 32) {
          const int t = t0 + lane;
          const bool hit = t < S && asg[t] == k;
          const unsigned m = __ballot_sync(0xffffffffu, hit);
          if (hit) members[base + __popc(m & ((1u << lane) - 1u))] = t;
          base += __popc(m);
        }
      }
      __syncthreads();
    }
    // 3. centre update: mean of members in token order
    {
      const int* asg = s_assign ? s_assign : assign;
      for (int e = tid; e < K * KD; e += nt) {
        const int k = e / KD, c = e % KD;
        float v;
        if (cnt[k] == 0) {
          v = X[(size_t)ridx[k] * KD + c];
        } else {
          float s = 0.f;
          if (members) {
            const int* ml = members + moff[k];
            const int m = cnt[k];
            if (XT) {
              const float* col = XT + c * SP1;
              for (int i = 0; i < m; ++i) s += col[ml[i]];
            } else {
              for (int i = 0; i < m; ++i) s += X[(size_t)ml[i] * KD + c];
            }
          } else if (XT) {
            const float* col = XT + c * SP1;
            for (int t = 0; t < S; ++t) if (asg[t] == k) s += col[t];
          } else {
            for (int t = 0; t < S; ++t) if (asg[t] == k) s += X[(size_t)t * KD + c];
          }
          v = s / (float)cnt[k];
        }
        C[e] = v;
      }
    }
    __syncthreads();
    // 4. centre shift = sum_k ||c_k - c_k_prev||
    for (int k = tid; k < K; k += nt) {
      float s = 0.f;
      for (int c = 0; c < KD; ++c) { const float df = C[k * KD + c] - Cprev[k * KD + c]; s = fmaf(df, df, s); }
      shiftk[k] = sqrtf(s);
    }
    __syncthreads();
    if (tid == 0) {
      float sh = 0.f;
      for (int k = 0; k < K; ++k) sh += shiftk[k];
      s_flag[2]++;
      s_flag[1] = (sh * sh < a.tol) || (a.iter_limit != 0 && s_flag[2] >= a.iter_limit);
    }
    __syncthreads();
    if (s_flag[1]) break;
  }
  // anchor pick: per cluster the member with the largest super-pixel (first maximum), anchor_gen.py:98-101
  float* hint = a.hint_mask + (size_t)n * S;
  for (int t = tid; t < S; t += nt) hint[t] = 0.f;
  __syncthreads();
  {
    const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    const int* asg = s_assign ? s_assign : assign;
    for (int k = warp; k < K; k += nw) {
      float best = -FLT_MAX; int bi = 0x7fffffff;
      for (int t = lane; t < S; t += 32) {
        const float sc = (asg[t] == k ? 1.0f : 0.0f) + a.sizes[(size_t)n * S + t] * 0.01f;
        if (sc > best) { best = sc; bi = t; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) ridx[k] = bi;
    }
  }
  __syncthreads();
  if (tid == 0) {
    for (int k = 0; k < K; ++k) hint[ridx[k]] += 1.0f;
    a.events[n] = s_flag[0];
    a.iters[n] = s_flag[2];
  }
  __syncthreads();
}

// mode 0: grid B, speculative (draw offset 0).  mode 1: grid 1, sequential fix-up in image order.
// Dynamic shared memory: [transposed tokens XT[64][S+1] when they fit] int assign[S], int members[S].
__global__ void __launch_bounds__(512) kmeans_anchor_kernel(const KmArgs a, int mode, int use_smem) {
  __shared__ float C[KMAX * KD];
  __shared__ float Cprev[KMAX * KD];
  __shared__ int cnt[KMAX];
  __shared__ int ridx[KMAX];
  __shared__ float shiftk[KMAX];
  __shared__ int s_flag[4];
  __shared__ int moff[KMAX];
  extern __shared__ float km_dyn[];
  float* XT = use_smem ? km_dyn : nullptr;
  int* s_assign = reinterpret_cast<int*>(km_dyn + (use_smem ? KD * (a.S + 1) : 0));
  int* members = s_assign + a.S;
  if (mode == 0) {
    kmeans_one_image(a, blockIdx.x, 0, C, Cprev, cnt, ridx, shiftk, s_flag, XT, s_assign, members, moff);
  } else {
    int running = 0;
    for (int n = 0; n < a.B; ++n) {
      const int ev = a.events[n];
      if (ev > 0 && running > 0) kmeans_one_image(a, n, running, C, Cprev, cnt, ridx, shiftk, s_flag, XT, s_assign, members, moff);
      __syncthreads();
      running += a.events[n];
      __syncthreads();
    }
    if (threadIdx.x == 0) a.events[a.B] = running;
  }
}

// ------------------------------------------------------------------------------------------
// labels: one thread per token; logits are read in the API layout [B,313,S] (coalesced over tokens)
// ------------------------------------------------------------------------------------------
// mode 0 (arg-max of 313 logits, the hot one): 8 threads per token split the classes (c = part, part + 8, ...) so that the
// 313 strided loads of a token are 40 rounds of latency instead of 313 (r2a capture: 48 us at batch 64, 6 % of the HBM
// peak, one thread per token); first maximum wins like torch.max: ties resolve to the lower class index.  blockDim = 128.
__global__ void token_labels_kernel(int mode, const float* __restrict__ src, const float* __restrict__ table, int B,
                                    int S, int32_t* __restrict__ labels, float* __restrict__ colors) {
  if (mode == 0) {
    // block = 16 consecutive tokens x 8 class partitions, token index fastest (16 consecutive floats per load instruction)
    __shared__ float sb[8][16];
    __shared__ int si[8][16];
    const int tl = threadIdx.x & 15, part = threadIdx.x >> 4;
    const int idx = blockIdx.x * 16 + tl;
    const bool ok = idx < B * S;
    const int n = ok ? idx / S : 0, t = ok ? idx % S : 0;
    float best = -FLT_MAX;
    int bi = 0x7fffffff;
    const float* col = src + (size_t)n * 313 * S + t;
    if (ok)
      for (int c = part; c < 313; c += 8) { const float v = col[(size_t)c * S]; if (v > best) { best = v; bi = c; } }
    sb[part][tl] = best;
    si[part][tl] = bi;
    __syncthreads();
    if (ok && part == 0) {
#pragma unroll
      for (int p = 1; p < 8; ++p) {
        const float ob = sb[p][tl];
        const int oi = si[p][tl];
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      labels[idx] = bi;
      if (colors) {
        colors[((size_t)n * 2 + 0) * S + t] = table[bi * 2] / 110.0f;
        colors[((size_t)n * 2 + 1) * S + t] = table[bi * 2 + 1] / 110.0f;
      }
    }
    return;
  }
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * S) return;
  const int n = idx / S, t = idx % S;
  int bi = 0;
  {
    float best = FLT_MAX;
    const float a0 = src[((size_t)n * 2 + 0) * S + t] * 110.0f, a1 = src[((size_t)n * 2 + 1) * S + t] * 110.0f;
    for (int c = 0; c < 313; ++c) {
      const float d0 = table[c * 2] - a0, d1 = table[c * 2 + 1] - a1;
      const float v = sqrtf(d0 * d0 + d1 * d1);
      if (v < best) { best = v; bi = c; }
    }
  }
  labels[idx] = bi;
}

// ------------------------------------------------------------------------------------------
// diverse anchor colours (sampled_T > 0): per token the 10 most probable bins (stable order: ties keep the lower
// bin index, like torch.sort on CPU), then T=0 top-1, T=1 the candidate farthest from top-1, T=2 the candidate
// maximising d(.,top-1) + d(.,T=1 pick)  (reference models/anchor_gen.py:54-90).  One thread per token.
// ------------------------------------------------------------------------------------------
__global__ void token_sample3_kernel(const float* __restrict__ logits, const float* __restrict__ table, int B, int S,
                                     int32_t* __restrict__ labels3, float* __restrict__ colors3) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * S) return;
  const int n = idx / S, t = idx % S;
  float tv[10];
  int ti[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) { tv[i] = -FLT_MAX; ti[i] = 0; }
  const float* col = logits + (size_t)n * 313 * S + t;
  for (int c = 0; c < 313; ++c) {
    const float v = col[(size_t)c * S];
    if (v > tv[9]) {
      float cv = v; int ci = c;
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        if (cv > tv[i]) { const float fv = tv[i]; const int fi = ti[i]; tv[i] = cv; ti[i] = ci; cv = fv; ci = fi; }
      }
    }
  }
  float ca[10], cb[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) { ca[i] = table[ti[i] * 2] / 110.0f; cb[i] = table[ti[i] * 2 + 1] / 110.0f; }
  float d0[10];
  int p1 = 0; float best = -1.f;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const float da = ca[i] - ca[0], db = cb[i] - cb[0];
    d0[i] = sqrtf(da * da + db * db);
    if (d0[i] > best) { best = d0[i]; p1 = i; }
  }
  int p2 = 0; best = -1.f;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const float da = ca[i] - ca[p1], db = cb[i] - cb[p1];
    const float sc = d0[i] + sqrtf(da * da + db * db);
    if (sc > best) { best = sc; p2 = i; }
  }
  const int pick[3] = {0, p1, p2};
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    const int k = pick[v];
    labels3[(size_t)v * B * S + idx] = ti[k];
    colors3[(((size_t)v * B + n) * 2 + 0) * S + t] = ca[k];
    colors3[(((size_t)v * B + n) * 2 + 1) * S + t] = cb[k];
  }
}

}  // namespace

extern "C" int disco_linear(disco_handle* h, const disco_linear_desc* d, void* stream) {
  DISCO_CHECK_ARG(h && d && d->X && d->W && d->Y, "linear: null pointer");
  DISCO_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0 && d->K % 16 == 0, "linear: K must be a positive multiple of 16 (got %d)", d->K);
  DISCO_CHECK_ARG(!d->ln_gamma || d->N == 64, "linear: LayerNorm epilogue needs N == 64");
  DISCO_CHECK_ARG(!d->pos || (d->S > 0 && d->pos_cols % 64 == 0), "linear: pos needs S > 0 and pos_cols %% 64 == 0");
  DISCO_CHECK_ARG(!d->hint_mask || (d->labels && d->emb), "linear: hint embedding needs labels and emb");
  DiscoDeviceGuard guard(h);
  dim3 grid((d->M + LT - 1) / LT, (d->N + LT - 1) / LT);
  linear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*d);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}

extern "C" int disco_encoder_tail(disco_handle* h, const float* att, const float* x, float* y, int M, const float* wo,
                                  const float* bo, const float* ln1_g, const float* ln1_b, const float* w1, const float* b1,
                                  const float* w2, const float* b2, const float* ln2_g, const float* ln2_b, void* stream) {
  DISCO_CHECK_ARG(h && att && x && y && wo && bo && ln1_g && ln1_b && w1 && b1 && w2 && b2 && ln2_g && ln2_b && M > 0,
                  "encoder_tail: bad argument");
  const EncTailArgs a{att, x, y, M, wo, bo, ln1_g, ln1_b, w1, b1, w2, b2, ln2_g, ln2_b};
  const int smem = ((64 + 256 + 32) * 68 + 64 * 65) * (int)sizeof(float);
  DiscoDeviceGuard guard(h);
  if (int rc = disco_ensure_smem(h, (const void*)encoder_tail_kernel, smem)) return rc;
  encoder_tail_kernel<<<(M + 63) / 64, 256, smem, (cudaStream_t)stream>>>(a);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}

extern "C" int disco_attention(disco_handle* h, const float* qkv, int batch, int S, float* out, void* stream) {
  DISCO_CHECK_ARG(h && qkv && out && batch > 0 && S > 0, "attention: bad argument");
  const size_t smem = (size_t)((S + 3) & ~3) * 16 * sizeof(float);
  DISCO_CHECK_ARG(smem <= 200 * 1024, "attention: S=%d too large for the shared-memory K/V stage", S);
  DiscoDeviceGuard guard(h);
  if (int rc = disco_ensure_smem(h, (const void*)attention_kernel, (int)smem)) return rc;
  // two queries per thread, 128-thread CTAs (measured best of 32/64/128/256 at both S = 256 and S = 1024: enough CTAs
  // for an even spread over the SMs, K/V stage shared by 256 queries)
  int threads = 128;
  if (const char* e = getenv("DISCO_ATT_THREADS")) threads = atoi(e);
  dim3 grid(batch * 8, (S + 2 * threads - 1) / (2 * threads));
  attention_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(qkv, S, out);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}

extern "C" int disco_kmeans_anchor(disco_handle* h, const float* X, const int32_t* init_idx, const int32_t* draws,
                                   int n_draws, const float* sizes, int batch, int S, int K, int iter_limit, float tol,
                                   int32_t* assign, float* hint_mask, int32_t* events, int32_t* iters, void* stream) {
  DISCO_CHECK_ARG(h && X && init_idx && draws && sizes && assign && hint_mask && events && iters, "kmeans: null pointer");
  DISCO_CHECK_ARG(K >= 1 && K <= KMAX, "kmeans: K must be in [1,%d] (got %d)", KMAX, K);
  DISCO_CHECK_ARG(K <= S, "kmeans: n_clusters (%d) exceeds the number of tokens (%d)", K, S);
  KmArgs a{X, init_idx, draws, n_draws, sizes, batch, S, K, iter_limit, tol, assign, hint_mask, events, iters};
  DiscoDeviceGuard guard(h);
  cudaStream_t st = (cudaStream_t)stream;
  DISCO_CUDA(cudaMemsetAsync(events, 0, sizeof(int32_t) * (batch + 2), st));
  // dynamic smem: [XT (64 x (S+1) floats) when it fits] + assign[S] + members[S]
  const size_t dyn_full = ((size_t)KD * (S + 1) + 2 * (size_t)S) * sizeof(float);
  const int use_smem = dyn_full <= 160 * 1024;
  const size_t dyn = use_smem ? dyn_full : 2 * (size_t)S * sizeof(int);
  DISCO_CHECK_ARG(dyn <= 160 * 1024, "kmeans: S=%d too large", S);
  if (int rc = disco_ensure_smem(h, (const void*)kmeans_anchor_kernel, (int)dyn)) return rc;
  kmeans_anchor_kernel<<<batch, 512, dyn, st>>>(a, 0, use_smem);
  DISCO_LAUNCH_CHECK(h);
  kmeans_anchor_kernel<<<1, 512, dyn, st>>>(a, 1, use_smem);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}

extern "C" int disco_token_sample3(disco_handle* h, const float* logits, const float* q_to_ab, int batch, int S,
                                   int32_t* labels3, float* colors3, void* stream) {
  DISCO_CHECK_ARG(h && logits && q_to_ab && labels3 && colors3, "token_sample3: null pointer");
  DiscoDeviceGuard guard(h);
  token_sample3_kernel<<<(batch * S + 127) / 128, 128, 0, (cudaStream_t)stream>>>(logits, q_to_ab, batch, S, labels3, colors3);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}

extern "C" int disco_token_labels(disco_handle* h, int mode, const float* src, const float* q_to_ab, int batch, int S,
                                  int32_t* labels, float* colors, void* stream) {
  DISCO_CHECK_ARG(h && src && q_to_ab && labels, "token_labels: null pointer");
  DISCO_CHECK_ARG(mode == 0 || mode == 1, "token_labels: mode must be 0 or 1");
  DiscoDeviceGuard guard(h);
  const long long blocks = mode == 0 ? ((long long)batch * S + 15) / 16 : ((long long)batch * S + 127) / 128;
  token_labels_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(mode, src, q_to_ab, batch, S, labels, colors);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}
