// EXPERIMENT (off by default, DISCO_TS=1): 64 -> 64 channel 3x3 stride-1 convolution with the A operand in TENSOR MEMORY
// and row-sliding accumulators (reference layers: the plain 64-channel convolutions at full resolution of ColorProbNet /
// HourGlass2, models/network.py:10-28,83-101,152-201).  Correct (tests/test_gpu_conv_tc.py::test_conv_64_to_64_* pass with
// DISCO_TS=1) but not faster than the resident-weight kernel of conv_tc.cu: 0.33 ms against 0.28 ms per layer at batch 64,
// 256 x 256 -- kept as the measured answer to "would A from tensor memory lift the N = 64 layers off the shared-memory port?".
//
// Idea.  On the shared-memory operand path an N = 64 MMA (128 x 64 x 16) reads 4 KB of A and 2 KB of B per 32 clk of math:
// 48 clk at the 128 B/clk read port -> these layers sit at 67 % of the tensor pipe.  The A tile of tap (ky, kx)
// for output row y is the input row y + ky - 1 shifted by kx - 1: the SAME tile serves the three ky of the output rows
// y+1, y, y-1.  Here
//   * a 128-pixel tile is 128 consecutive pixels of ONE image row (TMEM lane = pixel);
//   * one TMA box per input row (130 pixels x 128 B, 128B-swizzled, zero fill outside the image) lands in a shared-memory
//     ring; the MMA thread copies it to tensor memory three times, shifted by kx pixels, with tcgen05.cp.128x256b (the
//     descriptor an SS MMA would use for that K = 16 slice) -- tcgen05.cp and tcgen05.mma execute in issue order, so no
//     barrier separates a copy from the MMAs that consume it;
//   * per (input row, kx) the same thread issues three groups of four tcgen05.mma (A in TMEM, B = the resident
//     64 x 64 weight block of tap (ky, kx) in shared memory) into the live accumulators of the three output rows;
//   * four accumulators (4 x 64 TMEM columns) rotate over the output rows; an epilogue warp group drains a finished row
//     (bias, residual, activation, post affine) while the next rows accumulate.
// CTAs are persistent over (image, 128-pixel column strip, chunk of rows).
//
// What the probes (DISCO_TS_MODE bits: 1 no epilogue math/stores, 2 no copies, 4 no MMAs, 8 no TMA, 16 no tcgen05.ld) measured
// per layer (235 input rows per SM): MMAs alone 0.21 ms = ~44 clk per 128 x 64 x 16 MMA with A in TMEM (32 at the pipe's
// peak), copies alone 0.147 ms = ~92 clk per 4 KB tcgen05.cp, both 0.265 ms (partly overlapped), this file's epilogue alone
// 0.28 ms (one pixel = 128 B per thread: 16-byte stores at a 128-byte lane stride), everything 0.33 ms.  Even with a free
// epilogue the copy + MMA stream (0.265 ms) is within 5 % of the shared-memory-operand kernel: the copies cost what the
// MMA operand reads saved.  Lessons kept: issue from `elect.sync` (with `lane == 0` every tcgen05 instruction became an
// ELECT loop: 0.50 -> 0.33 ms), per-thread 128-byte global rows into tcgen05.st are 32 L1 wavefronts per load (first
// version, 0.75 ms).
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <algorithm>

namespace {

constexpr int TS_TMA_WARP = 0, TS_MMA_WARP = 1, TS_EPI_WARP0 = 4;   // warps 2, 3 idle: epilogue warp w drains lane quarter w % 4
constexpr int TS_THREADS = 256;
constexpr int TS_NR = 4;                         // input-row slots in shared memory
constexpr int TS_ROW_PX = 130;                   // 128 pixels + one halo pixel on each side
constexpr int TS_ROW_BYTES = TS_ROW_PX * 128;    // TMA transaction size
constexpr int TS_ROW_SLOT = 17 * 1024;           // slot pitch (1024-aligned: the swizzle phase follows the address)
constexpr int TS_NA = 3;                         // A tiles in tensor memory = the three kx of a row (ordering by the tcgen05 pipe, no barriers)
constexpr int TS_NACC = 4;                       // accumulator slots
constexpr int TS_ACOLS = 32;                     // TMEM columns of an A tile: 64 bf16 per lane
constexpr int TS_ACC0 = 0, TS_A0 = TS_NACC * 64; // column map: accumulators, then the A tiles
constexpr int TS_TMEM_COLS = 512;
constexpr int TS_WBYTES = 9 * 64 * 128;          // resident weights: 9 taps x 64 rows (co) x 128 B (64 ci), 128B-swizzled
constexpr int TS_SMEM = 1024 + TS_WBYTES + TS_NR * TS_ROW_SLOT + 3 * 64 * 4 + 256;
constexpr long long kTsSpin = 4000000000ll;

struct TsParams {
  CUtensorMap tm;                // input as (64 ch, W, B*H), box (64, 130, 1), 128B swizzle
  const __nv_bfloat16* x;        // [B, H, W, 64]
  const uint16_t* w;             // [9][64 co][64 ci] bf16
  const float* bias;
  const float* post_scale;       // or nullptr
  const float* post_shift;
  const __nv_bfloat16* res;      // or nullptr
  __nv_bfloat16* y;
  int B, H, W;
  int act;
  float slope;
  int strips, chunks, rows_per_chunk, items;
  int mode;                      // probes (DISCO_TS_MODE): 1 no epilogue math/stores, 2 no tcgen05.cp, 4 no MMAs
  int32_t* error_flag;
};

__device__ __forceinline__ uint32_t ts_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ts_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ts_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ts_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ts_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ts_mbar_wait(uint64_t* bar, uint32_t parity, int32_t* error_flag) {
  const uint32_t addr = ts_smem_u32(bar);
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if ((it & 1023) == 1023) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kTsSpin) {
        if (error_flag) atomicExch(error_flag, 1);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void ts_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ts_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ts_tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          ts_smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(ts_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 128 rows x 256 bits (one K = 16 slice of a K-major operand tile) from shared memory to 8 TMEM columns
__device__ __forceinline__ void ts_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// one lane of a converged warp; unlike `lane == 0` the compiler knows the branch is warp-uniform and keeps the operands of
// the tcgen05 / TMA instructions in uniform registers (with `lane == 0` every such instruction became an ELECT loop)
__device__ __forceinline__ bool ts_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void ts_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ts_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ts_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ts_smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void ts_mma(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void ts_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// K-major 64-element (128 B) rows, 128B swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t ts_bdesc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// c_format f32, a/b bf16, K-major, N = 64, M = 128
constexpr uint32_t kTsIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct TsItem { int n, x0, y0, y1; };   // output rows [y0, y1), 128 pixels from x0
__device__ __forceinline__ TsItem ts_item(const TsParams& P, int item) {
  TsItem it;
  const int chunk = item % P.chunks, r = item / P.chunks;
  it.x0 = (r % P.strips) * 128;
  it.n = r / P.strips;
  it.y0 = chunk * P.rows_per_chunk;
  it.y1 = min(it.y0 + P.rows_per_chunk, P.H);
  return it;
}

__global__ void __launch_bounds__(TS_THREADS, 1) conv64_ts_kernel(const __grid_constant__ TsParams P) {
  extern __shared__ uint8_t ts_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ts_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;                                                  // 9 x 8 KB, 1024-aligned
  uint8_t* rows = smem + TS_WBYTES;                                     // TS_NR row slots
  float* epi = reinterpret_cast<float*>(rows + TS_NR * TS_ROW_SLOT);    // bias | post_scale | post_shift
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + 3 * 64);
  uint64_t* row_full = bars;                 // [TS_NR]
  uint64_t* row_empty = bars + TS_NR;        // [TS_NR]
  uint64_t* acc_full = bars + 2 * TS_NR;     // [TS_NACC]
  uint64_t* acc_empty = acc_full + TS_NACC;  // [TS_NACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + TS_NACC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = P.H, W = P.W;

  // resident weights: [tap][co][ci] -> swizzled rows (16-byte chunk c of row co at chunk c ^ (co & 7))
  for (int i = tid; i < 9 * 64 * 8; i += TS_THREADS) {
    const int c = i & 7, row = i >> 3;                                  // row = tap * 64 + co
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(P.w) + i);
    *reinterpret_cast<uint4*>(wsm + row * 128 + ((c ^ (row & 7)) << 4)) = v;
  }
  if (tid < 64) {
    epi[tid] = P.bias[tid];
    epi[64 + tid] = P.post_scale ? P.post_scale[tid] : 1.0f;
    epi[128 + tid] = P.post_shift ? P.post_shift[tid] : 0.0f;
  }
  if (tid == 0) {
    for (int i = 0; i < TS_NR; ++i) { ts_mbar_init(&row_full[i], 1); ts_mbar_init(&row_empty[i], 1); }
    for (int i = 0; i < TS_NACC; ++i) { ts_mbar_init(&acc_full[i], 1); ts_mbar_init(&acc_empty[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TS_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ts_smem_u32(tmem_slot)),
                 "r"((uint32_t)TS_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // generic-proxy writes of the weights must be visible to the tensor-core (async) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ts_fence_before();
  __syncthreads();
  ts_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == TS_TMA_WARP) {
    // ---------------- TMA producer: one box per valid input row
    if (ts_elect_one()) {
      int rowc = 0;
      for (int item = blockIdx.x; item < P.items; item += gridDim.x) {
        const TsItem it = ts_item(P, item);
        const int r0 = max(it.y0 - 1, 0), r1 = min(it.y1, H - 1);         // valid input rows [r0, r1]
        for (int r = r0; r <= r1; ++r, ++rowc) {
          const int rs = rowc % TS_NR;
          ts_mbar_wait(&row_empty[rs], ((rowc / TS_NR) & 1) ^ 1, P.error_flag);
          if (P.mode & 8) { ts_mbar_arrive(&row_full[rs]); continue; }
          ts_mbar_expect_tx(&row_full[rs], TS_ROW_BYTES);
          ts_tma_load_3d(rows + rs * TS_ROW_SLOT, &P.tm, &row_full[rs], 0, it.x0 - 1, it.n * H + r);
        }
      }
    }
  } else if (warp == TS_MMA_WARP) {
    // ---------------- copy + MMA issuer (one lane).  The loop is instruction-bound (one thread issues 12 copies and 36 MMAs
    // per row: the first version spent ~4800 clk per row on descriptor and index arithmetic): descriptors are built once,
    // the issuing lane is chosen with elect.sync.
    if (ts_elect_one()) {
      const uint32_t wbase = ts_smem_u32(wsm), rbase = ts_smem_u32(rows);
      const uint64_t bd0 = ts_bdesc(wbase), ad_base = ts_bdesc(rbase);   // tap t: + t * 512 (8 KB >> 4); row slot rs: + rs * 1088
      const uint32_t acc_t = tmem_base + TS_ACC0, a_base = tmem_base + TS_A0;
      const bool do_cp = !(P.mode & 2), do_mma = !(P.mode & 4);
      int rowc = 0, seq0 = 0;
      for (int item = blockIdx.x; item < P.items; item += gridDim.x) {
        const TsItem it = ts_item(P, item);
        const int r0 = max(it.y0 - 1, 0), r1 = min(it.y1, H - 1);
        for (int r = r0; r <= r1; ++r, ++rowc) {
          const int rs = rowc % TS_NR;
          ts_mbar_wait(&row_full[rs], (rowc / TS_NR) & 1, P.error_flag);
          const uint64_t ad0 = ad_base + (uint64_t)(rs * (TS_ROW_SLOT >> 4));
          if (r > it.y0 && r + 1 < it.y1) {
            // ---- interior row: feeds y = r+1 (its first contribution), r, and r-1 (its last); no per-tap decisions
            const int seq2 = seq0 + (r + 1 - it.y0);
            const uint32_t s2 = seq2 & (TS_NACC - 1), s1 = (seq2 - 1) & (TS_NACC - 1), s0 = (seq2 - 2) & (TS_NACC - 1);
            ts_mbar_wait(&acc_empty[s2], ((seq2 / TS_NACC) & 1) ^ 1, P.error_flag);
            ts_fence_after();
            const uint32_t d2 = acc_t + s2 * 64, d1 = acc_t + s1 * 64, d0 = acc_t + s0 * 64;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              static_assert(TS_NA == 3, "one A tile per kx");
              const uint32_t a_t = a_base + kx * TS_ACOLS;
#pragma unroll
              for (int k = 0; k < 4; ++k) if (do_cp) ts_cp_128x256b(a_t + k * 8, ad0 + (uint64_t)(kx * 8 + k * 2));
              if (do_mma) {
#pragma unroll
                for (int k = 0; k < 4; ++k) ts_mma(d2, a_t + k * 8, bd0 + (uint64_t)((0 * 3 + kx) * 512 + k * 2), kTsIdesc, (kx == 0 && k == 0) ? 0u : 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) ts_mma(d1, a_t + k * 8, bd0 + (uint64_t)((1 * 3 + kx) * 512 + k * 2), kTsIdesc, 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) ts_mma(d0, a_t + k * 8, bd0 + (uint64_t)((2 * 3 + kx) * 512 + k * 2), kTsIdesc, 1u);
              }
            }
            ts_commit(&acc_full[s0]);
            ts_commit(&row_empty[rs]);
            continue;
          }
          ts_fence_after();
          // ---- rows at the chunk / image edges: output rows fed by this input row are y = r + 1 - ky; validity,
          //      accumulator slot and first / last flags per ky
          int asl[3];
          bool ok[3], fst[3], lst[3];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int y = r + 1 - ky;
            ok[ky] = y >= it.y0 && y < it.y1;
            const int seq = seq0 + (y - it.y0);
            asl[ky] = seq & (TS_NACC - 1);
            fst[ky] = r == max(y - 1, 0);                               // with kx == 0: first contribution to row y
            lst[ky] = r == min(y + 1, H - 1);                           // with kx == 2: last contribution
            if (ok[ky] && fst[ky]) {
              ts_mbar_wait(&acc_empty[asl[ky]], ((seq / TS_NACC) & 1) ^ 1, P.error_flag);
              ts_fence_after();
            }
          }
#pragma unroll 1
          for (int kx = 0; kx < 3; ++kx) {
            const uint32_t a_t = a_base + kx * TS_ACOLS;
#pragma unroll
            for (int k = 0; k < 4; ++k) if (do_cp) ts_cp_128x256b(a_t + k * 8, ad0 + (uint64_t)(kx * 8 + k * 2));
#pragma unroll 1
            for (int ky = 0; ky < 3; ++ky) {
              if (!ok[ky]) continue;
              const uint32_t d_t = acc_t + asl[ky] * 64;
              const uint64_t bd = bd0 + (uint64_t)((ky * 3 + kx) * 512);
              if (do_mma) {
                ts_mma(d_t, a_t, bd, kTsIdesc, (kx == 0 && fst[ky]) ? 0u : 1u);
#pragma unroll
                for (int k = 1; k < 4; ++k) ts_mma(d_t, a_t + k * 8, bd + (uint64_t)(k * 2), kTsIdesc, 1u);
              }
              if (kx == 2 && lst[ky]) ts_commit(&acc_full[asl[ky]]);
            }
          }
          ts_commit(&row_empty[rs]);                                     // the three copies of this row have been read
        }
        seq0 += it.y1 - it.y0;
      }
    }
  } else if (warp < TS_EPI_WARP0) {
    // idle warps
  } else {
    // ---------------- epilogue: warp quarter q drains TMEM lanes 32q .. 32q+31 (pixel p of the tile)
    const int q = warp & 3, p = q * 32 + lane;
    int seq0 = 0;
    for (int item = blockIdx.x; item < P.items; item += gridDim.x) {
      const TsItem it = ts_item(P, item);
      for (int y = it.y0; y < it.y1; ++y) {
        const int seq = seq0 + (y - it.y0), as = seq % TS_NACC;
        ts_mbar_wait(&acc_full[as], (seq / TS_NACC) & 1, P.error_flag);
        ts_fence_after();
        uint32_t acc[2][32];
        const uint32_t t = tmem_base + ((uint32_t)(q * 32) << 16) + TS_ACC0 + as * 64;
        if (!(P.mode & 16)) {
          ts_tmem_ld32(t, acc[0]);
          ts_tmem_ld32(t + 32, acc[1]);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        ts_fence_before();
        ts_mbar_arrive(&acc_empty[as]);
        const int x = it.x0 + p;
        if (x < W && !(P.mode & 1)) {
          const size_t pix = (((size_t)it.n * H + y) * W + x) * 64;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {                               // 8 channels = one 16-byte store
              const int c0 = hh * 32 + j * 8;
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(acc[hh][j * 8 + e]) + epi[c0 + e];
              if (P.res) {
                const uint4 rr = *reinterpret_cast<const uint4*>(P.res + pix + c0);
                const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  v[2 * e] += __uint_as_float(rw[e] << 16);
                  v[2 * e + 1] += __uint_as_float(rw[e] & 0xffff0000u);
                }
              }
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float a = v[2 * e], b = v[2 * e + 1];
                if (P.act != DISCO_ACT_NONE) { a = fmaxf(a, a * P.slope); b = fmaxf(b, b * P.slope); }
                a = a * epi[64 + c0 + 2 * e] + epi[128 + c0 + 2 * e];
                b = b * epi[64 + c0 + 2 * e + 1] + epi[128 + c0 + 2 * e + 1];
                __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                o[e] = *reinterpret_cast<uint32_t*>(&h2);
              }
              *reinterpret_cast<uint4*>(P.y + pix + c0) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
        }
      }
      seq0 += it.y1 - it.y0;
    }
  }
  ts_fence_before();
  __syncthreads();
  if (warp == TS_MMA_WARP) {
    ts_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TS_TMEM_COLS) : "memory");
  }
}

int ts_mode() {                                   // DISCO_TS: 0 off, 1 on
  static int m = -1;
  if (m < 0) {
    const char* e = getenv("DISCO_TS");
    m = e ? atoi(e) : 0;
  }
  return m;
}

}  // namespace

bool conv_ts_match(const disco_conv_desc* d) {
  if (!d || ts_mode() == 0 || d->dtype != DISCO_BF16 || d->kind != DISCO_CONV3 || d->stride != 1 || d->n_src != 1) return false;
  if (d->head != DISCO_HEAD_NONE || d->Cout != 64) return false;
  const disco_conv_src& s = d->src[0];
  if (s.is_f32 || s.up2 || s.C != 64 || s.H != d->Ho || s.W != d->Wo || !s.ptr) return false;
  if ((d->post_scale == nullptr) != (d->post_shift == nullptr)) return false;
  if (d->act == DISCO_ACT_LRELU && !(d->slope >= 0.f && d->slope < 1.f)) return false;
  if (d->batch <= 0 || d->Ho < 1 || d->Wo < 64) return false;            // narrow maps: the tcgen05 tile kernels fit better
  return true;
}

int64_t conv_ts_weight_elems(const disco_conv_desc*) { return 9 * 64 * 64; }

// fp32 [tap][ci][co] -> bf16 [tap][co][ci]
int conv_ts_pack(const disco_conv_desc* d, const float* w32, uint16_t* out) {
  const float* wb = w32 + d->src[0].w_off;
  for (int tap = 0; tap < 9; ++tap)
    for (int co = 0; co < 64; ++co)
      for (int ci = 0; ci < 64; ++ci) {
        uint32_t u;
        memcpy(&u, &wb[((size_t)tap * 64 + ci) * 64 + co], 4);
        out[((size_t)tap * 64 + co) * 64 + ci] =
            (u & 0x7fffffffu) > 0x7f800000u ? (uint16_t)0x7fff : (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
      }
  return DISCO_OK;
}

int conv_ts_launch(disco_handle* h, const disco_conv_desc* d, cudaStream_t st) {
  DISCO_CHECK_ARG(d->weights && d->bias && d->out, "conv_ts: null pointer");
  TsParams P;
  P.x = reinterpret_cast<const __nv_bfloat16*>(d->src[0].ptr);
  P.w = reinterpret_cast<const uint16_t*>(d->weights);
  P.bias = d->bias;
  P.post_scale = d->post_scale;
  P.post_shift = d->post_shift;
  P.res = reinterpret_cast<const __nv_bfloat16*>(d->residual);
  P.y = reinterpret_cast<__nv_bfloat16*>(d->out);
  P.B = d->batch; P.H = d->Ho; P.W = d->Wo;
  P.act = d->act;
  P.slope = d->act == DISCO_ACT_RELU ? 0.f : d->slope;
  P.strips = (P.W + 127) / 128;
  // chunks of rows: enough work items to balance the SMs, few enough that the one halo row per chunk edge stays cheap
  int rows = 32;
  while (rows > 8 && (long long)P.B * P.strips * ((P.H + rows - 1) / rows) < 4ll * h->sm_count) rows >>= 1;
  P.rows_per_chunk = rows;
  P.chunks = (P.H + rows - 1) / rows;
  const long long items = (long long)P.B * P.strips * P.chunks;
  DISCO_CHECK_ARG(items < (1ll << 30), "conv_ts: too many work items");
  P.items = (int)items;
  P.error_flag = h->error_flag;
  P.mode = getenv("DISCO_TS_MODE") ? atoi(getenv("DISCO_TS_MODE")) : 0;
  {
    if (!h->tmap_encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult q;
      DISCO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
      DISCO_CHECK_ARG(fn && q == cudaDriverEntryPointSuccess, "conv_ts: cuTensorMapEncodeTiled not available from the driver");
      h->tmap_encode = fn;
    }
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    const cuuint64_t dims[3] = {64, (cuuint64_t)P.W, (cuuint64_t)P.B * P.H};
    const cuuint64_t strides[2] = {128, (cuuint64_t)P.W * 128};
    const cuuint32_t box[3] = {64, (cuuint32_t)TS_ROW_PX, 1}, estr[3] = {1, 1, 1};
    const CUresult r = reinterpret_cast<EncodeTiledFn>(h->tmap_encode)(
        &P.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(P.x), dims, strides, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DISCO_CHECK_ARG(r == CUDA_SUCCESS, "conv_ts: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  }
  if (int rc = disco_ensure_smem(h, (const void*)conv64_ts_kernel, TS_SMEM)) return rc;
  const int grid = (int)std::min<long long>(items, h->sm_count);
  conv64_ts_kernel<<<grid, TS_THREADS, TS_SMEM, st>>>(P);
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}
