// tcgen05 implicit-GEMM convolution for sm_100a: bf16 NHWC activations, bf16 weights, fp32 accumulation in
// tensor memory, fused epilogue (bias, residual, ReLU/LeakyReLU, post-activation BN affine, softmax9 / tanh2
// heads, fp32 L-channel side input).
//
// GEMM view.  M = output pixels (a CTA tile is a TH x TW patch of NB images = 128 rows = the 128 TMEM lanes),
// N = output channels (BN <= 256 columns per tile), K = (filter tap, input-channel chunk of KC).
// The A operand is never materialised (no im2col): for K-block (tap, chunk) the producer issues ONE tiled TMA
// load of the box {KC channels, TW, TH, NB} from the NHWC tensor at the tap's (dy, dx) offset.  Hardware
// out-of-bounds zero fill implements the zero padding, and the box lands in shared memory exactly in the
// K-major, 8-row-group, hardware-swizzled layout the UMMA shared-memory descriptor expects (row = pixel,
// row pitch = KC*2 bytes = swizzle span).  Weights are pre-packed [K-block][Cout][KC] so a B tile is one 2-D box.
//
// Variants handled by the tap table built on the host (build_plan):
//   stride 1           taps (dy,dx) in {-1,0,1}^2 on the 4-D map (C, W, H, N)
//   stride 2           the source is viewed as a 5-D tensor (2C [x-parity, c], W/2, 2 [y-parity], H/2, N), so
//                      "every second pixel" is a plain box with unit traversal strides
//   nn.Upsample(x2)+conv and ConvTranspose2d(4,2,1): decomposed into 4 output-parity phases, each a 2x2-tap
//                      convolution on the LOW-resolution source (weights pre-summed per parity); 2.25x fewer MACs
//                      than convolving the up-sampled tensor and no up-sampled tensor in HBM
//   two sources        (skip concat, conv8up + conv3short8): extra taps accumulate into the same TMEM tile
//   fp32 1-channel source (the L channel beside full_feats in enhanceNet.inConv): added in the epilogue
//
// Warp roles (320 threads, persistent CTAs, static round-robin tile schedule):
//   warp 0   TMA producer (one elected lane), `stages`-deep full/empty mbarrier ring
//   warp 1   MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16 per instruction;
//            tcgen05.commit releases smem slots and publishes finished accumulators
//   warps 2-9 epilogue: tcgen05.ld (32 lanes x 16/32 columns per instruction), fused math, 16-byte vector stores.
//            Two TMEM accumulator stages (2*BN columns) overlap the epilogue of tile i with the MMAs of tile i+1.
#include "common.cuh"
#include <cuda.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

constexpr int kMaxTaps = 24;
constexpr int kThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kThreadsRes = 352;  // resident kernel: + warp 10 = second MMA issuer
constexpr long long kSpinCycles = 6000000000ll;   // mbarrier wait watchdog (~3 s) -> trap instead of hanging the GPU

struct Tap {
  int8_t src;      // source index
  int8_t mode;     // 0: 4-D map (c, x, y, n); 1: 5-D stride-2 map (pc, X, py, Y, n)
  int8_t oy, ox;   // offset added to the tile origin in the source's (Y, X) index space
  int8_t py, px;   // parities (mode 1)
  int16_t nchunks; // input-channel chunks of KC
  int32_t wkb0;    // first K-block row-group in the packed weight matrix
  int32_t c_base;  // channel offset of chunk 0 inside the inner dimension (mode 1: px*C)
};

// One pipeline step of the resident-weight kernel: ONE TMA load of an activation sub-tile that serves up to three
// filter taps (the taps of one filter column: same dx, dy = -1,0,+1 -> the box carries TH+2 rows and each tap's
// A operand starts `dy*TW` rows further down, which keeps the start address on a swizzle-atom boundary).
struct Step {
  int8_t src, mode, ox, oy, py, px, ntap;
  int8_t nch;          // grouped streaming kernel: input-channel chunks this step iterates over (resident kernel: 1)
  int32_t c0;          // inner-dimension start coordinate (channel chunk, + px*C for the stride-2 view)
  uint32_t bytes;      // bytes the load delivers
  uint32_t a_sbo16;    // byte distance between 8-row groups of the A operand, >> 4 (halo tiles: (TW+2) pixels)
  uint32_t mt_off16;   // byte distance between the two 128-pixel sub-tiles of a 256-pixel tile inside the box, >> 4
  uint32_t tap[9];     // (A start offset in bytes >> 4) | (resident weight block index << 16)
};
constexpr int kMaxSteps = 16;

struct TcParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  const uint32_t* tables;     // device copy of {Step steps[4][kMaxSteps]; Tap taps[4][kMaxTaps]} (staged to smem at start;
                              // keeping them out of the parameter block keeps it < 1 KB -> constant-cache resident)
  int32_t nsteps[4];
  int32_t wkb_phase0[4];      // first weight K-block of each phase in the packed matrix
  int32_t res_stages, res_a_stage_bytes, res_b_bytes;
  int32_t grp_b_stages;       // grouped streaming kernel: weight-tile ring depth (res_stages = activation-box ring depth)
  int32_t res_dual;           // 1: two MMA issuers alternate tiles (only when their pipeline-stage sets are disjoint)
  int32_t halo_bo_mask;       // 0: A descriptors carry base_offset 0; else mask applied to (start_addr >> 7)
  int32_t ntaps[4];
  int32_t kblocks[4];         // K-blocks per phase
  int32_t n_phase, os;        // phases (1 | 4), output stride of a phase grid (1 | 2)
  int32_t B, Hg, Wg;          // phase-grid extent
  int32_t TW, TH, NB;         // tile: TW*TH*NB == 128 (powers of two)
  int32_t tw_log2, th_log2;
  int32_t tiles_x, tiles_y, tiles_b, tiles_n, tiles_total;
  int32_t cout_pad;           // rows per K-block in the packed weight matrix
  // epilogue
  int32_t Ho, Wo, Cout, act, head;
  float slope;
  const float* bias;
  const float* post_scale;
  const float* post_shift;
  const __nv_bfloat16* residual;
  void* out;
  // optional fp32 single-channel 3x3 side input
  const float* gray;
  const float* gray_w;        // [9][Cout]
  int32_t* error_flag;
  // resident kernels, when the caller supplies host copies of the epilogue parameters: bias | post_scale | post_shift
  // travel in the kernel-parameter block, so the epilogue's FADD/FFMA read them through uniform registers / the
  // constant bank instead of the per-warp shared-memory cache (measured: 7..12 % faster 16/32-channel layers, neutral at
  // 64 channels; the L-channel side-input layer keeps the shared-memory path: 9 x 32 more constants per chunk made it slower).
  float epi_c[3][64];
  int32_t const_params;
  int32_t direct_store;       // 1: epilogue lanes store their 64-byte channel run with two 32-byte stores (no smem staging)
  int32_t dbg_mode;           // experiments: 1 = epilogue skips TMEM loads + stores, 2 = skips stores only, 3 = producer loads once
  long long* dbg;             // optional timeline buffer (DISCO_TC_DEBUG=1): [role][tile][slot] clock64 stamps of CTA 0
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int32_t* error_flag) {
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if ((it & 1023) == 1023) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kSpinCycles) {
        if (error_flag) atomicExch(error_flag, 1);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants: loads land in the executing CTA's shared memory, the transaction bytes are
// signalled on the LEADER CTA's mbarrier (`bar_cluster` = shared::cluster address obtained with mapa)
// Programmatic dependent launch: every tensor-core kernel is launched with programmaticStreamSerialization, so its CTAs
// may become resident (barrier setup, tensor-memory allocation, weight loads) while the tail of the previous kernel is
// still draining; nothing that depends on the previous kernel's output may be touched before pdl_wait().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_5d(void* dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1, int c2,
                                             int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void lds2x64(uint32_t saddr, f32x2& a, f32x2& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(saddr));
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
// 32-byte (one full sector) global store
__device__ __forceinline__ void stg256(void* p, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}
constexpr int kDbgTiles = 48, kDbgSlots = 8;
__device__ __forceinline__ void dbg_stamp(const TcParams& P, int role, int tile_i, int slot) {
#ifdef DISCO_TC_TIMELINE   // compile-time only: the checks cost issue slots in the single-thread role loops
  if (P.dbg && blockIdx.x == 0 && tile_i < kDbgTiles && slot < kDbgSlots)
    P.dbg[(role * kDbgTiles + tile_i) * kDbgSlots + slot] = clock64();
#endif
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// cta_group::2: one instruction multiplies the 256 x K operand held as two 128-row halves (one per CTA of the pair) with
// the N x K operand held as two N/2-row halves; each CTA's tensor memory receives its own 128 accumulator rows.
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N]) {
  if constexpr (N == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
}

// K-major operand tile in shared memory, rows of `KC*2` bytes = one swizzle span, 8-row groups contiguous.
template <int KC>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  constexpr uint64_t layout = KC == 64 ? 2 : (KC == 32 ? 4 : 6);   // SWIZZLE_128B / 64B / 32B
  constexpr uint64_t sbo = (8 * KC * 2) >> 4;                      // byte distance between 8-row groups
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

template <int BN, int M = 128>
__host__ __device__ constexpr uint32_t instr_desc() {
  // c_format f32 [4,6)=1, a/b format bf16 [7,10)/[10,13)=1, K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN, int KC, int MT = 1, bool PAIR = false>
struct Cfg {
  // MT = 128-pixel sub-tiles per CTA tile.  MT = 2 (narrow N <= 128 tiles only): one TMA box brings 256 pixels and
  // both halves are multiplied against the SAME weight tile, which halves the weight traffic per MAC -- N = 128
  // tiles at 125 B/clk/SM of operand fill are otherwise bound by the L2->SM fabric (~107 B/clk/SM measured).
  static constexpr int A_BYTES = 128 * MT * KC * 2;
  // PAIR: the two CTAs of a cluster run cta_group::2 MMAs (M = 256) on two horizontally adjacent pixel tiles; each
  // CTA stages only HALF of the weight tile (BN/2 rows), which halves the weight L2->SMEM fill and the B-operand
  // shared-memory reads per MMA (N = 128 tiles are otherwise bound by the 128 B/clk shared-memory read port).
  static constexpr int B_BYTES = BN * KC * 2 / (PAIR ? 2 : 1);
  // K-blocks per pipeline stage: with MT = 1 narrow tiles finish a K-block's MMAs in <= 256 cycles, which does not
  // cover an mbarrier round trip, so two K-blocks share one stage / one full-empty handshake
  static constexpr int KB = (BN <= 128 && MT == 1) ? 2 : 1;
  static constexpr int STAGE_BYTES = KB * (A_BYTES + B_BYTES);
  static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = 2 * MT * BN < 32 ? 32 : 2 * MT * BN;
  // per-epilogue-warp parameter cache: bias | post_scale | post_shift (+ 9 x fp32 L-channel weights when BN <= 64)
  static constexpr int EPI_COLS = (BN >= 32 && MT == 1) ? BN / 2 : BN;
  static constexpr int EPI_FLOATS = 3 * EPI_COLS + (BN <= 64 ? 9 * EPI_COLS : 0);
  static constexpr int EPI_BYTES = 8 * EPI_FLOATS * 4 + 8 * 2048;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + EPI_BYTES;
};

// Mixed-radix tile counter (n fastest, then x, y, image, phase) advanced by gridDim.x per iteration without
// integer division in the loop.
struct TileIter {
  int nt, xt, yt, bt, phase;
  int dn, dx, dy, db, dp;
  __device__ __forceinline__ void init(const TcParams& P, int tile, int step) {
    int t = tile;
    nt = t % P.tiles_n; t /= P.tiles_n;
    xt = t % P.tiles_x; t /= P.tiles_x;
    yt = t % P.tiles_y; t /= P.tiles_y;
    bt = t % P.tiles_b; phase = t / P.tiles_b;
    t = step;
    dn = t % P.tiles_n; t /= P.tiles_n;
    dx = t % P.tiles_x; t /= P.tiles_x;
    dy = t % P.tiles_y; t /= P.tiles_y;
    db = t % P.tiles_b; dp = t / P.tiles_b;
  }
  __device__ __forceinline__ void next(int tn, int tX, int tY, int tB) {
    nt += dn; int c = nt >= tn; nt -= c ? tn : 0;
    xt += dx + c; c = xt >= tX; xt -= c ? tX : 0;
    yt += dy + c; c = yt >= tY; yt -= c ? tY : 0;
    bt += db + c; c = bt >= tB; bt -= c ? tB : 0;
    phase += dp + c;
  }
  __device__ __forceinline__ void next(const TcParams& P) { next(P.tiles_n, P.tiles_x, P.tiles_y, P.tiles_b); }
};


// ------------------------------------------------------------------------------------------------ epilogue
// Warps 2..9 (8 warps): TMEM -> registers -> fused math -> global.  Shared by both kernels.  Warp w reads TMEM
// lane quarter (w & 3) (a hardware restriction) and, when BN >= 32, the column half ((w - 2) >> 2): two warps per
// SM sub-partition keep the epilogue's issue rate up (one warp alone runs at IPC ~0.2 on dependent fp32 math).
template <class T> struct CVal { static constexpr int value = 0; };
template <int V> struct CVal<std::integral_constant<int, V>> { static constexpr int value = V; };
__device__ __forceinline__ int cval_rt(int v) { return v; }
template <int V> __device__ __forceinline__ int cval_rt(std::integral_constant<int, V>) { return V; }

// ALT (resident kernel): the two warps of a lane quarter take ALTERNATE tiles and all BN columns each, instead of
// splitting the columns of every tile.  Timeline stamps showed ~2300 cycles per tile and warp for 64-channel tiles, of
// which ~1000 are per-tile fixed cost (barrier wait, tile bookkeeping, address math): the epilogue, not the tensor pipe,
// paced these layers.  Alternating halves the fixed cost per tile.
template <int BN, int MT = 1, bool PAIR = false, int NACC = 2, bool CP = false, bool ALT = false>
__device__ __forceinline__ void epilogue_role(const TcParams& P, uint32_t tmem_base, uint64_t* tfull, uint64_t* tempty,
                                              float* epi_params, int warp, int lane) {
  // PAIR: this CTA owns pixel-tile column 2*xt + rank of the cluster's tile pair; accumulator-free signals go to the
  // leader CTA's barrier (the only MMA issuer of the pair)
  const int crank = PAIR ? (int)cluster_ctarank() : 0;
  const int bid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int gdim = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t tempty_leader = PAIR ? mapa_u32(smem_u32(tempty), 0) : 0;
  pdl_wait();                                   // residual / L-channel reads and all stores come after the previous kernel
  constexpr int EPI_STAGE = 2048;               // per-warp output staging patch: 32 pixels x 64 B
  constexpr int CH = BN >= 64 ? 32 : 16;        // accumulator columns per tcgen05.ld
  // MT == 1: the two warps of a lane quarter split the columns.  MT == 2: they take one 128-pixel sub-tile each.
  static_assert(!ALT || MT == 1, "alternating epilogue: single 128-pixel tiles only");
  static_assert(!CP || ALT, "constant-bank parameters are used by the resident kernel only");
  constexpr int NHALF = ALT ? 1 : ((BN >= 32 && MT == 1) ? 2 : 1);   // column halves
  constexpr int COLS = BN / NHALF;              // columns this warp owns
  // per-warp cache: bias|scale|shift(|9 gray taps); the resident kernel sizes it at run time (host mirrors this)
  const int EPI_FLOATS = ALT ? (3 * COLS + (P.gray ? 9 * COLS : 0)) : (3 * COLS + (BN <= 64 ? 9 * COLS : 0));
  const int q = warp & 3;
  const int half = (warp - 2) >> 2;
  const bool active = ALT ? true : half < NHALF * MT;
  const int sub = MT == 2 ? half : 0;           // 128-pixel sub-tile of this warp
  const int m = sub * 128 + q * 32 + lane;      // pixel within the tile
  const int tx = m & (P.TW - 1), ty = (m >> P.tw_log2) & (P.TH - 1), nb = m >> (P.tw_log2 + P.th_log2);
  const uint32_t wp = smem_u32(epi_params + (warp - 2) * EPI_FLOATS);   // this warp's private parameter cache
  const uint32_t stg = smem_u32(epi_params + 8 * EPI_FLOATS) + (warp - 2) * EPI_STAGE;   // output staging patch
  const int col0 = (MT == 2 || ALT) ? 0 : half * COLS;
  // hot parameters hoisted out of the tile loop (the parameter block is several KB: re-reading it through the
  // constant cache inside the loop stalls on misses)
  const int pTW = P.TW, pTH = P.TH, pNB = P.NB, pWg = P.Wg, pHg = P.Hg, pB = P.B, pos = P.os, pHo = P.Ho, pWo = P.Wo;
  const int pCout = P.Cout, pact = P.act, phead = P.head, ptotal = P.tiles_total;
  const float pslope = P.slope;
  const bool has_post = P.post_scale != nullptr;
  const bool pdirect = P.direct_store != 0;
  const __nv_bfloat16* const pres = P.residual;
  __nv_bfloat16* const pout = reinterpret_cast<__nv_bfloat16*>(P.out);
  int32_t* const perr = P.error_flag;
  int cached_nt = -1;
  uint32_t as = 0, aph = 0;
  TileIter it;
  it.init(P, bid, gdim);
  const int tn = P.tiles_n, tX = P.tiles_x, tY = P.tiles_y, tB = P.tiles_b;
  int li = 0;                                   // local tile index
  for (int tile = bid; tile < ptotal; tile += gdim, it.next(tn, tX, tY, tB), ++li) {
    if constexpr (ALT) {
      if ((li & 1) != half) continue;           // the other warp of this lane quarter owns this tile
      as = (uint32_t)li & (NACC - 1);
      aph = ((uint32_t)li / NACC) & 1;
    }
    const int phase = it.phase, nt = it.nt;
    const int X = (PAIR ? 2 * it.xt + crank : it.xt) * pTW + tx, Y = it.yt * pTH + ty, b = it.bt * pNB + nb;
    const bool valid = active && X < pWg && Y < pHg && b < pB;
    const int oy = Y * pos + (phase >> 1), ox = X * pos + (phase & 1);
    const size_t pix = ((size_t)b * pHo + oy) * pWo + ox;
    const int n0 = nt * BN;
    if (!CP && nt != cached_nt) {
      for (int j = lane; j < COLS; j += 32) {
        const int n = n0 + col0 + j;
        const bool ok = n < pCout;
        sts32(wp + 4 * j, ok ? P.bias[n] : 0.f);
        sts32(wp + 4 * (COLS + j), (ok && P.post_scale) ? P.post_scale[n] : 1.f);
        sts32(wp + 4 * (2 * COLS + j), (ok && P.post_shift) ? P.post_shift[n] : 0.f);
        if constexpr (BN <= 64) {
          if (P.gray) {
#pragma unroll
            for (int t = 0; t < 9; ++t) sts32(wp + 4 * ((3 + t) * COLS + j), ok ? P.gray_w[t * pCout + n] : 0.f);
          }
        }
      }
      cached_nt = nt;
      __syncwarp();
    }
    float g[9];
    if constexpr (BN <= 64) {
      if (P.gray) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int gy = oy + t / 3 - 1, gx = ox + t % 3 - 1;
          g[t] = (valid && gy >= 0 && gy < pHo && gx >= 0 && gx < pWo) ? P.gray[((size_t)b * pHo + gy) * pWo + gx] : 0.f;
        }
      }
    }
    // residual of the first column chunk: issued before the accumulator wait so its latency hides behind the MMAs
    uint4 rpre[CH / 8];
    if (pres != nullptr && valid && n0 + col0 < pCout) {
      const uint4* rp = reinterpret_cast<const uint4*>(pres + pix * pCout + n0 + col0);
#pragma unroll
      for (int u = 0; u < CH / 8; ++u) rpre[u] = __ldg(rp + u);
    }
    if (lane == 0 && (warp == 2 || warp == 6)) dbg_stamp(P, warp == 2 ? 2 : 3, (tile - bid) / gdim, 0);
    mbar_wait(&tfull[as], aph, perr);
    tc_fence_after();
    if (lane == 0 && (warp == 2 || warp == 6)) dbg_stamp(P, warp == 2 ? 2 : 3, (tile - bid) / gdim, 1);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * (MT * BN) + sub * BN;
    if (phead == DISCO_HEAD_NONE) {
      // CP: the warp's column half is a compile-time constant inside each instantiation of this lambda, so the
      // parameter reads below are constant-bank operands with immediate offsets (P.epi_c[k][HALF * COLS + j])
      // one accumulator chunk of CH columns starting at column c0v (a compile-time constant in CP mode, so that the
      // parameter reads are constant-bank operands with immediate offsets)
      auto do_chunk = [&](auto c0v) {
        const int c0 = cval_rt(c0v);
        {
          uint32_t r[CH];
          tmem_ld<CH>(taddr + c0, r);
          if (lane == 0 && (warp == 2 || warp == 6)) dbg_stamp(P, warp == 2 ? 2 : 3, li, 3 + (c0 != col0 ? 2 : 0));
          if (valid && n0 + c0 < pCout) {
            float v[CH];
            if constexpr (CP) {
#pragma unroll
              for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]) + P.epi_c[0][CVal<decltype(c0v)>::value + j];
            } else {
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              const float4 bv = lds128(wp + 4 * (c0 - col0 + j));
              v[j] = __uint_as_float(r[j]) + bv.x; v[j + 1] = __uint_as_float(r[j + 1]) + bv.y;
              v[j + 2] = __uint_as_float(r[j + 2]) + bv.z; v[j + 3] = __uint_as_float(r[j + 3]) + bv.w;
            }
            if constexpr (BN <= 64) {
              if (P.gray) {
                // fp32 L-channel taps as packed fp32 pairs (FFMA2): 9 x CH/2 instead of 9 x CH FMAs per pixel, the
                // weight pairs come straight out of LDS.128 as two 64-bit operands
                f32x2 v2[CH / 2];
#pragma unroll
                for (int j = 0; j < CH / 2; ++j) v2[j] = pack2(v[2 * j], v[2 * j + 1]);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                  const f32x2 G = pack2(g[t], g[t]);
#pragma unroll
                  for (int j = 0; j < CH; j += 4) {
                    f32x2 w01, w23;
                    lds2x64(wp + 4 * ((3 + t) * COLS + c0 - col0 + j), w01, w23);
                    v2[j / 2] = ffma2(G, w01, v2[j / 2]);
                    v2[j / 2 + 1] = ffma2(G, w23, v2[j / 2 + 1]);
                  }
                }
#pragma unroll
                for (int j = 0; j < CH / 2; ++j) unpack2(v2[j], v[2 * j], v[2 * j + 1]);
              }
            }
            }
            if (pres) {
              const uint4* rp = reinterpret_cast<const uint4*>(pres + pix * pCout + n0 + c0);
#pragma unroll
              for (int u = 0; u < CH / 8; ++u) {
                const uint4 rr = (c0 == col0) ? rpre[u] : __ldg(rp + u);
                const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[j]);
                  v[u * 8 + 2 * j] += __low2float(h2);
                  v[u * 8 + 2 * j + 1] += __high2float(h2);
                }
              }
            }
            if (pact == DISCO_ACT_RELU) {
#pragma unroll
              for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], 0.f);
            } else if (pact == DISCO_ACT_LRELU) {
#pragma unroll
              for (int j = 0; j < CH; ++j) v[j] = fmaxf(v[j], v[j] * pslope);   // slope in [0,1)
            }
            if (has_post && CP) {
#pragma unroll
              for (int j = 0; j < CH; ++j)
                v[j] = fmaf(v[j], P.epi_c[1][CVal<decltype(c0v)>::value + j], P.epi_c[2][CVal<decltype(c0v)>::value + j]);
            } else if (has_post) {
#pragma unroll
              for (int j = 0; j < CH; j += 4) {
                const float4 sv = lds128(wp + 4 * (COLS + c0 - col0 + j));
                const float4 hv = lds128(wp + 4 * (2 * COLS + c0 - col0 + j));
                v[j] = fmaf(v[j], sv.x, hv.x); v[j + 1] = fmaf(v[j + 1], sv.y, hv.y);
                v[j + 2] = fmaf(v[j + 2], sv.z, hv.z); v[j + 3] = fmaf(v[j + 3], sv.w, hv.w);
              }
            }
            if constexpr (CH == 16) {
              uint4* op = reinterpret_cast<uint4*>(pout + pix * pCout + n0 + c0);
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[u * 8 + 2 * j], v[u * 8 + 2 * j + 1]);
                  w[j] = *reinterpret_cast<uint32_t*>(&h2);
                }
                op[u] = make_uint4(w[0], w[1], w[2], w[3]);
              }
            } else if (pdirect) {
              // the shared-memory port is what bounds narrow-N tiles (operand reads of the MMAs): keep the epilogue off it.
              // Each lane owns 64 contiguous bytes of its pixel = two full 32-byte sectors.
              uint32_t w[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                w[j] = *reinterpret_cast<uint32_t*>(&h2);
              }
              if (lane == 0 && (warp == 2 || warp == 6)) dbg_stamp(P, warp == 2 ? 2 : 3, li, 4 + (c0 != col0 ? 2 : 0));
              if (P.dbg_mode != 2) {
                __nv_bfloat16* op = pout + pix * pCout + n0 + c0;
                stg256(op, w);
                stg256(op + 16, w + 8);
              }
            } else {
              // stage this lane's 64 B (32 channels) in the warp's patch, XOR-swizzled on the 16-byte piece index
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[u * 8 + 2 * j], v[u * 8 + 2 * j + 1]);
                  w[j] = *reinterpret_cast<uint32_t*>(&h2);
                }
                sts128(stg + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4), w[0], w[1], w[2], w[3]);
              }
            }
          }
          if (CH == 32 && !pdirect) {
            // coalesced write-out: 4 lanes cover one pixel's 64 B, a warp instruction writes 8 pixels x 64 B
            __syncwarp();
            const bool chunk_ok = n0 + c0 < pCout;
            const __nv_bfloat16* mybase = pout + pix * pCout + n0 + c0;
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2) {
              const int p = (lane >> 2) + 8 * i2, piece = lane & 3;
              const unsigned long long a = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)mybase, p);
              const int ok = __shfl_sync(0xffffffffu, (int)valid, p);
              const uint4 val = lds128u(stg + p * 64 + ((piece ^ ((p >> 1) & 3)) << 4));
              if (ok && chunk_ok && P.dbg_mode != 2) *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(a) + piece * 16) = val;
            }
            __syncwarp();
          }
        }
      };
      if (active && P.dbg_mode != 1 && P.dbg_mode != 5) {
        if constexpr (CP) {
          static_assert(COLS <= 2 * CH, "at most two chunks per warp");
          do_chunk(std::integral_constant<int, 0>{});
          if constexpr (COLS > CH) do_chunk(std::integral_constant<int, CH>{});
        } else {
#pragma unroll 1
          for (int c0 = col0; c0 < col0 + COLS; c0 += CH) do_chunk(c0);
        }
      }
    } else if (active) {
      uint32_t r[16];
      tmem_ld16(taddr, r);
      if (valid) {
        float* outp = reinterpret_cast<float*>(P.out);
        const size_t plane = (size_t)pHo * pWo;
        const size_t base = (size_t)b * pCout * plane + (size_t)oy * pWo + ox;
        if (phead == DISCO_HEAD_SOFTMAX9) {
          float v[9], mx = -3.4e38f, s = 0.f;
#pragma unroll
          for (int j = 0; j < 9; ++j) { v[j] = __uint_as_float(r[j]) + (CP ? P.epi_c[0][j] : lds32(wp + 4 * j)); mx = fmaxf(mx, v[j]); }
#pragma unroll
          for (int j = 0; j < 9; ++j) { v[j] = expf(v[j] - mx); s += v[j]; }
          const float inv = 1.0f / s;
#pragma unroll
          for (int j = 0; j < 9; ++j) outp[base + j * plane] = v[j] * inv;
        } else {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float y = __uint_as_float(r[j]) + (CP ? P.epi_c[0][j] : lds32(wp + 4 * j));
            outp[base + j * plane] = phead == DISCO_HEAD_TANH2 ? tanhf(y) : y;      // DISCO_HEAD_RAW2: pre-tanh map
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if constexpr (PAIR) mbar_arrive_cluster(tempty_leader + 8 * as);
      else mbar_arrive(&tempty[as]);
    }
    if (lane == 0 && (warp == 2 || warp == 6)) dbg_stamp(P, warp == 2 ? 2 : 3, (tile - bid) / gdim, 2);
    if constexpr (!ALT) {
      if (++as == NACC) { as = 0; aph ^= 1; }
    }
  }
}

template <int BN, int KC, int MT, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ TcParams P) {
  using C = Cfg<BN, KC, MT, PAIR>;
  // PAIR: launched as clusters of two CTAs; cluster c works on tile pairs c, c + #clusters, ... (P.tiles_x and
  // P.tiles_total count PAIRS of horizontally adjacent pixel tiles)
  const int crank = PAIR ? (int)cluster_ctarank() : 0;
  const int bid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int gdim = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                        // [stage][KB] A tiles
  uint8_t* smem_b = smem + C::STAGES * C::KB * C::A_BYTES;       // [stage][KB] B tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tfull = bars + 2 * C::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* epi_params = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ Tap s_taps[4 * kMaxTaps];
  {
    const uint32_t* src = P.tables + sizeof(Step) * 4 * kMaxSteps / 4;
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_taps);
    for (int i = threadIdx.x; i < (int)(sizeof(Tap) * 4 * kMaxTaps / 4); i += blockDim.x) dst[i] = src[i];
  }

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], PAIR ? 16 : 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      // issued by the same warp of both CTAs; both receive the same column range
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)C::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)C::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();     // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      pdl_wait();
      uint32_t stage = 0, ph = 0;
      const uint32_t full_leader = PAIR ? mapa_u32(smem_u32(full), 0) : 0;
      const bool skip_a = P.dbg_mode == 3, skip_b = P.dbg_mode == 4;     // experiments: operand fill switched off
      const uint32_t stage_tx = (skip_a ? 0 : C::A_BYTES) + (skip_b ? 0 : C::B_BYTES);
      TileIter it;
      it.init(P, bid, gdim);
      for (int tile = bid; tile < P.tiles_total; tile += gdim, it.next(P)) {
        const int phase = it.phase, nt = it.nt;
        const int x0 = (PAIR ? 2 * it.xt + crank : it.xt) * P.TW, y0 = it.yt * P.TH, b0 = it.bt * P.NB;
        const int ntap = P.ntaps[phase];
        int sub = 0;                        // K-blocks already issued into the current stage
        int left = P.kblocks[phase];        // K-blocks of this tile not yet issued
        for (int t = 0; t < ntap; ++t) {
          const Tap tp = s_taps[phase * kMaxTaps + t];
          for (int ch = 0; ch < tp.nchunks; ++ch) {
            if (sub == 0) {
              mbar_wait(&empty[stage], ph ^ 1, P.error_flag);
              const int n = left < C::KB ? left : C::KB;
              // PAIR: the leader's barrier collects the bytes of both CTAs' loads
              if (!PAIR || crank == 0) mbar_expect_tx(&full[stage], (uint32_t)((PAIR ? 2 : 1) * n) * stage_tx);
            }
            void* da = smem_a + (stage * C::KB + sub) * C::A_BYTES;
            void* db = smem_b + (stage * C::KB + sub) * C::B_BYTES;
            if constexpr (PAIR) {
              const uint32_t fb = full_leader + 8 * stage;
              if (skip_a) {
              } else if (tp.mode == 0)
                tma2_load_4d(da, &P.tmA[tp.src], fb, tp.c_base + ch * KC, x0 + tp.ox, y0 + tp.oy, b0);
              else
                tma2_load_5d(da, &P.tmA[tp.src], fb, tp.c_base + ch * KC, x0 + tp.ox, tp.py, y0 + tp.oy, b0);
              if (!skip_b) tma2_load_2d(db, &P.tmB, fb, 0, (tp.wkb0 + ch) * P.cout_pad + nt * BN + crank * (BN / 2));
            } else {
              if (skip_a) {
              } else if (tp.mode == 0)
                tma_load_4d(da, &P.tmA[tp.src], &full[stage], tp.c_base + ch * KC, x0 + tp.ox, y0 + tp.oy, b0);
              else
                tma_load_5d(da, &P.tmA[tp.src], &full[stage], tp.c_base + ch * KC, x0 + tp.ox, tp.py, y0 + tp.oy, b0);
              if (!skip_b) tma_load_2d(db, &P.tmB, &full[stage], 0, (tp.wkb0 + ch) * P.cout_pad + nt * BN);
            }
            --left;
            if (++sub == C::KB || left == 0) {
              sub = 0;
              if (++stage == C::STAGES) { stage = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (PAIR: leader CTA only)
    if ((!PAIR || crank == 0) && elect_one()) {
      constexpr uint32_t idesc = instr_desc<BN, PAIR ? 256 : 128>();
      constexpr int NK = KC / 16;
      const uint64_t desc0 = make_smem_desc<KC>(0);
      const uint32_t a_base16 = (smem_u32(smem_a) & 0x3FFFF) >> 4, b_base16 = (smem_u32(smem_b) & 0x3FFFF) >> 4;
      uint32_t stage = 0, ph = 0, as = 0, aph = 0;
      const int total = P.tiles_total, tn = P.tiles_n, tX = P.tiles_x, tY = P.tiles_y, tB = P.tiles_b;
      const int kbs[4] = {P.kblocks[0], P.kblocks[1], P.kblocks[2], P.kblocks[3]};
      int32_t* const perr = P.error_flag;
      TileIter it;
      it.init(P, bid, gdim);
      for (int tile = bid; tile < total; tile += gdim, it.next(tn, tX, tY, tB)) {
        const int nkb = kbs[it.phase];
        mbar_wait(&tempty[as], aph ^ 1, perr);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (MT * BN);
        for (int kb = 0; kb < nkb; kb += C::KB) {
          mbar_wait(&full[stage], ph, perr);
          tc_fence_after();
#pragma unroll
          for (int sb = 0; sb < C::KB; ++sb) {
            if (kb + sb < nkb) {
              const uint64_t ad = desc0 + (uint64_t)((a_base16 + (stage * C::KB + sb) * (C::A_BYTES >> 4)) & 0x3fffu);
              const uint64_t bd = desc0 + (uint64_t)((b_base16 + (stage * C::KB + sb) * (C::B_BYTES >> 4)) & 0x3fffu);
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                  if constexpr (PAIR)
                    umma2_bf16(d_tmem + mt * BN, ad + mt * (128 * KC * 2 >> 4) + 2 * k, bd + 2 * k, idesc, (kb | sb | k) != 0);
                  else
                    umma_bf16(d_tmem + mt * BN, ad + mt * (128 * KC * 2 >> 4) + 2 * k, bd + 2 * k, idesc, (kb | sb | k) != 0);
                }
              }
            }
          }
          if constexpr (PAIR) umma2_commit(&empty[stage]); else umma_commit(&empty[stage]);
          if (++stage == (uint32_t)C::STAGES) { stage = 0; ph ^= 1; }
        }
        if constexpr (PAIR) umma2_commit(&tfull[as]); else umma_commit(&tfull[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else {
    epilogue_role<BN, MT, PAIR>(P, tmem_base, tfull, tempty, epi_params, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();     // neither CTA leaves (or frees tensor memory) while the pair is still working
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Resident-weight kernel for layers whose per-phase weights fit in shared memory (Cout <= 64 at full
// resolution, heads, the 16/32-channel SpixelNet layers).  Differences from the streaming kernel above:
//   * the weight matrix of the current phase is loaded ONCE per CTA (re-loaded only at a phase change) and
//     every tile's MMAs read it in place -> no per-K-block weight traffic;
//   * activation loads are grouped by filter column: one TMA box with TH+2 rows feeds the three taps
//     dy = -1,0,+1, so a 3x3 layer costs 3 activation loads per tile instead of 9 (2.4x less L2->SMEM traffic)
//     and 3 instead of 9 mbarrier round trips, which is what bounds N <= 64 tiles (their MMAs are short).
// Shared-memory layout: [stages x A stage][resident weights][barriers][epilogue parameter caches].
// ------------------------------------------------------------------------------------------------
// PAIR: clusters of two CTAs on horizontally adjacent tiles run cta_group::2 MMAs (M = 256): each CTA keeps only half of
// every weight block resident (N/2 rows) -- 5 KB instead of 6 KB of shared-memory operand reads per N = 64 MMA (the read
// port is what paces them) -- and ONE thread issues the MMAs of both SMs, which halves the per-tile issue cost.
template <int BN, int KC, bool CP, bool PAIR>
__global__ void __launch_bounds__(kThreadsRes, 1) conv_tc_res_kernel(const __grid_constant__ TcParams P) {
  constexpr int B_BYTES = BN * KC * 2 / (PAIR ? 2 : 1);
  const int crank = PAIR ? (int)cluster_ctarank() : 0;
  const int bid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int gdim = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int MAX_ST = 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = P.res_stages, a_stage = P.res_a_stage_bytes;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + stages * a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + P.res_b_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_ST;
  // four accumulator stages (narrow tiles leave tensor memory mostly empty): the MMA issuers run up to three tiles
  // ahead of the epilogue, so its latency (TMEM load, math, stores) never stalls the tensor pipe
  constexpr int NACC = 4;
  uint64_t* tfull = bars + 2 * MAX_ST;
  uint64_t* tempty = tfull + NACC;
  uint64_t* bfull = tempty + NACC;
  uint64_t* bempty = bfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bempty + 1);
  float* epi_params = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
  constexpr int TMEM_COLS = NACC * BN < 32 ? 32 : NACC * BN;
  // step tables in shared memory: indexed reads of the multi-KB kernel-parameter block go through the constant
  // cache and cost ~60+ dependent cycles per tap inside the single-thread issue loops
  __shared__ Step s_steps[4 * kMaxSteps];
  __shared__ int s_nsteps[4];
  {
    const uint32_t* src = P.tables;
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_steps);
    for (int i = threadIdx.x; i < (int)(sizeof(Step) * 4 * kMaxSteps / 4); i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x < 4) s_nsteps[threadIdx.x] = P.nsteps[threadIdx.x];
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < NACC; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], PAIR ? 8 : 4); }   // 4 warps (per CTA) own a tile
    mbar_init(bfull, 1);
    mbar_init(bempty, P.res_dual ? 2 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_phase = P.tiles_n * P.tiles_x * P.tiles_y * P.tiles_b;
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      uint32_t stage = 0, ph = 0, beph = 0;
      int cur_phase = -1;
      const bool lead = !PAIR || crank == 0;
      const uint32_t mult = PAIR ? 2u : 1u;
      const uint32_t full_leader = PAIR ? mapa_u32(smem_u32(full), 0) : 0;
      const uint32_t bfull_leader = PAIR ? mapa_u32(smem_u32(bfull), 0) : 0;
      TileIter it;
      it.init(P, bid, gdim);
      for (int tile = bid; tile < P.tiles_total; tile += gdim, it.next(P)) {
        const int phase = it.phase, bt = it.bt;
        if (phase != cur_phase) {
          if (cur_phase >= 0) { mbar_wait(bempty, beph, P.error_flag); beph ^= 1; }   // MMAs of the old phase are done
          const int nkb = P.kblocks[phase];
          if (lead) mbar_expect_tx(bfull, mult * (uint32_t)(nkb * B_BYTES));
          for (int kb = 0; kb < nkb; ++kb) {
            const int row = (P.wkb_phase0[phase] + kb) * P.cout_pad + (PAIR ? crank * (BN / 2) : 0);
            if constexpr (PAIR) tma2_load_2d(smem_b + kb * B_BYTES, &P.tmB, bfull_leader, 0, row);
            else tma_load_2d(smem_b + kb * B_BYTES, &P.tmB, bfull, 0, row);
          }
          cur_phase = phase;
        }
        const int x0 = (PAIR ? 2 * it.xt + crank : it.xt) * P.TW, y0 = it.yt * P.TH;
        const int ns = s_nsteps[phase];
        if (tile == bid) pdl_wait();     // the weights above are static; activations come from the previous kernel
        for (int i = 0; i < ns; ++i) {
          const Step& sp = s_steps[phase * kMaxSteps + i];
          mbar_wait(&empty[stage], ph ^ 1, P.error_flag);
          const bool skip_a = P.dbg_mode == 3 || P.dbg_mode == 5;     // experiment: activation fill switched off
          if (lead) mbar_expect_tx(&full[stage], skip_a ? 0u : mult * sp.bytes);
          void* da = smem_a + stage * a_stage;
          if (skip_a) {
          } else if constexpr (PAIR) {
            if (sp.mode == 0)
              tma2_load_4d(da, &P.tmA[sp.src], full_leader + 8 * stage, sp.c0, x0 + sp.ox, y0 + sp.oy, bt);
            else
              tma2_load_5d(da, &P.tmA[sp.src], full_leader + 8 * stage, sp.c0, x0 + sp.ox, sp.py, y0 + sp.oy, bt);
          } else if (sp.mode == 0)
            tma_load_4d(da, &P.tmA[sp.src], &full[stage], sp.c0, x0 + sp.ox, y0 + sp.oy, bt);
          else
            tma_load_5d(da, &P.tmA[sp.src], &full[stage], sp.c0, x0 + sp.ox, sp.py, y0 + sp.oy, bt);
          dbg_stamp(P, 0, (tile - bid) / gdim, i);
          if (++stage == (uint32_t)stages) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ===================================================================== MMA issuers (two, alternating tiles)
    // Issuer r owns the CTA's tiles with local index parity r (TMEM accumulator stage = local index mod NACC).  While one issuer sits in
    // its mbarrier waits / loop bookkeeping (~1000 cycles per tile, comparable to the 1700-cycle MMA batch of an
    // N = 64 tile) the other one's MMAs keep the tensor pipe busy.
    const int r = warp == 1 ? 0 : 1;
    const bool dual = P.res_dual != 0;
    // Two issuers are only safe when they never share a pipeline stage (an mbarrier waiter must not run a whole
    // phase ahead of the barrier): the host enables `res_dual` iff stages % (2 * steps_per_tile) == 0.
    if ((dual || r == 0) && (!PAIR || crank == 0) && elect_one()) {
      constexpr uint32_t idesc = instr_desc<BN, PAIR ? 256 : 128>();
      constexpr int NK = KC / 16;
      const uint64_t desc_hi_lo = make_smem_desc<KC>(0);          // all fields except the start address
      const uint32_t a_base = smem_u32(smem_a) & 0x3FFFF, b_base = smem_u32(smem_b) & 0x3FFFF;
      const uint32_t desc_hi_fixed = (uint32_t)(desc_hi_lo >> 32) & ~0x3fffu;   // version + swizzle mode
      const uint32_t b_hi = (uint32_t)(desc_hi_lo >> 32);
      const uint32_t b_lo0 = (b_base >> 4) | 0x10000u;
      uint32_t stage = 0, ph = 0, bfph = 0;
      int cur_phase = -1;
      const int total = P.tiles_total, tn = P.tiles_n, tX = P.tiles_x, tY = P.tiles_y, tB = P.tiles_b;
      int32_t* const perr = P.error_flag;
      TileIter it;
      it.init(P, bid, gdim);
      int li = 0;                                                   // local tile index
      for (int tile = bid; tile < total; tile += gdim, it.next(tn, tX, tY, tB), ++li) {
        const int phase = it.phase;
        const int ns = s_nsteps[phase];
        const bool mine = dual ? ((li & 1) == r) : true;
        if (phase != cur_phase) {
          mbar_wait(bfull, bfph, perr);
          bfph ^= 1;
          cur_phase = phase;
        }
        if (mine) {
          const int as = li & (NACC - 1);
          const uint32_t apar = (uint32_t)((li / NACC) & 1);
          dbg_stamp(P, 1, li, 0);
          mbar_wait(&tempty[as], apar ^ 1, perr);
          tc_fence_after();
          dbg_stamp(P, 1, li, 1);
          const uint32_t d_tmem = tmem_base + as * BN;
          uint32_t acc = 0;
          for (int i = 0; i < ns; ++i) {
            const Step& sp = s_steps[phase * kMaxSteps + i];
            mbar_wait(&full[stage], ph, perr);
            tc_fence_after();
            dbg_stamp(P, 1, li, 2 + (i < 2 ? i : 2));
            // Descriptors are assembled from 32-bit halves with one add per tap: the hi words (stride, version,
            // swizzle mode) are loop invariants, the lo word is (smem address >> 4) | LBO.
            const uint32_t sa_lo = ((a_base + stage * a_stage) >> 4) | 0x10000u;
            const uint32_t a_hi = (sp.a_sbo16 & 0x3fffu) | desc_hi_fixed;
            const int ntap = sp.ntap;
            uint32_t tw[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) tw[t] = sp.tap[t];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              if (t < ntap) {
                const uint32_t a_lo = sa_lo + (tw[t] & 0xffffu);
                const uint32_t b_lo = b_lo0 + (tw[t] >> 16) * (uint32_t)(B_BYTES >> 4);
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                  if constexpr (PAIR)
                    umma2_bf16(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 2 * k), ((uint64_t)b_hi << 32) | (b_lo + 2 * k), idesc, acc);
                  else
                    umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 2 * k), ((uint64_t)b_hi << 32) | (b_lo + 2 * k), idesc, acc);
                  acc = 1;
                }
              }
            }
            if constexpr (PAIR) umma2_commit(&empty[stage]); else umma_commit(&empty[stage]);
            if (++stage == (uint32_t)stages) { stage = 0; ph ^= 1; }
          }
          if constexpr (PAIR) umma2_commit(&tfull[as]); else umma_commit(&tfull[as]);
          dbg_stamp(P, 1, li, 6);
        } else {
          // the other issuer's tile: only track which pipeline stages it consumes
          for (int i = 0; i < ns; ++i)
            if (++stage == (uint32_t)stages) { stage = 0; ph ^= 1; }
        }
        const int next = tile + gdim;
        if (next < total && next >= (phase + 1) * tiles_per_phase) {                      // both issuers (count 2)
          if constexpr (PAIR) umma2_commit(bempty); else umma_commit(bempty);
        }
      }
    }
  } else if (warp < 10) {
    epilogue_role<BN, 1, PAIR, NACC, CP, true>(P, tmem_base, tfull, tempty, epi_params, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Grouped streaming kernel for wide layers (Cout >= 128).  The plain streaming kernel above fills 16 KB of activations
// per (tap, chunk) K-block; measured on B200 the L2->SM fill (64..80 B/clk/SM for these tiles), not the tensor pipe,
// bounds it (switching the A fill off makes the same MMA stream 25..50 % faster).  Here ONE activation box with
// a row halo (TH+2 rows; optionally also a column halo) serves the 3 (or all 9) taps that read it: the tap's A
// descriptor simply starts `dy*row_pitch (+dx)` pixels further into the box.  Activations and weights travel in
// separate rings (a box is consumed by several weight tiles): per step 1 box load + ntap weight-tile loads.
//   PAIR: clusters of two CTAs, cta_group::2 MMAs (M = 256), each CTA stages half of every weight tile.
// ------------------------------------------------------------------------------------------------
template <int BN, int KC, int MT, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_grp_kernel(const __grid_constant__ TcParams P) {
  constexpr int B_BYTES = BN * KC * 2 / (PAIR ? 2 : 1);
  constexpr int MAX_A = 4, MAX_B = 8;
  constexpr int TMEM_COLS = 2 * MT * BN < 32 ? 32 : 2 * MT * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int na = P.res_stages, a_stage = P.res_a_stage_bytes, nb = P.grp_b_stages;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + na * a_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + nb * B_BYTES);
  uint64_t* afull = bars;
  uint64_t* aempty = bars + MAX_A;
  uint64_t* bfull = bars + 2 * MAX_A;
  uint64_t* bempty = bfull + MAX_B;
  uint64_t* tfull = bempty + MAX_B;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* epi_params = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
  const int crank = PAIR ? (int)cluster_ctarank() : 0;
  const int bid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int gdim = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  __shared__ Step s_steps[4 * kMaxSteps];
  __shared__ int s_nsteps[4];
  {
    const uint32_t* src = P.tables;
    uint32_t* dst = reinterpret_cast<uint32_t*>(s_steps);
    for (int i = threadIdx.x; i < (int)(sizeof(Step) * 4 * kMaxSteps / 4); i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x < 4) s_nsteps[threadIdx.x] = P.nsteps[threadIdx.x];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < na; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
    for (int i = 0; i < nb; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], PAIR ? 16 : 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      pdl_wait();
      uint32_t as = 0, aph = 0, bs = 0, bph = 0;
      const uint32_t afull_leader = PAIR ? mapa_u32(smem_u32(afull), 0) : 0;
      const uint32_t bfull_leader = PAIR ? mapa_u32(smem_u32(bfull), 0) : 0;
      const bool lead = !PAIR || crank == 0;
      const uint32_t mult = PAIR ? 2u : 1u;
      const int cout_pad = P.cout_pad;
      int32_t* const perr = P.error_flag;
      TileIter it;
      it.init(P, bid, gdim);
      for (int tile = bid; tile < P.tiles_total; tile += gdim, it.next(P)) {
        const int phase = it.phase, bt = it.bt;
        const int x0 = (PAIR ? 2 * it.xt + crank : it.xt) * P.TW, y0 = it.yt * P.TH;
        const int brow0 = P.wkb_phase0[phase] * cout_pad + it.nt * BN + (PAIR ? crank * (BN / 2) : 0);
        const int ns = s_nsteps[phase];
        for (int i = 0; i < ns; ++i) {
          const Step& sp = s_steps[phase * kMaxSteps + i];
          const int nch = sp.nch, ntap = sp.ntap;
          for (int ch = 0; ch < nch; ++ch) {
            mbar_wait(&aempty[as], aph ^ 1, perr);
            if (lead) mbar_expect_tx(&afull[as], mult * sp.bytes);
            void* da = smem_a + as * a_stage;
            if constexpr (PAIR) {
              if (sp.mode == 0)
                tma2_load_4d(da, &P.tmA[sp.src], afull_leader + 8 * as, sp.c0 + ch * KC, x0 + sp.ox, y0 + sp.oy, bt);
              else
                tma2_load_5d(da, &P.tmA[sp.src], afull_leader + 8 * as, sp.c0 + ch * KC, x0 + sp.ox, sp.py, y0 + sp.oy, bt);
            } else {
              if (sp.mode == 0)
                tma_load_4d(da, &P.tmA[sp.src], &afull[as], sp.c0 + ch * KC, x0 + sp.ox, y0 + sp.oy, bt);
              else
                tma_load_5d(da, &P.tmA[sp.src], &afull[as], sp.c0 + ch * KC, x0 + sp.ox, sp.py, y0 + sp.oy, bt);
            }
            if (++as == (uint32_t)na) { as = 0; aph ^= 1; }
            for (int t = 0; t < ntap; ++t) {
              mbar_wait(&bempty[bs], bph ^ 1, perr);
              if (lead) mbar_expect_tx(&bfull[bs], mult * (uint32_t)B_BYTES);
              const int row = brow0 + ((int)(sp.tap[t] >> 16) + ch) * cout_pad;
              if constexpr (PAIR) tma2_load_2d(smem_b + bs * B_BYTES, &P.tmB, bfull_leader + 8 * bs, 0, row);
              else tma_load_2d(smem_b + bs * B_BYTES, &P.tmB, &bfull[bs], 0, row);
              if (++bs == (uint32_t)nb) { bs = 0; bph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (PAIR: leader CTA only)
    if ((!PAIR || crank == 0) && elect_one()) {
      constexpr uint32_t idesc = instr_desc<BN, PAIR ? 256 : 128>();
      constexpr int NK = KC / 16;
      const uint64_t desc_hi_lo = make_smem_desc<KC>(0);
      const uint32_t a_base = smem_u32(smem_a) & 0x3FFFF, b_base = smem_u32(smem_b) & 0x3FFFF;
      const uint32_t desc_hi_fixed = (uint32_t)(desc_hi_lo >> 32) & ~0x3fffu;
      const uint32_t b_hi = (uint32_t)(desc_hi_lo >> 32);
      uint32_t as = 0, aph = 0, bs = 0, bph = 0, acs = 0, acph = 0;
      const int total = P.tiles_total, tn = P.tiles_n, tX = P.tiles_x, tY = P.tiles_y, tB = P.tiles_b;
      int32_t* const perr = P.error_flag;
      TileIter it;
      it.init(P, bid, gdim);
      for (int tile = bid; tile < total; tile += gdim, it.next(tn, tX, tY, tB)) {
        const int phase = it.phase;
        const int ns = s_nsteps[phase];
        mbar_wait(&tempty[acs], acph ^ 1, perr);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acs * (MT * BN);
        uint32_t acc = 0;
        for (int i = 0; i < ns; ++i) {
          const Step& sp = s_steps[phase * kMaxSteps + i];
          const int nch = sp.nch, ntap = sp.ntap;
          const uint32_t a_hi = (sp.a_sbo16 & 0x3fffu) | desc_hi_fixed;
          const uint32_t mt_off = sp.mt_off16;
          for (int ch = 0; ch < nch; ++ch) {
            mbar_wait(&afull[as], aph, perr);
            tc_fence_after();
            const uint32_t sa_lo = ((a_base + as * a_stage) >> 4) | 0x10000u;
            for (int t = 0; t < ntap; ++t) {
              const uint32_t a_lo = sa_lo + (sp.tap[t] & 0xffffu);
              const uint32_t b_lo = ((b_base + bs * B_BYTES) >> 4) | 0x10000u;
              mbar_wait(&bfull[bs], bph, perr);
              tc_fence_after();
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                for (int k = 0; k < NK; ++k) {
                  const uint64_t ad = ((uint64_t)a_hi << 32) | (a_lo + mt * mt_off + 2 * k);
                  const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo + 2 * k);
                  if constexpr (PAIR) umma2_bf16(d_tmem + mt * BN, ad, bd, idesc, acc | (uint32_t)k);
                  else umma_bf16(d_tmem + mt * BN, ad, bd, idesc, acc | (uint32_t)k);
                }
              }
              acc = 1;
              if constexpr (PAIR) umma2_commit(&bempty[bs]); else umma_commit(&bempty[bs]);
              if (++bs == (uint32_t)nb) { bs = 0; bph ^= 1; }
            }
            if constexpr (PAIR) umma2_commit(&aempty[as]); else umma_commit(&aempty[as]);
            if (++as == (uint32_t)na) { as = 0; aph ^= 1; }
          }
        }
        if constexpr (PAIR) umma2_commit(&tfull[acs]); else umma_commit(&tfull[acs]);
        if (++acs == 2) { acs = 0; acph ^= 1; }
      }
    }
  } else {
    epilogue_role<BN, MT, PAIR>(P, tmem_base, tfull, tempty, epi_params, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host plan
struct Plan {
  bool ok = false;
  int BN = 0, KC = 0;
  int n_phase = 1, os = 1;
  int gray_src = -1;                 // index of the fp32 1-channel side source, or -1
  int nkb_total = 0;                 // K-block row-groups in the packed weight matrix
  int cout_pad = 0;
  struct WTap { int src; int phase; int nchunks; int wkb0; std::vector<int> ktaps; };   // ktaps: reference taps summed
  std::vector<WTap> wtaps;
  Tap taps[4][kMaxTaps];
  int ntaps[4] = {0, 0, 0, 0};
  int kblocks[4] = {0, 0, 0, 0};
  int wkb_phase0[4] = {0, 0, 0, 0};
  // tile shape (phase grid): TW*TH*NB == 128
  int Hg = 0, Wg = 0, TW = 0, TH = 0, NB = 0;
  int MT = 1;                        // 128-pixel sub-tiles per CTA tile (streaming kernel)
  bool pair = false;                 // streaming kernel as CTA pairs (cta_group::2, M = 256, weight tile split in halves)
  bool grouped = false;              // grouped streaming kernel (activation boxes with halo rows shared by several taps)
  int grp_b_stages = 0, grp_smem_bytes = 0;
  bool grp_failed = false;           // the grouped plan did not fit: re-plan without it
  // resident-weight variant
  bool resident = false;
  Step steps[4][kMaxSteps];
  int nsteps[4] = {0, 0, 0, 0};
  int span[2] = {0, 0};              // extra box rows of each source's mode-0 loads
  int spanx[2] = {0, 0};             // extra box columns (halo mode)
  bool halo = false;
  int res_stages = 0, res_a_stage_bytes = 0, res_b_bytes = 0, res_smem_bytes = 0;
  bool res_dual = false;
};

int pow2_ceil_(int v) { int p = 1; while (p < v) p *= 2; return p; }
bool g_allow_resident = true;
bool g_allow_mt2 = true;
bool g_allow_dual = true;
bool g_allow_pair = true;
bool g_allow_grp = true;
int g_direct_default = 1;
bool g_allow_res_pair = false;     // measured slower than single-CTA tiles (0.316 vs 0.278 ms at 64->64 @256x256): kept as an experiment switch
bool g_allow_const = true;
bool g_grp_mixed = false;             // group layers that mix shareable (mode 0) and stride-2-sampled (mode 1) taps
int g_grp_halo = 1;                  // 1: grouped streaming kernel uses 8-pixel-wide tiles with a full (x and y) halo box
int g_pair_min_kb = 16;              // short-K tiles are epilogue-paced: pairing only adds cross-CTA handshakes (measured)
int g_halo_mode = 1;                 // 0 off, 1 halo tiles (descriptor base_offset 0: the hardware swizzle is a function of
                                     // the absolute smem address -- verified on B200), 2 = experiment: base_offset from addr bits (wrong)

int pick_kc(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : (c % 16 == 0 ? 16 : 0)); }

Plan build_plan_impl(const disco_conv_desc* d, bool allow_grp) {
  Plan p;
  if (d->dtype != DISCO_BF16) return p;
  if (d->n_src < 1 || d->n_src > 2) return p;
  int kc = 64;
  bool any_up2 = false;
  int n_mma_src = 0;
  for (int s = 0; s < d->n_src; ++s) {
    const disco_conv_src& src = d->src[s];
    if (src.is_f32) {
      if (src.C != 1 || d->kind != DISCO_CONV3 || d->stride != 1 || src.up2 || p.gray_src >= 0) return p;
      p.gray_src = s;
      continue;
    }
    const int k = pick_kc(src.C);
    if (k == 0) return p;
    kc = k < kc ? k : kc;
    any_up2 |= src.up2 != 0;
    n_mma_src++;
  }
  if (n_mma_src == 0) return p;
  if (p.gray_src >= 0 && (d->head != DISCO_HEAD_NONE || any_up2)) return p;
  if (d->head == DISCO_HEAD_NONE && d->Cout % 16 != 0) return p;
  if (d->kind == DISCO_DECONV4 && d->n_src != 1) return p;
  if (d->kind == DISCO_CONV3 && d->stride == 2 && any_up2) return p;
  p.KC = kc;
  const int cout16 = (d->Cout + 15) / 16 * 16;
  p.BN = cout16 % 256 == 0 ? 256 : (cout16 % 128 == 0 ? 128 : (cout16 % 64 == 0 ? 64 : (cout16 % 32 == 0 ? 32 : 16)));
  if (p.gray_src >= 0 && p.BN > 64) return p;
  p.cout_pad = (d->Cout + p.BN - 1) / p.BN * p.BN;
  const bool phased = any_up2 || d->kind == DISCO_DECONV4;
  p.n_phase = phased ? 4 : 1;
  p.os = phased ? 2 : 1;
  int wkb = 0;
  for (int ph = 0; ph < p.n_phase; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    int nt = 0;
    p.wkb_phase0[ph] = wkb;
    for (int s = 0; s < d->n_src; ++s) {
      if (s == p.gray_src) continue;
      const disco_conv_src& src = d->src[s];
      const int nch = src.C / kc;
      auto add = [&](int mode, int oy, int ox, int ppy, int ppx, int cbase, std::vector<int> ktaps) {
        if (nt >= kMaxTaps) { p.ok = false; nt = kMaxTaps + 1; return; }
        Tap& t = p.taps[ph][nt++];
        t.src = (int8_t)s; t.mode = (int8_t)mode; t.oy = (int8_t)oy; t.ox = (int8_t)ox; t.py = (int8_t)ppy; t.px = (int8_t)ppx;
        t.nchunks = (int16_t)nch; t.wkb0 = wkb; t.c_base = cbase;
        p.wtaps.push_back({s, ph, nch, wkb, std::move(ktaps)});
        wkb += nch;
        p.kblocks[ph] += nch;
      };
      if (d->kind == DISCO_DECONV4) {
        // out (2i+py): py=0 <- ky=1 (row i), ky=3 (row i-1);  py=1 <- ky=0 (row i+1), ky=2 (row i)
        const int kys[2][2] = {{1, 3}, {0, 2}}, offs[2][2] = {{0, -1}, {1, 0}};
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b) add(0, offs[py][a], offs[px][b], 0, 0, 0, {kys[py][a] * 4 + kys[px][b]});
      } else if (src.up2) {
        // out (2i+py) <- up-sampled rows 2i+py+dy-1: py=0 -> {i-1: dy0}, {i: dy1,dy2}; py=1 -> {i: dy0,dy1}, {i+1: dy2}
        const std::vector<int> grp[2][2] = {{{0}, {1, 2}}, {{0, 1}, {2}}};
        const int offs[2][2] = {{-1, 0}, {0, 1}};
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b) {
            std::vector<int> kt;
            for (int dy : grp[py][a]) for (int dx : grp[px][b]) kt.push_back(dy * 3 + dx);
            add(0, offs[py][a], offs[px][b], 0, 0, 0, kt);
          }
      } else if (phased || d->stride == 2) {
        // direct source sampled on every second pixel: v = 2i + o, o = (py|0) + dy - 1
        for (int dy = 0; dy < 3; ++dy)
          for (int dx = 0; dx < 3; ++dx) {
            const int o_y = (phased ? py : 0) + dy - 1, o_x = (phased ? px : 0) + dx - 1;
            add(1, o_y >> 1, o_x >> 1, o_y & 1, o_x & 1, (o_x & 1) * src.C, {dy * 3 + dx});
          }
      } else {
        for (int dy = 0; dy < 3; ++dy)
          for (int dx = 0; dx < 3; ++dx) add(0, dy - 1, dx - 1, 0, 0, 0, {dy * 3 + dx});
      }
    }
    if (nt > kMaxTaps) return p;
    p.ntaps[ph] = nt;
  }
  p.nkb_total = wkb;
  p.ok = true;

  // ---- tile shape on the phase grid
  p.Hg = d->Ho / p.os; p.Wg = d->Wo / p.os;
  int TW = (p.Wg % 16 == 0) ? 16 : (p.Wg % 8 == 0 ? 8 : (p.Wg % 4 == 0 ? 4 : 16));
  if (TW > pow2_ceil_(p.Wg)) TW = pow2_ceil_(p.Wg);
  int TH = 128 / TW;
  if (TH > pow2_ceil_(p.Hg)) TH = pow2_ceil_(p.Hg);
  p.TW = TW; p.TH = TH; p.NB = 128 / (TW * TH);

  // ---- resident-weight variant: one n-tile, whole-image tiles, weights of a phase fit in shared memory
  int max_kb = 0;
  for (int ph = 0; ph < p.n_phase; ++ph) max_kb = p.kblocks[ph] > max_kb ? p.kblocks[ph] : max_kb;
  const int b_bytes = max_kb * p.BN * kc * 2;
  const bool want_res = g_allow_resident && p.cout_pad == p.BN && p.BN <= 64 && p.NB == 1 && TW >= 8 && b_bytes <= 112 * 1024;
  bool halo = false;
  if (!want_res) {
    // streaming kernel: 256-pixel tiles for narrow N when a tile stays inside one image
    const int th_base = TH;
    if (g_allow_mt2 && p.BN <= 128 && p.NB == 1 && TW * TH == 128 && p.Hg >= 2 * TH) { p.MT = 2; p.TH = 2 * TH; }
    // CTA pairs: two horizontally adjacent pixel tiles share one weight tile (wide N only: that is where the weight
    // fill and the B-operand reads matter); needs an even number of tile columns
    const int tiles_x = (p.Wg + p.TW - 1) / p.TW;
    int min_kb = 1 << 30;
    for (int ph = 0; ph < p.n_phase; ++ph) min_kb = p.kblocks[ph] < min_kb ? p.kblocks[ph] : min_kb;
    const bool wide = (p.BN == 256 && p.MT == 1) || (p.BN == 128 && p.MT == 2);
    p.pair = g_allow_pair && kc == 64 && tiles_x % 2 == 0 && min_kb >= g_pair_min_kb && wide;
    // grouped variant: wide N, whole tiles inside one image, at least one tap family that can share a box
    bool any_mode0 = false;
    for (int ph = 0; ph < p.n_phase; ++ph)
      for (int t = 0; t < p.ntaps[ph]; ++t) any_mode0 |= p.taps[ph][t].mode == 0;
    const bool want_grp = g_allow_grp && allow_grp && kc == 64 && wide && p.NB == 1 && TW == 16 && th_base == 8 && p.Wg % 16 == 0 &&
                          any_mode0 && min_kb >= g_pair_min_kb;
    if (!want_grp) return p;
    bool any_mode1 = false;
    for (int ph = 0; ph < p.n_phase; ++ph)
      for (int t = 0; t < p.ntaps[ph]; ++t) any_mode1 |= p.taps[ph][t].mode == 1;
    if (any_mode1 && !g_grp_mixed) return p;
    // 8-pixel-wide halo tiles make the stride-2-sampled boxes of skip sources costlier (measured): row halo only there
    halo = g_grp_halo != 0 && p.Hg >= 16 * p.MT && !any_mode1;
    if (halo) { TW = 8; TH = 16 * p.MT; } else { TW = 16; TH = 8 * p.MT; }
    p.TW = TW; p.TH = TH; p.NB = 1;
    p.pair = g_allow_pair && ((p.Wg / TW) % 2 == 0);
    p.grouped = true;
  } else {
  // halo mode: 8-pixel-wide tiles so that an 8-row MMA group is 8 consecutive pixels of one image row; ONE box
  // with the full (TW+2) x (TH+2) halo then serves all taps of a source (A descriptors start at arbitrary pixel
  // offsets inside the box; the swizzle phase is carried by the address / descriptor base offset)
  halo = g_halo_mode != 0 && p.Wg % 8 == 0 && pow2_ceil_(p.Hg) >= 16;
  if (halo) {
    // steps per tile with column grouping vs. with halo boxes (worst phase); halo tiles are 8 pixels wide, which makes
    // the stride-2-sampled taps of skip sources (one step each in both schemes) slightly costlier, so require a real gain
    int worst_grouped = 0, worst_halo = 0;
    for (int ph = 0; ph < p.n_phase; ++ph) {
      int grouped = 0, with_halo = 0;
      for (int s = 0; s < d->n_src; ++s) {
        bool seen_ox[8] = {false};
        int n0 = 0, n1 = 0, nch = 0;
        for (int t = 0; t < p.ntaps[ph]; ++t) {
          const Tap& tp = p.taps[ph][t];
          if (tp.src != s) continue;
          nch = tp.nchunks;
          if (tp.mode == 0) { if (!seen_ox[tp.ox + 4]) { seen_ox[tp.ox + 4] = true; n0++; } }
          else n1++;
        }
        grouped += (n0 + n1) * nch;
        with_halo += ((n0 ? 1 : 0) + n1) * nch;
      }
      worst_grouped = grouped > worst_grouped ? grouped : worst_grouped;
      worst_halo = with_halo > worst_halo ? with_halo : worst_halo;
    }
    halo = worst_halo * 2 <= worst_grouped;
  }
  if (halo) { TW = 8; TH = 16; p.TW = TW; p.TH = TH; p.NB = 1; }
  }
  p.halo = halo;
  // per-source extents of the mode-0 tap offsets (box = tile + span)
  int omin_y[4][2], omin_x[4][2];
  for (int s = 0; s < 2; ++s) { p.span[s] = 0; p.spanx[s] = 0; }
  for (int ph = 0; ph < p.n_phase; ++ph)
    for (int s = 0; s < d->n_src; ++s) {
      int ylo = 127, yhi = -127, xlo = 127, xhi = -127;
      for (int t = 0; t < p.ntaps[ph]; ++t) {
        const Tap& tp = p.taps[ph][t];
        if (tp.src != s || tp.mode != 0) continue;
        ylo = tp.oy < ylo ? tp.oy : ylo; yhi = tp.oy > yhi ? tp.oy : yhi;
        xlo = tp.ox < xlo ? tp.ox : xlo; xhi = tp.ox > xhi ? tp.ox : xhi;
      }
      omin_y[ph][s] = ylo; omin_x[ph][s] = xlo;
      if (yhi >= ylo) {
        p.span[s] = (yhi - ylo) > p.span[s] ? (yhi - ylo) : p.span[s];
        if (halo) p.spanx[s] = (xhi - xlo) > p.spanx[s] ? (xhi - xlo) : p.spanx[s];
      }
    }
  int a_stage = TH * TW * kc * 2;
  for (int s = 0; s < d->n_src; ++s) {
    const int b = (TH + p.span[s]) * (TW + p.spanx[s]) * kc * 2;
    a_stage = b > a_stage ? b : a_stage;
  }
  a_stage = (a_stage + 1023) / 1024 * 1024;
  for (int ph = 0; ph < p.n_phase; ++ph) {
    int ns = 0;
    bool used[kMaxTaps] = {false};
    for (int t = 0; t < p.ntaps[ph]; ++t) {
      if (used[t]) continue;
      const Tap& tp = p.taps[ph][t];
      std::vector<int> members;
      if (tp.mode == 0) {
        for (int u = t; u < p.ntaps[ph]; ++u) {
          const Tap& o = p.taps[ph][u];
          if (!used[u] && o.src == tp.src && o.mode == 0 && (halo || o.ox == tp.ox)) { members.push_back(u); used[u] = true; }
        }
      } else {
        members.push_back(t);
        used[t] = true;
      }
      if ((int)members.size() > 9) { p.grp_failed = p.grouped; return p; }
      const int oy_min = tp.mode == 0 ? omin_y[ph][tp.src] : tp.oy;
      const int ox_min = (tp.mode == 0 && halo) ? omin_x[ph][tp.src] : tp.ox;
      const int row_px = (tp.mode == 0) ? TW + p.spanx[tp.src] : TW;        // pixels per box row
      const int n_expand = p.grouped ? 1 : tp.nchunks;     // the grouped streaming kernel loops over chunks itself
      for (int ch = 0; ch < n_expand; ++ch) {
        if (ns >= kMaxSteps) { p.grp_failed = p.grouped; return p; }
        Step& st = p.steps[ph][ns++];
        memset(&st, 0, sizeof(st));
        st.src = tp.src; st.mode = tp.mode; st.ox = (int8_t)ox_min; st.oy = (int8_t)oy_min; st.py = tp.py; st.px = tp.px;
        st.ntap = (int8_t)members.size();
        st.nch = (int8_t)(p.grouped ? tp.nchunks : 1);
        st.mt_off16 = (uint32_t)((TH / 2) * row_px * kc * 2) >> 4;
        st.c0 = tp.c_base + ch * kc;
        st.bytes = (uint32_t)((tp.mode == 0 ? (TH + p.span[tp.src]) * row_px : TH * TW) * kc * 2);
        st.a_sbo16 = (uint32_t)(((halo && tp.mode == 0) ? row_px : 8) * kc * 2) >> 4;
        for (size_t i = 0; i < members.size(); ++i) {
          const Tap& o = p.taps[ph][members[i]];
          const uint32_t off = (uint32_t)(((o.oy - oy_min) * row_px + (o.ox - ox_min)) * kc * 2) >> 4;
          const uint32_t wblk = (uint32_t)(o.wkb0 - p.wkb_phase0[ph] + ch);
          st.tap[i] = off | (wblk << 16);
        }
      }
    }
    p.nsteps[ph] = ns;
  }
  if (p.grouped) {
    const int epi_cols_g = p.MT == 1 ? p.BN / 2 : p.BN;
    const int epi_bytes_g = 8 * (3 * epi_cols_g) * 4 + 8 * 2048;
    const int tile_b = p.BN * kc * 2 / (p.pair ? 2 : 1);
    const int budget = 227 * 1024 - 1024 - 256 - epi_bytes_g - 4096 /* static tables */;
    int na = 3, nbs = (budget - na * a_stage) / tile_b;
    if (nbs > 8) { nbs = 8; if ((budget - nbs * tile_b) / a_stage >= 4) na = 4; }
    if (nbs < 3) { p.grp_failed = true; return p; }
    p.res_stages = na; p.res_a_stage_bytes = a_stage; p.grp_b_stages = nbs;
    p.grp_smem_bytes = na * a_stage + nbs * tile_b + 256 + epi_bytes_g + 1024;
    return p;
  }
  const int epi_cols = p.BN;       // alternating epilogue: every warp owns all columns of its tiles
  const int epi_bytes = 8 * (3 * epi_cols + (p.gray_src >= 0 ? 9 * epi_cols : 0)) * 4 + 8 * 2048;
  // CTA pairs (cta_group::2) for the 64-channel layers: each CTA keeps half of every weight block
  p.pair = g_allow_pair && g_allow_res_pair && p.BN == 64 && ((p.Wg + TW - 1) / TW) % 2 == 0;
  const int b_res = p.pair ? b_bytes / 2 : b_bytes;
  const int fixed = b_res + 256 + epi_bytes + 1024;
  int stages = (225 * 1024 - fixed) / a_stage;
  if (stages > 8) stages = 8;
  if (stages < 2) return p;
  // dual MMA issuers: all phases must have the same number of steps per tile and the stage ring must split evenly
  {
    int ns0 = p.nsteps[0];
    bool same = true;
    for (int ph = 1; ph < p.n_phase; ++ph) same = same && p.nsteps[ph] == ns0;
    p.res_dual = false;
    if (g_allow_dual && same && ns0 >= 1 && stages >= 2 * ns0) {
      stages = stages / (2 * ns0) * (2 * ns0);
      p.res_dual = true;
    }
  }
  p.resident = true;
  p.res_stages = stages; p.res_a_stage_bytes = a_stage; p.res_b_bytes = b_res;
  p.res_smem_bytes = stages * a_stage + fixed;
  return p;
}

Plan build_plan(const disco_conv_desc* d) {
  Plan p = build_plan_impl(d, true);
  if (p.grp_failed) p = build_plan_impl(d, false);
  return p;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode(disco_handle* h, CUtensorMap* tm, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
           const cuuint32_t* box, int kc) {
  if (!h->tmap_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    DISCO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) {
      disco_set_error("cuTensorMapEncodeTiled not available from the driver");
      return DISCO_ERR_CUDA;
    }
    h->tmap_encode = fn;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = reinterpret_cast<EncodeTiledFn>(h->tmap_encode)(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims,
                                                              strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    disco_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, box0 %u)", (int)r, rank, box[0]);
    return DISCO_ERR_CUDA;
  }
  return DISCO_OK;
}

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

struct Cached {
  TcParams params;
  Plan plan;
  int grid;
  int device = -1;
  void* tables_dev = nullptr;
};
constexpr size_t kMaxCachedPlans = 2048;   // ~70 descriptors per (batch, H, W) workspace of one forward
std::mutex g_mu;
std::map<std::string, Cached> g_cache;
long long* g_dbg = nullptr;

bool g_allow_pdl = true;
// Every tensor-core launch: optional 2-CTA clusters + programmatic stream serialization (see pdl_wait()).
template <typename Kern>
int launch_tc(Kern kern, const TcParams& P, int grid, int threads, size_t smem, int cluster, cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_allow_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)na;
  DISCO_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  return DISCO_OK;
}

template <int BN, int KC, int MT>
int launch_cfg(disco_handle* h, const TcParams& P, int grid, cudaStream_t st) {
  using C = Cfg<BN, KC, MT>;
  if (int rc = disco_ensure_smem(h, (const void*)conv_tc_kernel<BN, KC, MT, false>, C::SMEM_BYTES)) return rc;
  return launch_tc(conv_tc_kernel<BN, KC, MT, false>, P, grid, kThreads, C::SMEM_BYTES, 1, st);
}

// CTA-pair variant: clusters of two CTAs (one TPC), `grid` is even
template <int BN, int KC, int MT>
int launch_pair_cfg(disco_handle* h, const TcParams& P, int grid, cudaStream_t st) {
  using C = Cfg<BN, KC, MT, true>;
  if (int rc = disco_ensure_smem(h, (const void*)conv_tc_kernel<BN, KC, MT, true>, C::SMEM_BYTES)) return rc;
  return launch_tc(conv_tc_kernel<BN, KC, MT, true>, P, grid, kThreads, C::SMEM_BYTES, 2, st);
}

template <int BN, int KC, bool CP, bool PAIR>
int launch_res_cfg2(disco_handle* h, const TcParams& P, int grid, int smem_bytes, cudaStream_t st) {
  if (int rc = disco_ensure_smem(h, (const void*)conv_tc_res_kernel<BN, KC, CP, PAIR>, smem_bytes)) return rc;
  return launch_tc(conv_tc_res_kernel<BN, KC, CP, PAIR>, P, grid, kThreadsRes, (size_t)smem_bytes, PAIR ? 2 : 1, st);
}
template <int BN, int KC>
int launch_res_cfg(disco_handle* h, const TcParams& P, int grid, int smem_bytes, bool pair, cudaStream_t st) {
  if constexpr (BN == 64) {
    if (pair)
      return P.const_params ? launch_res_cfg2<BN, KC, true, true>(h, P, grid, smem_bytes, st)
                            : launch_res_cfg2<BN, KC, false, true>(h, P, grid, smem_bytes, st);
  }
  return P.const_params ? launch_res_cfg2<BN, KC, true, false>(h, P, grid, smem_bytes, st)
                        : launch_res_cfg2<BN, KC, false, false>(h, P, grid, smem_bytes, st);
}

template <int BN, int KC, int MT, bool PAIR>
int launch_grp_cfg(disco_handle* h, const TcParams& P, int grid, int smem_bytes, cudaStream_t st) {
  if (int rc = disco_ensure_smem(h, (const void*)conv_tc_grp_kernel<BN, KC, MT, PAIR>, smem_bytes)) return rc;
  return launch_tc(conv_tc_grp_kernel<BN, KC, MT, PAIR>, P, grid, kThreads, (size_t)smem_bytes, PAIR ? 2 : 1, st);
}

int launch_grp(disco_handle* h, const Plan& pl, const TcParams& P, int grid, cudaStream_t st) {
  if (pl.BN == 256 && pl.MT == 1)
    return pl.pair ? launch_grp_cfg<256, 64, 1, true>(h, P, grid, pl.grp_smem_bytes, st)
                   : launch_grp_cfg<256, 64, 1, false>(h, P, grid, pl.grp_smem_bytes, st);
  if (pl.BN == 128 && pl.MT == 2)
    return pl.pair ? launch_grp_cfg<128, 64, 2, true>(h, P, grid, pl.grp_smem_bytes, st)
                   : launch_grp_cfg<128, 64, 2, false>(h, P, grid, pl.grp_smem_bytes, st);
  disco_set_error("conv_tc: unsupported grouped configuration BN %d MT %d", pl.BN, pl.MT);
  return DISCO_ERR_INVALID;
}

template <int KC>
int launch_res_bn(disco_handle* h, int BN, const TcParams& P, int grid, int smem_bytes, bool pair, cudaStream_t st) {
  switch (BN) {
    case 16: return launch_res_cfg<16, KC>(h, P, grid, smem_bytes, pair, st);
    case 32: return launch_res_cfg<32, KC>(h, P, grid, smem_bytes, pair, st);
    case 64: return launch_res_cfg<64, KC>(h, P, grid, smem_bytes, pair, st);
  }
  disco_set_error("conv_tc: unsupported resident BN %d", BN);
  return DISCO_ERR_INVALID;
}

template <int KC>
int launch_bn(disco_handle* h, int BN, int MT, const TcParams& P, int grid, cudaStream_t st) {
  if (MT == 2) {
    switch (BN) {
      case 16: return launch_cfg<16, KC, 2>(h, P, grid, st);
      case 32: return launch_cfg<32, KC, 2>(h, P, grid, st);
      case 64: return launch_cfg<64, KC, 2>(h, P, grid, st);
      case 128: return launch_cfg<128, KC, 2>(h, P, grid, st);
    }
  } else {
    switch (BN) {
      case 16: return launch_cfg<16, KC, 1>(h, P, grid, st);
      case 32: return launch_cfg<32, KC, 1>(h, P, grid, st);
      case 64: return launch_cfg<64, KC, 1>(h, P, grid, st);
      case 128: return launch_cfg<128, KC, 1>(h, P, grid, st);
      case 256: return launch_cfg<256, KC, 1>(h, P, grid, st);
    }
  }
  disco_set_error("conv_tc: unsupported BN %d / MT %d", BN, MT);
  return DISCO_ERR_INVALID;
}

}  // namespace

bool conv_tc_supported(const disco_conv_desc* d) { return conv_narrow_match(d) || conv_ts_match(d) || build_plan(d).ok; }

void conv_tc_cache_clear(disco_handle* h) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto it = g_cache.begin(); it != g_cache.end();) {
    if (it->second.device == h->device) {
      cudaFree(it->second.tables_dev);         // synchronises the device
      it = g_cache.erase(it);
    } else {
      ++it;
    }
  }
}

extern "C" int disco_conv_tc_cache_clear(disco_handle* h) {
  DISCO_CHECK_ARG(h != nullptr, "conv_tc_cache_clear: null handle");
  DiscoDeviceGuard guard(h);
  conv_tc_cache_clear(h);
  return DISCO_OK;
}

extern "C" int disco_conv_tc_supported(disco_handle* h, const disco_conv_desc* d) {
  return h && d && h->use_tc && (conv_narrow_match(d) || conv_ts_match(d) || build_plan(d).ok) ? 1 : 0;
}

extern "C" int disco_set_tensor_core(disco_handle* h, int enable) {
  DISCO_CHECK_ARG(h != nullptr, "set_tensor_core: null handle");
  h->use_tc = enable != 0;
  return DISCO_OK;
}

extern "C" int64_t disco_conv_tc_weight_elems(const disco_conv_desc* d) {
  if (!d) return 0;
  if (conv_narrow_match(d)) return conv_narrow_weight_elems(d);
  if (conv_ts_match(d)) return conv_ts_weight_elems(d);
  Plan p = build_plan(d);
  return p.ok ? (int64_t)p.nkb_total * p.cout_pad * p.KC : 0;
}

// fp32 blocks [tap][cin_s][cout] per source (host, same layout as the CUDA-core path) -> bf16 [K-block][cout_pad][KC]
extern "C" int disco_conv_tc_pack_weights(const disco_conv_desc* d, const float* w32, uint16_t* out) {
  DISCO_CHECK_ARG(d && w32 && out, "tc_pack: null pointer");
  if (conv_narrow_match(d)) return conv_narrow_pack(d, w32, out);
  if (conv_ts_match(d)) return conv_ts_pack(d, w32, out);
  Plan p = build_plan(d);
  DISCO_CHECK_ARG(p.ok, "tc_pack: descriptor not supported by the tensor-core kernel");
  const size_t total = (size_t)p.nkb_total * p.cout_pad * p.KC;
  memset(out, 0, total * sizeof(uint16_t));
  // blocked transpose: w32 is [tap][ci][co] (co contiguous), the packing is [k-block][co][c] (c contiguous); 64 x 64 tiles
  // keep both sides in L1 (a plain co-outer / c-inner loop read with a stride of Cout floats: 34 ms for a 512 x 512 layer,
  // 0.4 s per engine; the CLI builds one engine per run)
  constexpr int kTile = 64;
  std::vector<float> tile((size_t)p.KC * kTile);
  for (const Plan::WTap& wt : p.wtaps) {
    const disco_conv_src& src = d->src[wt.src];
    const float* wb = w32 + src.w_off;
    for (int ch = 0; ch < wt.nchunks; ++ch)
      for (int co0 = 0; co0 < d->Cout; co0 += kTile) {
        const int nco = std::min(kTile, d->Cout - co0);
        std::fill(tile.begin(), tile.end(), 0.f);
        for (int kt : wt.ktaps)                              // same summation order per element as before: taps outermost
          for (int c = 0; c < p.KC; ++c) {
            const float* row = wb + ((size_t)kt * src.C + ch * p.KC + c) * d->Cout + co0;
            float* t = tile.data() + (size_t)c * kTile;
            for (int j = 0; j < nco; ++j) t[j] += row[j];
          }
        for (int j = 0; j < nco; ++j) {
          uint16_t* o = out + ((size_t)(wt.wkb0 + ch) * p.cout_pad + co0 + j) * p.KC;
          for (int c = 0; c < p.KC; ++c) {                    // round-to-nearest-even bf16 (== __float2bfloat16_rn; NaN kept quiet)
            uint32_t u;
            memcpy(&u, &tile[(size_t)c * kTile + j], 4);
            o[c] = (u & 0x7fffffffu) > 0x7f800000u ? (uint16_t)0x7fff : (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
          }
        }
      }
  }
  return DISCO_OK;
}

// debug: copies the timeline of the last DISCO_TC_DEBUG launch to `out` (4*48*8 int64); returns element count
extern "C" int disco_debug_timeline(long long* out) {
  if (!g_dbg) return 0;
  cudaDeviceSynchronize();
  cudaMemcpy(out, g_dbg, sizeof(long long) * 4 * kDbgTiles * kDbgSlots, cudaMemcpyDeviceToHost);
  return 4 * kDbgTiles * kDbgSlots;
}

int conv_tc_launch(disco_handle* h, const disco_conv_desc* d, cudaStream_t st) {
  if (conv_narrow_match(d)) return conv_narrow_launch(h, d, st);
  if (conv_ts_match(d)) return conv_ts_launch(h, d, st);
  static bool env_read = false;
  if (!env_read) {
    const char* e = getenv("DISCO_TC_RESIDENT");
    if (e && e[0] == '0') g_allow_resident = false;
    const char* du = getenv("DISCO_TC_DUAL");
    if (du && du[0] == '0') g_allow_dual = false;
    const char* m2 = getenv("DISCO_TC_MT2");
    if (m2 && m2[0] == '0') g_allow_mt2 = false;
    const char* pr = getenv("DISCO_TC_PAIR");
    if (pr && pr[0] == '0') g_allow_pair = false;
    if (getenv("DISCO_TC_PAIR_MIN_KB")) g_pair_min_kb = atoi(getenv("DISCO_TC_PAIR_MIN_KB"));
    const char* rp = getenv("DISCO_TC_RES_PAIR");
    if (rp) g_allow_res_pair = rp[0] != '0';
    const char* pd = getenv("DISCO_TC_PDL");
    if (pd && pd[0] == '0') g_allow_pdl = false;
    const char* cp = getenv("DISCO_TC_CONST");
    if (cp && cp[0] == '0') g_allow_const = false;
    const char* gr = getenv("DISCO_TC_GRP");
    if (gr && gr[0] == '0') g_allow_grp = false;
    if (getenv("DISCO_TC_GRP_HALO")) g_grp_halo = atoi(getenv("DISCO_TC_GRP_HALO"));
    if (getenv("DISCO_TC_GRP_MIXED")) g_grp_mixed = atoi(getenv("DISCO_TC_GRP_MIXED")) != 0;
    const char* hm = getenv("DISCO_TC_HALO");
    if (hm) g_halo_mode = atoi(hm);
    env_read = true;
  }
  std::string key(reinterpret_cast<const char*>(d), sizeof(*d));
  key.append(reinterpret_cast<const char*>(&h->device), sizeof(int));
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_cache.size() > kMaxCachedPlans) {
    // plans of other devices are freed on their own device
    int cur = -1;
    cudaGetDevice(&cur);
    for (auto& kv : g_cache) {
      if (kv.second.device != cur) cudaSetDevice(kv.second.device), cur = kv.second.device;
      cudaFree(kv.second.tables_dev);          // cudaFree synchronises: no launch still reads the tables
    }
    if (cur != h->device) cudaSetDevice(h->device);
    g_cache.clear();
  }
  auto it = g_cache.find(key);
  if (it == g_cache.end()) {
    Cached c;
    c.device = h->device;
    c.plan = build_plan(d);
    DISCO_CHECK_ARG(c.plan.ok, "conv_tc: descriptor not supported");
    const Plan& pl = c.plan;
    TcParams& P = c.params;
    memset(&P, 0, sizeof(P));
    P.error_flag = h->error_flag;
    if (getenv("DISCO_TC_MODE")) P.dbg_mode = atoi(getenv("DISCO_TC_MODE"));
    {
      // bit 0: resident kernel, bit 1: streaming kernels
      const int dm = getenv("DISCO_TC_DIRECT") ? atoi(getenv("DISCO_TC_DIRECT")) : g_direct_default;
      P.direct_store = (pl.resident ? (dm & 1) : (dm & 2)) ? 1 : 0;
    }
    if (getenv("DISCO_TC_DEBUG")) {
      if (!g_dbg) {
        DISCO_CUDA(cudaMalloc(&g_dbg, sizeof(long long) * 4 * kDbgTiles * kDbgSlots));
      }
      DISCO_CUDA(cudaMemset(g_dbg, 0, sizeof(long long) * 4 * kDbgTiles * kDbgSlots));
      P.dbg = g_dbg;
    }
    P.n_phase = pl.n_phase; P.os = pl.os;
    P.B = d->batch; P.Hg = pl.Hg; P.Wg = pl.Wg;
    const int TW = pl.TW, TH = pl.TH, NB = pl.NB;
    {
      std::vector<uint8_t> host(sizeof(pl.steps) + sizeof(pl.taps));
      memcpy(host.data(), pl.steps, sizeof(pl.steps));
      memcpy(host.data() + sizeof(pl.steps), pl.taps, sizeof(pl.taps));
      void* dev = nullptr;
      DISCO_CUDA(cudaMalloc(&dev, host.size()));
      DISCO_CUDA(cudaMemcpy(dev, host.data(), host.size(), cudaMemcpyHostToDevice));
      P.tables = reinterpret_cast<const uint32_t*>(dev);
      c.tables_dev = dev;
    }
    memcpy(P.nsteps, pl.nsteps, sizeof(P.nsteps));
    memcpy(P.wkb_phase0, pl.wkb_phase0, sizeof(P.wkb_phase0));
    P.res_stages = pl.res_stages; P.res_a_stage_bytes = pl.res_a_stage_bytes; P.res_b_bytes = pl.res_b_bytes;
    P.res_dual = pl.res_dual ? 1 : 0;
    P.grp_b_stages = pl.grp_b_stages;
    P.halo_bo_mask = (pl.halo && g_halo_mode == 2) ? (pl.KC == 64 ? 7 : (pl.KC == 32 ? 3 : 1)) : 0;
    P.TW = TW; P.TH = TH; P.NB = NB; P.tw_log2 = ilog2(TW); P.th_log2 = ilog2(TH);
    P.tiles_x = (P.Wg + TW - 1) / TW; P.tiles_y = (P.Hg + TH - 1) / TH; P.tiles_b = (P.B + NB - 1) / NB;
    if (pl.pair) P.tiles_x /= 2;      // the kernel iterates over PAIRS of tile columns
    P.tiles_n = pl.cout_pad / pl.BN;
    P.tiles_total = P.tiles_x * P.tiles_y * P.tiles_b * P.tiles_n * pl.n_phase;
    P.cout_pad = pl.cout_pad;
    memcpy(P.ntaps, pl.ntaps, sizeof(P.ntaps));
    memcpy(P.kblocks, pl.kblocks, sizeof(P.kblocks));
    P.Ho = d->Ho; P.Wo = d->Wo; P.Cout = d->Cout; P.act = d->act; P.head = d->head; P.slope = d->slope;
    P.bias = d->bias; P.post_scale = d->post_scale; P.post_shift = d->post_shift;
    P.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
    P.out = d->out;
    if (pl.gray_src >= 0) {
      P.gray = reinterpret_cast<const float*>(d->src[pl.gray_src].ptr);
      P.gray_w = reinterpret_cast<const float*>(d->gray_weights);
      DISCO_CHECK_ARG(P.gray_w != nullptr, "conv_tc: gray source needs desc.gray_weights (fp32 [9][Cout])");
    }
    // tensor maps: modes used per source
    for (int s = 0; s < d->n_src; ++s) {
      if (s == pl.gray_src) continue;
      const disco_conv_src& src = d->src[s];
      int mode = -1;
      for (int ph = 0; ph < pl.n_phase; ++ph)
        for (int t = 0; t < pl.ntaps[ph]; ++t)
          if (pl.taps[ph][t].src == s) mode = pl.taps[ph][t].mode;
      const cuuint64_t Cc = src.C, Wd = src.W, Hd = src.H, Bd = d->batch;
      int rc;
      if (mode == 0) {
        cuuint64_t dims[4] = {Cc, Wd, Hd, Bd};
        cuuint64_t str[3] = {Cc * 2, Wd * Cc * 2, Hd * Wd * Cc * 2};
        const bool boxed = pl.resident || pl.grouped;
        cuuint32_t box[4] = {(cuuint32_t)pl.KC, (cuuint32_t)(boxed ? TW + pl.spanx[s] : TW),
                             (cuuint32_t)(boxed ? TH + pl.span[s] : TH), (cuuint32_t)NB};
        rc = encode(h, &P.tmA[s], const_cast<void*>(src.ptr), 4, dims, str, box, pl.KC);
      } else {
        DISCO_CHECK_ARG(src.H % 2 == 0 && src.W % 2 == 0, "conv_tc: stride-2 source must have even H, W");
        cuuint64_t dims[5] = {2 * Cc, Wd / 2, 2, Hd / 2, Bd};
        cuuint64_t str[4] = {2 * Cc * 2, Wd * Cc * 2, 2 * Wd * Cc * 2, Hd * Wd * Cc * 2};
        cuuint32_t box[5] = {(cuuint32_t)pl.KC, (cuuint32_t)TW, 1, (cuuint32_t)TH, (cuuint32_t)NB};
        rc = encode(h, &P.tmA[s], const_cast<void*>(src.ptr), 5, dims, str, box, pl.KC);
      }
      if (rc != DISCO_OK) return rc;
    }
    {
      cuuint64_t dims[2] = {(cuuint64_t)pl.KC, (cuuint64_t)pl.nkb_total * pl.cout_pad};
      cuuint64_t str[1] = {(cuuint64_t)pl.KC * 2};
      cuuint32_t box[2] = {(cuuint32_t)pl.KC, (cuuint32_t)(pl.pair ? pl.BN / 2 : pl.BN)};
      int rc = encode(h, &P.tmB, const_cast<void*>(d->weights), 2, dims, str, box, pl.KC);
      if (rc != DISCO_OK) return rc;
    }
    c.grid = P.tiles_total < h->sm_count ? P.tiles_total : h->sm_count;
    if (pl.pair) c.grid = 2 * (P.tiles_total < h->sm_count / 2 ? P.tiles_total : h->sm_count / 2);
    it = g_cache.emplace(key, c).first;
  }
  Cached& c = it->second;
  if (c.plan.resident) {
    // constant-bank epilogue parameters: host copies are re-read at EVERY launch (the descriptor cache must not
    // freeze parameter values)
    TcParams& P = c.params;
    const bool have = g_allow_const && d->bias_host != nullptr && (d->post_scale == nullptr) == (d->post_scale_host == nullptr) &&
                      (d->post_shift == nullptr) == (d->post_shift_host == nullptr) &&
                      c.plan.gray_src < 0 && d->Cout <= 64;
    P.const_params = have ? 1 : 0;
    if (have) {
      for (int j = 0; j < 64; ++j) {
        const bool ok = j < d->Cout;
        P.epi_c[0][j] = ok ? d->bias_host[j] : 0.f;
        P.epi_c[1][j] = (ok && d->post_scale_host) ? d->post_scale_host[j] : 1.f;
        P.epi_c[2][j] = (ok && d->post_shift_host) ? d->post_shift_host[j] : 0.f;
      }
    }
  }
  int rc;
  if (c.plan.resident) {
    switch (c.plan.KC) {
      case 64: rc = launch_res_bn<64>(h, c.plan.BN, c.params, c.grid, c.plan.res_smem_bytes, c.plan.pair, st); break;
      case 32: rc = launch_res_bn<32>(h, c.plan.BN, c.params, c.grid, c.plan.res_smem_bytes, c.plan.pair, st); break;
      case 16: rc = launch_res_bn<16>(h, c.plan.BN, c.params, c.grid, c.plan.res_smem_bytes, c.plan.pair, st); break;
      default: disco_set_error("conv_tc: bad KC"); return DISCO_ERR_INVALID;
    }
  } else if (c.plan.grouped) {
    rc = launch_grp(h, c.plan, c.params, c.grid, st);
  } else if (c.plan.pair) {
    if (c.plan.BN == 256) rc = launch_pair_cfg<256, 64, 1>(h, c.params, c.grid, st);
    else rc = launch_pair_cfg<128, 64, 2>(h, c.params, c.grid, st);
  } else {
    switch (c.plan.KC) {
      case 64: rc = launch_bn<64>(h, c.plan.BN, c.plan.MT, c.params, c.grid, st); break;
      case 32: rc = launch_bn<32>(h, c.plan.BN, c.plan.MT, c.params, c.grid, st); break;
      case 16: rc = launch_bn<16>(h, c.plan.BN, c.plan.MT, c.params, c.grid, st); break;
      default: disco_set_error("conv_tc: bad KC"); return DISCO_ERR_INVALID;
    }
  }
  if (rc != DISCO_OK) return rc;
  DISCO_LAUNCH_CHECK(h);
  return DISCO_OK;
}
