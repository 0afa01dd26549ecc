// Micro-benchmark: issue cost of tcgen05.mma (kind::f16, cta_group::1) for several tile shapes, operands in
// shared memory (SS) or A in tensor memory (TS).  One CTA per SM, one issuing thread, NREP back-to-back MMAs
// followed by one commit; reports cycles per MMA.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // K-major SW128, SBO = 1024 B
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int M, int N, bool A_TMEM, int KSTEPS>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int nrep) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (128 * 128 + 256 * 128) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tslot;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (warp == 1 && elect_one()) {
    const uint64_t ad = make_desc(smem_u32(smem)), bd = make_desc(smem_u32(smem + 128 * 128));
    const uint32_t a_t = tm + 256;   // A operand columns (TS variant): garbage contents are fine for timing
    long long t0 = clock64();
    for (int r = 0; r < nrep; ++r) {
#pragma unroll
      for (int k = 0; k < KSTEPS; ++k) {
        if (A_TMEM)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::
                           "r"(tm), "r"(a_t + k * 8), "l"(bd + 2 * k), "r"(idesc), "r"(1));
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::
                           "r"(tm), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(1));
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    long long t1 = clock64();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)));
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

template <int M, int N, bool A_TMEM>
void run(const char* name, long long* dout) {
  const int nrep = 2000, smem = 128 * 128 + 256 * 128 + 2048;
  auto k = bench<M, N, A_TMEM, 4>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 148}) {
    k<<<grid, 128, smem>>>(dout, nrep);
    k<<<grid, 128, smem>>>(dout, nrep);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s grid=%3d  issue %.1f cyc/MMA  complete %.1f cyc/MMA  (%s)\n", name, grid, (double)h[0] / (nrep * 4),
           (double)h[1] / (nrep * 4), cudaGetErrorString(e));
  }
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 64);
  run<128, 256, false>("SS M=128 N=256", dout);
  run<128, 128, false>("SS M=128 N=128", dout);
  run<128, 64, false>("SS M=128 N=64", dout);
  run<128, 32, false>("SS M=128 N=32", dout);
  run<128, 16, false>("SS M=128 N=16", dout);
  run<64, 256, false>("SS M=64  N=256", dout);
  run<64, 128, false>("SS M=64  N=128", dout);
  run<64, 64, false>("SS M=64  N=64", dout);
  run<128, 256, true>("TS M=128 N=256 (A in TMEM)", dout);
  run<128, 128, true>("TS M=128 N=128 (A in TMEM)", dout);
  run<128, 64, true>("TS M=128 N=64 (A in TMEM)", dout);
  run<128, 16, true>("TS M=128 N=16 (A in TMEM)", dout);
  return 0;
}
