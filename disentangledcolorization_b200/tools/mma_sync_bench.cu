// Microbenchmark: peak rate of the legacy warp-level tensor path (mma.sync bf16) on sm_100a, as used by the fused token
// stack and the fused narrow-layer SpixelNet kernels.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_bench.bin
// Prints MACs/clk/SM for m16n8k16 and m16n8k8 with W warps per SM and ILP independent accumulators per warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}

template <int ILP, bool K8>
__global__ void bench(float* out, int iters, long long* cycles) {
  float c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (K8) mma1688(c[i], a0, a1, b0); else mma16816(c[i], a0, a1, a2, a3, b0, b1);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP, bool K8>
void run(int warps, float* out, long long* cyc) {
  const int iters = 4096;
  bench<ILP, K8><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  bench<ILP, K8><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double macs = (double)warps * iters * ILP * (K8 ? 1024.0 : 2048.0);
  printf("%s warps/SM=%2d ILP=%d : %8.1f MAC/clk/SM  (%.2f clk per MMA per SM)\n", K8 ? "m16n8k8 " : "m16n8k16", warps, ILP,
         macs / (double)h, (double)h / ((double)warps * iters * ILP));
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  for (int w : {4, 8, 16}) {
    run<1, false>(w, out, cyc);
    run<4, false>(w, out, cyc);
    run<8, false>(w, out, cyc);
    run<8, true>(w, out, cyc);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
