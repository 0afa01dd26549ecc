// Shared host/device helpers for libdisco_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include "../../include/disco_b200.h"

struct disco_handle {
  int device;
  int sm_count;
  int64_t launches;
  void* tmap_encode;   // cuTensorMapEncodeTiled entry point (resolved lazily)
  bool use_tc;         // route supported bf16 descriptors to the tcgen05 kernel
  // Per-device state (one handle per (process, device)): the opt-in dynamic shared-memory size of a kernel is a
  // per-device attribute, and the watchdog flag the tensor-core kernels raise lives in this device's memory.
  void* smem_attr;     // std::map<const void*, int>*: largest MaxDynamicSharedMemorySize set per kernel on this device
  int32_t* error_flag; // device int32, 0 = ok (set by the mbarrier watchdog before it traps)
};

void disco_set_error(const char* fmt, ...);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel, size high-water mark)
int disco_ensure_smem(disco_handle* h, const void* func, int bytes);

// Every entry point runs on the handle's device whatever the caller's current device is (the reference CLI wraps the
// model in DataParallel when several GPUs are visible, so one process may drive several handles).
struct DiscoDeviceGuard {
  int prev = -1;
  bool changed = false;
  explicit DiscoDeviceGuard(const disco_handle* h) {
    if (h && cudaGetDevice(&prev) == cudaSuccess && prev != h->device) changed = cudaSetDevice(h->device) == cudaSuccess;
  }
  ~DiscoDeviceGuard() { if (changed) cudaSetDevice(prev); }
};

#define DISCO_CHECK_ARG(cond, ...)                 \
  do {                                             \
    if (!(cond)) {                                 \
      disco_set_error(__VA_ARGS__);                \
      return DISCO_ERR_INVALID;                    \
    }                                              \
  } while (0)

#define DISCO_CUDA(call)                                                                   \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      disco_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return DISCO_ERR_CUDA;                                                               \
    }                                                                                      \
  } while (0)

#define DISCO_LAUNCH_CHECK(h)                      \
  do {                                             \
    (h)->launches++;                               \
    DISCO_CUDA(cudaGetLastError());                \
  } while (0)

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Packed fp32 pairs (Blackwell FFMA2): one instruction = two fp32 FMAs, full fp32 precision.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == DISCO_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == DISCO_ACT_LRELU) return v > 0.f ? v : v * slope;
  return v;
}

// internal entry points implemented in the individual .cu files
int conv_simt_launch(disco_handle* h, const disco_conv_desc* d, cudaStream_t st);
int conv_tc_launch(disco_handle* h, const disco_conv_desc* d, cudaStream_t st);
bool conv_tc_supported(const disco_conv_desc* d);
// narrow-layer mma.sync kernels (conv_narrow.cu): claim a bf16 descriptor before the tcgen05 plan is consulted
bool conv_narrow_match(const disco_conv_desc* d);
int64_t conv_narrow_weight_elems(const disco_conv_desc* d);
int conv_narrow_pack(const disco_conv_desc* d, const float* w32_host, uint16_t* out_host);
int conv_narrow_launch(disco_handle* h, const disco_conv_desc* d, cudaStream_t st);
// 64 -> 64 stride-1 layers with the A operand in tensor memory (conv_ts.cu); DISCO_TS selects it
bool conv_ts_match(const disco_conv_desc* d);
int64_t conv_ts_weight_elems(const disco_conv_desc* d);
int conv_ts_pack(const disco_conv_desc* d, const float* w32_host, uint16_t* out_host);
int conv_ts_launch(disco_handle* h, const disco_conv_desc* d, cudaStream_t st);
void conv_tc_cache_clear(disco_handle* h);   // frees the cached plans of h->device (caller holds the device guard)
