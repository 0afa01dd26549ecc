// C-ABI plumbing of libdisco_b200.so: handle lifetime, error reporting, conv dispatch.
#include "common.cuh"
#include <cstddef>
#include <cstring>
#include <map>
#include <new>
#include <vector>

static thread_local char g_err[512] = "";

void disco_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int disco_version(void) { return 101; }
// ABI self-description for bindings: 0 = sizeof(disco_conv_src), 1 = sizeof(disco_conv_desc), 2 = sizeof(disco_linear_desc),
// 3 = offsetof(disco_conv_desc, out), 4 = offsetof(disco_conv_desc, bias_host)
extern "C" int disco_abi_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(disco_conv_src);
    case 1: return (int)sizeof(disco_conv_desc);
    case 2: return (int)sizeof(disco_linear_desc);
    case 3: return (int)offsetof(disco_conv_desc, out);
    case 4: return (int)offsetof(disco_conv_desc, bias_host);
  }
  return -1;
}

extern "C" const char* disco_last_error(void) { return g_err; }

extern "C" int disco_create(disco_handle** out, int device) {
  DISCO_CHECK_ARG(out != nullptr, "disco_create: out is NULL");
  int count = 0;
  DISCO_CUDA(cudaGetDeviceCount(&count));
  DISCO_CHECK_ARG(device >= 0 && device < count, "disco_create: device %d out of range (%d visible)", device, count);
  cudaDeviceProp prop;
  DISCO_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    disco_set_error("disco_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    return DISCO_ERR_UNSUPPORTED;
  }
  disco_handle* h = new (std::nothrow) disco_handle();
  DISCO_CHECK_ARG(h != nullptr, "disco_create: out of host memory");
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->launches = 0;
  h->tmap_encode = nullptr;
  h->use_tc = true;
  h->smem_attr = new (std::nothrow) std::map<const void*, int>();
  h->error_flag = nullptr;
  {
    DiscoDeviceGuard guard(h);
    if (cudaMalloc(&h->error_flag, sizeof(int32_t)) != cudaSuccess || cudaMemset(h->error_flag, 0, sizeof(int32_t)) != cudaSuccess) {
      disco_set_error("disco_create: cannot allocate the device error flag on device %d", device);
      delete reinterpret_cast<std::map<const void*, int>*>(h->smem_attr);
      delete h;
      return DISCO_ERR_CUDA;
    }
  }
  *out = h;
  return DISCO_OK;
}

extern "C" int disco_destroy(disco_handle* h) {
  if (h) {
    DiscoDeviceGuard guard(h);
    conv_tc_cache_clear(h);
    if (h->error_flag) cudaFree(h->error_flag);
    delete reinterpret_cast<std::map<const void*, int>*>(h->smem_attr);
  }
  delete h;
  return DISCO_OK;
}

int disco_ensure_smem(disco_handle* h, const void* func, int bytes) {
  auto* m = reinterpret_cast<std::map<const void*, int>*>(h->smem_attr);
  DISCO_CHECK_ARG(m != nullptr, "handle has no attribute table");
  int& have = (*m)[func];
  if (bytes > have) {
    DISCO_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    have = bytes;
  }
  return DISCO_OK;
}

extern "C" int64_t disco_launch_count(disco_handle* h) { return h ? h->launches : -1; }
extern "C" void disco_reset_launch_count(disco_handle* h) { if (h) h->launches = 0; }
extern "C" void disco_add_launch_count(disco_handle* h, int64_t n) { if (h) h->launches += n; }

extern "C" int disco_conv(disco_handle* h, const disco_conv_desc* d, void* stream) {
  DISCO_CHECK_ARG(h && d, "conv: null handle/descriptor");
  DISCO_CHECK_ARG(d->n_src >= 1 && d->n_src <= 2, "conv: n_src must be 1 or 2 (got %d)", d->n_src);
  DiscoDeviceGuard guard(h);
  DISCO_CHECK_ARG(d->out && d->weights && d->bias, "conv: null out/weights/bias");
  DISCO_CHECK_ARG(d->batch > 0 && d->Ho > 0 && d->Wo > 0 && d->Cout > 0, "conv: bad output shape");
  DISCO_CHECK_ARG(d->kind == DISCO_CONV3 || d->kind == DISCO_DECONV4, "conv: unknown kind %d", d->kind);
  DISCO_CHECK_ARG(d->kind == DISCO_DECONV4 || d->stride == 1 || d->stride == 2, "conv: stride must be 1 or 2");
  DISCO_CHECK_ARG(d->dtype == DISCO_F32 || d->dtype == DISCO_BF16, "conv: unknown dtype %d", d->dtype);
  DISCO_CHECK_ARG(d->head == DISCO_HEAD_NONE || (d->head == DISCO_HEAD_SOFTMAX9 && d->Cout == 9) ||
                      ((d->head == DISCO_HEAD_TANH2 || d->head == DISCO_HEAD_RAW2) && d->Cout == 2),
                  "conv: head/Cout mismatch");
  for (int s = 0; s < d->n_src; ++s) {
    const disco_conv_src& src = d->src[s];
    DISCO_CHECK_ARG(src.ptr && src.H > 0 && src.W > 0 && src.C > 0, "conv: bad source %d", s);
    if (d->kind == DISCO_DECONV4) {
      DISCO_CHECK_ARG(src.H * 2 == d->Ho && src.W * 2 == d->Wo && !src.up2, "conv: deconv source %d shape mismatch", s);
    } else {
      const int vh = src.H << src.up2, vw = src.W << src.up2;
      DISCO_CHECK_ARG(vh == d->Ho * d->stride && vw == d->Wo * d->stride,
                      "conv: source %d is %dx%d (up2=%d) but output %dx%d stride %d", s, src.H, src.W, src.up2, d->Ho,
                      d->Wo, d->stride);
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == DISCO_BF16 && h->use_tc && conv_tc_supported(d)) return conv_tc_launch(h, d, st);
  return conv_simt_launch(h, d, st);
}

// ------------------------------------------------------------------------------------------------ host RNG helper
namespace {
// MT19937 exactly as numpy's legacy bit generator steps it (state = key[624] + pos; refill when pos == 624)
struct Mt {
  uint32_t* key;
  int pos;
  void refill() {
    constexpr int N = 624, M = 397;
    constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
    int i;
    uint32_t y;
    for (i = 0; i < N - M; i++) {
      y = (key[i] & UPPER) | (key[i + 1] & LOWER);
      key[i] = key[i + M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    for (; i < N - 1; i++) {
      y = (key[i] & UPPER) | (key[i + 1] & LOWER);
      key[i] = key[i + (M - N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    y = (key[N - 1] & UPPER) | (key[0] & LOWER);
    key[N - 1] = key[M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    pos = 0;
  }
  uint32_t next() {
    if (pos == 624) refill();
    uint32_t y = key[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
};
}  // namespace

extern "C" int disco_host_choice_rows(uint32_t* mt_key, int32_t* mt_pos, int S, int K, int rows, int keep_lo, int keep_hi,
                                      int32_t* out) {
  DISCO_CHECK_ARG(mt_key && mt_pos && out, "host_choice_rows: null pointer");
  DISCO_CHECK_ARG(S >= 1 && K >= 1 && K <= S && rows >= 0, "host_choice_rows: need 1 <= K <= S (got K=%d, S=%d)", K, S);
  DISCO_CHECK_ARG(*mt_pos >= 0 && *mt_pos <= 624, "host_choice_rows: bad MT19937 position %d", *mt_pos);
  DISCO_CHECK_ARG(keep_lo >= 0 && keep_lo <= keep_hi && keep_hi <= rows, "host_choice_rows: bad keep range");
  Mt mt{mt_key, *mt_pos};
  std::vector<int32_t> arr((size_t)S);
  for (int r = 0; r < rows; ++r) {
    for (int i = 0; i < S; ++i) arr[i] = i;
    for (int i = S - 1; i >= 1; --i) {          // numpy _shuffle_raw: j = random_interval(i); swap(arr[i], arr[j])
      uint32_t mask = (uint32_t)i;
      mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
      uint32_t v;
      while ((v = mt.next() & mask) > (uint32_t)i) {}
      const int32_t t = arr[i]; arr[i] = arr[v]; arr[v] = t;
    }
    if (r >= keep_lo && r < keep_hi)
      for (int k = 0; k < K; ++k) out[(size_t)(r - keep_lo) * K + k] = arr[k];
  }
  *mt_pos = mt.pos;
  return DISCO_OK;
}
