// C-ABI plumbing of libdisco_b200.so: handle lifetime, error reporting, conv dispatch.
#include "common.cuh"
#include <cstddef>
#include <cstring>
#include <new>

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
  *out = h;
  return DISCO_OK;
}

extern "C" int disco_destroy(disco_handle* h) {
  delete h;
  return DISCO_OK;
}

extern "C" int64_t disco_launch_count(disco_handle* h) { return h ? h->launches : -1; }
extern "C" void disco_reset_launch_count(disco_handle* h) { if (h) h->launches = 0; }
extern "C" void disco_add_launch_count(disco_handle* h, int64_t n) { if (h) h->launches += n; }

extern "C" int disco_conv(disco_handle* h, const disco_conv_desc* d, void* stream) {
  DISCO_CHECK_ARG(h && d, "conv: null handle/descriptor");
  DISCO_CHECK_ARG(d->out && d->weights && d->bias, "conv: null out/weights/bias");
  DISCO_CHECK_ARG(d->batch > 0 && d->Ho > 0 && d->Wo > 0 && d->Cout > 0, "conv: bad output shape");
  DISCO_CHECK_ARG(d->kind == DISCO_CONV3 || d->kind == DISCO_DECONV4, "conv: unknown kind %d", d->kind);
  DISCO_CHECK_ARG(d->kind == DISCO_DECONV4 || d->stride == 1 || d->stride == 2, "conv: stride must be 1 or 2");
  DISCO_CHECK_ARG(d->dtype == DISCO_F32 || d->dtype == DISCO_BF16, "conv: unknown dtype %d", d->dtype);
  DISCO_CHECK_ARG(d->head == DISCO_HEAD_NONE || (d->head == DISCO_HEAD_SOFTMAX9 && d->Cout == 9) ||
                      (d->head == DISCO_HEAD_TANH2 && d->Cout == 2),
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
