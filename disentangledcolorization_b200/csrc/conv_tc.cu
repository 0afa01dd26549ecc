// tcgen05 implicit-GEMM convolution (bf16 activations/weights, fp32 TMEM accumulators).
// Placeholder until the tensor-core kernel lands: nothing is claimed supported, so disco_conv
// routes every descriptor to the CUDA-core kernel.
#include "common.cuh"

bool conv_tc_supported(const disco_conv_desc*) { return false; }
int conv_tc_launch(disco_handle*, const disco_conv_desc*, cudaStream_t) {
  disco_set_error("conv_tc: not built");
  return DISCO_ERR_UNSUPPORTED;
}
