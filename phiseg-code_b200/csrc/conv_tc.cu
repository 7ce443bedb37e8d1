// placeholder until the tcgen05 kernels land (next milestone)
#include "common.cuh"
int conv2d_tc(const phs_tensor*, const void*, const float*, const phs_tensor*, int, int, int, float*, cudaStream_t) {
  phs_set_error("tensor-core convolution not built into this library yet");
  return -2;
}
int conv2d_wgrad_tc(const phs_tensor*, const phs_tensor*, float*, float*, int, int, cudaStream_t) {
  phs_set_error("tensor-core filter gradient not built into this library yet");
  return -2;
}
