// Host-side helpers shared by the tensor-core convolution translation units: cached TMA tensor maps, launch helpers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

// 4-D tensor map over an NHWC bf16 channel slice: dims (C, W, H, N), box (bc, bw, bh, bn); swizzle = bc * 2 bytes (64 | 128)
int activation_map(const phs_tensor* t, int bc, int bw, int bh, int bn, CUtensorMap* out);
// 2-D tensor map over a K-major bf16 filter shadow [rows][K]: box (bk, box_rows)
int filter_map(const void* w, int K, int rows, int bk, CUtensorMap* out);
int filter_map_rows(const void* w, int K, int rows, int bk, int box_rows, CUtensorMap* out);
int num_sms();

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// dynamic shared memory opt-in: 227 KB per CTA minus the kernel's static shared memory
constexpr int SMEM_OPTIN = 227 * 1024 - 2048;
template <typename K>
int allow_big_smem(K kernel, bool* done) {
  if (*done) return 0;
  // (a kernel with more than 2 KB of static shared memory gets what is left of the 227 KB)
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
  int limit = SMEM_OPTIN;
  if (e == cudaSuccess && 227 * 1024 - (int)fa.sharedSizeBytes < limit) limit = 227 * 1024 - (int)fa.sharedSizeBytes;
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, limit);
  if (e != cudaSuccess) {
    phs_set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize): %s", cudaGetErrorString(e));
    cudaGetLastError();      // (do not leave the error behind for the caller's next runtime call)
    return (int)e;
  }
  *done = true;
  return 0;
}
