// Shared device/host helpers for libphiseg_sm100.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/phiseg_sm100.h"

typedef __nv_bfloat16 bf16;

void phs_set_error(const char* fmt, ...);
int phs_check_launch(const char* what);

#define PHS_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      phs_set_error(__VA_ARGS__);     \
      return -1;                      \
    }                                 \
  } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// A step is ~1100 dependent launches; between two dependent CUDA-graph nodes the device idles for 1.2-2.5 us (launch +
// CTA scheduling + the kernel's own prologue).  Kernels launched through phs_launch carry the programmatic-stream-
// serialization attribute: the next kernel of the stream (graph branch) may be SCHEDULED while its predecessor still runs,
// executes its prologue (barrier init, TMEM allocation, index arithmetic), and blocks in griddepcontrol.wait until the
// predecessor has completed and its memory is visible.  Rule: a kernel launched through phs_launch must execute
// PHS_PDL_PROLOGUE() before its first global-memory access (tests/test_cabi_symbols.py checks the sources for it); it
// then also releases ITS successor (launch_dependents), so at most one kernel per stream is staged ahead.
// PHS_PDL=0 in the environment turns the attribute off (the instructions are no-ops then).
int phs_pdl_enabled();
#define PHS_PDL_PROLOGUE()                                          \
  do {                                                              \
    asm volatile("griddepcontrol.wait;" ::: "memory");              \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)

// Kernels that allocate TENSOR MEMORY split the prologue: PHS_PDL_WAIT() where PHS_PDL_PROLOGUE() would stand, and
// PHS_PDL_TRIGGER() only after the CTA holds its tensor memory (behind the barrier that publishes the TMEM base address).
// tcgen05.alloc blocks without a time limit when the SM's 512 columns are taken, and launch_dependents counts per CTA as soon
// as ONE thread issues it: with the trigger in front of the allocation, the other warps of a CTA released the successor
// kernel while its MMA warp was still waiting for columns; the successor's CTAs - resident early, allocating in their own
// prologue, then parked in griddepcontrol.wait until this kernel completes - could take exactly those columns: a cycle
// that never resolves (round 2: two training runs in ~20 stalled for minutes inside a step; all other waits in these
// kernels are bounded).  For the same reason tcgen05 kernels are not launched early THEMSELVES (phs_launch_tc: no
// programmatic attribute unless PHS_PDL_TC=1): a CTA that holds tensor memory never sits in griddepcontrol.wait.
#define PHS_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define PHS_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
int phs_pdl_tc_enabled();

template <typename... KArgs, typename... Args>
static inline void phs_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = phs_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// tensor-memory kernels (see PHS_PDL_WAIT above): the programmatic attribute only on request
template <typename... KArgs, typename... Args>
static inline void phs_launch_tc(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (phs_pdl_enabled() && phs_pdl_tc_enabled()) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// same, as clusters of two CTAs (cta_group::2 pairs: the two CTAs land on the two SMs of one TPC); grid.x must be even
template <typename... KArgs, typename... Args>
static inline void phs_launch_cluster2(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                       Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (phs_pdl_enabled() && phs_pdl_tc_enabled()) ? 2 : 1;      // (only tensor-memory kernels run as clusters)
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Division of a 31-bit index by a launch-time constant as multiply-high + shift (a 64-bit `/` or `%` costs ~50-100
// instructions, enough to make a 16-byte-per-thread streaming kernel ALU bound).  Valid for n < 2^31.
struct fdiv_t { uint32_t d, m, s; };
static inline fdiv_t fdiv_make(uint32_t d) {
  fdiv_t f;
  f.d = d;
  f.s = 0;
  while ((1u << f.s) < d) ++f.s;
  f.m = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << f.s) - d)) / d + 1);
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const fdiv_t& f) { return (__umulhi(n, f.m) + n) >> f.s; }
// flat index -> (vector, w, h, n) of an [N][H][W][nvec] iteration space
struct idx4_t { fdiv_t nvec, W, H; };
static inline idx4_t idx4_make(int nvec, int W, int H) {
  idx4_t r;
  r.nvec = fdiv_make((uint32_t)nvec); r.W = fdiv_make((uint32_t)W); r.H = fdiv_make((uint32_t)H);
  return r;
}
__device__ __forceinline__ void idx4_decode(uint32_t i, const idx4_t& f, int& cv, int& w, int& h, int& n) {
  const uint32_t pix = fdiv(i, f.nvec);
  cv = (int)(i - pix * f.nvec.d);
  const uint32_t row = fdiv(pix, f.W);
  w = (int)(pix - row * f.W.d);
  const uint32_t img = fdiv(row, f.H);
  h = (int)(row - img * f.H.d);
  n = (int)img;
}

// ---- dtype-generic scalar/vector access ------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }

template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// V consecutive channels; V==8 => one 16-byte access for bf16, two for float; V==4 likewise; V==1 scalar
template <typename T, int V>
__device__ __forceinline__ void ldv(const T* p, float* out) {
  if constexpr (V == 1) {
    out[0] = ldf<T>(p);
  } else if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
      out[i] = t.x; out[i + 1] = t.y; out[i + 2] = t.z; out[i + 3] = t.w;
    }
  } else {
    static_assert(V == 8 || V == 4, "bf16 vector width");
    if constexpr (V == 8) {
      uint4 t = *reinterpret_cast<const uint4*>(p);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        out[2 * i] = f.x; out[2 * i + 1] = f.y;
      }
    } else {
      uint2 t = *reinterpret_cast<const uint2*>(p);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        out[2 * i] = f.x; out[2 * i + 1] = f.y;
      }
    }
  }
}

template <typename T, int V>
__device__ __forceinline__ void stv(T* p, const float* in) {
  if constexpr (V == 1) {
    stf<T>(p, in[0]);
  } else if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int i = 0; i < V; i += 4)
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = make_float4(in[i], in[i + 1], in[i + 2], in[i + 3]);
  } else {
    if constexpr (V == 8) {
      uint4 t;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
      *reinterpret_cast<uint4*>(p) = t;
    } else {
      uint2 t;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
      for (int i = 0; i < 2; ++i) h[i] = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
      *reinterpret_cast<uint2*>(p) = t;
    }
  }
}

// 8 consecutive channels as loaded (16 bytes of bf16 / 32 bytes of float): unpacked only where they are consumed, so that
// several vectors per thread can be in flight without the unpacked copies filling the register file
template <typename T>
struct Raw8;
template <>
struct Raw8<bf16> {
  uint4 r;
  __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void unpack(float* v) const {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <>
struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void unpack(float* v) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// ---- normalisation arithmetic shared by every kernel that applies batch_norm / group_norm2D + ReLU -----------------
// (tfwrapper/normalisation.py:17-36,145-163, tfwrapper/layers.py:134-135).  The activation a = act(gamma*(y-mean)*rstd +
// beta) is produced in three places - the stand-alone norm_act kernels, the operand transform of the fused convolution
// (conv_halo.cu, PRE) and the re-materialisation in norm_bwd_reduce - and all three must give the SAME bits, so the
// roundings are spelled out (no FMA contraction left to the compiler).
__device__ __forceinline__ void norm_moments(double s, double q, double cnt, float eps, double* m, double* var, double* r) {
  const double mm = s / cnt;
  double v = __dsub_rn(q / cnt, __dmul_rn(mm, mm));
  if (v < 0) v = 0;
  *m = mm;
  *var = v;
  *r = 1.0 / sqrt(__dadd_rn(v, (double)eps));
}
__device__ __forceinline__ void norm_scale_shift(float gamma, float beta, float mean, float rstd, float* sc, float* sh) {
  const float s = __fmul_rn(gamma, rstd);
  *sc = s;
  *sh = __fsub_rn(beta, __fmul_rn(mean, s));
}
__device__ __forceinline__ float norm_act1(float y, float sc, float sh, int relu) {
  const float r = fmaf(y, sc, sh);
  return (relu && r < 0.f) ? 0.f : r;
}
// batch-norm moving averages (decay 0.99, Bessel-corrected variance: tf.contrib.layers.batch_norm, normalisation.py:27-34)
__device__ __forceinline__ void bn_moving_update(float* moving_mean, float* moving_var, int c, float decay, double m,
                                                 double var, double cnt) {
  const double unb = var * (cnt / (cnt > 1 ? cnt - 1 : 1));
  moving_mean[c] = __fadd_rn(__fmul_rn(decay, moving_mean[c]), __fmul_rn(1.f - decay, (float)m));
  moving_var[c] = __fadd_rn(__fmul_rn(decay, moving_var[c]), __fmul_rn(1.f - decay, (float)unb));
}

// widest vector (in elements) usable for a channel-slice tensor
static inline int pick_vec(const phs_tensor* t) {
  int es = t->dtype == PHS_BF16 ? 2 : 4;
  uintptr_t a = (uintptr_t)t->ptr;
  if (t->C % 8 == 0 && t->ld % 8 == 0 && a % (8 * es > 16 ? 16 : 8 * es) == 0) return 8;
  if (t->C % 4 == 0 && t->ld % 4 == 0 && a % (4 * es > 16 ? 16 : 4 * es) == 0) return 4;
  return 1;
}
static inline int min_vec(int a, int b) { return a < b ? a : b; }

#define PHS_DISPATCH_DTYPE(dt, T, ...)                  \
  if ((dt) == PHS_F32) { typedef float T; __VA_ARGS__; } \
  else { typedef bf16 T; __VA_ARGS__; }

#define PHS_DISPATCH_VEC(v, V, ...)                        \
  if ((v) == 8) { constexpr int V = 8; __VA_ARGS__; }        \
  else if ((v) == 4) { constexpr int V = 4; __VA_ARGS__; }   \
  else { constexpr int V = 1; __VA_ARGS__; }
