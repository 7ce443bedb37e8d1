// tcgen05 implicit-GEMM convolutions for sm_100a (tf.nn.conv2d SAME stride 1 and its two gradients,
// tfwrapper/layers.py:123 + the Conv2DBackpropInput/Filter nodes optimizer.minimize adds, phiseg_model.py:141).
//
// Forward / input gradient  (conv_tc_kernel):
//   GEMM  D[pixel][cout] = sum_{tap, ci} X[pixel + tap][ci] * Wt[cout][tap*Cin + ci]
//   - a CTA owns 128 output pixels (a TN x TH x TW brick of the NHWC tensor) and ALL output channels (N = Cout <= 256);
//   - the A tile of one (tap, 64-channel chunk) is ONE TMA box of the 4-D activation tensor map taken at the brick
//     origin shifted by the tap; out-of-bounds rows/columns are zero-filled by the TMA unit, which IS the SAME padding;
//     the box lands as 128 rows x 128 B, 128B-swizzled = the canonical K-major UMMA operand layout;
//   - the B tile is a 2-D box [Cout][64] of the K-major bf16 filter shadow (phs_weight_prep);
//   - warp 0 = TMA producer, warp 1 = MMA issuer (one thread, tcgen05.mma M=128, N=Cout, K=16), warps 2..5 = epilogue
//     (tcgen05.ld -> +bias -> bf16/f32 -> global); accumulators are double-buffered in TMEM (2 x 256 columns) so the
//     epilogue of tile i overlaps the main loop of tile i+1; the kernel is persistent (grid = #SMs).
// Filter gradient (wgrad_tc_kernel):
//   GEMM  D[ci][co] = sum_{pixel} X[pixel + tap][ci] * dY[pixel][co]   (reduction over pixels)
//   - both operands are MN-major in shared memory (channels contiguous), again plain TMA boxes of the NHWC tensors;
//   - a CTA owns (tap, 128-channel block of Cin, a slice of the pixel range); partial sums are added to the fp32 HWIO
//     gradient with vector red.global.add (split-K).
#include <stdlib.h>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>
#include <string.h>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_host.cuh"

using namespace tc;

namespace {

// ------------------------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

typedef std::tuple<const void*, int, int, int, int, int, int, int, int, int, int> MapKey;
std::map<MapKey, CUtensorMap> g_maps;
std::mutex g_maps_mu;

}  // namespace

// 4-D map over an NHWC bf16 channel slice: dims (C, W, H, N), box (bc, bw, bh, bn); swz_bytes = bc * 2 in {64, 128}
int activation_map(const phs_tensor* t, int bc, int bw, int bh, int bn, CUtensorMap* out) {
  MapKey key(t->ptr, t->N, t->H, t->W, t->C, t->ld, bc, bw, bh, bn, 4);
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) {
    *out = it->second;
    return 0;
  }
  EncodeTiledFn enc = encode_fn();
  PHS_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[4] = {(cuuint64_t)t->C, (cuuint64_t)t->W, (cuuint64_t)t->H, (cuuint64_t)t->N};
  cuuint64_t strides[3] = {(cuuint64_t)t->ld * 2, (cuuint64_t)t->W * t->ld * 2, (cuuint64_t)t->H * t->W * t->ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle swz = bc * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t->ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PHS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation C=%d W=%d H=%d N=%d ld=%d box %d,%d,%d,%d) failed: %d",
              t->C, t->W, t->H, t->N, t->ld, bc, bw, bh, bn, (int)r);
  g_maps[key] = *out;
  return 0;
}

// 2-D map over the K-major filter shadow [rows][K] bf16: box (bk, box_rows)
int filter_map_rows(const void* w, int K, int rows, int bk, int box_rows, CUtensorMap* out) {
  MapKey key(w, K, rows, bk, box_rows, 0, 0, 0, 0, 0, 2);
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) {
    *out = it->second;
    return 0;
  }
  EncodeTiledFn enc = encode_fn();
  PHS_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUtensorMapSwizzle swz = bk * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PHS_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(filter K=%d rows=%d bk=%d) failed: %d", K, rows, bk, (int)r);
  g_maps[key] = *out;
  return 0;
}
int filter_map(const void* w, int K, int rows, int bk, CUtensorMap* out) { return filter_map_rows(w, K, rows, bk, rows, out); }

namespace {

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// a brick of P pixels (P a power of two): TW x TH x TN, widest along W first
struct Brick {
  int TW, TH, TN, tilesW, tilesH, tilesN, num;
};
Brick make_brick(int N, int H, int W, int P) {
  Brick b;
  b.TW = next_pow2(W) < P ? next_pow2(W) : P;
  int rest = P / b.TW;
  b.TH = next_pow2(H) < rest ? next_pow2(H) : rest;
  b.TN = rest / b.TH;
  b.tilesW = (W + b.TW - 1) / b.TW;
  b.tilesH = (H + b.TH - 1) / b.TH;
  b.tilesN = (N + b.TN - 1) / b.TN;
  b.num = b.tilesW * b.tilesH * b.tilesN;
  return b;
}

}  // namespace

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

namespace {

// ------------------------------------------------------------------------------------------------------------
// forward / dgrad kernel
// ------------------------------------------------------------------------------------------------------------
constexpr int MAX_STAGES = 16;

struct ConvParams {
  int N, H, W, Cin, Cout;
  int taps;  // 1 or 9
  int TW, TH, TN, tilesW, tilesH, num_tiles;
  int kchunks;  // Cin / BK
  int stages;
  void* y;
  int y_ld, y_f32;
  const float* bias;
  int accumulate;
  int post_on;          // inference-mode batch norm + ReLU of this layer folded into the epilogue (phs_conv2d_post)
  phs_norm_pre post;
};

// post16: (scale, shift) pairs of the 16 channels, or null
__device__ __forceinline__ void post_affine16(float* v, const float* post16, int relu) {
  if (post16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = norm_act1(v[i], post16[2 * i], post16[2 * i + 1], relu);
  }
}

template <typename T>
__device__ __forceinline__ void store_row16(T* dst, const uint32_t* r, const float* bias16, bool accumulate,
                                            const float* post16 = nullptr, int relu = 0);

template <>
__device__ __forceinline__ void store_row16<bf16>(bf16* dst, const uint32_t* r, const float* bias16, bool accumulate,
                                                  const float* post16, int relu) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + (bias16 ? bias16[i] : 0.f);
  post_affine16(v, post16, relu);
  if (accumulate) {
    float o[16];
    ldv<bf16, 8>(dst, o);
    ldv<bf16, 8>(dst + 8, o + 8);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += o[i];
  }
  stv<bf16, 8>(dst, v);
  stv<bf16, 8>(dst + 8, v + 8);
}
template <>
__device__ __forceinline__ void store_row16<float>(float* dst, const uint32_t* r, const float* bias16, bool accumulate,
                                                   const float* post16, int relu) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + (bias16 ? bias16[i] : 0.f);
  post_affine16(v, post16, relu);
  if (accumulate) {
    float o[16];
    ldv<float, 8>(dst, o);
    ldv<float, 8>(dst + 8, o + 8);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += o[i];
  }
  stv<float, 8>(dst, v);
  stv<float, 8>(dst + 8, v + 8);
}

template <int BK>
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 4];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[256];
  __shared__ float post_s[512];      // (scale, shift) per output channel when p.post_on (one CTA per SM: room to spare)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t A_BYTES = 128 * BK * 2;
  const uint32_t B_BYTES = (uint32_t)p.Cout * BK * 2;
  const uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * MAX_STAGES + 2 + a); };

  if (warp == W_PROD && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(smem_u32(&tmem_base_s), 512);
  PHS_PDL_WAIT();
  for (int c = threadIdx.x; c < 256; c += blockDim.x) bias_s[c] = (p.bias && c < p.Cout) ? p.bias[c] : 0.f;
  if (p.post_on) {
    for (int c = threadIdx.x; c < p.Cout; c += blockDim.x)
      norm_scale_shift(p.post.gamma[c], p.post.beta[c], p.post.moving_mean[c], rsqrtf(p.post.moving_var[c] + p.post.eps),
                       &post_s[2 * c], &post_s[2 * c + 1]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  PHS_PDL_TRIGGER();      // this CTA holds its tensor memory: the successor kernel may be scheduled now
  const int kiters = p.taps * p.kchunks;

  // lean, warp-uniform issue loops: ring index / phase are counters, descriptors advance by adds (see conv_halo.cu)
  const int stages = p.stages, kchunks = p.kchunks, taps = p.taps, Cin = p.Cin;
  if (warp == W_PROD) {
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int w0 = (tile % p.tilesW) * p.TW;
      const int h0 = ((tile / p.tilesW) % p.tilesH) * p.TH;
      const int n0 = (tile / (p.tilesW * p.tilesH)) * p.TN;
      int dh = taps == 9 ? -1 : 0, dw = dh;
#pragma unroll 1
      for (int tap = 0; tap < taps; ++tap) {
#pragma unroll 1
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(empty_bar(s), ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(full_bar(s), STAGE_BYTES);
            const uint32_t a_s = smem0 + s * STAGE_BYTES;
            tma_load_4d(a_s, &tmA, full_bar(s), kc * BK, w0 + dw, h0 + dh, n0);
            tma_load_2d(a_s + A_BYTES, &tmB, full_bar(s), tap * Cin + kc * BK, 0);
          }
          __syncwarp();
          if (++s == stages) { s = 0; ph ^= 1; }
        }
        if (++dw == 2) { dw = -1; ++dh; }
      }
    }
  } else if (warp == W_MMA) {
    const uint32_t idesc = idesc_bf16(128, p.Cout, 0, 0);
    constexpr uint64_t LAYOUT = BK == 64 ? LAYOUT_SW128 : LAYOUT_SW64;
    constexpr uint32_t SBO = BK == 64 ? 1024 : 512;
    const uint32_t hi = desc_hi(SBO, LAYOUT);
    const uint32_t a0_lo = desc_lo(smem0, 16), b0_lo = desc_lo(smem0 + A_BYTES, 16), stage16 = STAGE_BYTES >> 4;
    int s = 0;
    uint32_t ph = 0, acc = 0, aph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), aph ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * 256;
#pragma unroll 1
      for (int kit = 0; kit < kiters; ++kit) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a0_lo + s * stage16, b_lo = b0_lo + s * stage16;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_lohi(d, a_lo + 2 * k, hi, b_lo + 2 * k, hi, idesc, k ? 1u : (kit != 0 ? 1u : 0u));
          umma_commit(empty_bar(s));
          if (kit == kiters - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) aph ^= 1;
    }
  } else {
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
      const int w0 = (tile % p.tilesW) * p.TW;
      const int h0 = ((tile / p.tilesW) % p.tilesH) * p.TH;
      const int n0 = (tile / (p.tilesW * p.tilesH)) * p.TN;
      const int m = q * 32 + lane;
      const int w = w0 + m % p.TW, h = h0 + (m / p.TW) % p.TH, n = n0 + m / (p.TW * p.TH);
      const bool valid = w < p.W && h < p.H && n < p.N;
      const size_t pix = ((size_t)n * p.H + h) * p.W + w;
      mbar_wait(tfull_bar(acc), aph);
      tc_fence_after();
      const uint32_t t0 = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(t0 + c0, r);
        tmem_ld_wait();
        if (valid) {
          const float* post16 = p.post_on ? post_s + 2 * c0 : nullptr;
          if (p.y_f32)
            store_row16<float>((float*)p.y + pix * p.y_ld + c0, r, p.bias ? bias_s + c0 : nullptr, p.accumulate, post16,
                               p.post.relu);
          else
            store_row16<bf16>((bf16*)p.y + pix * p.y_ld + c0, r, p.bias ? bias_s + c0 : nullptr, p.accumulate, post16,
                              p.post.relu);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------
// filter-gradient kernel
// ------------------------------------------------------------------------------------------------------------
struct WgradParams {
  int N, H, W, Cin, Cout;
  int taps;
  int TW, TH, TN, tilesW, tilesH, num_ptiles;  // bricks of 64 pixels
  int mblocks;                                 // ceil(Cin / 128)
  int tiles_per_split;
  int slabA, nslabA;  // channels per A slab (64 | 32), slabs per 128-row block
  int slabB, nslabB;  // same for dY; nslabB * slabB == Cout
  int stages;
  int tmem_cols;
  float* dw;
};

constexpr int KT = 64;  // pixels per pipeline stage

__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 1];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t slabA_bytes = KT * p.slabA * 2, slabB_bytes = KT * p.slabB * 2;
  const uint32_t A_BYTES = slabA_bytes * p.nslabA, B_BYTES = slabB_bytes * p.nslabB;
  const uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * MAX_STAGES);

  const int tap = blockIdx.y % p.taps, mb = blockIdx.y / p.taps;
  const int t_begin = blockIdx.x * p.tiles_per_split;
  const int t_end = min(p.num_ptiles, t_begin + p.tiles_per_split);

  if (warp == W_PROD && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(smem_u32(&tmem_base_s), p.tmem_cols);
  PHS_PDL_WAIT();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  PHS_PDL_TRIGGER();      // this CTA holds its tensor memory: the successor kernel may be scheduled now

  const int stages = p.stages;
  if (warp == W_PROD) {
    const int dh = p.taps == 9 ? tap / 3 - 1 : 0;
    const int dw = p.taps == 9 ? tap % 3 - 1 : 0;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int w0 = (tile % p.tilesW) * p.TW;
      const int h0 = ((tile / p.tilesW) % p.tilesH) * p.TH;
      const int n0 = (tile / (p.tilesW * p.tilesH)) * p.TN;
      mbar_wait(empty_bar(s), ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), STAGE_BYTES);
        const uint32_t a_s = smem0 + s * STAGE_BYTES;
        for (int j = 0; j < p.nslabA; ++j)
          tma_load_4d(a_s + j * slabA_bytes, &tmX, full_bar(s), mb * 128 + j * p.slabA, w0 + dw, h0 + dh, n0);
        for (int j = 0; j < p.nslabB; ++j)
          tma_load_4d(a_s + A_BYTES + j * slabB_bytes, &tmDY, full_bar(s), j * p.slabB, w0, h0, n0);
      }
      __syncwarp();
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == W_MMA) {
    const uint32_t idesc = idesc_bf16(128, p.Cout, 1, 1);
    const uint64_t layA = p.slabA == 64 ? LAYOUT_SW128 : LAYOUT_SW64;
    const uint64_t layB = p.slabB == 64 ? LAYOUT_SW128 : LAYOUT_SW64;
    const uint32_t rowA = p.slabA * 2, rowB = p.slabB * 2;  // bytes per pixel row inside a slab
    // MN-major: LBO = distance between 64(32)-channel slabs, SBO = distance between groups of 8 pixels
    const uint32_t a_hi = desc_hi(8 * rowA, layA), b_hi = desc_hi(8 * rowB, layB);
    const uint32_t a_k = (16 * rowA) >> 4, b_k = (16 * rowB) >> 4;
    int s = 0;
    uint32_t ph = 0, first = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_s = smem0 + s * STAGE_BYTES;
        const uint32_t a_lo = desc_lo(a_s, slabA_bytes), b_lo = desc_lo(a_s + A_BYTES, slabB_bytes);
#pragma unroll
        for (int k = 0; k < KT / 16; ++k)
          umma_bf16_lohi(tmem_base, a_lo + k * a_k, a_hi, b_lo + k * b_k, b_hi, idesc, k ? 1u : first);
        umma_commit(empty_bar(s));
        if (tile == t_end - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
      first = 1;
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else if (t_begin < t_end) {
    const int q = warp & 3;
    const int ci = mb * 128 + q * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16);
    float* dst = p.dw + ((size_t)tap * p.Cin + ci) * p.Cout;
    for (int c0 = 0; c0 < p.Cout; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(t0 + c0, r);
      tmem_ld_wait();
      if (ci < p.Cin) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + i), "f"(__uint_as_float(r[i])),
                       "f"(__uint_as_float(r[i + 1])), "f"(__uint_as_float(r[i + 2])), "f"(__uint_as_float(r[i + 3]))
                       : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// per-channel sum of dY (bias gradient) for the rare un-normalised tensor-core layer
__global__ void __launch_bounds__(256) bias_grad_bf16_kernel(const bf16* __restrict__ dy, int ld, int C, int64_t M,
                                                             float* __restrict__ db) {
  const int c = blockIdx.y * 32 + (threadIdx.x & 31);
  const int rows = blockDim.x >> 5, row = threadIdx.x >> 5;
  float a = 0.f;
  if (c < C)
    for (int64_t m = (int64_t)blockIdx.x * rows + row; m < M; m += (int64_t)gridDim.x * rows)
      a += __bfloat162float(dy[m * ld + c]);
  __shared__ float sm[256];
  sm[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x < 32 && c < C) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += sm[r * 32 + threadIdx.x];
    atomicAdd(db + c, s);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------------------
bool conv_halo_eligible(const phs_tensor* x, const phs_tensor* y, int ksize);
int chan_stats_run(const phs_tensor* y, double* stats, bool with_totals, bool zero_first, cudaStream_t st);
int conv2d_halo(const phs_tensor* x, const void* w, const float* bias, const phs_tensor* y, int accumulate, double* stats,
                cudaStream_t st);
int conv2d_halo_pre(const phs_tensor* x, const phs_norm_pre* pre, const void* w, const float* bias, const phs_tensor* y,
                    int accumulate_flags, double* stats, cudaStream_t st);
bool wgrad_halo_eligible(const phs_tensor* x, const phs_tensor* dy, int ksize);
int conv2d_wgrad_halo(const phs_tensor* x, const phs_tensor* dy, float* dw, cudaStream_t st);

constexpr int STATS_MIN_HW_DEFAULT = 2048;   // pixels per image below which phs_conv2d_stats_acc does not fuse the statistics
// (in-step A/B, profiles/step_ab_stats_split_r02.txt: 0 -> 12.27-12.30 ms, 512 -> 12.25, 2048 -> 12.20, 8192 -> 12.19)

static int stats_min_hw() {
  const char* e = getenv("PHS_STATS_MIN_HW");      // read per call: tools/step_ab.py switches it inside one process
  return e ? atoi(e) : STATS_MIN_HW_DEFAULT;
}

// phs_conv2d_pre (include/phiseg_sm100.h): 3x3 forward convolution of act(norm(yprev)), the normalisation applied to the
// operand tile in shared memory; stats (optional) follow the phs_conv2d_stats_acc contract, including the separate
// statistics pass for few-tile layers
extern "C" int phs_conv2d_pre(const phs_tensor* x, const phs_norm_pre* pre, const void* w, const float* bias,
                              const phs_tensor* y, double* stats, void* stream) {
  PHS_REQUIRE(x && pre && w && y && x->ptr && y->ptr, "phs_conv2d_pre: null argument");
  PHS_REQUIRE(x->N == y->N && x->H == y->H && x->W == y->W, "phs_conv2d_pre: SAME stride-1 needs equal N,H,W");
  PHS_REQUIRE(pre->gamma && pre->beta, "phs_conv2d_pre: gamma / beta required");
  PHS_REQUIRE(pre->mode == PHS_NORM_BN_INFER ? (pre->moving_mean && pre->moving_var) : pre->stats != nullptr,
              "phs_conv2d_pre: mode %d needs %s", pre->mode, pre->mode == PHS_NORM_BN_INFER ? "moving statistics" : "stats");
  PHS_REQUIRE(pre->mode == PHS_NORM_BN_TRAIN || pre->mode == PHS_NORM_BN_INFER || pre->mode == PHS_NORM_GN,
              "phs_conv2d_pre: unknown mode %d", pre->mode);
  PHS_REQUIRE(!pre->mean == !pre->rstd, "phs_conv2d_pre: mean and rstd go together");
  if (x->dtype != PHS_BF16 || y->dtype != PHS_BF16 || !conv_halo_eligible(x, y, 3)) return -3;
  cudaStream_t st = (cudaStream_t)stream;
  const bool split_stats = stats && y->H * y->W < stats_min_hw();
  int rc = conv2d_halo_pre(x, pre, w, bias, y, 2, split_stats ? nullptr : stats, st);
  if (rc == 0 && split_stats) return chan_stats_run(y, stats, true, false, st);
  return rc;
}

int conv2d_halo_post(const phs_tensor* x, const void* w, const float* bias, const phs_norm_pre* post, const phs_tensor* y,
                     cudaStream_t st);
static int conv2d_tc_impl(const phs_tensor* x, const void* w, const float* bias, const phs_tensor* y, int ksize, int dgrad,
                          int accumulate, double* stats, cudaStream_t st, const phs_norm_pre* post);

int conv2d_tc(const phs_tensor* x, const void* w, const float* bias, const phs_tensor* y, int ksize, int dgrad,
              int accumulate, double* stats, cudaStream_t st) {
  return conv2d_tc_impl(x, w, bias, y, ksize, dgrad, accumulate, stats, st, nullptr);
}

// phs_conv2d_post (include/phiseg_sm100.h): forward convolution + inference-mode batch norm + ReLU in one launch, any
// tensor-core shape (halo kernel or shifted-box kernel)
extern "C" int phs_conv2d_post(const phs_tensor* x, const void* w, const float* bias, const phs_norm_pre* post,
                               const phs_tensor* a, int ksize, void* stream) {
  PHS_REQUIRE(x && w && post && a && x->ptr && a->ptr, "phs_conv2d_post: null argument");
  PHS_REQUIRE(ksize == 1 || ksize == 3, "phs_conv2d_post: ksize=%d", ksize);
  PHS_REQUIRE(x->N == a->N && x->H == a->H && x->W == a->W, "phs_conv2d_post: SAME stride-1 needs equal N,H,W");
  PHS_REQUIRE(post->mode == PHS_NORM_BN_INFER && post->gamma && post->beta && post->moving_mean && post->moving_var,
              "phs_conv2d_post: inference-mode batch norm with gamma / beta / moving statistics required");
  return conv2d_tc_impl(x, w, bias, a, ksize, 0, 0, nullptr, (cudaStream_t)stream, post);
}

static int conv2d_tc_impl(const phs_tensor* x, const void* w, const float* bias, const phs_tensor* y, int ksize, int dgrad,
                          int accumulate, double* stats, cudaStream_t st, const phs_norm_pre* post) {
  // the gradient w.r.t. the input is the same GEMM on the dgrad filter shadow
  PHS_REQUIRE(x->dtype == PHS_BF16, "conv2d_tc: input must be bf16");
  const int stats_prezeroed = accumulate & 2;   // bit 1: the caller zeroed stats (phs_conv2d_stats_acc)
  accumulate &= 1;
  PHS_REQUIRE(x->C % 32 == 0 && y->C % 16 == 0 && y->C >= 16,
              "conv2d_tc: Cin=%d must be a multiple of 32 and Cout=%d a multiple of 16", x->C, y->C);
  if (y->C > 256) {
    // one UMMA covers at most 256 output channels: split the GEMM's N (e.g. the 384-channel input gradient of
    // likelihood/post_c_3_1) into equal parts, each a channel slice of y and a row block of the filter shadow
    const int nparts = (y->C + 255) / 256;
    const int part = ((y->C + nparts - 1) / nparts + 15) / 16 * 16;
    const int es = y->dtype == PHS_F32 ? 4 : 2;
    for (int c0 = 0; c0 < y->C; c0 += part) {
      phs_tensor ys = *y;
      ys.ptr = (char*)y->ptr + (size_t)c0 * es;
      ys.C = y->C - c0 < part ? y->C - c0 : part;
      const bf16* ws = (const bf16*)w + (size_t)c0 * ksize * ksize * x->C;
      PHS_REQUIRE(stats == nullptr, "conv2d_tc: fused statistics need Cout <= 256");
      phs_norm_pre ps;
      if (post) {
        ps = *post;
        ps.gamma += c0; ps.beta += c0; ps.moving_mean += c0; ps.moving_var += c0;
      }
      int rc = conv2d_tc_impl(x, ws, bias ? bias + c0 : nullptr, &ys, ksize, dgrad, accumulate, nullptr, st, post ? &ps : nullptr);
      if (rc) return rc;
    }
    return 0;
  }
  PHS_REQUIRE(x->ld % 8 == 0 && aligned16(x->ptr) && aligned16(w), "conv2d_tc: input / filter not 16-byte aligned");
  const int yes = y->dtype == PHS_F32 ? 4 : 2;
  PHS_REQUIRE(((size_t)y->ld * yes) % 16 == 0 && aligned16(y->ptr), "conv2d_tc: output not 16-byte aligned");
  if (conv_halo_eligible(x, y, ksize) && !getenv("PHS_NO_HALO")) {
    // Few-tile layers (a CTA holds one or two tiles): the fused statistics epilogue cannot hide behind the next tile's
    // MMAs and is pure tail (16x16x192: 47 us with, 25 us without), while one pass over the few-MB output costs ~5 us:
    // below PHS_STATS_MIN_HW pixels per image the statistics come from a separate launch (the layout is the same).
    const bool split_stats = stats && stats_prezeroed && y->H * y->W < stats_min_hw();
    // (few-tile forward layers whose statistics come from the separate pass keep the single-CTA kernel, PHS_SPLIT_PAIR=1
    // lets them use CTA pairs: with one or two tiles per CTA pairs bought ~0.03 ms per step, and the stalls at the end of
    // round 2 appeared only after this - the newest - use of cluster launches was added)
    static const bool split_pair = getenv("PHS_SPLIT_PAIR") != nullptr;
    const int no_pair = (split_stats && !split_pair) ? 4 : 0;
    int rc = post ? conv2d_halo_post(x, w, bias, post, y, st)
                  : conv2d_halo(x, w, bias, y, accumulate | stats_prezeroed | no_pair, split_stats ? nullptr : stats, st);
    if (rc == 0 && split_stats) return chan_stats_run(y, stats, true, false, st);
    if (rc != -3) return rc;
  }
  if (stats) {
    // shapes the halo kernel does not take: plain convolution, then a separate statistics pass over y
    int rc = conv2d_tc(x, w, bias, y, ksize, dgrad, accumulate, nullptr, st);
    if (rc) return rc;
    return chan_stats_run(y, stats, stats_prezeroed != 0, !stats_prezeroed, st);
  }
  const int BK = x->C % 64 == 0 ? 64 : 32;
  const int taps = ksize * ksize;
  Brick b = make_brick(x->N, x->H, x->W, 128);
  CUtensorMap tmA, tmB;
  int rc = activation_map(x, BK, b.TW, b.TH, b.TN, &tmA);
  if (rc) return rc;
  rc = filter_map(w, taps * x->C, y->C, BK, &tmB);
  if (rc) return rc;
  ConvParams p;
  p.N = x->N; p.H = x->H; p.W = x->W; p.Cin = x->C; p.Cout = y->C;
  p.taps = taps;
  p.TW = b.TW; p.TH = b.TH; p.TN = b.TN; p.tilesW = b.tilesW; p.tilesH = b.tilesH; p.num_tiles = b.num;
  p.kchunks = x->C / BK;
  const int stage_bytes = 128 * BK * 2 + y->C * BK * 2;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  p.stages = stages;
  p.y = y->ptr; p.y_ld = y->ld; p.y_f32 = y->dtype == PHS_F32;
  p.bias = bias;
  p.accumulate = accumulate;
  p.post_on = post != nullptr;
  if (post) p.post = *post;
  else memset(&p.post, 0, sizeof(p.post));
  const int smem = stages * stage_bytes + 1024;
  const int grid = b.num < num_sms() ? b.num : num_sms();
  if (BK == 64) {
    static bool attr = false;
    if ((rc = allow_big_smem(conv_tc_kernel<64>, &attr))) return rc;
    phs_launch_tc(conv_tc_kernel<64>, grid, 192, smem, st, tmA, tmB, p);
  } else {
    static bool attr = false;
    if ((rc = allow_big_smem(conv_tc_kernel<32>, &attr))) return rc;
    phs_launch_tc(conv_tc_kernel<32>, grid, 192, smem, st, tmA, tmB, p);
  }
  return phs_check_launch("conv_tc_kernel");
}

static int bias_grad_tc(const phs_tensor* dy, float* db, cudaStream_t st) {
  const int64_t M = (int64_t)dy->N * dy->H * dy->W;
  int blocks = (int)(ceil_div64(M, 64) < 296 ? ceil_div64(M, 64) : 296);
  bias_grad_bf16_kernel<<<dim3(blocks, (dy->C + 31) / 32), 256, 0, st>>>((const bf16*)dy->ptr, dy->ld, dy->C, M, db);
  return phs_check_launch("bias_grad_bf16");
}

int conv2d_wgrad_tc(const phs_tensor* x, const phs_tensor* dy, float* dw, float* db, int ksize, int accumulate,
                    cudaStream_t st) {
  PHS_REQUIRE(x->dtype == PHS_BF16 && dy->dtype == PHS_BF16, "conv2d_wgrad_tc: x and dy must be bf16");
  PHS_REQUIRE(x->C % 32 == 0 && dy->C % 32 == 0 && dy->C <= 256,
              "conv2d_wgrad_tc: Cin=%d and Cout=%d must be multiples of 32, Cout <= 256", x->C, dy->C);
  PHS_REQUIRE(x->ld % 8 == 0 && dy->ld % 8 == 0 && aligned16(x->ptr) && aligned16(dy->ptr) && aligned16(dw),
              "conv2d_wgrad_tc: operands not 16-byte aligned");
  const int taps = ksize * ksize;
  const size_t nw = (size_t)taps * x->C * dy->C;
  if (!accumulate) {
    cudaMemsetAsync(dw, 0, nw * sizeof(float), st);
    if (db) cudaMemsetAsync(db, 0, dy->C * sizeof(float), st);
  }
  if (wgrad_halo_eligible(x, dy, ksize) && !getenv("PHS_NO_HALO")) {
    int rc = conv2d_wgrad_halo(x, dy, dw, st);
    if (rc) return rc;
    return db ? bias_grad_tc(dy, db, st) : 0;
  }
  WgradParams p;
  p.N = x->N; p.H = x->H; p.W = x->W; p.Cin = x->C; p.Cout = dy->C;
  p.taps = taps;
  Brick b = make_brick(x->N, x->H, x->W, KT);
  p.TW = b.TW; p.TH = b.TH; p.TN = b.TN; p.tilesW = b.tilesW; p.tilesH = b.tilesH; p.num_ptiles = b.num;
  p.mblocks = (x->C + 127) / 128;
  p.slabA = x->C % 64 == 0 ? 64 : 32;
  p.nslabA = 128 / p.slabA;
  p.slabB = dy->C % 64 == 0 ? 64 : 32;
  p.nslabB = dy->C / p.slabB;
  p.dw = dw;
  p.tmem_cols = dy->C <= 32 ? 32 : dy->C <= 64 ? 64 : dy->C <= 128 ? 128 : 256;
  const int stage_bytes = KT * 128 * 2 + KT * dy->C * 2;
  int stages = (160 * 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  p.stages = stages;
  const int items = taps * p.mblocks;
  // ~160 KB of shared memory per CTA: one CTA per SM, so the grid is one wave of at most num_sms CTAs
  int splits = num_sms() / items;
  if (splits > b.num) splits = b.num;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (b.num + splits - 1) / splits;
  splits = (b.num + p.tiles_per_split - 1) / p.tiles_per_split;
  CUtensorMap tmX, tmDY;
  int rc = activation_map(x, p.slabA, b.TW, b.TH, b.TN, &tmX);
  if (rc) return rc;
  rc = activation_map(dy, p.slabB, b.TW, b.TH, b.TN, &tmDY);
  if (rc) return rc;
  static bool attr = false;
  if ((rc = allow_big_smem(wgrad_tc_kernel, &attr))) return rc;
  const int smem = stages * stage_bytes + 1024;
  phs_launch_tc(wgrad_tc_kernel, dim3(splits, items), 192, smem, st, tmX, tmDY, p);
  rc = phs_check_launch("wgrad_tc_kernel");
  if (rc) return rc;
  return db ? bias_grad_tc(dy, db, st) : 0;
}
