// 3x3 SAME convolution forward / input gradient on tcgen05, "halo" formulation (tf.nn.conv2d, tfwrapper/layers.py:123,
// and Conv2DBackpropInput).  Same GEMM as conv_tc.cu,
//
//   D[pixel][cout] = sum_{tap, ci} X[pixel + tap][ci] * Wt[cout][tap*Cin + ci],
//
// but the activation operand is loaded ONCE per 64-channel chunk instead of once per tap: a CTA owns a super-tile of
// 16 rows x 8*S columns of one image (S sub-tiles of 128 pixels, one TMEM accumulator each), loads the
// 18 x (8*S+2) halo tile with one TMA box (out-of-bounds rows/columns zero-filled = SAME padding) and issues the MMAs of
// all 9 taps and S sub-tiles from it through SHIFTED K-major descriptors: tap (kh,kw), sub-tile s start at halo pixel
// (kh, 8*s + kw); the 8 pixels of one image row are contiguous 128-byte rows, the next image row is SBO = (8*S+2) rows
// further.  tcgen05.mma applies the swizzle to absolute shared-memory address bits (tools/umma_probe.cu), so these
// unaligned starts read the TMA-written tile correctly.  The TMA unit's cost is per 128-byte row: this cuts the
// activation rows per MMA by ~6x and shares every filter tile between S sub-tiles.
//
//   warp 4: TMA producer (A ring: halo tiles; B ring: one [Cout][64] filter tile per (chunk, tap); when the whole
//           filter fits it is loaded once and stays resident)
//   warp 5: MMA issuer, accumulators double-buffered in TMEM (2 x S x Cout <= 512 columns); highest warp index = first
//           pick of the issue arbiter (tc_ptx.cuh)
//   warps 0-3: epilogue: tcgen05.ld -> +bias -> bf16/f32 store, and the per-(sample, channel) sum / sum of squares the
//           following batch_norm / group_norm2D needs (tfwrapper/normalisation.py:27-34,156) from the fp32 accumulators:
//           32-lane butterfly transpose-reduce per 16 columns, accumulated in registers while the CTA stays inside one
//           image (CTAs own contiguous tile ranges), then one red.global.add per (channel, quantity).
//   warps 6-7 (PRE only): operand transform.  The activation operand is the RAW output of the previous convolution; these
//           warps apply its batch_norm / group_norm2D + ReLU (tfwrapper/layers.py:123-135) to every landed halo tile in
//           place, between the TMA load and the MMAs, so the normalised activation never exists in HBM.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_host.cuh"

using namespace tc;

namespace {

constexpr int HB_MAX_A = 6, HB_MAX_B = 40;
constexpr int TILE_H = 16, SUB_W = 8;

struct HaloParams {
  int N, H, W, Cin, Cout;
  int S;                      // sub-tiles per super-tile
  int tilesW, tilesH, num_tiles;
  int kchunks;
  int na, nb, b_resident;
  int acc_stages, acc_cols, tmem_cols;   // TMEM: acc_stages accumulator sets of acc_cols columns
  uint32_t a_stage_bytes;     // rounded up to 1024
  void* y;
  int y_ld, y_f32;
  const float* bias;
  int accumulate;
  double* stats;              // [N][Cout][2] or null (fp64: see flush())
  double* totals;             // [Cout][2] batch totals (phs_conv2d_stats_acc layout) or null
  int stage_g;                // 0: direct register -> global stores; 32 | 64: channels per smem-staged TMA store group
  long long* trace;           // PHS_HALO_TRACE: per-role clock64 stamps of the first CTAs (tools/trace_halo.py)
  int dbg;                    // profiling switches (PHS_HALO_DBG): 1 = no TMA, 2 = no MMA, 4 = no epilogue stores
  phs_norm_pre pre;           // PRE: normalisation of the producer layer, applied to the activation operand
                              // POST: THIS layer's inference-mode batch norm + ReLU, applied in the epilogue
  uint32_t pre_tab;           // PRE / POST: byte offset (from the aligned dynamic shared memory base) of the (scale, shift) table
};

// mean / rstd of channel c of sample n from the statistics the producer's epilogue left behind (phs_conv2d_stats_acc
// layout: per-sample sums [N][C][2], then the batch totals [C][2]) - the expressions of norm_act_fwd_stats_kernel
__device__ __forceinline__ void pre_moments(const phs_norm_pre& pre, int N, int HW, int C, int n, int c, double* m,
                                            double* var, double* r, double* cnt) {
  if (pre.mode == PHS_NORM_GN) {
    const int G = max(2, C / 16), cpg = C / G;
    const double* gs = pre.stats + ((size_t)n * C + (size_t)(c / cpg) * cpg) * 2;
    double s = 0.0, q = 0.0;
    for (int i = 0; i < cpg; ++i) { s += gs[2 * i]; q += gs[2 * i + 1]; }
    *cnt = (double)HW * cpg;
    norm_moments(s, q, *cnt, pre.eps, m, var, r);
  } else {
    const double* totals = pre.stats + (size_t)N * C * 2;
    *cnt = (double)HW * N;
    norm_moments(totals[2 * c], totals[2 * c + 1], *cnt, pre.eps, m, var, r);
  }
}

__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t* w) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ld_shared_v4f(uint32_t addr, float* w) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w[0]), "=f"(w[1]), "=f"(w[2]), "=f"(w[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void st_shared_v2f(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}

template <typename T>
__device__ __forceinline__ void store16(T* dst, const float* v, bool accumulate) {
  float o[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = v[i];
  if (accumulate) {
    float p[16];
    ldv<T, 8>(dst, p);
    ldv<T, 8>(dst + 8, p + 8);
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] += p[i];
  }
  stv<T, 8>(dst, o);
  stv<T, 8>(dst + 8, o + 8);
}

// sum over the 32 lanes of each of 16 per-lane values; lane L ends up with the total of column col16(L)
__device__ __forceinline__ int col16(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}
__device__ __forceinline__ float transpose_reduce16(const float* v, int lane) {
  float a8[8], a4[4], a2[2];
  bool up = lane & 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float send = up ? v[i] : v[i + 8];
    float keep = up ? v[i + 8] : v[i];
    a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  up = lane & 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float send = up ? a8[i] : a8[i + 4];
    float keep = up ? a8[i + 4] : a8[i];
    a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  up = lane & 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float send = up ? a4[i] : a4[i + 2];
    float keep = up ? a4[i + 2] : a4[i];
    a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  up = lane & 2;
  float send = up ? a2[0] : a2[1];
  float keep = up ? a2[1] : a2[0];
  float a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  return a1 + __shfl_xor_sync(0xffffffffu, a1, 1);
}

// 192 threads are launched; the bound of 256 caps the kernel at 128 registers (no spills; ptxas takes 168 when allowed),
// which leaves 16 K registers of the SM to the register-only kernels of the other lanes.  (Measured: step time unchanged
// at 128 and at 96 registers - kernels of different lanes already interleave at CTA-retire granularity.)
// PAIR: the CTA is one half of a cta_group::2 pair (cluster of 2 on one TPC).  Both CTAs own a tile of their own (halo
// tile, accumulators, epilogue), but every [Cout][BK] filter tile is split between them - each stages Cout/2 rows - and
// the leader (cluster rank 0) issues ONE M=256 MMA per (tap, k) that reads both halves.  Per SM that halves the filter
// bytes streamed from L2 and the shared memory a filter stage takes: the wide layers were bound by exactly that ring
// (3 stages of 16 KB per CTA, one stage per ~640 clk against 256 clk of MMA work).
// SW: four extra "statistics warps" (warps 4-7; the producer and the MMA issuer move to warps 8 and 9).  They read the same
// accumulators as the epilogue warps (warp w and warp w+4 share TMEM lane quadrant w) and do nothing but the per-channel
// sum / sum-of-squares reduction, which costs twice the instructions of the whole store path: with both jobs on the same
// four warps every layer with fused statistics was epilogue-bound (+11 % at 128->128, +50 % at 32->192).
// PRE: two transform warps (warps 6-7, 256 threads) normalise + ReLU every landed halo tile in place before the MMA
// issuer may read it (a_full -> transform -> a_ready).  In-bounds pixels only: the TMA unit zero-filled the out-of-image
// halo, and SAME padding pads the ACTIVATION, so those rows must stay zero.  Not combined with PAIR (the peer's tile
// arrival is signalled on the leader's barrier only) or SW.
// POST: inference-mode batch norm (moving statistics: known before the convolution runs) + ReLU of THIS layer folded into
// the epilogue - a = act(gamma * (conv + bias - moving_mean) * rsqrt(moving_var + eps) + beta) leaves the kernel directly,
// the raw convolution output is never written (sampling / validation programs: phs_norm_finalize + phs_norm_act_fwd and
// two tensor passes per layer disappear).  The affine map is applied to the fp32 accumulator, i.e. one bf16 rounding
// fewer than the two-launch path.
template <int BK, bool PAIR, bool SW, bool PRE, bool POST = false>
__global__ void __launch_bounds__(SW ? 320 : 256, 2)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmY, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  static_assert(!PRE || (!PAIR && !SW), "operand transform: single-CTA kernel without statistics warps");
  static_assert(!POST || (!PRE && !SW), "epilogue normalisation: inference only (no statistics, no operand transform)");
  __shared__ __align__(8) uint64_t bars[2 * HB_MAX_A + 2 * HB_MAX_B + 4 + (PRE ? HB_MAX_A : 0)];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_PROD = SW ? 8 : tc::W_PROD, W_MMA = SW ? 9 : tc::W_MMA;     // (shadow the 192-thread role table)
  // optional timeline of the first 8 CTAs: trace[(cta * 3 + role) * 256 + k], role 0 producer / 1 MMA / 2 epilogue warp 2
  const int trole = warp == W_PROD ? 0 : (warp == W_MMA ? 1 : (warp == 0 ? 2 : -1));
  long long* trc = (p.trace && blockIdx.x < 8 && lane == 0 && trole >= 0) ? p.trace + (blockIdx.x * 3 + trole) * 256 : nullptr;
  int tri = 0;
  auto stamp = [&]() {
    if (trc && tri < 256) trc[tri++] = clock64();
  };
  stamp();
  constexpr uint32_t ROW = BK * 2;
  const int HW_ = SUB_W * p.S + 2;                 // halo width in pixels
  const uint32_t a_bytes = (uint32_t)(TILE_H + 2) * HW_ * ROW;
  const uint32_t b_bytes = (uint32_t)(PAIR ? p.Cout / 2 : p.Cout) * ROW;   // this CTA's part of a filter tile
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_b0 = smem0 + p.na * p.a_stage_bytes;
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (HB_MAX_A + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * HB_MAX_A + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (2 * HB_MAX_A + HB_MAX_B + s); };
  auto tfull = [&](int a) { return bar0 + 8u * (2 * HB_MAX_A + 2 * HB_MAX_B + a); };
  auto tempty = [&](int a) { return bar0 + 8u * (2 * HB_MAX_A + 2 * HB_MAX_B + 2 + a); };
  auto a_ready = [&](int s) { return bar0 + 8u * (2 * HB_MAX_A + 2 * HB_MAX_B + 4 + s); };   // PRE: tile s is transformed

  if (warp == W_PROD && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.stage_g) tma_prefetch_desc(&tmY);
    for (int s = 0; s < p.na; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
      if (PRE) mbar_init(a_ready(s), 2);      // one arrival per transform warp
    }
    for (int s = 0; s < p.nb; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull(a), 1);
      mbar_init(tempty(a), (SW ? 8 : 4) * (PAIR ? 2 : 1));   // every warp that reads the accumulators, of both CTAs of a pair
    }
    fence_barrier_init();
  }
  if (warp == W_MMA) {
    if (PAIR) tmem_alloc_pair(smem_u32(&tmem_base_s), p.tmem_cols);
    else tmem_alloc(smem_u32(&tmem_base_s), p.tmem_cols);
  }
  PHS_PDL_WAIT();     // everything above touches only shared / tensor memory and kernel parameters
  for (int c = threadIdx.x; c < 256; c += blockDim.x) bias_s[c] = (p.bias && c < p.Cout) ? p.bias[c] : 0.f;
  if (POST) {
    // (scale, shift) of every output channel: the expressions of norm_finalize_kernel (inference) + norm_act_fwd_kernel
    for (int c = threadIdx.x; c < p.Cout; c += blockDim.x) {
      float sc, sh;
      norm_scale_shift(p.pre.gamma[c], p.pre.beta[c], p.pre.moving_mean[c], rsqrtf(p.pre.moving_var[c] + p.pre.eps), &sc, &sh);
      st_shared_v2f(smem0 + p.pre_tab + (uint32_t)c * 8u, sc, sh);
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();     // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  PHS_PDL_TRIGGER();      // this CTA holds its tensor memory: the successor kernel may be scheduled now
  stamp();

  // contiguous tile range of this CTA (pair: of the pair, which walks it two tiles at a time - rank r takes tiles
  // t_begin + 2i + r; an odd range leaves the last tile of rank 1 empty: it loads zero-filled boxes and stores nothing)
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, nunits = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int t_begin = (int)((int64_t)p.num_tiles * unit / nunits);
  const int t_end = (int)((int64_t)p.num_tiles * (unit + 1) / nunits);
  const int iters = PAIR ? (t_end - t_begin + 1) >> 1 : t_end - t_begin;
  const int tstep = PAIR ? 2 : 1, tfirst = t_begin + (int)rank;
  const int tiles_per_img = p.tilesW * p.tilesH;
  // mbarriers of the pair's leader as seen from this CTA (the leader's own when this is the leader)
  auto lead = [&](uint32_t bar) { return PAIR ? mapa_u32(bar, 0) : bar; };

  // The two issue loops below run on one warp each and every instruction in them is on the critical path of the
  // tensor pipe (the UTCHMMA instructions themselves never stall): ring indices and phase bits are counters (no
  // runtime division), the tap loop is unrolled so tap offsets are immediates, kernel parameters sit in registers.
  const int S = p.S, kchunks = p.kchunks, na = p.na, nb = p.nb, resident = p.b_resident, Cout = p.Cout, Cin = p.Cin;
  const int tilesW = p.tilesW, acc_stages = p.acc_stages, dbg = p.dbg;
  const uint32_t a_stage_bytes = p.a_stage_bytes, acc_cols = p.acc_cols;
  if (warp == W_PROD) {
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    for (int it = 0, tile = tfirst; it < iters; ++it, tile += tstep) {
      // (an empty tile of rank 1 asks for image N: every box row is out of bounds and arrives zero-filled)
      const int n = tile < t_end ? tile / tiles_per_img : p.N;
      const int r = tile < t_end ? tile - n * tiles_per_img : 0;
      const int th = r / tilesW;
      const int h0 = th * TILE_H, w0 = (r - th * tilesW) * SUB_W * S;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(a_empty(sa), pha ^ 1);
        if (elect_one()) {
          if (dbg & 1) {
            if (!PAIR || rank == 0) mbar_arrive(a_full(sa));
          } else if (PAIR) {
            // the leader expects both CTAs' bytes on its barrier; the peer's load signals that barrier too
            if (rank == 0) mbar_expect_tx(a_full(sa), 2 * a_bytes);
            tma_load_4d_pair(smem0 + sa * a_stage_bytes, &tmA, lead(a_full(sa)), kc * BK, w0 - 1, h0 - 1, n);
          } else {
            mbar_expect_tx(a_full(sa), a_bytes);
            tma_load_4d(smem0 + sa * a_stage_bytes, &tmA, a_full(sa), kc * BK, w0 - 1, h0 - 1, n);
          }
        }
        __syncwarp();
        stamp();
        if (++sa == na) { sa = 0; pha ^= 1; }
        if (resident && it != 0) continue;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int sbi = resident ? kc * 9 + tap : sb;
          if (!resident) mbar_wait(b_empty(sbi), phb ^ 1);
          if (elect_one()) {
            if (dbg & 1) {
              if (!PAIR || rank == 0) mbar_arrive(b_full(sbi));
            } else if (PAIR) {
              if (rank == 0) mbar_expect_tx(b_full(sbi), 2 * b_bytes);
              tma_load_2d_pair(smem_b0 + sbi * b_bytes, &tmB, lead(b_full(sbi)), tap * Cin + kc * BK, (int)rank * (Cout / 2));
            } else {
              mbar_expect_tx(b_full(sbi), b_bytes);
              tma_load_2d(smem_b0 + sbi * b_bytes, &tmB, b_full(sbi), tap * Cin + kc * BK, 0);
            }
          }
          __syncwarp();
          if (!resident && ++sb == nb) { sb = 0; phb ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA && (!PAIR || rank == 0)) {
    const uint32_t idesc = idesc_bf16(PAIR ? 256 : 128, Cout, 0, 0);
    constexpr uint64_t LAYOUT = BK == 64 ? LAYOUT_SW128 : LAYOUT_SW64;
    const uint32_t a_hi = desc_hi(HW_ * ROW, LAYOUT);   // SBO: next image row of the sub-tile inside the halo tile
    const uint32_t b_hi = desc_hi(8 * ROW, LAYOUT);
    const uint32_t hw16 = (HW_ * ROW) >> 4;             // one halo row, in descriptor (16-byte) units
    const uint32_t b_base_lo = desc_lo(smem_b0, 16), b_step = b_bytes >> 4;
    int sa = 0, sb = 0, acc = 0;
    uint32_t pha = 0, phb = 0, aph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(tempty(acc), aph ^ 1);
      tc_fence_after();
      stamp();
      const uint32_t d0 = tmem_base + acc * acc_cols;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(PRE ? a_ready(sa) : a_full(sa), pha);
        tc_fence_after();
        if (kc == 0) stamp();
        const uint32_t a_base_lo = desc_lo(smem0 + sa * a_stage_bytes, 16);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int sbi = resident ? kc * 9 + tap : sb;
          mbar_wait(b_full(sbi), resident ? 0u : phb);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_lo = b_base_lo + sbi * b_step;
            uint32_t a_lo = a_base_lo + (tap / 3) * hw16 + (tap % 3) * (ROW >> 4);
            uint32_t d = d0;
            const uint32_t first = tap != 0 ? 1u : (kc != 0 ? 1u : 0u);
            for (int sub = 0; sub < S; ++sub, a_lo += (SUB_W * ROW) >> 4, d += Cout) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                if (!(dbg & 2)) {
                  if (PAIR) umma_bf16_lohi_pair(d, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, k ? 1u : first);
                  else umma_bf16_lohi(d, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, k ? 1u : first);
                }
            }
            // pair: the commits arrive on the barrier at the same offset in BOTH CTAs (each producer / epilogue waits locally)
            if (!resident) { if (PAIR) umma_commit_pair(b_empty(sbi)); else umma_commit(b_empty(sbi)); }
            if (tap == 8) {
              if (PAIR) umma_commit_pair(a_empty(sa)); else umma_commit(a_empty(sa));
              if (kc == kchunks - 1) { if (PAIR) umma_commit_pair(tfull(acc)); else umma_commit(tfull(acc)); }
            }
          }
          __syncwarp();
          if (!resident && ++sb == nb) { sb = 0; phb ^= 1; }
        }
        if (++sa == na) { sa = 0; pha ^= 1; }
      }
      stamp();
      if (++acc == acc_stages) { acc = 0; aph ^= 1; }
    }
  } else if (PRE && warp >= 6) {
    // ---- operand transform: a = act(gamma * (y - mean) * rstd + beta) on the landed halo tile, in place ----
    const int tt = (int)threadIdx.x - 192;            // 0..63
    constexpr int CPR = BK / 8;                       // 16-byte chunks (8 channels) per pixel row
    constexpr int RPI = 64 / CPR;                     // pixel rows per pass of the 64 threads
    constexpr int U = 4;                              // rows in flight per thread
    const int j = tt % CPR, r0 = tt / CPR;            // this thread's (logical) chunk column and first row
    const uint32_t tab = smem0 + p.pre_tab;           // (scale, shift) per input channel of the current sample
    const int rows = (TILE_H + 2) * HW_;
    const uint32_t inv_hw = ((1u << 20) + HW_ - 1) / HW_;   // row / HW_ = (row * inv_hw) >> 20 for row < 2^10 * ... (rows <= 1188)
    const int HWpix = p.H * p.W;
    int sa = 0, cur_n = -1;
    uint32_t pha = 0;
    for (int it = 0, tile = tfirst; it < iters; ++it, tile += tstep) {
      const int n = tile / tiles_per_img;
      const int r = tile - n * tiles_per_img;
      const int th = r / tilesW;
      const int h0 = th * TILE_H - 1, w0 = (r - th * tilesW) * SUB_W * S - 1;      // image coordinates of halo pixel (0, 0)
      if (n != cur_n) {
        // coefficients of sample n (batch norm: the same for every sample, but the owner of a sample's first tile also
        // publishes mean / rstd of that sample for the backward kernels, and the owner of tile 0 updates the moving averages)
        named_bar_sync(2, 64);                        // nobody still reads the previous table
        const bool owner = r == 0;
        for (int c = tt; c < Cin; c += 64) {
          float mf, rf;
          if (p.pre.mode == PHS_NORM_BN_INFER) {
            mf = p.pre.moving_mean[c];
            rf = rsqrtf(p.pre.moving_var[c] + p.pre.eps);
          } else {
            double m, var, rr, cnt;
            pre_moments(p.pre, p.N, HWpix, Cin, n, c, &m, &var, &rr, &cnt);
            mf = (float)m;
            rf = (float)rr;
            if (p.pre.mode == PHS_NORM_BN_TRAIN && p.pre.moving_mean && tile == 0)
              bn_moving_update(p.pre.moving_mean, p.pre.moving_var, c, p.pre.decay, m, var, cnt);
          }
          if (owner && p.pre.mean) {
            p.pre.mean[(size_t)n * Cin + c] = mf;
            p.pre.rstd[(size_t)n * Cin + c] = rf;
          }
          float sc, sh;
          norm_scale_shift(p.pre.gamma[c], p.pre.beta[c], mf, rf, &sc, &sh);
          st_shared_v2f(tab + (uint32_t)c * 8u, sc, sh);
        }
        named_bar_sync(2, 64);
        cur_n = n;
      }
      for (int kc = 0; kc < kchunks; ++kc) {
        float cf[16];                                 // (scale, shift) of this thread's 8 channels of the chunk
#pragma unroll
        for (int i = 0; i < 4; ++i) ld_shared_v4f(tab + (uint32_t)(kc * BK + j * 8) * 8u + 16u * i, cf + 4 * i);
        mbar_wait(a_full(sa), pha);
        const uint32_t base = smem0 + sa * a_stage_bytes;
        for (int row = r0; row < rows; row += U * RPI) {
          uint32_t addr[U], v[U][4];
          bool ok[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int rw = row + u * RPI;
            const int hr = (int)(((uint32_t)rw * inv_hw) >> 20);
            const int wc = rw - hr * HW_;
            ok[u] = rw < rows && (unsigned)(h0 + hr) < (unsigned)p.H && (unsigned)(w0 + wc) < (unsigned)p.W;
            // TMA swizzle (absolute address bits, stage bases are 1024-aligned): 128-byte rows XOR the chunk index with
            // (row & 7), 64-byte rows with ((row >> 1) & 3)
            const uint32_t x = BK == 64 ? (uint32_t)(rw & 7) : (uint32_t)((rw >> 1) & 3);
            addr[u] = base + (uint32_t)rw * ROW + ((((uint32_t)j) ^ x) << 4);
            if (ok[u]) ld_shared_v4(addr[u], v[u]);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (ok[u]) {
              uint32_t o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float lo = norm_act1(__uint_as_float(v[u][i] << 16), cf[4 * i], cf[4 * i + 1], p.pre.relu);
                const float hi = norm_act1(__uint_as_float(v[u][i] & 0xffff0000u), cf[4 * i + 2], cf[4 * i + 3], p.pre.relu);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
                o[i] = *reinterpret_cast<uint32_t*>(&h2);
              }
              st_shared_v4(addr[u], o[0], o[1], o[2], o[3]);
            }
          }
        }
        fence_proxy_async();        // the generic-proxy writes above are visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready(sa));
        if (++sa == na) { sa = 0; pha ^= 1; }
      }
    }
  } else if (SW && warp >= 4 && warp < 8) {
    // ---- statistics warps: per-(sample, channel) sum and sum of squares of the fp32 accumulators (+ bias) ----
    const int q = warp & 3;
    uint32_t tcount = 0;
    float st_s[16], st_q[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) st_s[j] = st_q[j] = 0.f;
    int st_n = -1;
    auto flush = [&]() {
      if (st_n >= 0) {
        double* dst = p.stats + (size_t)st_n * p.Cout * 2;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j * 16 < p.Cout) {
            const int c = j * 16 + col16(lane);
            const double v = (double)((lane & 1) ? st_q[j] : st_s[j]);
            atomicAdd(dst + c * 2 + (lane & 1), v);
            if (p.totals) atomicAdd(p.totals + c * 2 + (lane & 1), v);
            st_s[j] = st_q[j] = 0.f;
          }
      }
    };
    for (int it = 0, tile = tfirst; it < iters; ++it, tile += tstep, ++tcount) {
      const uint32_t acc = tcount % p.acc_stages, aph = (tcount / p.acc_stages) & 1;
      const bool live = tile < t_end;
      const int n = tile / tiles_per_img;
      if (live && n != st_n) {
        flush();
        st_n = n;
      }
      mbar_wait(tfull(acc), aph);
      tc_fence_after();
      for (int s = 0; live && s < p.S; ++s) {
        const uint32_t t0 = tmem_base + acc * p.acc_cols + s * p.Cout + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          if (jj * 32 < p.Cout) {
            uint32_t rr[32];
            tmem_ld32(t0 + jj * 32, rr);
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rr[i]) + bias_s[jj * 32 + i];
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              float sq[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) sq[i] = v[hlf * 16 + i] * v[hlf * 16 + i];
              st_s[2 * jj + hlf] += transpose_reduce16(v + hlf * 16, lane);
              st_q[2 * jj + hlf] += transpose_reduce16(sq, lane);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(tempty(acc), 0));
        else mbar_arrive(tempty(acc));
      }
    }
    flush();
  } else if (warp < 4) {
    double* const stats_here = SW ? nullptr : p.stats;      // with statistics warps the epilogue warps only store
    double* const totals_here = SW ? nullptr : p.totals;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    uint32_t tcount = 0;
    // running per-channel statistics of the current image: chunk j holds column 16*j + col16(lane)
    float st_s[16], st_q[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) st_s[j] = st_q[j] = 0.f;
    int st_n = -1;
    auto flush = [&]() {
      if (stats_here && st_n >= 0) {
        // fp64 accumulators: the partials are fp32 (24-bit) values of similar magnitude, so their fp64 sum is exact in
        // (almost) any order - the statistics, and with them every bf16 rounding downstream, do not depend on the order in
        // which the CTAs arrive (fp32 atomics made the whole forward pass irreproducible run to run)
        double* dst = stats_here + (size_t)st_n * p.Cout * 2;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j * 16 < p.Cout) {
            const int c = j * 16 + col16(lane);
            const double v = (double)((lane & 1) ? st_q[j] : st_s[j]);
            atomicAdd(dst + c * 2 + (lane & 1), v);
            if (totals_here) atomicAdd(totals_here + c * 2 + (lane & 1), v);
            st_s[j] = st_q[j] = 0.f;
          }
      }
    };
    if (p.stage_g) {
      // Outputs leave through shared memory and the TMA unit: a thread owns one pixel row of the sub-tile, so direct
      // stores are 16-byte pieces 2*ld bytes apart (one L2 request each, ~0.45 requests/clk/SM: measured 7 B/clk/SM).
      // Instead every thread writes its bf16 row into a staging tile (the TMA swizzle keeps the 16-byte st.shared
      // conflict-free) and the tile leaves as a cp.async.bulk.tensor store: full 64/128-byte rows, no LSU traffic.
      // Every WARP stages and stores on its own: its 32 TMEM lanes are 4 image rows x 8 pixels of the sub-tile = one
      // (G, 8, 4, 1) box, two staging buffers per warp, one elected lane issues the store.  No barrier couples the four
      // epilogue warps (round 1 ran them in lockstep - named barrier + one issuing thread waiting for the previous store's
      // read - and with every pipeline ablated the kernel still took 75-90 % of its time: that chain was the bound of all
      // narrow layers and of every layer with fused statistics).
      const uint32_t G = p.stage_g, rowb = 2 * G;
      const uint32_t wbuf_bytes = 32 * rowb;
      const uint32_t stage_all = smem_b0 + nb * b_bytes;
      const uint32_t stage0 = stage_all + q * 2 * wbuf_bytes;
      const uint32_t xr = G == 64 ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
      const uint32_t my_row = lane * rowb;
      uint32_t sg = 0;
      float tot[4] = {0.f, 0.f, 0.f, 0.f};      // this thread's share of the CTA's batch totals (entries m, m+128, ...)
      // statistics leave the CTA once per image: the four warps' partials are combined through the (drained) staging
      // buffers in a fixed order, 2*Cout fp64 atomics per CTA and image instead of 8*Cout; the batch totals accumulate
      // in registers and leave once per CTA (they made every CTA hammer the same 2*Cout addresses at every image change)
      auto flush_staged = [&]() {
        if (!stats_here || st_n < 0) return;
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j * 16 < p.Cout) {
            const int c = j * 16 + col16(lane);
            st_shared_f32(stage0 + (uint32_t)(c * 2 + (lane & 1)) * 4u, (lane & 1) ? st_q[j] : st_s[j]);
            st_s[j] = st_q[j] = 0.f;
          }
        named_bar_sync(1, 128);
        double* dst = stats_here + (size_t)st_n * p.Cout * 2;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int e = m + 128 * k;
          if (e < 2 * p.Cout) {
            float a = ld_shared_f32(stage_all + (uint32_t)e * 4u);
#pragma unroll
            for (int w2 = 1; w2 < 4; ++w2) a += ld_shared_f32(stage_all + w2 * 2 * wbuf_bytes + (uint32_t)e * 4u);
            atomicAdd(dst + e, (double)a);
            tot[k] += a;
          }
        }
        named_bar_sync(1, 128);
      };
      for (int it = 0, tile = tfirst; it < iters; ++it, tile += tstep, ++tcount) {
        const uint32_t acc = tcount % p.acc_stages, aph = (tcount / p.acc_stages) & 1;
        const bool live = tile < t_end;          // (pair: rank 1 may own an empty last tile)
        const int n = tile / tiles_per_img;
        const int r = tile - n * tiles_per_img;
        const int h0 = (r / p.tilesW) * TILE_H + 4 * q;
        const int w0 = (r % p.tilesW) * SUB_W * p.S;
        if (live && n != st_n) {
          flush_staged();
          st_n = n;
        }
        mbar_wait(tfull(acc), aph);
        tc_fence_after();
        stamp();
        for (int s = 0; live && s < p.S; ++s) {
          const uint32_t t0 = tmem_base + acc * p.acc_cols + s * p.Cout + ((uint32_t)(q * 32) << 16);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            if (jj * 32 < p.Cout) {
              uint32_t rr[32];
              tmem_ld32(t0 + jj * 32, rr);
              if ((((uint32_t)jj * 32) & (G - 1)) == 0) {
                // first chunk of a group: the buffer it goes to was handed to the TMA unit two groups ago
                if (lane == 0) bulk_wait_read<1>();
                __syncwarp();
              }
              tmem_ld_wait();
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rr[i]) + bias_s[jj * 32 + i];
              if (POST) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  float cf[4];
                  ld_shared_v4f(smem0 + p.pre_tab + (uint32_t)(jj * 32 + i) * 8u, cf);
                  v[i] = norm_act1(v[i], cf[0], cf[1], p.pre.relu);
                  v[i + 1] = norm_act1(v[i + 1], cf[2], cf[3], p.pre.relu);
                }
              }
              const uint32_t dst = stage0 + (sg & 1) * wbuf_bytes + my_row;
              const uint32_t slot0 = ((uint32_t)(jj * 32) & (G - 1)) >> 3;
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                uint32_t w4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[t * 8 + 2 * i], v[t * 8 + 2 * i + 1]);
                  w4[i] = *reinterpret_cast<uint32_t*>(&h2);
                }
                st_shared_v4(dst + (((slot0 + t) ^ xr) << 4), w4[0], w4[1], w4[2], w4[3]);
              }
              if ((((uint32_t)(jj + 1) * 32) & (G - 1)) == 0) {   // the group is complete: hand it to the TMA unit
                fence_proxy_async();
                __syncwarp();
                if (lane == 0 && !(p.dbg & 4)) {
                  tma_store_4d(&tmY, stage0 + (sg & 1) * wbuf_bytes, (int)(((uint32_t)(jj * 32)) & ~(G - 1)), w0 + s * SUB_W, h0, n);
                  bulk_commit();
                }
                ++sg;
              }
              if (stats_here) {
#pragma unroll
                for (int hlf = 0; hlf < 2; ++hlf) {
                  float sq[16];
#pragma unroll
                  for (int i = 0; i < 16; ++i) sq[i] = v[hlf * 16 + i] * v[hlf * 16 + i];
                  st_s[2 * jj + hlf] += transpose_reduce16(v + hlf * 16, lane);
                  st_q[2 * jj + hlf] += transpose_reduce16(sq, lane);
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(tempty(acc), 0));
          else mbar_arrive(tempty(acc));
        }
        stamp();
      }
      flush_staged();
      if (totals_here) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int e = m + 128 * k;
          if (e < 2 * p.Cout) atomicAdd(totals_here + e, (double)tot[k]);
        }
      }
      st_n = -1;      // (the generic flush() below has nothing left to do)
      if (lane == 0) bulk_wait<0>();
    } else {
      for (int it = 0, tile = tfirst; it < iters; ++it, tile += tstep, ++tcount) {
        const uint32_t acc = tcount % p.acc_stages, aph = (tcount / p.acc_stages) & 1;
        const bool live = tile < t_end;
        const int n = tile / tiles_per_img;
        const int r = tile - n * tiles_per_img;
        const int h = (r / p.tilesW) * TILE_H + m / SUB_W;
        const int wbase = (r % p.tilesW) * SUB_W * p.S + m % SUB_W;
        if (live && n != st_n) {
          flush();
          st_n = n;
        }
        mbar_wait(tfull(acc), aph);
        tc_fence_after();
        stamp();
        for (int s = 0; live && s < p.S; ++s) {
          const size_t pix = ((size_t)n * p.H + h) * p.W + wbase + s * SUB_W;
          const uint32_t t0 = tmem_base + acc * p.acc_cols + s * p.Cout + ((uint32_t)(q * 32) << 16);
  #pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j * 16 < p.Cout) {
              uint32_t rr[16];
              tmem_ld16(t0 + j * 16, rr);
              tmem_ld_wait();
              float v[16];
  #pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rr[i]) + bias_s[j * 16 + i];
              if (POST) {
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                  float cf[4];
                  ld_shared_v4f(smem0 + p.pre_tab + (uint32_t)(j * 16 + i) * 8u, cf);
                  v[i] = norm_act1(v[i], cf[0], cf[1], p.pre.relu);
                  v[i + 1] = norm_act1(v[i + 1], cf[2], cf[3], p.pre.relu);
                }
              }
              if (p.dbg & 4) {
              } else if (p.y_f32) store16<float>((float*)p.y + pix * p.y_ld + j * 16, v, p.accumulate);
              else store16<bf16>((bf16*)p.y + pix * p.y_ld + j * 16, v, p.accumulate);
              if (stats_here) {
                float sq[16];
  #pragma unroll
                for (int i = 0; i < 16; ++i) sq[i] = v[i] * v[i];
                st_s[j] += transpose_reduce16(v, lane);
                st_q[j] += transpose_reduce16(sq, lane);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(tempty(acc), 0));
          else mbar_arrive(tempty(acc));
        }
        stamp();
      }
    }
    flush();
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();     // neither CTA leaves (or frees tensor memory) while the other may still signal it
  else __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace

bool conv_halo_eligible(const phs_tensor* x, const phs_tensor* y, int ksize) {
  return ksize == 3 && x->H % TILE_H == 0 && x->W % SUB_W == 0 && x->C % 32 == 0 && y->C % 16 == 0 && y->C >= 16 &&
         y->C <= 256;
}

// plan_out != nullptr: only choose the geometry and report it (phs_conv_halo_plan), nothing is launched
// pre != nullptr: x is the raw output of the previous convolution and *pre its normalisation (phs_conv2d_pre)
// post != nullptr: *post is THIS layer's inference-mode batch norm, applied in the epilogue (phs_conv2d_post)
static int conv_halo_impl(const phs_tensor* x, const void* w, const float* bias, const phs_tensor* y, int accumulate_flags,
                          double* stats, cudaStream_t st, int* plan_out, const phs_norm_pre* pre = nullptr,
                          const phs_norm_pre* post = nullptr) {
  const int accumulate = accumulate_flags & 1;
  const bool stats_prezeroed = (accumulate_flags & 2) != 0;   // the caller cleared stats (one fill for the whole program)
  const bool no_pair = (accumulate_flags & 4) != 0;           // the caller rules CTA pairs out for this launch
  const int BK = x->C % 64 == 0 ? 64 : 32;
  const int ROW = BK * 2;
  HaloParams p;
  p.N = x->N; p.H = x->H; p.W = x->W; p.Cin = x->C; p.Cout = y->C;
  p.kchunks = x->C / BK;
  const int subs_w = x->W / SUB_W;
  const int64_t total_subs = (int64_t)x->N * (x->H / TILE_H) * subs_w;
  // Geometry.  The filter tiles are re-streamed from L2 for every super-tile, and L2 -> SM bandwidth (~42 B/clk per SM)
  // is what bounds the wide layers: a [Cout][64] tile feeds only 4*S MMAs, so S (sub-tiles sharing one filter tile) has
  // to be as large as TMEM allows.  Two regimes:
  //   ctas == 2: two CTAs per SM (<= 256 TMEM columns, ~110 KB each) cover each other's epilogue / barrier latency;
  //   ctas == 1: one CTA per SM with all 512 TMEM columns and ~220 KB: S up to 4 for 128 channels, a deep filter ring.
  const char* e_ctas = getenv("PHS_HALO_CTAS");
  const char* e_s = getenv("PHS_HALO_S");
  const char* e_acc = getenv("PHS_HALO_ACC");
  const char* e_g = getenv("PHS_HALO_G");
  // at most one 128-pixel tile per SM (the 16x16 levels at batch 64): a second CTA slot would stay empty, so the one
  // CTA gets the whole shared memory = a filter ring deep enough to cover the TMA round trip
  // (round 2: inside the step the small layers of the prior and the posterior run side by side on two lanes; giving each
  // kernel whole SMs serialised them - two CTAs per SM everywhere is 0.05-0.09 ms per step faster, PHS_HALO_CTAS=1 / the
  // old rule "one CTA per SM when there is at most one tile per SM" stay selectable)
  const int ctas_per_sm = e_ctas ? atoi(e_ctas) : (getenv("PHS_HALO_1CTA") ? (total_subs <= num_sms() ? 1 : 2) : 2);
  // dynamic shared memory per CTA: 228 KB per SM, 1 KB reserved + ~1.8 KB static per CTA, 1 KB alignment slack
  constexpr int PRE_TAB_BYTES = 2048;      // (scale, shift) of up to 256 input channels
  if (pre && x->C > PRE_TAB_BYTES / 8) return -3;
  if (post && (pre || stats || accumulate)) return -1;
  const int budget_all = (ctas_per_sm == 1 ? SMEM_OPTIN - 2048 : 112896) - ((pre || post) ? PRE_TAB_BYTES : 0);
  const int max_cols = ctas_per_sm == 1 ? 512 : 256;
  const int min_tiles = ctas_per_sm * num_sms();
  // CTA pairs (cta_group::2): each CTA stages half of every filter tile.  Correct (tests/test_gpu_conv_tc.py::
  // test_conv_halo_cta_pairs: bit-identical to the single-CTA kernel) but only ~5 % faster on the wide layers without
  // fused statistics and slower with them (round 2, B=64, us: 128x128 128->128 247 -> 232 / 273 -> 317 with statistics,
  // 64x64 192->192 129 -> 124 / 163 -> 140, 128x128 32->192 172 -> 187): those layers run at 71 % tensor-pipe occupancy,
  // i.e. at 75-89 % of what cuBLAS reaches on this part, and the filter ring is no longer what holds them back.
  // INSIDE the training step, however, pairs win (tools/step_ab.py, ms per step: no pairs 12.47, pairs for launches
  // without fused statistics 12.32, pairs everywhere 12.23): fewer filter bytes per SM leave more L2 -> SM bandwidth to
  // the kernels of the other lanes.  Measured best (2): pairs for the launches WITHOUT fused statistics - with statistics the
  // two CTAs' long epilogues gate one shared accumulator hand-back and the kernel alone is slower.  PHS_HALO_PAIR=1:
  // everywhere, 0: nowhere (the default since the end of round 2, see below).
  const char* e_pair = getenv("PHS_HALO_PAIR");
  const bool pair_ok = y->C % 32 == 0 && y->C >= 32 && total_subs >= 2;
  // PHS_HALO_PAIR: 1 = wherever possible, 2 = only launches without fused statistics, 3 = only with;
  // PHS_HALO_PAIR_MINC / _MINCIN: smallest Cout / Cin that uses pairs
  // default since the end of round 2: 0 (no cluster launches; PHS_HALO_PAIR=2 is the measured-best setting, -0.15 ms per
  // step) - same reason as PHS_PDL in api.cu: two stalled training runs whose cause could not be isolated any more, and
  // the pair protocol (remote arrives, multicast commits, cluster barriers) is the newest code with unbounded waits
  const int pair_mode = e_pair ? atoi(e_pair) : 0;
  const int pair_minc = getenv("PHS_HALO_PAIR_MINC") ? atoi(getenv("PHS_HALO_PAIR_MINC")) : 32;
  const int pair_mincin = getenv("PHS_HALO_PAIR_MINCIN") ? atoi(getenv("PHS_HALO_PAIR_MINCIN")) : 32;
  bool pair = pair_ok && y->C >= pair_minc && x->C >= pair_mincin && !pre && !no_pair &&
              (pair_mode == 1 || (pair_mode == 2 && !stats) || (pair_mode == 3 && stats));
  const int b_bytes_full = y->C * ROW;
  int b_bytes = pair ? b_bytes_full / 2 : b_bytes_full;
  // staged TMA-store epilogue: bf16 outputs that are not accumulated onto, 32-channel granularity
  const bool can_stage = !accumulate && y->dtype == PHS_BF16 && y->C % 32 == 0 && y->ld % 8 == 0 && aligned16(y->ptr) &&
                         !(e_g && atoi(e_g) == 0);
  int na_pref = getenv("PHS_HALO_NA") ? atoi(getenv("PHS_HALO_NA")) : 2;   // activation (halo) stages when the filter streams
  auto geometry = [&](int G) -> bool {
    const int budget = budget_all - (G ? 2 * 128 * 2 * G : 0);
    int S = 1;
    if (e_s) {
      S = atoi(e_s);
      while (S > 1 && (S * y->C > max_cols || subs_w % S != 0)) S /= 2;
    } else {
      while (S * 2 * y->C <= max_cols && subs_w % (S * 2) == 0 &&
             na_pref * (((TILE_H + 2) * (SUB_W * S * 2 + 2) * ROW + 1023) / 1024 * 1024) + 4 * b_bytes <= budget &&
             total_subs / (S * 2) >= min_tiles)
        S *= 2;
    }
    p.S = S;
    p.stage_g = G;
    p.tilesW = subs_w / S;
    p.tilesH = x->H / TILE_H;
    p.num_tiles = p.tilesW * p.tilesH * x->N;
    const int a_bytes = (TILE_H + 2) * (SUB_W * S + 2) * ROW;
    p.a_stage_bytes = (a_bytes + 1023) / 1024 * 1024;
    const int cols = S * y->C;
    p.acc_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    // double-buffer the accumulators whenever the CTA's TMEM share allows it (the epilogue of tile i then overlaps
    // the MMAs of tile i+1 inside the CTA as well)
    p.acc_stages = 2 * p.acc_cols <= max_cols ? 2 : 1;
    if (e_acc && atoi(e_acc) == 1) p.acc_stages = 1;
    p.tmem_cols = p.acc_cols * p.acc_stages;
    // filter resident?
    const int b_all = p.kchunks * 9 * b_bytes;
    if (p.kchunks * 9 <= HB_MAX_B && b_all + 2 * (int)p.a_stage_bytes <= budget) {
      p.b_resident = 1;
      p.nb = p.kchunks * 9;
      p.na = (budget - b_all) / (int)p.a_stage_bytes;
      if (p.na > HB_MAX_A) p.na = HB_MAX_A;
      return true;
    }
    p.b_resident = 0;
    p.na = na_pref;
    p.nb = (budget - p.na * (int)p.a_stage_bytes) / b_bytes;
    if (p.nb > 12) p.nb = 12;
    return p.nb >= 2;
  };
  // The staging tiles come out of the filter ring.  A [Cout][64] filter stage feeds S*4 MMAs (~max(32, Cout/2) clk each)
  // and a ring refill is a ~2000 clk round trip, so narrow-output layers need a deep ring more than they need fast
  // stores: staging is dropped when it would leave the ring both short of that and less than half as deep as without it.
  int s_plain = 1;
  auto ring_ok = [&](int nb_plain) {
    if (p.S < s_plain) return false;   // never trade filter-tile reuse for staging
    if (p.b_resident) return true;
    const int per_stage = p.S * 4 * (y->C / 2 > 32 ? y->C / 2 : 32);
    const int need = (2000 + per_stage - 1) / per_stage;
    return p.nb >= need || 2 * p.nb > nb_plain;
  };
  bool ok = false;
  if (can_stage) {
    int nb_plain = geometry(0) ? (p.b_resident ? 1 << 20 : p.nb) : 0;
    s_plain = p.S;
    int G = e_g ? atoi(e_g) : (pair ? 32 : 64);
    if (G != 32 && (G != 64 || y->C % 64 != 0)) G = 32;
    ok = geometry(G) && (p.b_resident || p.nb >= 3 || G == 32) && ring_ok(nb_plain);
    if (!ok && G == 64) ok = geometry(32) && ring_ok(nb_plain);
    if (e_g && atoi(e_g) != 0 && !ok) ok = geometry(atoi(e_g) == 64 && y->C % 64 == 0 ? 64 : 32);
  }
  if (!ok) ok = geometry(0);
  // <= 64 output channels with a streamed filter: a [Cout][64] filter tile feeds only 4 short MMAs per sub-tile, so
  // doubling S (one halo stage instead of two pays for it) beats prefetching the next halo tile (measured: 64x64 64->64
  // 52 -> 40 us; 128-channel outputs: no gain)
  if (ok && !p.b_resident && p.S == 1 && y->C <= 64 && na_pref == 2 && !e_s && !pre) {
    const HaloParams keep = p;
    na_pref = 1;
    const bool ok1 = geometry(keep.stage_g) && p.S == 2 && !p.b_resident && p.nb >= keep.nb;
    if (!ok1) { p = keep; na_pref = 2; }
  }
  if (!ok) return -3;   // caller falls back to the shifted-box kernel
  if (pair && p.num_tiles < 2) return -3;
  p.y = y->ptr; p.y_ld = y->ld; p.y_f32 = y->dtype == PHS_F32;
  p.bias = bias;
  p.accumulate = accumulate;
  p.stats = stats;
  p.totals = (stats && stats_prezeroed) ? stats + (size_t)x->N * y->C * 2 : nullptr;
  {
    const char* e = getenv("PHS_HALO_DBG");
    p.dbg = e ? atoi(e) : 0;
    const char* t = getenv("PHS_HALO_TRACE");
    p.trace = t ? (long long*)strtoull(t, nullptr, 0) : nullptr;
  }
  p.pre_tab = (uint32_t)(p.na * (int)p.a_stage_bytes + p.nb * b_bytes + (p.stage_g ? 2 * 128 * 2 * p.stage_g : 0));
  if (pre) p.pre = *pre;
  else if (post) p.pre = *post;
  else memset(&p.pre, 0, sizeof(p.pre));
  const int smem = (int)p.pre_tab + ((pre || post) ? PRE_TAB_BYTES : 0) + 1024;
  const int ctas = ctas_per_sm * num_sms();
  int grid = p.num_tiles < ctas ? p.num_tiles : ctas;
  if (pair) {
    const int pairs = (p.num_tiles + 1) / 2 < ctas / 2 ? (p.num_tiles + 1) / 2 : ctas / 2;
    grid = 2 * pairs;
  }
  if (plan_out) {
    const int v[12] = {ctas_per_sm, p.S, p.na, p.nb, p.b_resident, p.stage_g, p.acc_stages, p.tmem_cols, smem, grid,
                       p.num_tiles, BK};
    for (int i = 0; i < 12; ++i) plan_out[i] = v[i];
    return 0;
  }
  CUtensorMap tmA, tmB;
  int rc = activation_map(x, BK, SUB_W * p.S + 2, TILE_H + 2, 1, &tmA);
  if (rc) return rc;
  rc = pair ? filter_map_rows(w, 9 * x->C, y->C, BK, y->C / 2, &tmB) : filter_map(w, 9 * x->C, y->C, BK, &tmB);
  if (rc) return rc;
  CUtensorMap tmY = tmA;
  if (p.stage_g && (rc = activation_map(y, p.stage_g, SUB_W, 4, 1, &tmY))) return rc;   // one epilogue warp's rows
  if (stats && !stats_prezeroed) cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)x->N * y->C, st);
  // statistics warps (PHS_HALO_SW=1): measured SLOWER than letting the four epilogue warps do both jobs (round 2, B=64:
  // 128x128 64->128 218 vs 182 us, 32->32 59 vs 52 us, step 12.77 vs 12.50 ms; ten warps at 96 registers and a second
  // pass over tensor memory cost more than the shuffles they take off the store path), so they stay opt-in
  const char* e_sw = getenv("PHS_HALO_SW");
  const bool sw = stats != nullptr && e_sw && atoi(e_sw) == 1 && !pre;
#define PHS_HALO_LAUNCH(BKV, PAIRV, SWV)                                                                  \
  do {                                                                                                    \
    static bool attr = false;                                                                             \
    if ((rc = allow_big_smem(conv_halo_kernel<BKV, PAIRV, SWV, false>, &attr))) return rc;                \
    if (PAIRV) phs_launch_cluster2(conv_halo_kernel<BKV, PAIRV, SWV, false>, grid, SWV ? 320 : 192, smem, st, tmA, tmB, tmY, p); \
    else phs_launch_tc(conv_halo_kernel<BKV, PAIRV, SWV, false>, grid, SWV ? 320 : 192, smem, st, tmA, tmB, tmY, p); \
  } while (0)
#define PHS_HALO_LAUNCH_PRE(BKV)                                                                          \
  do {                                                                                                    \
    static bool attr = false;                                                                             \
    if ((rc = allow_big_smem(conv_halo_kernel<BKV, false, false, true>, &attr))) return rc;               \
    phs_launch_tc(conv_halo_kernel<BKV, false, false, true>, grid, 256, smem, st, tmA, tmB, tmY, p);         \
  } while (0)
#define PHS_HALO_LAUNCH_POST(BKV, PAIRV)                                                                  \
  do {                                                                                                    \
    static bool attr = false;                                                                             \
    if ((rc = allow_big_smem(conv_halo_kernel<BKV, PAIRV, false, false, true>, &attr))) return rc;        \
    if (PAIRV) phs_launch_cluster2(conv_halo_kernel<BKV, PAIRV, false, false, true>, grid, 192, smem, st, tmA, tmB, tmY, p); \
    else phs_launch_tc(conv_halo_kernel<BKV, PAIRV, false, false, true>, grid, 192, smem, st, tmA, tmB, tmY, p); \
  } while (0)
  if (post) {
    if (BK == 64) { if (pair) PHS_HALO_LAUNCH_POST(64, true); else PHS_HALO_LAUNCH_POST(64, false); }
    else { if (pair) PHS_HALO_LAUNCH_POST(32, true); else PHS_HALO_LAUNCH_POST(32, false); }
  } else if (pre) {
    if (BK == 64) PHS_HALO_LAUNCH_PRE(64);
    else PHS_HALO_LAUNCH_PRE(32);
  } else if (BK == 64) {
    if (pair) { if (sw) PHS_HALO_LAUNCH(64, true, true); else PHS_HALO_LAUNCH(64, true, false); }
    else { if (sw) PHS_HALO_LAUNCH(64, false, true); else PHS_HALO_LAUNCH(64, false, false); }
  } else {
    if (pair) { if (sw) PHS_HALO_LAUNCH(32, true, true); else PHS_HALO_LAUNCH(32, true, false); }
    else { if (sw) PHS_HALO_LAUNCH(32, false, true); else PHS_HALO_LAUNCH(32, false, false); }
  }
#undef PHS_HALO_LAUNCH
#undef PHS_HALO_LAUNCH_PRE
#undef PHS_HALO_LAUNCH_POST
  return phs_check_launch("conv_halo_kernel");
}

// phs_conv2d_pre: conv(act(norm(yprev))) with the normalisation applied to the operand tile in shared memory
int conv2d_halo_pre(const phs_tensor* x, const phs_norm_pre* pre, const void* w, const float* bias, const phs_tensor* y,
                    int accumulate_flags, double* stats, cudaStream_t st) {
  return conv_halo_impl(x, w, bias, y, accumulate_flags, stats, st, nullptr, pre);
}

// phs_conv2d_post: act(bn_infer(conv(x))) with the normalisation folded into the epilogue
int conv2d_halo_post(const phs_tensor* x, const void* w, const float* bias, const phs_norm_pre* post, const phs_tensor* y,
                     cudaStream_t st) {
  return conv_halo_impl(x, w, bias, y, 0, nullptr, st, nullptr, nullptr, post);
}

extern "C" int phs_conv2d_pre_plan(const phs_tensor* x, const phs_tensor* y, int with_stats, int* plan) {
  PHS_REQUIRE(x && y && plan, "phs_conv2d_pre_plan: null argument");
  if (!conv_halo_eligible(x, y, 3) || x->dtype != PHS_BF16 || y->dtype != PHS_BF16) return 0;
  static double dummy_stats;
  static phs_norm_pre dummy_pre;
  int rc = conv_halo_impl(x, nullptr, nullptr, y, 2, with_stats ? &dummy_stats : nullptr, nullptr, plan, &dummy_pre);
  return rc == 0 ? 1 : (rc == -3 ? 0 : rc);
}

int conv2d_halo(const phs_tensor* x, const void* w, const float* bias, const phs_tensor* y, int accumulate_flags,
                double* stats, cudaStream_t st) {
  return conv_halo_impl(x, w, bias, y, accumulate_flags, stats, st, nullptr);
}

// Host-only: the launch geometry conv2d_halo would choose for this layer (no device work; usable without a GPU, the SM
// count then defaults to 148).  plan[12] = {CTAs per SM, S, halo stages, filter stages, filter resident, staging group,
// accumulator stages, TMEM columns, dynamic shared memory, grid, tiles, BK}.  Returns 1 if the halo kernel takes the
// layer, 0 if it does not (other kernels do), <0 on bad arguments.
extern "C" int phs_conv_halo_plan(const phs_tensor* x, const phs_tensor* y, int accumulate, int with_stats, int* plan) {
  PHS_REQUIRE(x && y && plan, "phs_conv_halo_plan: null argument");
  if (!conv_halo_eligible(x, y, 3)) return 0;
  static double dummy_stats;
  int rc = conv_halo_impl(x, nullptr, nullptr, y, accumulate, with_stats ? &dummy_stats : nullptr, nullptr, plan);
  return rc == 0 ? 1 : (rc == -3 ? 0 : rc);
}
