// Filter gradient of a 3x3 SAME convolution on tcgen05, "halo" formulation (Conv2DBackpropFilter, the node
// optimizer.minimize adds for tfwrapper/layers.py:123; phiseg_model.py:141).
//
//   dW[kh][kw][ci][co] = sum_{n,h,w} X[n, h+kh-1, w+kw-1, ci] * dY[n, h, w, co]
//
// The TMA unit is the scarce resource of the shifted-box kernel in conv_tc.cu (its cost is per 64/128-byte row, and
// every tap re-loads the activation tile).  Here a CTA loads, per brick of 16 rows x 8 columns of one image, ONE halo
// tile of X (16 x 10 pixels for its filter row kh) and one tile of dY, and feeds the three taps kw = 0,1,2 from the same
// shared-memory tile through SHIFTED UMMA descriptors: both operands are MN-major (channels contiguous, one pixel per
// 64/128-byte row), so a tap shift is a whole-row offset of the descriptor start address.  tools/umma_probe.cu shows
// that tcgen05.mma applies the 64B/128B swizzle to absolute shared-memory address bits, so start addresses that are
// not aligned to the swizzle pattern, and 8-row groups 10 rows apart (SBO = 10 rows), address the TMA-written tile
// correctly with base_offset = 0.
//
//   M (128 rows of D)   = input channels.  Cin >= 128: a 128-channel block as two 64-channel halo boxes (LBO = box
//                         stride), one accumulator per kw.  Cin = 64 / 32: the whole Cin for 2 / 4 CONSECUTIVE kw shifts
//                         stacked along M, read from the same box with LBO = one pixel row (the 4th shift of a 3-wide
//                         filter is computed and discarded).
//   N                   = a block of output channels (<= 128 columns per accumulator, <= 3 accumulators in TMEM)
//   K                   = pixels, 16 per MMA (two image rows of the brick), 8 MMAs per brick and accumulator
//
// A CTA owns (kh, ci block, co block, a slice of the bricks); partial sums go to the fp32 HWIO gradient with
// red.global.add.v4.f32.
#include <stdlib.h>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_host.cuh"

using namespace tc;

namespace {

constexpr int WG_MAX_STAGES = 12;
constexpr int HALO_W = 10;  // 8 + 2 columns
constexpr int BR_H = 16, BR_W = 8;

struct WgradHaloParams {
  int N, H, W, Cin, Cout;
  int bricksW, bricksH, num_bricks, bricks_per_split;
  int slabw;       // channels per X box: 64 (128B swizzle) or 32 (64B swizzle)
  int a_boxes;     // X boxes per stage (2 or 4 channel slabs, or 1 when kw shifts are stacked along M)
  int a_lbo;       // bytes between the M slabs as the MMA sees them
  int nkh;         // filter rows per CTA: 1 (kh from the work item) or 3 (18-row halo tile, narrow inputs)
  int n_acc;       // accumulators = kw groups (per filter row)
  int kw_step;     // kw distance between accumulators
  int kw_per_acc;  // kw shifts stacked in one accumulator (1, 2 or 4)
  int nb;          // output channels per CTA (N of the MMA)
  int slabB, nslabB;
  int ci_blocks, co_blocks;
  int stages, tmem_cols;
  float* dw;
};

__global__ void __launch_bounds__(192, 2)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                  const WgradHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * WG_MAX_STAGES + 1];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rowA = p.slabw * 2, rowB = p.slabB * 2;
  const int halo_h = p.nkh == 3 ? BR_H + 2 : BR_H;
  const uint32_t boxA_bytes = halo_h * HALO_W * rowA, boxB_bytes = BR_H * BR_W * rowB;
  // every TMA box starts on a 1024-byte boundary (swizzle pattern anchor); TX bytes are the boxes' own sizes
  const uint32_t boxA_stride = (boxA_bytes + 1023u) & ~1023u;
  const uint32_t A_BYTES = boxA_stride * p.a_boxes, B_BYTES = boxB_bytes * p.nslabB;
  const uint32_t TX_BYTES = boxA_bytes * p.a_boxes + B_BYTES;
  const uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (WG_MAX_STAGES + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * WG_MAX_STAGES);

  // work item: kh fastest so the three CTAs sharing the same bricks run together
  int item = blockIdx.x;
  const int kh = p.nkh == 3 ? 0 : item % 3;
  if (p.nkh != 3) item /= 3;
  const int cib = item % p.ci_blocks;
  const int cob = item / p.ci_blocks;
  const int t_begin = blockIdx.y * p.bricks_per_split;
  const int t_end = min(p.num_bricks, t_begin + p.bricks_per_split);

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

  // lean, warp-uniform issue loops (see conv_halo.cu): counters instead of divisions, descriptors advanced by adds
  const int stages = p.stages, a_boxes = p.a_boxes, nslabB = p.nslabB, n_acc = p.n_acc, nb = p.nb, nkh = p.nkh;
  if (warp == W_PROD) {
    int s = 0;
    uint32_t ph = 0;
    const int bpi = p.bricksW * p.bricksH;
    int n = t_begin / bpi;
    int r = t_begin - n * bpi;
    const int c_a = cib * 128, c_b = cob * nb;
    for (int t = t_begin; t < t_end; ++t) {
      const int bh = r / p.bricksW;
      const int w0 = (r - bh * p.bricksW) * BR_W, h0 = bh * BR_H;
      mbar_wait(empty_bar(s), ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), TX_BYTES);
        const uint32_t a_s = smem0 + s * STAGE_BYTES;
        for (int j = 0; j < a_boxes; ++j)
          tma_load_4d(a_s + j * boxA_stride, &tmX, full_bar(s), c_a + j * p.slabw, w0 - 1, h0 + (nkh == 3 ? 0 : kh) - 1, n);
        for (int j = 0; j < nslabB; ++j)
          tma_load_4d(a_s + A_BYTES + j * boxB_bytes, &tmDY, full_bar(s), c_b + j * p.slabB, w0, h0, n);
      }
      __syncwarp();
      if (++s == stages) { s = 0; ph ^= 1; }
      if (++r == bpi) { r = 0; ++n; }
    }
  } else if (warp == W_MMA) {
    const uint32_t idesc = idesc_bf16(128, nb, 1, 1);
    const uint64_t layA = p.slabw == 64 ? LAYOUT_SW128 : LAYOUT_SW64;
    const uint64_t layB = p.slabB == 64 ? LAYOUT_SW128 : LAYOUT_SW64;
    // MN-major: LBO = distance between channel slabs (or one pixel row when kw shifts are stacked along M),
    // SBO = distance between groups of 8 pixels (HALO_W rows in the X tile, 8 rows in the dY tile)
    const uint32_t a_hi = desc_hi(HALO_W * rowA, layA), b_hi = desc_hi(BR_W * rowB, layB);
    const uint32_t a_kstep = (2 * HALO_W * rowA) >> 4, b_kstep = (2 * BR_W * rowB) >> 4;
    const uint32_t a_gstep = (p.kw_step * rowA) >> 4;
    int s = 0;
    uint32_t ph = 0;
    uint32_t first = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_s = smem0 + s * STAGE_BYTES;
        uint32_t a_row = desc_lo(a_s, p.a_lbo);
        const uint32_t b_lo = desc_lo(a_s + A_BYTES, boxB_bytes);
        uint32_t d = tmem_base;
        for (int khi = 0; khi < nkh; ++khi, a_row += (HALO_W * rowA) >> 4) {   // filter row = halo rows khi..khi+15
          uint32_t a_lo = a_row;
          for (int g = 0; g < n_acc; ++g, a_lo += a_gstep, d += nb) {
#pragma unroll
            for (int k = 0; k < BR_H / 2; ++k)
              umma_bf16_lohi(d, a_lo + k * a_kstep, a_hi, b_lo + k * b_kstep, b_hi, idesc, k ? 1u : first);
          }
        }
        umma_commit(empty_bar(s));
        if (t == t_end - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
      first = 1;
      if (++s == stages) { s = 0; ph ^= 1; }
    }
    if (t_begin >= t_end && elect_one()) umma_commit(tfull_bar);
  } else if (t_begin < t_end) {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    for (int ga = 0; ga < p.nkh * p.n_acc; ++ga) {
      const int g = ga % p.n_acc;
      const int khe = p.nkh == 3 ? ga / p.n_acc : kh;
      int kw, ci;
      if (p.kw_per_acc == 1) {
        kw = g;
        ci = cib * 128 + m;
      } else {
        kw = g * p.kw_step + m / p.slabw;
        ci = m % p.slabw;
      }
      const bool valid = kw < 3 && ci < p.Cin;
      float* dst = p.dw + ((size_t)((khe * 3 + (valid ? kw : 0)) * p.Cin + (valid ? ci : 0))) * p.Cout + cob * p.nb;
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + ga * p.nb;
      for (int c0 = 0; c0 < p.nb; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(t0 + c0, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + i), "f"(__uint_as_float(r[i])),
                         "f"(__uint_as_float(r[i + 1])), "f"(__uint_as_float(r[i + 2])), "f"(__uint_as_float(r[i + 3]))
                         : "memory");
        }
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

}  // namespace

bool wgrad_halo_eligible(const phs_tensor* x, const phs_tensor* dy, int ksize) {
  return ksize == 3 && x->H % BR_H == 0 && x->W % BR_W == 0 && x->C % 32 == 0 && dy->C % 32 == 0 && dy->C <= 256;
}

// dw must already hold the values to accumulate onto (the caller zeroes it when accumulate == 0)
// plan_out != nullptr: only choose the geometry and report it (phs_wgrad_halo_plan), nothing is launched
static int wgrad_halo_impl(const phs_tensor* x, const phs_tensor* dy, float* dw, cudaStream_t st, int* plan_out) {
  WgradHaloParams p;
  p.N = x->N; p.H = x->H; p.W = x->W; p.Cin = x->C; p.Cout = dy->C;
  p.bricksW = x->W / BR_W;
  p.bricksH = x->H / BR_H;
  p.num_bricks = p.bricksW * p.bricksH * x->N;
  if (x->C == 32 || x->C == 64) {
    p.slabw = x->C;
    p.a_boxes = 1;
    p.a_lbo = x->C * 2;            // next M slab = the same tile one pixel further
    p.kw_per_acc = 128 / x->C;
    p.kw_step = p.kw_per_acc;
    p.n_acc = (3 + p.kw_per_acc - 1) / p.kw_per_acc;
    p.ci_blocks = 1;
  } else {
    p.slabw = x->C % 64 == 0 ? 64 : 32;
    p.a_boxes = 128 / p.slabw;
    p.a_lbo = 0;                   // set below: the (1024-aligned) stride between the channel-slab boxes
    p.kw_per_acc = 1;
    p.kw_step = 1;
    p.n_acc = 3;
    p.ci_blocks = (x->C + 127) / 128;
  }
  // narrow inputs (kw shifts stacked along M): one CTA takes all three filter rows from an 18-row halo tile, so X and
  // dY are read once instead of three times; wide inputs keep one filter row per CTA (N would drop to 32 otherwise)
  // (measured for 64-channel inputs: does not pay, N would shrink to 64).  Wide inputs with 32 output channels are the
  // other case where it fits (9 accumulators of 32 columns): X, the big operand, is then read once instead of three times.
  p.nkh = p.kw_per_acc == 4 ? 3 : 1;
  {
    const char* e = getenv("PHS_WGRAD_NKH3");
    if (p.kw_per_acc == 1 && dy->C == 32 && !(e && atoi(e) == 0)) p.nkh = 3;
  }
  // output-channel block: nkh * n_acc * nb TMEM columns <= 512, nb <= 128 keeps the stage small
  const int max_nb = 512 / (p.nkh * p.n_acc) >= 128 ? 128 : (512 / (p.nkh * p.n_acc)) / 32 * 32;
  p.co_blocks = (dy->C + max_nb - 1) / max_nb;
  p.nb = dy->C / p.co_blocks;
  if (p.nb % 32 || p.nb > max_nb) {
    p.co_blocks = dy->C / 32;
    p.nb = 32;
  }
  p.slabB = p.nb % 64 == 0 ? 64 : 32;
  p.nslabB = p.nb / p.slabB;
  int cols = p.nkh * p.n_acc * p.nb;
  p.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  p.dw = dw;
  const int halo_h = p.nkh == 3 ? BR_H + 2 : BR_H;
  const int box_stride = (halo_h * HALO_W * p.slabw * 2 + 1023) / 1024 * 1024;
  if (p.kw_per_acc == 1) p.a_lbo = box_stride;
  const int stage_bytes = box_stride * p.a_boxes + BR_H * BR_W * p.nb * 2;
  // Residency decides the grid: with <= 256 TMEM columns and >= 4 stages in half the shared memory two CTAs share an SM
  // (one CTA's atomics epilogue / pipeline fill overlaps the other's MMAs); otherwise one CTA owns the SM.  The grid is
  // then exactly ONE wave of resident CTAs (297 CTAs on 148 one-CTA SMs used to run as two waves plus a straggler).
  const char* e_res = getenv("PHS_WGRAD_CTAS");
  int resident = (p.tmem_cols <= 256 && (110 * 1024) / stage_bytes >= 4) ? 2 : 1;
  if (e_res) resident = atoi(e_res) == 2 && p.tmem_cols <= 256 ? 2 : 1;
  int stages = ((resident == 2 ? 112 * 1024 : SMEM_OPTIN) - 2048) / stage_bytes;
  if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const int items = (p.nkh == 3 ? 1 : 3) * p.ci_blocks * p.co_blocks;
  int splits = (resident * num_sms()) / items;
  if (splits > p.num_bricks) splits = p.num_bricks;
  if (splits < 1) splits = 1;
  p.bricks_per_split = (p.num_bricks + splits - 1) / splits;
  splits = (p.num_bricks + p.bricks_per_split - 1) / p.bricks_per_split;
  if (plan_out) {
    const int v[12] = {resident, p.nkh, p.n_acc, p.nb, p.ci_blocks, p.co_blocks, stages, p.tmem_cols,
                       stages * stage_bytes + 2048, items, splits, p.bricks_per_split};
    for (int i = 0; i < 12; ++i) plan_out[i] = v[i];
    return 0;
  }
  CUtensorMap tmX, tmDY;
  int rc = activation_map(x, p.slabw, HALO_W, halo_h, 1, &tmX);
  if (rc) return rc;
  rc = activation_map(dy, p.slabB, BR_W, BR_H, 1, &tmDY);
  if (rc) return rc;
  static bool attr = false;
  if ((rc = allow_big_smem(wgrad_halo_kernel, &attr))) return rc;
  const int smem = stages * stage_bytes + 2048;
  phs_launch_tc(wgrad_halo_kernel, dim3(items, splits), 192, smem, st, tmX, tmDY, p);
  return phs_check_launch("wgrad_halo_kernel");
}

int conv2d_wgrad_halo(const phs_tensor* x, const phs_tensor* dy, float* dw, cudaStream_t st) {
  return wgrad_halo_impl(x, dy, dw, st, nullptr);
}

// Host-only (no device work): the geometry conv2d_wgrad_halo would choose.  plan[12] = {resident CTAs per SM, filter rows
// per CTA, accumulators per filter row, output channels per CTA, input-channel blocks, output-channel blocks, pipeline
// stages, TMEM columns, dynamic shared memory, work items (grid.x), brick splits (grid.y), bricks per split}.
extern "C" int phs_wgrad_halo_plan(const phs_tensor* x, const phs_tensor* dy, int* plan) {
  PHS_REQUIRE(x && dy && plan, "phs_wgrad_halo_plan: null argument");
  if (!wgrad_halo_eligible(x, dy, 3)) return 0;
  return wgrad_halo_impl(x, dy, nullptr, nullptr, plan) == 0 ? 1 : -1;
}
