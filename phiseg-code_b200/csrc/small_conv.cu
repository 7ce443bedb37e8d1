// Memory-bound convolutions with a tiny channel count on one side (tf.nn.conv2d SAME stride 1, tfwrapper/layers.py:123):
//   - network inputs  (Cin in {1,2,3,5} -> 32..192 channels, 3x3): posterior/prior z0_pre_1, z*_ups_to_*_c_1, z*_post_1
//   - heads           (32..192 channels -> Cout in {2,4,6}):        z*_mu, z*_sigma, y_lvl*, prediction, pre_mu/sigma
// and the gradients of both.  These are dot products of a few hundred terms per pixel: no tensor cores, the job is to
// stream the wide tensor once with 16-byte accesses.  fp32 master filters (HWIO), fp32 accumulation, any mix of
// float32 / bfloat16 activations.
#include <stdlib.h>
#include "common.cuh"

namespace {

struct SmallGeom {
  int N, H, W;
  int Cin, Cout;  // channels of this launch's input / output tensors (roles swapped under dgrad)
  int ks, dgrad;
  int ldx, ldy;
};

// filter element seen by this launch: input channel a, output channel b, tap t (of THIS launch's correlation)
__device__ __forceinline__ float wsel(const float* __restrict__ w, const SmallGeom& g, int t, int a, int b) {
  const int taps = g.ks * g.ks;
  if (!g.dgrad) return w[((size_t)t * g.Cin + a) * g.Cout + b];
  return w[((size_t)(taps - 1 - t) * g.Cout + b) * g.Cin + a];   // dgrad: in = dy (co), out = dx (ci), taps flipped
}

// ---- wide input -> NOUT <= 8 outputs ----------------------------------------------------------------------------
// TPP threads share a pixel (each takes every TPP-th 8-channel vector), partial dot products are combined by shuffles.
// ksize 1: a thread's channels never change, so its filter slice lives in registers (MAXV vectors per thread);
// ksize 3: compact [tap][Cin][NOUT] filter copy in shared memory.
// The ksize-1 heads are pure streaming: what bounds them is bytes in flight (round 2, ncu: 64x64 192->2 at 1.5 TB/s with 95
// registers = 2 blocks per SM and MAXV loads per thread outstanding, long_scoreboard 10 warps per issue).  A thread
// therefore takes UP pixel groups per iteration and issues all their loads first (6-8 outstanding 16-byte loads), the grid
// is one wave of two resident blocks per SM.
constexpr int small_cout_up(int maxv, int nout) {      // pixel groups per iteration that fit next to the filter registers
  const int wr = maxv * 8 * nout, per = maxv * 8;
  const int up = per > 0 && wr < 96 ? (96 - wr) / per : 1;
  return up < 1 ? 1 : (up > 4 ? 4 : up);
}
template <typename TI, typename TO, int TPP, int NOUT, int MAXV>
__global__ void __launch_bounds__(256, MAXV > 0 ? (MAXV * 8 * NOUT <= 64 ? 2 : 1) : (NOUT <= 2 ? 3 : 2))
    small_cout_kernel(const TI* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                      TO* __restrict__ y, SmallGeom g, int accumulate) {
  PHS_PDL_PROLOGUE();
  extern __shared__ float ws[];  // ksize 3 only: [tap][Cin][NOUT]
  const int taps = g.ks * g.ks;
  const int nvec = g.Cin / 8;
  const int sub = threadIdx.x % TPP;
  constexpr int MV = MAXV > 0 ? MAXV : 1;
  float wr[MV][8][NOUT];
  if (MAXV > 0) {
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int cv = sub + j * TPP;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int o = 0; o < NOUT; ++o)
          wr[j][i][o] = (cv < nvec && o < g.Cout) ? wsel(w, g, 0, cv * 8 + i, o) : 0.f;
    }
  } else {
    // [tap][vector][8 channels x NOUT + 4 floats of padding]: the TPP threads of a pixel read consecutive vectors with
    // 16-byte loads; the padded stride (80 / 144 / 272 bytes) puts them on distinct banks (round 2, ncu of the unpadded
    // scalar reads: 6.9 M bank conflicts in 9.4 M shared-memory wavefronts, mio_throttle the dominant stall)
    constexpr int VS = 8 * NOUT + 4;
    for (int i = threadIdx.x; i < taps * g.Cin * NOUT; i += blockDim.x) {
      int o = i % NOUT, a = (i / NOUT) % g.Cin, t = (i / NOUT) / g.Cin;
      ws[(size_t)(t * nvec + a / 8) * VS + (a % 8) * NOUT + o] = o < g.Cout ? wsel(w, g, t, a, o) : 0.f;
    }
    __syncthreads();
  }
  const int64_t M = (int64_t)g.N * g.H * g.W;
  const int pad = g.ks / 2;
  if (MAXV > 0) {
    constexpr int UP = small_cout_up(MAXV, NOUT);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // (block-uniform loop bound: every lane takes part in the shuffles)
    // unconditional loads: a vector index beyond the tensor is clamped (its filter registers are zero), a pixel beyond the
    // end is clamped and not stored - no branch separates the loads, so all UP * MAXV of them are issued back to back
    int cvoff[MV];
#pragma unroll
    for (int j = 0; j < MAXV; ++j) cvoff[j] = (sub + j * TPP < nvec ? sub + j * TPP : nvec - 1) * 8;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < M * TPP; base += UP * stride) {
      Raw8<TI> raw[UP][MV];
      int64_t pixs[UP];
      bool lives[UP];
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        const int64_t pix_raw = (base + u * stride + threadIdx.x) / TPP;
        lives[u] = pix_raw < M;
        pixs[u] = lives[u] ? pix_raw : M - 1;
        const TI* px = x + pixs[u] * g.ldx;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) raw[u][j].load(px + cvoff[j]);
      }
#pragma unroll
      for (int u = 0; u < UP; ++u) {
        float acc[NOUT];
#pragma unroll
        for (int o = 0; o < NOUT; ++o) acc[o] = 0.f;
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
          float v[8];
          raw[u][j].unpack(v);
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int o = 0; o < NOUT; ++o) acc[o] = fmaf(v[i], wr[j][i][o], acc[o]);
        }
#pragma unroll
        for (int off = TPP / 2; off > 0; off >>= 1)
#pragma unroll
          for (int o = 0; o < NOUT; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
        if (sub == 0 && lives[u]) {
          TO* py = y + pixs[u] * g.ldy;
#pragma unroll
          for (int o = 0; o < NOUT; ++o)
            if (o < g.Cout) {
              float r = acc[o] + (bias ? bias[o] : 0.f);
              if (accumulate) r += ldf<TO>(py + o);
              stf<TO>(py + o, r);
            }
        }
      }
    }
    return;
  }
  // the loop bound is block-uniform: every lane takes part in the shuffles below
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < M * TPP; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix_raw = (base + threadIdx.x) / TPP;
    const bool live = pix_raw < M;
    const int64_t pix = live ? pix_raw : M - 1;
    float acc[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) acc[o] = 0.f;
    if (MAXV > 0) {
      const TI* px = x + pix * g.ldx;
      float v[MV][8];
#pragma unroll
      for (int j = 0; j < MAXV; ++j)
        if (sub + j * TPP < nvec) ldv<TI, 8>(px + (sub + j * TPP) * 8, v[j]);
#pragma unroll
      for (int j = 0; j < MAXV; ++j)
        if (sub + j * TPP < nvec) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int o = 0; o < NOUT; ++o) acc[o] = fmaf(v[j][i], wr[j][i][o], acc[o]);
        }
    } else {
      const int wq = (int)(pix % g.W);
      const int hq = (int)((pix / g.W) % g.H);
      for (int t = 0; t < taps; ++t) {
        const int hh = hq + t / g.ks - pad, ww = wq + t % g.ks - pad;
        if (hh < 0 || hh >= g.H || ww < 0 || ww >= g.W) continue;
        const TI* px = x + (pix + (int64_t)(hh - hq) * g.W + (ww - wq)) * g.ldx;
        constexpr int VS = 8 * NOUT + 4;
        const float* wt = ws + (size_t)t * nvec * VS;
        for (int cv0 = sub; cv0 < nvec; cv0 += 4 * TPP) {
          Raw8<TI> raw[4];        // up to four of this thread's vectors in flight (index clamped: no branch between the loads)
#pragma unroll
          for (int j = 0; j < 4; ++j) raw[j].load(px + (cv0 + j * TPP < nvec ? cv0 + j * TPP : nvec - 1) * 8);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (cv0 + j * TPP < nvec) {
              float v[8];
              raw[j].unpack(v);
              const float4* wv = reinterpret_cast<const float4*>(wt + (size_t)(cv0 + j * TPP) * VS);
              float wreg[8 * NOUT];
#pragma unroll
              for (int q = 0; q < 2 * NOUT; ++q) {
                const float4 f = wv[q];
                wreg[4 * q] = f.x; wreg[4 * q + 1] = f.y; wreg[4 * q + 2] = f.z; wreg[4 * q + 3] = f.w;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int o = 0; o < NOUT; ++o) acc[o] = fmaf(v[i], wreg[i * NOUT + o], acc[o]);
            }
        }
      }
    }
#pragma unroll
    for (int off = TPP / 2; off > 0; off >>= 1)
#pragma unroll
      for (int o = 0; o < NOUT; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
    if (sub == 0 && live) {
      TO* py = y + pix * g.ldy;
#pragma unroll
      for (int o = 0; o < NOUT; ++o)
        if (o < g.Cout) {
          float r = acc[o] + (bias ? bias[o] : 0.f);
          if (accumulate) r += ldf<TO>(py + o);
          stf<TO>(py + o, r);
        }
    }
  }
}

// ---- at most 8 input channels -> wide output --------------------------------------------------------------------
// one thread per (pixel, 8-channel output vector)
// KS / CIN > 0: compile-time filter size and input channels (the z inputs of the latent hierarchy and the likelihood: 3x3,
// zdim_0 = 2 channels): the tap loop unrolls with constant offsets.  The generic form spent ~1500 instructions per thread
// on runtime divisions and address arithmetic around 144 FMAs (round 2, ncu: issue-bound at 70 %, 36 us for an 8 MB output).
template <typename TI, typename TO, int KS = 0, int CIN = 0>
__global__ void __launch_bounds__(256)
    small_cin_kernel(const TI* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                     TO* __restrict__ y, SmallGeom g, int accumulate, idx4_t ix, uint32_t total) {
  PHS_PDL_PROLOGUE();
  extern __shared__ float ws[];  // [tap][Cin][Cout]
  const int taps = g.ks * g.ks;
  for (int i = threadIdx.x; i < taps * g.Cin * g.Cout; i += blockDim.x) {
    int b = i % g.Cout, a = (i / g.Cout) % g.Cin, t = i / (g.Cout * g.Cin);
    ws[i] = wsel(w, g, t, a, b);
  }
  __syncthreads();
  const int pad = g.ks / 2;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int cv, wq, hq, n;
    idx4_decode(i, ix, cv, wq, hq, n);     // multiply-high index arithmetic: 64-bit divisions made this ALU bound
    const int64_t pix = ((int64_t)n * g.H + hq) * g.W + wq;
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = bias ? bias[cv * 8 + o] : 0.f;
    if (KS > 0) {
      const TI* pc = x + pix * g.ldx;
      const float* wc = ws + cv * 8;
#pragma unroll
      for (int t = 0; t < KS * KS; ++t) {
        const int dh = t / KS - KS / 2, dw = t % KS - KS / 2;
        const int hh = hq + dh, ww = wq + dw;
        if (hh < 0 || hh >= g.H || ww < 0 || ww >= g.W) continue;
        const TI* px = pc + (dh * g.W + dw) * g.ldx;
#pragma unroll
        for (int a = 0; a < CIN; ++a) {
          const float xv = ldf<TI>(px + a);
          const float* wt = wc + (t * CIN + a) * g.Cout;
          const float4 w0 = *reinterpret_cast<const float4*>(wt);
          const float4 w1 = *reinterpret_cast<const float4*>(wt + 4);
          acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
          acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
          acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
          acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
        }
      }
    } else
    for (int t = 0; t < taps; ++t) {
      const int hh = hq + t / g.ks - pad, ww = wq + t % g.ks - pad;
      if (hh < 0 || hh >= g.H || ww < 0 || ww >= g.W) continue;
      const TI* px = x + (pix + (int64_t)(hh - hq) * g.W + (ww - wq)) * g.ldx;
      for (int a = 0; a < g.Cin; ++a) {
        const float xv = ldf<TI>(px + a);
        const float* wt = ws + ((size_t)t * g.Cin + a) * g.Cout + cv * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(wt);
        const float4 w1 = *reinterpret_cast<const float4*>(wt + 4);
        acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
        acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
        acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
        acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
      }
    }
    TO* py = y + pix * g.ldy + cv * 8;
    if (accumulate) {
      float o[8];
      ldv<TO, 8>(py, o);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += o[k];
    }
    stv<TO, 8>(py, acc);
  }
}

// ---- 1x1, at most CIN <= 4 input channels -> wide output (the input gradient of the 1x1 heads) -----------------------
// A thread owns one 8-channel output vector for good: its filter rows live in registers, it walks pixels with four
// independent loads in flight and does nothing but FMAs and 16-byte stores (no shared memory, no index arithmetic).
template <typename TI, typename TO, int CIN>
__global__ void __launch_bounds__(256)
    small_cin_1x1_kernel(const TI* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                         TO* __restrict__ y, SmallGeom g, int accumulate, int64_t M) {
  PHS_PDL_PROLOGUE();
  const int nvec = g.Cout / 8;
  const int PL = blockDim.x / nvec;
  const int cv = threadIdx.x % nvec, lp = threadIdx.x / nvec;
  if (lp >= PL) return;
  float wr[CIN][8], b8[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    b8[o] = bias ? bias[cv * 8 + o] : 0.f;
#pragma unroll
    for (int a = 0; a < CIN; ++a) wr[a][o] = a < g.Cin ? wsel(w, g, 0, a, cv * 8 + o) : 0.f;
  }
  constexpr int U = 4;
  const int64_t stride = (int64_t)gridDim.x * PL;
  for (int64_t p = (int64_t)blockIdx.x * PL + lp; p < M; p += U * stride) {
    float xv[U][CIN];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = p + u * stride;
#pragma unroll
      for (int a = 0; a < CIN; ++a) xv[u][a] = (q < M && a < g.Cin) ? ldf<TI>(x + q * g.ldx + a) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = p + u * stride;
      if (q >= M) break;
      float acc[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        acc[o] = b8[o];
#pragma unroll
        for (int a = 0; a < CIN; ++a) acc[o] = fmaf(xv[u][a], wr[a][o], acc[o]);
      }
      TO* py = y + q * g.ldy + cv * 8;
      if (accumulate) {
        float o8[8];
        ldv<TO, 8>(py, o8);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += o8[k];
      }
      stv<TO, 8>(py, acc);
    }
  }
}

// ---- filter gradient with a tiny channel count on one side --------------------------------------------------------
// dW[t][cs][cw] (small side = x, shifted by the tap)   or   dW[cw][cs] (ksize 1, small side = dy)
//   = sum_p S[p + t][cs] * Wd[p][cw]
// thread <-> (pair = (t, cs, 8-channel vector of the wide tensor), pixel lane); a block walks a contiguous pixel range
// with PL lanes, 4 pixels in flight per thread; lanes are combined in shared memory, blocks with atomics.
template <typename TS, typename TW>
__global__ void __launch_bounds__(256)
    wgrad_small_kernel(const TS* __restrict__ s, int lds, int Cs, const TW* __restrict__ wd, int ldw, int Cw, int N,
                       int H, int W, int ks, int small_is_x, int64_t pix_per_block, int PB, float* __restrict__ dw) {
  PHS_PDL_PROLOGUE();
  extern __shared__ float red[];  // [PB][8]
  const int taps = ks * ks;
  const int nvec = Cw / 8;
  const int pairs = taps * Cs * nvec;
  const int PL = blockDim.x / PB;
  const int lp = threadIdx.x / PB;
  const int pr = blockIdx.y * PB + threadIdx.x % PB;
  for (int i = threadIdx.x; i < PB * 8; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const bool active = pr < pairs && lp < PL;
  const int prc = pr < pairs ? pr : 0;
  const int cv = prc % nvec, j = prc / nvec;
  const int cs = j % Cs, t = j / Cs;
  const int pad = ks / 2;
  const int dh = t / ks - pad, dwv = t % ks - pad;
  const int64_t M = (int64_t)N * H * W;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p1 = p0 + pix_per_block < M ? p0 + pix_per_block : M;
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0.f;
  if (active) {
    constexpr int U = 4;
    // each pixel lane owns a contiguous sub-range: (h, w) are tracked incrementally, no divisions in the loop
    const int64_t chunk = (pix_per_block + PL - 1) / PL;
    int64_t p = p0 + (int64_t)lp * chunk;
    const int64_t pe = p + chunk < p1 ? p + chunk : p1;
    int wq = (int)(p % W);
    int hq = (int)((p / W) % H);
    for (; p < pe; p += U) {
      float sv[U];
      float v[U][8];
      int ww = wq, hh = hq;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t q = p + u;
        sv[u] = 0.f;
        if (q < pe) {
          const int h2 = hh + dh, w2 = ww + dwv;
          if (h2 >= 0 && h2 < H && w2 >= 0 && w2 < W) sv[u] = ldf<TS>(s + (q + (int64_t)dh * W + dwv) * lds + cs);
          ldv<TW, 8>(wd + q * ldw + cv * 8, v[u]);
        } else {
#pragma unroll
          for (int o = 0; o < 8; ++o) v[u][o] = 0.f;
        }
        if (++ww == W) {
          ww = 0;
          if (++hh == H) hh = 0;
        }
      }
      wq = ww;
      hq = hh;
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] = fmaf(sv[u], v[u][o], acc[o]);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) atomicAdd(&red[(threadIdx.x % PB) * 8 + o], acc[o]);
  }
  __syncthreads();
  if (threadIdx.x < PB && pr < pairs) {
    if (small_is_x) {
      float* d = dw + ((size_t)t * Cs + cs) * Cw + cv * 8;   // [tap][ci = small][co = wide]
#pragma unroll
      for (int o = 0; o < 8; ++o) atomicAdd(d + o, red[threadIdx.x * 8 + o]);
    } else {
      float* d = dw + (size_t)(cv * 8) * Cs + cs;            // ksize 1: [ci = wide][co = small]
#pragma unroll
      for (int o = 0; o < 8; ++o) atomicAdd(d + (size_t)o * Cs, red[threadIdx.x * 8 + o]);
    }
  }
}

// ---- 3x3 filter gradient, small side = x with CS <= 2 channels (the z inputs: z*_post_1, z*_ups_to_*_c_1) --------------
//   dW[t][cs][cw] = sum_p x[p + t][cs] * dy[p][cw]
// thread <-> (8-channel vector of dy, cs, pixel lane) with ALL NINE taps in registers (72 accumulators): the wide vector is
// loaded once per (pixel, cs) instead of once per (pixel, cs, tap) - the generic kernel above re-read dy 18 times through
// L1 (round 2, ncu: 84 us for an 8 MB tensor, LSU / issue bound).  Two pixels per iteration, loads first.  The pixel lanes
// of a block are combined in shared memory one lane at a time (plain read-modify-write: shared fp32 atomics are CAS
// loops), blocks with global atomics.
template <typename TS, typename TW, int CS>
__global__ void __launch_bounds__(256, 2)
    wgrad_small3_kernel(const TS* __restrict__ s, int lds, const TW* __restrict__ wd, int ldw, int Cw, int N, int H, int W,
                        int64_t pix_per_block, float* __restrict__ dw) {
  PHS_PDL_PROLOGUE();
  extern __shared__ float red[];  // [nvec * CS][72]
  const int nvec = Cw / 8;
  const int pairs = nvec * CS;
  const int PL = blockDim.x / pairs;
  const int pr = threadIdx.x % pairs, lp = threadIdx.x / pairs;
  const int cv = pr % nvec, cs = pr / nvec;
  for (int i = threadIdx.x; i < pairs * 72; i += blockDim.x) red[i] = 0.f;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[t][o] = 0.f;
  const int64_t M = (int64_t)N * H * W;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p1 = p0 + pix_per_block < M ? p0 + pix_per_block : M;
  const bool active = lp < PL && p0 < p1;
  if (active) {
    // each pixel lane owns a contiguous sub-range: (h, w) are tracked incrementally, no divisions in the loop
    const int64_t chunk = (pix_per_block + PL - 1) / PL;
    int64_t p = p0 + (int64_t)lp * chunk;
    const int64_t pe = p + chunk < p1 ? p + chunk : p1;
    int wq = (int)(p % W);
    int hq = (int)((p / W) % H);
    for (; p < pe; p += 2) {
      Raw8<TW> raw[2];
      float sv[2][9];
      int ww = wq, hh = hq;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t q = p + u;
        const bool in = q < pe;
        const int64_t qc = in ? q : pe - 1;
        raw[u].load(wd + qc * ldw + cv * 8);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int dh = t / 3 - 1, dx = t % 3 - 1;
          const int h2 = hh + dh, w2 = ww + dx;
          const bool ok = in && h2 >= 0 && h2 < H && w2 >= 0 && w2 < W;
          sv[u][t] = ok ? ldf<TS>(s + (q + (int64_t)dh * W + dx) * lds + cs) : 0.f;
        }
        if (++ww == W) {
          ww = 0;
          if (++hh == H) hh = 0;
        }
      }
      wq = ww;
      hq = hh;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float v[8];
        raw[u].unpack(v);
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
          for (int o = 0; o < 8; ++o) acc[t][o] = fmaf(sv[u][t], v[o], acc[t][o]);
      }
    }
  }
  __syncthreads();
  for (int l = 0; l < PL; ++l) {          // one pixel lane at a time: every (pair, tap, channel) slot has one writer
    if (active && lp == l) {
      float* d = red + (size_t)pr * 72;
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int o = 0; o < 8; ++o) d[t * 8 + o] += acc[t][o];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < pairs * 72; i += blockDim.x) {
    const int o = i & 7, t = (i >> 3) % 9, pr2 = i / 72;
    const int cv2 = pr2 % nvec, cs2 = pr2 / nvec;
    atomicAdd(dw + ((size_t)t * CS + cs2) * Cw + cv2 * 8 + o, red[i]);       // [tap][ci = small][co = wide]
  }
}

// ---- filter gradient of a 1x1 head (wide x, NS <= 8 output channels): dW[ci][co] = sum_p x[p][ci] * dy[p][co] -------
// thread <-> (8-channel vector of x, pixel lane) and ALL output channels: the 16-byte x load is shared by NS FMAs x 8.
// (round 2, ncu: 128x128 128->2 at 1.9 TB/s with four loads per thread outstanding and 592 blocks on 444 resident slots:
// eight loads in flight for NS <= 2 and one wave of two resident blocks per SM)
template <typename TW, typename TS, int NS>
__global__ void __launch_bounds__(256, NS <= 4 ? 2 : 1)
    wgrad_head_kernel(const TW* __restrict__ x, int ldx, int Cw, const TS* __restrict__ dy, int lds, int Cs, int64_t M,
                      int64_t pix_per_block, float* __restrict__ dw) {
  PHS_PDL_PROLOGUE();
  extern __shared__ float red[];  // [nvec][NS][8]
  const int nvec = Cw / 8;
  const int PL = blockDim.x / nvec;
  const int cv = threadIdx.x % nvec, lp = threadIdx.x / nvec;
  for (int i = threadIdx.x; i < nvec * NS * 8; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float acc[NS][8];
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[s][o] = 0.f;
  if (lp < PL) {
    const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
    const int64_t p1 = p0 + pix_per_block < M ? p0 + pix_per_block : M;
    constexpr int U = NS <= 2 ? 8 : 4;
    for (int64_t p = p0 + lp; p < p1; p += (int64_t)U * PL) {
      // unconditional loads (a pixel beyond the range is clamped and its dy taken as zero): no branch between them
      Raw8<TW> raw[U];
      float d[U][NS];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t q = p + (int64_t)u * PL;
        const bool in = q < p1;
        const int64_t qc = in ? q : p1 - 1;
        raw[u].load(x + qc * ldx + cv * 8);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const float t = s < Cs ? ldf<TS>(dy + qc * lds + s) : 0.f;
          d[u][s] = in ? t : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float v[8];
        raw[u].unpack(v);
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
          for (int o = 0; o < 8; ++o) acc[s][o] = fmaf(d[u][s], v[o], acc[s][o]);
      }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int o = 0; o < 8; ++o) atomicAdd(&red[(cv * NS + s) * 8 + o], acc[s][o]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nvec * NS * 8; i += blockDim.x) {
    const int o = i & 7, s2 = (i >> 3) % NS, c2 = (i >> 3) / NS;
    if (s2 < Cs) atomicAdd(dw + (size_t)(c2 * 8 + o) * Cs + s2, red[i]);   // ksize 1: [ci][co]
  }
}

}  // namespace

// returns 1 if handled, 0 if the shape is not for these kernels, <0 / >0 on error
int small_conv_try(const phs_tensor* x, const float* w, const float* bias, const phs_tensor* y, int ksize, int dgrad,
                   int accumulate, cudaStream_t st) {
  SmallGeom g;
  g.N = x->N; g.H = x->H; g.W = x->W; g.Cin = x->C; g.Cout = y->C; g.ks = ksize; g.dgrad = dgrad;
  g.ldx = x->ld; g.ldy = y->ld;
  const int taps = ksize * ksize;
  const int64_t M = (int64_t)x->N * x->H * x->W;
  const int xes = x->dtype == PHS_BF16 ? 2 : 4, yes = y->dtype == PHS_BF16 ? 2 : 4;
  if (y->C <= 8 && x->C % 8 == 0 && x->C >= 8 && x->ld % 8 == 0 && ((uintptr_t)x->ptr % (8 * xes > 16 ? 16 : 8 * xes)) == 0) {
    const int nout = y->C <= 2 ? 2 : y->C <= 4 ? 4 : 8;
    const size_t smem = ksize == 1 ? 0 : (size_t)taps * (x->C / 8) * (8 * nout + 4) * sizeof(float);   // padded, see the kernel
    if (smem > 48 * 1024) return 0;
    const int nvec = x->C / 8;
    const int tpp = nvec >= 8 ? 8 : nvec >= 4 ? 4 : nvec >= 2 ? 2 : 1;
    const int maxv = (nvec + tpp - 1) / tpp;
    if (ksize == 1 && (maxv > 4 || (nout == 8 && maxv > 2))) return 0;
    int64_t blocks = (M * tpp + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (ksize == 1 && blocks > 148 * 2) blocks = 148 * 2;      // one wave of the two resident blocks per SM
    if (ksize != 1 && blocks > 148 * 6) blocks = 148 * 6;      // two waves of the three resident blocks
#define LAUNCH_K(TI, TO, TPPV, NOUTV, MAXVV) \
  phs_launch(small_cout_kernel<TI, TO, TPPV, NOUTV, MAXVV>, (int)blocks, 256, smem, st, (const TI*)x->ptr, w, bias, (TO*)y->ptr, g, accumulate)
#define LAUNCH_V(TI, TO, TPPV, NOUTV)                          \
  do {                                                         \
    if (ksize != 1) LAUNCH_K(TI, TO, TPPV, NOUTV, 0);          \
    else if (maxv == 1) LAUNCH_K(TI, TO, TPPV, NOUTV, 1);      \
    else if (maxv == 2) LAUNCH_K(TI, TO, TPPV, NOUTV, 2);      \
    else if (NOUTV < 8 && maxv == 3) LAUNCH_K(TI, TO, TPPV, (NOUTV < 8 ? NOUTV : 2), 3); \
    else LAUNCH_K(TI, TO, TPPV, (NOUTV < 8 ? NOUTV : 2), 4);   \
  } while (0)
#define LAUNCH_N(TI, TO, TPPV)                    \
  do {                                            \
    if (nout == 2) LAUNCH_V(TI, TO, TPPV, 2);     \
    else if (nout == 4) LAUNCH_V(TI, TO, TPPV, 4);\
    else LAUNCH_V(TI, TO, TPPV, 8);               \
  } while (0)
#define LAUNCH_SC(TI, TO)                      \
  do {                                         \
    if (tpp == 8) LAUNCH_N(TI, TO, 8);         \
    else if (tpp == 4) LAUNCH_N(TI, TO, 4);    \
    else if (tpp == 2) LAUNCH_N(TI, TO, 2);    \
    else LAUNCH_N(TI, TO, 1);                  \
  } while (0)
    if (x->dtype == PHS_F32 && y->dtype == PHS_F32) LAUNCH_SC(float, float);
    else if (x->dtype == PHS_F32) LAUNCH_SC(float, bf16);
    else if (y->dtype == PHS_F32) LAUNCH_SC(bf16, float);
    else LAUNCH_SC(bf16, bf16);
#undef LAUNCH_SC
#undef LAUNCH_N
#undef LAUNCH_V
#undef LAUNCH_K
    int rc = phs_check_launch("small_cout_kernel");
    return rc ? rc : 1;
  }
  if (x->C <= 8 && y->C % 8 == 0 && y->ld % 8 == 0 && ((uintptr_t)y->ptr % (8 * yes > 16 ? 16 : 8 * yes)) == 0) {
    if (ksize == 1 && x->C <= 4 && y->C / 8 <= 256) {
      const int nvec = y->C / 8, PL = 256 / nvec;
      int64_t blocks = (M + (int64_t)PL * 4 - 1) / ((int64_t)PL * 4);
      if (blocks > 148 * 8) blocks = 148 * 8;
      if (blocks < 1) blocks = 1;
#define LAUNCH_S1(TI, TO)                                                                                          \
  do {                                                                                                             \
    if (x->C <= 2) phs_launch(small_cin_1x1_kernel<TI, TO, 2>, (int)blocks, 256, 0, st, (const TI*)x->ptr, w, bias, (TO*)y->ptr, g, accumulate, M); \
    else phs_launch(small_cin_1x1_kernel<TI, TO, 4>, (int)blocks, 256, 0, st, (const TI*)x->ptr, w, bias, (TO*)y->ptr, g, accumulate, M);          \
  } while (0)
      if (x->dtype == PHS_F32 && y->dtype == PHS_F32) LAUNCH_S1(float, float);
      else if (x->dtype == PHS_F32) LAUNCH_S1(float, bf16);
      else if (y->dtype == PHS_F32) LAUNCH_S1(bf16, float);
      else LAUNCH_S1(bf16, bf16);
#undef LAUNCH_S1
      int rc1 = phs_check_launch("small_cin_1x1_kernel");
      return rc1 ? rc1 : 1;
    }
    const size_t smem = (size_t)taps * x->C * y->C * sizeof(float);
    if (smem > 48 * 1024) return 0;
    int64_t total = M * (y->C / 8);
    if (total >= (1ll << 31)) return 0;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const idx4_t ix = idx4_make(y->C / 8, x->W, x->H);
#define LAUNCH_SI(TI, TO)                                                                                              \
  do {                                                                                                                 \
    if (ksize == 3 && x->C == 2 && !dgrad)                                                                             \
      phs_launch(small_cin_kernel<TI, TO, 3, 2>, (int)blocks, 256, smem, st, (const TI*)x->ptr, w, bias, (TO*)y->ptr, g, accumulate, ix, (uint32_t)total); \
    else                                                                                                               \
      phs_launch(small_cin_kernel<TI, TO>, (int)blocks, 256, smem, st, (const TI*)x->ptr, w, bias, (TO*)y->ptr, g, accumulate, ix, (uint32_t)total); \
  } while (0)
    if (x->dtype == PHS_F32 && y->dtype == PHS_F32) LAUNCH_SI(float, float);
    else if (x->dtype == PHS_F32) LAUNCH_SI(float, bf16);
    else if (y->dtype == PHS_F32) LAUNCH_SI(bf16, float);
    else LAUNCH_SI(bf16, bf16);
#undef LAUNCH_SI
    int rc = phs_check_launch("small_cin_kernel");
    return rc ? rc : 1;
  }
  return 0;
}

// dw must already hold the values to accumulate onto.  returns 1 if handled, 0 if not applicable.
int small_wgrad_try(const phs_tensor* x, const phs_tensor* dy, float* dw, int ksize, cudaStream_t st) {
  const int64_t M = (int64_t)x->N * x->H * x->W;
  const phs_tensor *s, *wd;
  int small_is_x;
  if (x->C <= 8 && dy->C % 8 == 0) {
    s = x; wd = dy; small_is_x = 1;
  } else if (dy->C <= 8 && x->C % 8 == 0 && ksize == 1) {
    s = dy; wd = x; small_is_x = 0;
  } else {
    return 0;
  }
  const int wes = wd->dtype == PHS_BF16 ? 2 : 4;
  if (wd->ld % 8 != 0 || ((uintptr_t)wd->ptr % (8 * wes > 16 ? 16 : 8 * wes)) != 0) return 0;
  if (!small_is_x && wd->C / 8 <= 32) {
    // 1x1 head: one thread per (x vector, pixel lane) for all output channels
    const int nvec = wd->C / 8;
    int64_t splits = s->C <= 4 ? 148 * 2 : 148;
    if (splits > (M + 255) / 256) splits = (M + 255) / 256;
    if (splits < 1) splits = 1;
    const int64_t ppb = (M + splits - 1) / splits;
    splits = (M + ppb - 1) / ppb;
    const int ns = s->C <= 2 ? 2 : s->C <= 4 ? 4 : 8;
    const size_t smem = (size_t)nvec * ns * 8 * sizeof(float);
#define LAUNCH_WH(TW, TS, NSV)                                                                                     \
  phs_launch(wgrad_head_kernel<TW, TS, NSV>, (unsigned)splits, 256, smem, st, (const TW*)wd->ptr, wd->ld, wd->C, (const TS*)s->ptr, \
                                                                      s->ld, s->C, M, ppb, dw)
#define LAUNCH_WH_N(TW, TS)              \
  do {                                   \
    if (ns == 2) LAUNCH_WH(TW, TS, 2);   \
    else if (ns == 4) LAUNCH_WH(TW, TS, 4); \
    else LAUNCH_WH(TW, TS, 8);           \
  } while (0)
    if (wd->dtype == PHS_F32 && s->dtype == PHS_F32) LAUNCH_WH_N(float, float);
    else if (wd->dtype == PHS_F32) LAUNCH_WH_N(float, bf16);
    else if (s->dtype == PHS_F32) LAUNCH_WH_N(bf16, float);
    else LAUNCH_WH_N(bf16, bf16);
#undef LAUNCH_WH_N
#undef LAUNCH_WH
    int rc = phs_check_launch("wgrad_head_kernel");
    return rc ? rc : 1;
  }
  if (small_is_x && ksize == 3 && (s->C == 1 || s->C == 2) && s->C * (wd->C / 8) <= 128 && !getenv("PHS_NO_WGRAD_SMALL3")) {
    // all nine taps per thread (wgrad_small3_kernel)
    const int pairs3 = s->C * (wd->C / 8);
    const int PL = 256 / pairs3;
    int64_t splits = 148 * 2;
    if (splits > (M + PL * 8 - 1) / (PL * 8)) splits = (M + PL * 8 - 1) / (PL * 8);     // >= 8 pixels per pixel lane
    if (splits < 1) splits = 1;
    const int64_t ppb = (M + splits - 1) / splits;
    splits = (M + ppb - 1) / ppb;
    const size_t smem3 = (size_t)pairs3 * 72 * sizeof(float);
#define LAUNCH_W3(TS, TW)                                                                                              \
  do {                                                                                                                 \
    if (s->C == 1) phs_launch(wgrad_small3_kernel<TS, TW, 1>, (unsigned)splits, 256, smem3, st, (const TS*)s->ptr, s->ld, (const TW*)wd->ptr, wd->ld, wd->C, x->N, x->H, x->W, ppb, dw); \
    else phs_launch(wgrad_small3_kernel<TS, TW, 2>, (unsigned)splits, 256, smem3, st, (const TS*)s->ptr, s->ld, (const TW*)wd->ptr, wd->ld, wd->C, x->N, x->H, x->W, ppb, dw); \
  } while (0)
    if (s->dtype == PHS_F32 && wd->dtype == PHS_F32) LAUNCH_W3(float, float);
    else if (s->dtype == PHS_F32) LAUNCH_W3(float, bf16);
    else if (wd->dtype == PHS_F32) LAUNCH_W3(bf16, float);
    else LAUNCH_W3(bf16, bf16);
#undef LAUNCH_W3
    int rc3 = phs_check_launch("wgrad_small3_kernel");
    return rc3 ? rc3 : 1;
  }
  const int taps = ksize * ksize;
  const int pairs = taps * s->C * (wd->C / 8);
  // PB pairs per block (32 or 64); the other 256 / PB thread groups are pixel lanes over the block's pixel range
  const int PB = pairs <= 32 ? 32 : 64;
  const int gy = (pairs + PB - 1) / PB;
  int64_t splits = (148 * 4 + gy - 1) / gy;
  if (splits > (M + 127) / 128) splits = (M + 127) / 128;
  if (splits < 1) splits = 1;
  const int64_t ppb = (M + splits - 1) / splits;
  splits = (M + ppb - 1) / ppb;
  dim3 grid((unsigned)splits, gy);
  const size_t smem = (size_t)PB * 8 * sizeof(float);
#define LAUNCH_WS(TS, TW)                                                                                             \
  phs_launch(wgrad_small_kernel<TS, TW>, grid, 256, smem, st, (const TS*)s->ptr, s->ld, s->C, (const TW*)wd->ptr, wd->ld, wd->C, \
                                                      x->N, x->H, x->W, ksize, small_is_x, ppb, PB, dw)
  if (s->dtype == PHS_F32 && wd->dtype == PHS_F32) LAUNCH_WS(float, float);
  else if (s->dtype == PHS_F32) LAUNCH_WS(float, bf16);
  else if (wd->dtype == PHS_F32) LAUNCH_WS(bf16, float);
  else LAUNCH_WS(bf16, bf16);
#undef LAUNCH_WS
  int rc = phs_check_launch("wgrad_small_kernel");
  return rc ? rc : 1;
}
