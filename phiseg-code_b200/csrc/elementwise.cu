// Memory-bound kernels of the PHiSeg hot path: normalisation (batch_norm / group_norm2D) forward and backward,
// 2x2 average pool, TF1-legacy bilinear x2 up-sampling, their adjoints, and small helpers.
// All tensors NHWC with a pixel pitch (ld) so that channel slices of concat buffers are addressed in place.
#include <stdlib.h>
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// per-(sample, channel) reductions.  grid = (chunks, N); each block walks a contiguous run of pixels of one
// sample with CW channel-vector lanes x PL pixel lanes, reduces the pixel lanes through shared memory and
// issues one atomicAdd per (n, c, quantity).
// ---------------------------------------------------------------------------------------------------------
constexpr int RED_THREADS = 256;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}


template <typename T, int V>
__global__ void __launch_bounds__(RED_THREADS) chan_stats_kernel(const T* __restrict__ y, int HW, int C, int ld,
                                                                 int pix_per_block, double* __restrict__ stats,
                                                                 double* __restrict__ totals) {
  PHS_PDL_PROLOGUE();
  const int n = blockIdx.y;
  const int nvec = C / V;
  const int CW = nvec < RED_THREADS ? nvec : RED_THREADS;
  const int PL = RED_THREADS / CW;
  const int lane_c = threadIdx.x % CW, lane_p = threadIdx.x / CW;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  extern __shared__ float sm[];  // [PL][C][2]
  const T* base = y + (size_t)n * HW * ld;
  for (int cv = lane_c; cv < nvec; cv += CW) {
    float s[V], q[V];
#pragma unroll
    for (int i = 0; i < V; ++i) s[i] = q[i] = 0.f;
    if (lane_p < PL) {
      constexpr int U = 4;   // independent 16-byte loads in flight per thread
      int p = p0 + lane_p;
      for (; p + (U - 1) * PL < p1; p += U * PL) {
        float v[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u) ldv<T, V>(base + (size_t)(p + u * PL) * ld + cv * V, v[u]);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int i = 0; i < V; ++i) { s[i] += v[u][i]; q[i] += v[u][i] * v[u][i]; }
      }
      for (; p < p1; p += PL) {
        float v[V];
        ldv<T, V>(base + (size_t)p * ld + cv * V, v);
#pragma unroll
        for (int i = 0; i < V; ++i) { s[i] += v[i]; q[i] += v[i] * v[i]; }
      }
      float* d = sm + ((size_t)lane_p * C + cv * V) * 2;
#pragma unroll
      for (int i = 0; i < V; ++i) { d[2 * i] = s[i]; d[2 * i + 1] = q[i]; }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += RED_THREADS) {
    float a = 0.f;
    for (int l = 0; l < PL; ++l) a += sm[(size_t)l * C * 2 + i];
    atomicAdd(&stats[(size_t)n * C * 2 + i], (double)a);   // fp64: order-independent in practice (conv_halo.cu flush)
    if (totals) atomicAdd(&totals[i], (double)a);
  }
}

// sums[n][c] = (sum g*mask, sum g*mask*xhat); per_sample == 0: sums[c] = the same summed over the batch
// REMAT: also write the activation (a_out).  Same grid and the same summation order as the plain variant (the partial
// sums, and with them every bit downstream, do not depend on which variant ran); two loads in flight instead of four keep
// it at three blocks per SM without spills.
// RAWL (V == 8, not with REMAT): the four (g, y) vector pairs stay packed in registers until used - eight loads in flight per
// thread instead of two (see norm_bwd_apply_kernel), two blocks per SM.
template <typename T, int V, bool REMAT, bool RAWL = false>
__global__ void __launch_bounds__(RED_THREADS, RAWL ? 2 : 3)
    norm_bwd_reduce_kernel(const T* __restrict__ g, int ldg, const T* __restrict__ y, int ldy, int HW, int C,
                           int pix_per_block, const float* __restrict__ mean, const float* __restrict__ rstd,
                           const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                           double* __restrict__ sums, int per_sample, T* __restrict__ a_out, int lda) {
  PHS_PDL_PROLOGUE();
  const int n = blockIdx.y;
  const int nvec = C / V;
  const int CW = nvec < RED_THREADS ? nvec : RED_THREADS;
  const int PL = RED_THREADS / CW;
  const int lane_c = threadIdx.x % CW, lane_p = threadIdx.x / CW;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  extern __shared__ float sm[];
  const T* gb = g + (size_t)n * HW * ldg;
  const T* yb = y + (size_t)n * HW * ldy;
  // a_out: the activation a = act(norm(y)) is re-materialised on the way (the forward pass of a fused conv -> norm -> ReLU
  // -> conv pair never wrote it, conv_halo.cu PRE; the filter gradient of the consumer reads it) - same bits as
  // norm_act_fwd would have written
  T* ab = REMAT ? a_out + (size_t)n * HW * lda : nullptr;
  for (int cv = lane_c; cv < nvec; cv += CW) {
    float s1[V], s2[V], mu[V], rs[V], ga[V], be[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      int c = cv * V + i;
      s1[i] = s2[i] = 0.f;
      mu[i] = mean[(size_t)n * C + c];
      rs[i] = rstd[(size_t)n * C + c];
      ga[i] = gamma[c];
      be[i] = beta[c];
    }
    if (lane_p < PL) {
      constexpr int U = REMAT ? 2 : 4;
      int p = p0 + lane_p;
      if constexpr (RAWL) {
        static_assert(V == 8 && !REMAT, "packed loads: 8-channel vectors, plain reduction");
        for (; p + (U - 1) * PL < p1; p += U * PL) {
          Raw8<T> graw[U], yraw[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            graw[u].load(gb + (size_t)(p + u * PL) * ldg + cv * V);
            yraw[u].load(yb + (size_t)(p + u * PL) * ldy + cv * V);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            float gv[V], yv[V];
            graw[u].unpack(gv);
            yraw[u].unpack(yv);
#pragma unroll
            for (int i = 0; i < V; ++i) {
              float xh = (yv[i] - mu[i]) * rs[i];
              float a = ga[i] * xh + be[i];
              float gm = (relu && a <= 0.f) ? 0.f : gv[i];
              s1[i] += gm;
              s2[i] += gm * xh;
            }
          }
        }
      }
      for (; p + (U - 1) * PL < p1; p += U * PL) {
        float gv[U][V], yv[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          ldv<T, V>(gb + (size_t)(p + u * PL) * ldg + cv * V, gv[u]);
          ldv<T, V>(yb + (size_t)(p + u * PL) * ldy + cv * V, yv[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            float xh = (yv[u][i] - mu[i]) * rs[i];
            float a = ga[i] * xh + be[i];
            float gm = (relu && a <= 0.f) ? 0.f : gv[u][i];
            s1[i] += gm;
            s2[i] += gm * xh;
          }
          if (REMAT) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
              float sc, sh;      // (recomputed per element: three flops, no registers held across the loop)
              norm_scale_shift(ga[i], be[i], mu[i], rs[i], &sc, &sh);
              yv[u][i] = norm_act1(yv[u][i], sc, sh, relu);
            }
            stv<T, V>(ab + (size_t)(p + u * PL) * lda + cv * V, yv[u]);
          }
        }
      }
      for (; p < p1; p += PL) {
        float gv[V], yv[V];
        ldv<T, V>(gb + (size_t)p * ldg + cv * V, gv);
        ldv<T, V>(yb + (size_t)p * ldy + cv * V, yv);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float xh = (yv[i] - mu[i]) * rs[i];
          float a = ga[i] * xh + be[i];
          float gm = (relu && a <= 0.f) ? 0.f : gv[i];
          s1[i] += gm;
          s2[i] += gm * xh;
        }
        if (REMAT) {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            float sc, sh;
            norm_scale_shift(ga[i], be[i], mu[i], rs[i], &sc, &sh);
            yv[i] = norm_act1(yv[i], sc, sh, relu);
          }
          stv<T, V>(ab + (size_t)p * lda + cv * V, yv);
        }
      }
      float* d = sm + ((size_t)lane_p * C + cv * V) * 2;
#pragma unroll
      for (int i = 0; i < V; ++i) { d[2 * i] = s1[i]; d[2 * i + 1] = s2[i]; }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += RED_THREADS) {
    float a = 0.f;
    for (int l = 0; l < PL; ++l) a += sm[(size_t)l * C * 2 + i];
    // per_sample == 0 (training-mode batch norm): only the batch totals [C][2] are needed downstream
    atomicAdd(&sums[(per_sample ? (size_t)n * C * 2 : 0) + i], (double)a);
  }
}

// chunks per sample for the grid = (chunks, N) streaming kernels: the largest count whose N * chunks blocks are all
// resident at once (`slots` = SMs x blocks per SM for the kernel's register use): a single wave, no straggler blocks
// PHS_NORM_BPS (tuning): blocks per SM assumed by the one-wave grids of the normalisation kernels (0 = each kernel's own)
static int norm_bps_override() {
  const char* e = getenv("PHS_NORM_BPS");
  return e ? atoi(e) : 0;
}

// PHS_NORM_RAW=0: the pre-round-2 load scheduling of the backward normalisation kernels (A/B switch, read per call)
static bool norm_raw_loads() {
  const char* e = getenv("PHS_NORM_RAW");
  return !(e && e[0] == '0');
}

static int pick_chunks(int N, int slots, int max_chunks) {
  if (norm_bps_override() > 0) slots = 148 * norm_bps_override();
  int chunks = slots / N;
  if (chunks > max_chunks) chunks = max_chunks;
  return chunks < 1 ? 1 : chunks;
}

static void red_geometry(int N, int HW, int C, int V, dim3* grid, int* ppb, size_t* smem, int blocks_per_sm = 3) {
  int nvec = C / V;
  int CW = nvec < RED_THREADS ? nvec : RED_THREADS;
  int PL = RED_THREADS / CW;
  // aim at >= ~4 blocks per SM overall while keeping >= 64 pixels per pixel-lane where possible
  int max_chunks = (HW + PL * 8 - 1) / (PL * 8);
  int chunks = pick_chunks(N, 148 * blocks_per_sm, max_chunks);     // 80 registers: three blocks per SM
  *ppb = (HW + chunks - 1) / chunks;
  chunks = (HW + *ppb - 1) / *ppb;
  *grid = dim3(chunks, N);
  *smem = (size_t)PL * C * 2 * sizeof(float);
}

// stats[N][C][2] (+)= per-sample sums; with_totals: the [C][2] batch totals that follow them are accumulated too
// (layout of phs_conv2d_stats_acc); zero_first: clear the per-sample part (the plain phs_chan_stats contract)
int chan_stats_run(const phs_tensor* y, double* stats, bool with_totals, bool zero_first, cudaStream_t st) {
  int HW = y->H * y->W;
  if (zero_first) cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)y->N * y->C, st);
  int v = pick_vec(y);
  dim3 grid; int ppb; size_t smem;
  red_geometry(y->N, HW, y->C, v, &grid, &ppb, &smem);
  PHS_REQUIRE(smem <= 48 * 1024, "phs_chan_stats: C=%d too large", y->C);
  double* totals = with_totals ? stats + (size_t)y->N * y->C * 2 : nullptr;
  PHS_DISPATCH_DTYPE(y->dtype, T, PHS_DISPATCH_VEC(v, V, (phs_launch(chan_stats_kernel<T, V>, grid, RED_THREADS, smem, st, 
                                                            (const T*)y->ptr, HW, y->C, y->ld, ppb, stats, totals))));
  return phs_check_launch("chan_stats");
}

int phs_chan_stats(const phs_tensor* y, double* stats, void* stream) {
  PHS_REQUIRE(y && y->ptr && stats, "phs_chan_stats: null argument");
  return chan_stats_run(y, stats, false, true, (cudaStream_t)stream);
}

static int norm_bwd_reduce_run(const char* fn, const phs_tensor* g, const phs_tensor* y, const float* mean, const float* rstd,
                               const float* gamma, const float* beta, int relu, double* sums, const phs_tensor* a,
                               void* stream) {
  PHS_REQUIRE(g && y && g->ptr && y->ptr && sums, "%s: null argument", fn);
  PHS_REQUIRE(g->dtype == y->dtype && g->N == y->N && g->H == y->H && g->W == y->W && g->C == y->C, "%s: g/y mismatch", fn);
  PHS_REQUIRE(!a || (a->ptr && a->dtype == y->dtype && a->N == y->N && a->H == y->H && a->W == y->W && a->C == y->C),
              "%s: a/y mismatch", fn);
  cudaStream_t st = (cudaStream_t)stream;
  int HW = y->H * y->W;
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)y->N * y->C, st);
  int v = min_vec(pick_vec(y), pick_vec(g));
  if (a) v = min_vec(v, pick_vec(a));
  dim3 grid; int ppb; size_t smem;
  const bool rawl = !a && v == 8 && norm_raw_loads();
  red_geometry(y->N, HW, y->C, v, &grid, &ppb, &smem, rawl ? 2 : 3);
  PHS_REQUIRE(smem <= 48 * 1024, "%s: C=%d too large", fn, y->C);
  if (rawl) {
    PHS_DISPATCH_DTYPE(y->dtype, T, (phs_launch(norm_bwd_reduce_kernel<T, 8, false, true>, grid, RED_THREADS, smem, st,
                                                (const T*)g->ptr, g->ld, (const T*)y->ptr, y->ld, HW, y->C, ppb, mean, rstd,
                                                gamma, beta, relu, sums, 1, (T*)nullptr, 0)));
  } else if (a) {
    PHS_DISPATCH_DTYPE(y->dtype, T,
                       PHS_DISPATCH_VEC(v, V, (phs_launch(norm_bwd_reduce_kernel<T, V, true>, grid, RED_THREADS, smem, st,
                                                  (const T*)g->ptr, g->ld, (const T*)y->ptr, y->ld, HW, y->C, ppb, mean,
                                                  rstd, gamma, beta, relu, sums, 1, (T*)a->ptr, a->ld))));
  } else {
    PHS_DISPATCH_DTYPE(y->dtype, T,
                       PHS_DISPATCH_VEC(v, V, (phs_launch(norm_bwd_reduce_kernel<T, V, false>, grid, RED_THREADS, smem, st,
                                                  (const T*)g->ptr, g->ld, (const T*)y->ptr, y->ld, HW, y->C, ppb, mean,
                                                  rstd, gamma, beta, relu, sums, 1, (T*)nullptr, 0))));
  }
  return phs_check_launch(fn);
}

int phs_norm_bwd_reduce(const phs_tensor* g, const phs_tensor* y, const float* mean, const float* rstd,
                        const float* gamma, const float* beta, int relu, double* sums, void* stream) {
  return norm_bwd_reduce_run("phs_norm_bwd_reduce", g, y, mean, rstd, gamma, beta, relu, sums, nullptr, stream);
}

int phs_norm_bwd_reduce_remat(const phs_tensor* g, const phs_tensor* y, const float* mean, const float* rstd,
                              const float* gamma, const float* beta, int relu, double* sums, const phs_tensor* a,
                              void* stream) {
  PHS_REQUIRE(a, "phs_norm_bwd_reduce_remat: null activation output");
  return norm_bwd_reduce_run("phs_norm_bwd_reduce_remat", g, y, mean, rstd, gamma, beta, relu, sums, a, stream);
}

int phs_norm_bwd_reduce_bn(const phs_tensor* g, const phs_tensor* y, const float* mean, const float* rstd,
                           const float* gamma, const float* beta, int relu, double* totals, void* stream) {
  PHS_REQUIRE(g && y && g->ptr && y->ptr && totals, "phs_norm_bwd_reduce_bn: null argument");
  PHS_REQUIRE(g->dtype == y->dtype && g->N == y->N && g->H == y->H && g->W == y->W && g->C == y->C,
              "phs_norm_bwd_reduce_bn: g/y mismatch");
  cudaStream_t st = (cudaStream_t)stream;
  int HW = y->H * y->W;
  int v = min_vec(pick_vec(y), pick_vec(g));
  dim3 grid; int ppb; size_t smem;
  red_geometry(y->N, HW, y->C, v, &grid, &ppb, &smem);
  PHS_REQUIRE(smem <= 48 * 1024, "phs_norm_bwd_reduce_bn: C=%d too large", y->C);
  PHS_DISPATCH_DTYPE(y->dtype, T,
                     PHS_DISPATCH_VEC(v, V, (phs_launch(norm_bwd_reduce_kernel<T, V, false>, grid, RED_THREADS, smem, st, 
                                                (const T*)g->ptr, g->ld, (const T*)y->ptr, y->ld, HW, y->C, ppb, mean,
                                                rstd, gamma, beta, relu, totals, 0, (T*)nullptr, 0))));
  return phs_check_launch("norm_bwd_reduce_bn");
}

__global__ void __launch_bounds__(128) norm_finalize_kernel(const double* __restrict__ stats, int N, int HW, int C,
                                                            int mode, float eps, float decay, float* moving_mean,
                                                            float* moving_var, float* __restrict__ mean,
                                                            float* __restrict__ rstd) {
  PHS_PDL_PROLOGUE();
  if (mode == PHS_NORM_GN) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * C) return;
    int n = idx / C, c = idx % C;
    int G = max(2, C / 16);
    int cpg = C / G;
    int c0 = (c / cpg) * cpg;
    double s = 0.0, q = 0.0;
    for (int i = 0; i < cpg; ++i) {
      s += stats[((size_t)n * C + c0 + i) * 2];
      q += stats[((size_t)n * C + c0 + i) * 2 + 1];
    }
    double cnt = (double)HW * cpg;
    double m = s / cnt;
    double var = q / cnt - m * m;
    if (var < 0) var = 0;
    mean[idx] = (float)m;
    rstd[idx] = (float)(1.0 / sqrt(var + (double)eps));
    return;
  }
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  float m, r;
  if (mode == PHS_NORM_BN_TRAIN) {
    double s = 0.0, q = 0.0;
    for (int n = lane; n < N; n += 32) {
      s += stats[((size_t)n * C + c) * 2];
      q += stats[((size_t)n * C + c) * 2 + 1];
    }
    s = warp_sum_d(s);
    q = warp_sum_d(q);
    double cnt = (double)HW * N;
    double mm = s / cnt;
    double var = q / cnt - mm * mm;
    if (var < 0) var = 0;
    m = (float)mm;
    r = (float)(1.0 / sqrt(var + (double)eps));
    if (moving_mean && lane == 0) {
      double unb = var * (cnt / (cnt > 1 ? cnt - 1 : 1));
      moving_mean[c] = decay * moving_mean[c] + (1.f - decay) * m;
      moving_var[c] = decay * moving_var[c] + (1.f - decay) * (float)unb;
    }
  } else {
    m = moving_mean[c];
    r = rsqrtf(moving_var[c] + eps);
  }
  for (int n = lane; n < N; n += 32) {
    mean[(size_t)n * C + c] = m;
    rstd[(size_t)n * C + c] = r;
  }
}

int phs_norm_finalize(const double* stats, int N, int HW, int C, int mode, float eps, float decay, float* moving_mean,
                      float* moving_var, float* mean, float* rstd, void* stream) {
  PHS_REQUIRE(mean && rstd, "phs_norm_finalize: null output");
  PHS_REQUIRE(mode == PHS_NORM_BN_INFER || stats, "phs_norm_finalize: stats required");
  PHS_REQUIRE(mode != PHS_NORM_BN_INFER || (moving_mean && moving_var), "phs_norm_finalize: moving stats required");
  PHS_REQUIRE(mode != PHS_NORM_GN || C % max(2, C / 16) == 0, "phs_norm_finalize: C=%d not divisible into groups", C);
  int blocks = mode == PHS_NORM_GN ? (N * C + 127) / 128 : (C + 3) / 4;
  phs_launch(norm_finalize_kernel, blocks, 128, 0, (cudaStream_t)stream, stats, N, HW, C, mode, eps, decay, moving_mean,
                                                                 moving_var, mean, rstd);
  return phs_check_launch("norm_finalize");
}

// backward finalize.  With g' = g*mask:  s1 = sum g', s2 = sum g'*xhat (per n,c).
//   dbeta[c] = sum_n s1, dgamma[c] = sum_n s2
//   BN: m1[c] = gamma*sum_n s1/(N*HW), m2[c] = gamma*sum_n s2/(N*HW)
//   GN: m1[n,grp] = sum_{c in grp} gamma_c*s1/(cpg*HW), m2 likewise
//   dx = rstd*(g'*gamma - m1 - xhat*m2);   dbias[c] = sum_{n,hw} dx  (closed form from the forward sums)
// one warp per channel, lanes stride over the samples
__global__ void __launch_bounds__(128)
    norm_bwd_finalize_kernel(const double* __restrict__ sums, const double* __restrict__ stats,
                             const float* __restrict__ mean, const float* __restrict__ rstd,
                             const float* __restrict__ gamma, int N, int HW, int C, int mode, float* __restrict__ coef,
                             float* dgamma, float* dbeta, float* dbias, int accumulate) {
  PHS_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  double a1 = 0.0, a2 = 0.0, db = 0.0;
  for (int n = lane; n < N; n += 32) {
    a1 += sums[((size_t)n * C + c) * 2];
    a2 += sums[((size_t)n * C + c) * 2 + 1];
  }
  a1 = warp_sum_d(a1);
  a2 = warp_sum_d(a2);
  const float ga = gamma[c];
  if (mode == PHS_NORM_GN) {
    const int G = max(2, C / 16);
    const int cpg = C / G;
    const int c0 = (c / cpg) * cpg;
    for (int n = lane; n < N; n += 32) {
      double m1 = 0.0, m2 = 0.0;
      for (int i = 0; i < cpg; ++i) {
        double gi = gamma[c0 + i];
        m1 += (double)gi * sums[((size_t)n * C + c0 + i) * 2];
        m2 += (double)gi * sums[((size_t)n * C + c0 + i) * 2 + 1];
      }
      double cnt = (double)HW * cpg;
      m1 /= cnt;
      m2 /= cnt;
      coef[((size_t)n * C + c) * 2] = (float)m1;
      coef[((size_t)n * C + c) * 2 + 1] = (float)m2;
      if (dbias) {
        double r = rstd[(size_t)n * C + c], mu = mean[(size_t)n * C + c];
        double sx = (stats[((size_t)n * C + c) * 2] - (double)HW * mu) * r;  // sum_hw xhat
        db += r * ((double)ga * sums[((size_t)n * C + c) * 2] - (double)HW * m1 - m2 * sx);
      }
    }
    db = warp_sum_d(db);
  } else {
    double cnt = (double)HW * N;
    double m1 = ga * a1 / cnt, m2 = ga * a2 / cnt;
    for (int n = lane; n < N; n += 32) {
      coef[((size_t)n * C + c) * 2] = (float)m1;
      coef[((size_t)n * C + c) * 2 + 1] = (float)m2;
    }
    db = 0.0;  // batch norm removes any per-channel offset: the bias gradient is exactly zero
  }
  if (lane == 0) {
    if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)a2;
    if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)a1;
    if (dbias) dbias[c] = (accumulate ? dbias[c] : 0.f) + (float)db;
  }
}

int phs_norm_bwd_finalize(const double* sums, const double* stats, const float* mean, const float* rstd,
                          const float* gamma, int N, int HW, int C, int mode, float* coef, float* dgamma, float* dbeta,
                          float* dbias, int accumulate, void* stream) {
  PHS_REQUIRE(sums && mean && rstd && gamma && coef, "phs_norm_bwd_finalize: null argument");
  PHS_REQUIRE(!dbias || stats, "phs_norm_bwd_finalize: dbias needs the forward stats");
  PHS_REQUIRE(mode == PHS_NORM_GN || mode == PHS_NORM_BN_TRAIN, "phs_norm_bwd_finalize: mode %d has no backward", mode);
  phs_launch(norm_bwd_finalize_kernel, (C + 3) / 4, 128, 0, (cudaStream_t)stream, sums, stats, mean, rstd, gamma, N, HW, C, mode,
                                                                        coef, dgamma, dbeta, dbias, accumulate);
  return phs_check_launch("norm_bwd_finalize");
}

// ---------------------------------------------------------------------------------------------------------
// streaming kernels: one thread per (pixel, channel-vector)
// ---------------------------------------------------------------------------------------------------------
// grid = (chunks, N); a block walks a contiguous run of pixels of ONE sample with CW channel-vector lanes x PL pixel
// lanes; every thread keeps the normalisation parameters of its V channels in registers and has U independent
// 16-byte loads in flight.
constexpr int STREAM_U = 4;

static void stream_geometry(int N, int HW, int C, int V, int blocks_per_sm, dim3* grid, int* ppb) {
  int nvec = C / V;
  int CW = nvec < 256 ? nvec : 256;
  int PL = 256 / CW;
  int max_chunks = (HW + PL * STREAM_U - 1) / (PL * STREAM_U);
  int chunks = pick_chunks(N, 148 * blocks_per_sm, max_chunks);
  *ppb = (HW + chunks - 1) / chunks;
  chunks = (HW + *ppb - 1) / *ppb;
  *grid = dim3(chunks, N);
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
    norm_act_fwd_kernel(const T* __restrict__ y, int ldy, T* __restrict__ a, int lda, int HW, int C, int pix_per_block,
                        const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ gamma, const float* __restrict__ beta, int relu) {
  PHS_PDL_PROLOGUE();
  const int n = blockIdx.y;
  const int nvec = C / V;
  const int CW = nvec < 256 ? nvec : 256;
  const int PL = 256 / CW;
  const int lane_c = threadIdx.x % CW, lane_p = threadIdx.x / CW;
  if (lane_p >= PL) return;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const T* yb = y + (size_t)n * HW * ldy;
  T* ab = a + (size_t)n * HW * lda;
  for (int cv = lane_c; cv < nvec; cv += CW) {
    float sc[V], sh[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      int c = cv * V + k;
      norm_scale_shift(gamma[c], beta[c], mean[(size_t)n * C + c], rstd[(size_t)n * C + c], &sc[k], &sh[k]);
    }
    int p = p0 + lane_p;
    for (; p + (STREAM_U - 1) * PL < p1; p += STREAM_U * PL) {
      float v[STREAM_U][V];
#pragma unroll
      for (int u = 0; u < STREAM_U; ++u) ldv<T, V>(yb + (size_t)(p + u * PL) * ldy + cv * V, v[u]);
#pragma unroll
      for (int u = 0; u < STREAM_U; ++u) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          v[u][k] = norm_act1(v[u][k], sc[k], sh[k], relu);
        }
        stv<T, V>(ab + (size_t)(p + u * PL) * lda + cv * V, v[u]);
      }
    }
    for (; p < p1; p += PL) {
      float v[V];
      ldv<T, V>(yb + (size_t)p * ldy + cv * V, v);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        v[k] = norm_act1(v[k], sc[k], sh[k], relu);
      }
      stv<T, V>(ab + (size_t)p * lda + cv * V, v);
    }
  }
}

// norm_finalize folded into norm_act_fwd: every thread derives mean / rstd of its V channels from the statistics the
// convolution epilogue left behind (batch norm: the [C][2] batch totals behind stats[N][C][2]; group norm: the sums of
// its group of sample n), so the forward chain of a layer is conv -> this kernel.  The blocks with blockIdx.x == 0 also
// write mean/rstd[N][C] for the backward kernels, block (0, 0) updates the batch-norm moving averages.
template <typename T, int V>
__global__ void __launch_bounds__(256)
    norm_act_fwd_stats_kernel(const T* __restrict__ y, int ldy, T* __restrict__ a, int lda, int HW, int C,
                              int pix_per_block, const double* __restrict__ stats, int mode, float eps, float decay,
                              float* moving_mean, float* moving_var, float* __restrict__ mean_out,
                              float* __restrict__ rstd_out, const float* __restrict__ gamma,
                              const float* __restrict__ beta, int relu) {
  PHS_PDL_PROLOGUE();
  const int n = blockIdx.y, N = gridDim.y;
  const int nvec = C / V;
  const int CW = nvec < 256 ? nvec : 256;
  const int PL = 256 / CW;
  const int lane_c = threadIdx.x % CW, lane_p = threadIdx.x / CW;
  if (lane_p >= PL) return;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const T* yb = y + (size_t)n * HW * ldy;
  T* ab = a + (size_t)n * HW * lda;
  const double* totals = stats + (size_t)N * C * 2;
  const int G = max(2, C / 16), cpg = C / G;
  for (int cv = lane_c; cv < nvec; cv += CW) {
    float sc[V], sh[V];
    int g_cached = -1;
    double g_m = 0.0, g_r = 0.0;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int c = cv * V + k;
      double m, r, var;
      if (mode == PHS_NORM_GN) {
        const int grp = c / cpg;
        if (grp != g_cached) {
          double s = 0.0, q = 0.0;
          const double* gs = stats + ((size_t)n * C + (size_t)grp * cpg) * 2;
          for (int i = 0; i < cpg; ++i) { s += gs[2 * i]; q += gs[2 * i + 1]; }
          norm_moments(s, q, (double)HW * cpg, eps, &g_m, &var, &g_r);
          g_cached = grp;
        }
        m = g_m; r = g_r;
      } else {
        const double cnt = (double)HW * N;
        norm_moments(totals[2 * c], totals[2 * c + 1], cnt, eps, &m, &var, &r);
        if (moving_mean && blockIdx.x == 0 && n == 0 && lane_p == 0) bn_moving_update(moving_mean, moving_var, c, decay, m, var, cnt);
      }
      if (blockIdx.x == 0 && lane_p == 0) {
        mean_out[(size_t)n * C + c] = (float)m;
        rstd_out[(size_t)n * C + c] = (float)r;
      }
      norm_scale_shift(gamma[c], beta[c], (float)m, (float)r, &sc[k], &sh[k]);
    }
    int p = p0 + lane_p;
    for (; p + (STREAM_U - 1) * PL < p1; p += STREAM_U * PL) {
      float v[STREAM_U][V];
#pragma unroll
      for (int u = 0; u < STREAM_U; ++u) ldv<T, V>(yb + (size_t)(p + u * PL) * ldy + cv * V, v[u]);
#pragma unroll
      for (int u = 0; u < STREAM_U; ++u) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          v[u][k] = norm_act1(v[u][k], sc[k], sh[k], relu);
        }
        stv<T, V>(ab + (size_t)(p + u * PL) * lda + cv * V, v[u]);
      }
    }
    for (; p < p1; p += PL) {
      float v[V];
      ldv<T, V>(yb + (size_t)p * ldy + cv * V, v);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        v[k] = norm_act1(v[k], sc[k], sh[k], relu);
      }
      stv<T, V>(ab + (size_t)p * lda + cv * V, v);
    }
  }
}

int phs_norm_act_fwd_stats(const phs_tensor* y, const double* stats, int mode, float eps, float decay, float* moving_mean,
                           float* moving_var, float* mean, float* rstd, const float* gamma, const float* beta, int relu,
                           const phs_tensor* a, void* stream) {
  PHS_REQUIRE(y && a && y->ptr && a->ptr && stats && mean && rstd && gamma && beta, "phs_norm_act_fwd_stats: null argument");
  PHS_REQUIRE(y->dtype == a->dtype && y->N == a->N && y->H == a->H && y->W == a->W && y->C == a->C,
              "phs_norm_act_fwd_stats: y/a mismatch");
  PHS_REQUIRE(mode == PHS_NORM_GN || mode == PHS_NORM_BN_TRAIN, "phs_norm_act_fwd_stats: mode %d needs phs_norm_finalize", mode);
  PHS_REQUIRE(mode != PHS_NORM_GN || y->C % max(2, y->C / 16) == 0, "phs_norm_act_fwd_stats: C=%d not divisible into groups", y->C);
  int v = min_vec(pick_vec(y), pick_vec(a));
  int HW = y->H * y->W;
  dim3 grid; int ppb;
  stream_geometry(y->N, HW, y->C, v, 3, &grid, &ppb);   // 77 registers: three blocks per SM
  PHS_DISPATCH_DTYPE(y->dtype, T,
                     PHS_DISPATCH_VEC(v, V, (phs_launch(norm_act_fwd_stats_kernel<T, V>, grid, 256, 0, (cudaStream_t)stream, 
                                                (const T*)y->ptr, y->ld, (T*)a->ptr, a->ld, HW, y->C, ppb, stats, mode, eps,
                                                decay, moving_mean, moving_var, mean, rstd, gamma, beta, relu))));
  return phs_check_launch("norm_act_fwd_stats");
}

int phs_norm_act_fwd(const phs_tensor* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                     int relu, const phs_tensor* a, void* stream) {
  PHS_REQUIRE(y && a && y->ptr && a->ptr && mean && rstd && gamma && beta, "phs_norm_act_fwd: null argument");
  PHS_REQUIRE(y->dtype == a->dtype && y->N == a->N && y->H == a->H && y->W == a->W && y->C == a->C,
              "phs_norm_act_fwd: y/a mismatch");
  int v = min_vec(pick_vec(y), pick_vec(a));
  int HW = y->H * y->W;
  dim3 grid; int ppb;
  stream_geometry(y->N, HW, y->C, v, 6, &grid, &ppb);
  PHS_DISPATCH_DTYPE(y->dtype, T,
                     PHS_DISPATCH_VEC(v, V, (phs_launch(norm_act_fwd_kernel<T, V>, grid, 256, 0, (cudaStream_t)stream, 
                                                (const T*)y->ptr, y->ld, (T*)a->ptr, a->ld, HW, y->C, ppb, mean, rstd,
                                                gamma, beta, relu))));
  return phs_check_launch("norm_act_fwd");
}

// batch-norm shortcut of norm_bwd_apply_kernel: batch totals [C][2] from phs_norm_bwd_reduce_bn instead of coef[N][C][2]
struct BnTotals {
  const double* totals;   // nullptr: use coef
  double count;           // N * H * W
  float* dgamma;
  float* dbeta;
  int accumulate;
};

// RAWL (bf16 / float, V == 8): the STREAM_U pixel vectors of g and y stay PACKED in registers until they are used, so all
// 2 * STREAM_U loads are issued back to back (with unpacked fp32 copies next to the 7 x V per-channel coefficients ptxas
// serialised the loads in pairs under the 85-register cap: 2 loads in flight per thread); two blocks per SM.
template <typename T, int V, bool RAWL>
__global__ void __launch_bounds__(256, RAWL ? 2 : 3)
    norm_bwd_apply_kernel(const T* __restrict__ g, int ldg, const T* __restrict__ y, int ldy, T* __restrict__ dy,
                          int lddy, int HW, int C, int pix_per_block, const float* __restrict__ mean,
                          const float* __restrict__ rstd, const float* __restrict__ gamma,
                          const float* __restrict__ beta, int relu, const float* __restrict__ coef,
                          BnTotals bn) {
  PHS_PDL_PROLOGUE();
  const int n = blockIdx.y;
  const int nvec = C / V;
  const int CW = nvec < 256 ? nvec : 256;
  const int PL = 256 / CW;
  const int lane_c = threadIdx.x % CW, lane_p = threadIdx.x / CW;
  if (lane_p >= PL) return;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const T* gb = g + (size_t)n * HW * ldg;
  const T* yb = y + (size_t)n * HW * ldy;
  T* db = dy + (size_t)n * HW * lddy;
  for (int cv = lane_c; cv < nvec; cv += CW) {
    // a = sc*y + sh (activation input, for the ReLU mask);  xhat = rs*y + xo;  dy = gr*g' - k1 - xhat*k2
    float sc[V], sh[V], rs[V], xo[V], gr[V], k1[V], k2[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      int c = cv * V + k;
      size_t nc = (size_t)n * C + c;
      rs[k] = rstd[nc];
      xo[k] = -mean[nc] * rs[k];
      sc[k] = gamma[c] * rs[k];
      sh[k] = beta[c] + gamma[c] * xo[k];
      gr[k] = gamma[c] * rs[k];
      if (bn.totals) {
        // training-mode batch norm: the finalize step is two multiplications per channel, done here from the batch
        // totals (same expressions and rounding as norm_bwd_finalize_kernel); block (0, 0) also owns dgamma / dbeta
        const double a1 = bn.totals[2 * c], a2 = bn.totals[2 * c + 1];
        k1[k] = rs[k] * (float)(gamma[c] * a1 / bn.count);
        k2[k] = rs[k] * (float)(gamma[c] * a2 / bn.count);
        if (blockIdx.x == 0 && blockIdx.y == 0 && lane_p == 0) {
          if (bn.dgamma) bn.dgamma[c] = (bn.accumulate ? bn.dgamma[c] : 0.f) + (float)a2;
          if (bn.dbeta) bn.dbeta[c] = (bn.accumulate ? bn.dbeta[c] : 0.f) + (float)a1;
        }
      } else {
        k1[k] = rs[k] * coef[nc * 2];
        k2[k] = rs[k] * coef[nc * 2 + 1];
      }
    }
    int p = p0 + lane_p;
    if constexpr (RAWL) {
      static_assert(V == 8, "packed loads are 8-channel vectors");
      for (; p + (STREAM_U - 1) * PL < p1; p += STREAM_U * PL) {
        Raw8<T> graw[STREAM_U], yraw[STREAM_U];
#pragma unroll
        for (int u = 0; u < STREAM_U; ++u) {
          graw[u].load(gb + (size_t)(p + u * PL) * ldg + cv * V);
          yraw[u].load(yb + (size_t)(p + u * PL) * ldy + cv * V);
        }
#pragma unroll
        for (int u = 0; u < STREAM_U; ++u) {
          float gv[V], yv[V];
          graw[u].unpack(gv);
          yraw[u].unpack(yv);
#pragma unroll
          for (int k = 0; k < V; ++k) {
            float xh = fmaf(yv[k], rs[k], xo[k]);
            float aa = fmaf(yv[k], sc[k], sh[k]);
            float gm = (relu && aa <= 0.f) ? 0.f : gv[k];
            gv[k] = gm * gr[k] - k1[k] - xh * k2[k];
          }
          stv<T, V>(db + (size_t)(p + u * PL) * lddy + cv * V, gv);
        }
      }
    }
    for (; p + (STREAM_U - 1) * PL < p1; p += STREAM_U * PL) {
      float gv[STREAM_U][V], yv[STREAM_U][V];
#pragma unroll
      for (int u = 0; u < STREAM_U; ++u) {
        ldv<T, V>(gb + (size_t)(p + u * PL) * ldg + cv * V, gv[u]);
        ldv<T, V>(yb + (size_t)(p + u * PL) * ldy + cv * V, yv[u]);
      }
#pragma unroll
      for (int u = 0; u < STREAM_U; ++u) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          float xh = fmaf(yv[u][k], rs[k], xo[k]);
          float aa = fmaf(yv[u][k], sc[k], sh[k]);
          float gm = (relu && aa <= 0.f) ? 0.f : gv[u][k];
          gv[u][k] = gm * gr[k] - k1[k] - xh * k2[k];
        }
        stv<T, V>(db + (size_t)(p + u * PL) * lddy + cv * V, gv[u]);
      }
    }
    for (; p < p1; p += PL) {
      float gv[V], yv[V];
      ldv<T, V>(gb + (size_t)p * ldg + cv * V, gv);
      ldv<T, V>(yb + (size_t)p * ldy + cv * V, yv);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        float xh = fmaf(yv[k], rs[k], xo[k]);
        float aa = fmaf(yv[k], sc[k], sh[k]);
        float gm = (relu && aa <= 0.f) ? 0.f : gv[k];
        gv[k] = gm * gr[k] - k1[k] - xh * k2[k];
      }
      stv<T, V>(db + (size_t)p * lddy + cv * V, gv);
    }
  }
}

static int norm_bwd_apply_run(const char* what, const phs_tensor* g, const phs_tensor* y, const float* mean,
                              const float* rstd, const float* gamma, const float* beta, int relu, const float* coef,
                              BnTotals bn, const phs_tensor* dy, void* stream) {
  PHS_REQUIRE(g && y && dy && g->ptr && y->ptr && dy->ptr && (coef || bn.totals), "%s: null argument", what);
  PHS_REQUIRE(g->dtype == y->dtype && dy->dtype == y->dtype && g->C == y->C && dy->C == y->C && g->N == y->N &&
                  g->H == y->H && g->W == y->W,
              "%s: tensor mismatch", what);
  int v = min_vec(min_vec(pick_vec(y), pick_vec(g)), pick_vec(dy));
  int HW = y->H * y->W;
  bn.count = (double)HW * y->N;
  dim3 grid; int ppb;
  const bool rawl = v == 8 && norm_raw_loads();
  stream_geometry(y->N, HW, y->C, v, rawl ? 2 : 3, &grid, &ppb);
  if (rawl) {
    PHS_DISPATCH_DTYPE(y->dtype, T, (phs_launch(norm_bwd_apply_kernel<T, 8, true>, grid, 256, 0, (cudaStream_t)stream,
                                                (const T*)g->ptr, g->ld, (const T*)y->ptr, y->ld, (T*)dy->ptr, dy->ld,
                                                HW, y->C, ppb, mean, rstd, gamma, beta, relu, coef, bn)));
  } else {
    PHS_DISPATCH_DTYPE(y->dtype, T,
                       PHS_DISPATCH_VEC(v, V, (phs_launch(norm_bwd_apply_kernel<T, V, false>, grid, 256, 0, (cudaStream_t)stream,
                                                  (const T*)g->ptr, g->ld, (const T*)y->ptr, y->ld, (T*)dy->ptr, dy->ld,
                                                  HW, y->C, ppb, mean, rstd, gamma, beta, relu, coef, bn))));
  }
  return phs_check_launch(what);
}

int phs_norm_bwd_apply(const phs_tensor* g, const phs_tensor* y, const float* mean, const float* rstd,
                       const float* gamma, const float* beta, int relu, const float* coef, const phs_tensor* dy,
                       void* stream) {
  return norm_bwd_apply_run("phs_norm_bwd_apply", g, y, mean, rstd, gamma, beta, relu, coef,
                            BnTotals{nullptr, 0.0, nullptr, nullptr, 0}, dy, stream);
}

int phs_norm_bwd_apply_bn(const phs_tensor* g, const phs_tensor* y, const float* mean, const float* rstd,
                          const float* gamma, const float* beta, int relu, const double* totals, const phs_tensor* dy,
                          float* dgamma, float* dbeta, int accumulate, void* stream) {
  return norm_bwd_apply_run("phs_norm_bwd_apply_bn", g, y, mean, rstd, gamma, beta, relu, nullptr,
                            BnTotals{totals, 0.0, dgamma, dbeta, accumulate}, dy, stream);
}

// ---- 2x2 average pool ------------------------------------------------------------------------------------
// All the resampling kernels below are one 16-byte vector per thread; their index arithmetic is multiply-high
// (idx4_decode) because 64-bit divisions made them ALU bound at ~2 TB/s.
static int stream_blocks(int64_t total) {
  int64_t b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
    avgpool2_fwd_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy, int Ho, int Wo, idx4_t ix,
                        uint32_t total) {
  PHS_PDL_PROLOGUE();
  const int Wi = Wo * 2;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int cv, wo, ho, n;
    idx4_decode(i, ix, cv, wo, ho, n);
    const T* p = x + (((int64_t)n * (2 * Ho) + 2 * ho) * Wi + 2 * wo) * (int64_t)ldx + cv * V;
    float a[V], b[V], c[V], d[V], o[V];
    ldv<T, V>(p, a);
    ldv<T, V>(p + ldx, b);
    ldv<T, V>(p + (int64_t)Wi * ldx, c);
    ldv<T, V>(p + (int64_t)Wi * ldx + ldx, d);
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = 0.25f * ((a[k] + b[k]) + (c[k] + d[k]));
    stv<T, V>(y + (((int64_t)n * Ho + ho) * Wo + wo) * (int64_t)ldy + cv * V, o);
  }
}

// one thread per pooled pixel: its gradient goes to the 2x2 input block
template <typename T, int V>
__global__ void __launch_bounds__(256)
    avgpool2_bwd_kernel(const T* __restrict__ dy, int lddy, T* __restrict__ dx, int lddx, int Ho, int Wo, idx4_t ix,
                        uint32_t total, int accumulate) {
  PHS_PDL_PROLOGUE();
  const int Wi = Wo * 2;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int cv, wo, ho, n;
    idx4_decode(i, ix, cv, wo, ho, n);
    float g[V];
    ldv<T, V>(dy + (((int64_t)n * Ho + ho) * Wo + wo) * (int64_t)lddy + cv * V, g);
#pragma unroll
    for (int k = 0; k < V; ++k) g[k] *= 0.25f;
    T* q = dx + (((int64_t)n * (2 * Ho) + 2 * ho) * Wi + 2 * wo) * (int64_t)lddx + cv * V;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      T* qq = q + ((r >> 1) * (int64_t)Wi + (r & 1)) * lddx;
      float o[V];
      if (accumulate) {
        ldv<T, V>(qq, o);
#pragma unroll
        for (int k = 0; k < V; ++k) o[k] += g[k];
      } else {
#pragma unroll
        for (int k = 0; k < V; ++k) o[k] = g[k];
      }
      stv<T, V>(qq, o);
    }
  }
}

int phs_avgpool2_fwd(const phs_tensor* x, const phs_tensor* y, void* stream) {
  PHS_REQUIRE(x && y && x->ptr && y->ptr, "phs_avgpool2_fwd: null argument");
  PHS_REQUIRE(x->dtype == y->dtype && x->N == y->N && x->C == y->C && x->H == 2 * y->H && x->W == 2 * y->W,
              "phs_avgpool2_fwd: shape mismatch (%d,%d,%d)->(%d,%d,%d)", x->H, x->W, x->C, y->H, y->W, y->C);
  int v = min_vec(pick_vec(x), pick_vec(y));
  int64_t total = (int64_t)y->N * y->H * y->W * (y->C / v);
  PHS_REQUIRE(total < (1ll << 31), "phs_avgpool2_fwd: tensor too large");
  const idx4_t ix = idx4_make(y->C / v, y->W, y->H);
  PHS_DISPATCH_DTYPE(x->dtype, T,
                     PHS_DISPATCH_VEC(v, V, (phs_launch(avgpool2_fwd_kernel<T, V>, stream_blocks(total), 256, 0, (cudaStream_t)stream, 
                                                (const T*)x->ptr, x->ld, (T*)y->ptr, y->ld, y->H, y->W, ix, (uint32_t)total))));
  return phs_check_launch("avgpool2_fwd");
}

int phs_avgpool2_bwd(const phs_tensor* dy, const phs_tensor* dx, int accumulate, void* stream) {
  PHS_REQUIRE(dx && dy && dx->ptr && dy->ptr, "phs_avgpool2_bwd: null argument");
  PHS_REQUIRE(dx->dtype == dy->dtype && dx->N == dy->N && dx->C == dy->C && dx->H == 2 * dy->H && dx->W == 2 * dy->W,
              "phs_avgpool2_bwd: shape mismatch");
  int v = min_vec(pick_vec(dx), pick_vec(dy));
  int64_t total = (int64_t)dy->N * dy->H * dy->W * (dy->C / v);
  PHS_REQUIRE(total < (1ll << 31), "phs_avgpool2_bwd: tensor too large");
  const idx4_t ix = idx4_make(dy->C / v, dy->W, dy->H);
  PHS_DISPATCH_DTYPE(dx->dtype, T,
                     PHS_DISPATCH_VEC(v, V, (phs_launch(avgpool2_bwd_kernel<T, V>, stream_blocks(total), 256, 0, (cudaStream_t)stream, 
                                                (const T*)dy->ptr, dy->ld, (T*)dx->ptr, dx->ld, dy->H, dy->W, ix, (uint32_t)total,
                                                accumulate))));
  return phs_check_launch("avgpool2_bwd");
}

// ---- TF1 legacy bilinear x2 ------------------------------------------------------------------------------
// forward: out[2k] = in[k]; out[2k+1] = 0.5*(in[k] + in[min(k+1,n-1)]) per axis.  One thread per INPUT pixel vector:
// four loads (the pixel, its right / lower / diagonal neighbours) make the 2x2 output block.
template <typename T, int V>
__global__ void __launch_bounds__(256)
    upsample2_fwd_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy, int Hi, int Wi, idx4_t ix,
                         uint32_t total) {
  PHS_PDL_PROLOGUE();
  const int Wo = 2 * Wi;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int cv, w0, h0, n;
    idx4_decode(i, ix, cv, w0, h0, n);
    const int h1 = min(h0 + 1, Hi - 1), w1 = min(w0 + 1, Wi - 1);
    const T* b = x + (int64_t)n * Hi * Wi * ldx + cv * V;
    float a[V], bb[V], c[V], d[V], o[V];
    ldv<T, V>(b + ((int64_t)h0 * Wi + w0) * ldx, a);
    ldv<T, V>(b + ((int64_t)h0 * Wi + w1) * ldx, bb);
    ldv<T, V>(b + ((int64_t)h1 * Wi + w0) * ldx, c);
    ldv<T, V>(b + ((int64_t)h1 * Wi + w1) * ldx, d);
    T* q = y + (((int64_t)n * (2 * Hi) + 2 * h0) * Wo + 2 * w0) * (int64_t)ldy + cv * V;
    // same operation order as the separable form: along w first (0.5*(l + r), exact when l == r), then along h
    stv<T, V>(q, a);
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = 0.5f * (a[k] + bb[k]);
    stv<T, V>(q + ldy, o);
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = 0.5f * (a[k] + c[k]);
    stv<T, V>(q + (int64_t)Wo * ldy, o);
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = 0.5f * (0.5f * (a[k] + bb[k]) + 0.5f * (c[k] + d[k]));
    stv<T, V>(q + (int64_t)Wo * ldy + ldy, o);
  }
}

// adjoint, gather form: per axis in[k] receives out[2k] (w 1), out[2k+1] (w .5, or 1 when k == n-1), out[2k-1] (w .5, k>=1)
template <typename T, int V>
__global__ void __launch_bounds__(256)
    upsample2_bwd_kernel(const T* __restrict__ dy, int lddy, T* __restrict__ dx, int lddx, int Hi, int Wi, idx4_t ix,
                         uint32_t total, int accumulate) {
  PHS_PDL_PROLOGUE();
  const int Ho = 2 * Hi, Wo = 2 * Wi;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int cv, w, h, n;
    idx4_decode(i, ix, cv, w, h, n);
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    const T* b = dy + (int64_t)n * Ho * Wo * lddy + cv * V;
#pragma unroll
    for (int dh = -1; dh <= 1; ++dh) {
      int ho = 2 * h + dh;
      if (ho < 0) continue;
      float wh = dh == 0 ? 1.f : (dh == 1 && h == Hi - 1 ? 1.f : 0.5f);
#pragma unroll
      for (int dw = -1; dw <= 1; ++dw) {
        int wo = 2 * w + dw;
        if (wo < 0) continue;
        float ww = dw == 0 ? 1.f : (dw == 1 && w == Wi - 1 ? 1.f : 0.5f);
        float g[V];
        ldv<T, V>(b + ((int64_t)ho * Wo + wo) * lddy, g);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += wh * ww * g[k];
      }
    }
    T* q = dx + (((int64_t)n * Hi + h) * Wi + w) * (int64_t)lddx + cv * V;
    if (accumulate) {
      float o[V];
      ldv<T, V>(q, o);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += o[k];
    }
    stv<T, V>(q, acc);
  }
}

int phs_upsample2_fwd(const phs_tensor* x, const phs_tensor* y, void* stream) {
  PHS_REQUIRE(x && y && x->ptr && y->ptr, "phs_upsample2_fwd: null argument");
  PHS_REQUIRE(x->dtype == y->dtype && x->N == y->N && x->C == y->C && y->H == 2 * x->H && y->W == 2 * x->W,
              "phs_upsample2_fwd: shape mismatch");
  int v = min_vec(pick_vec(x), pick_vec(y));
  int64_t total = (int64_t)x->N * x->H * x->W * (x->C / v);
  PHS_REQUIRE(total < (1ll << 31), "phs_upsample2_fwd: tensor too large");
  const idx4_t ix = idx4_make(x->C / v, x->W, x->H);
  PHS_DISPATCH_DTYPE(x->dtype, T,
                     PHS_DISPATCH_VEC(v, V, (phs_launch(upsample2_fwd_kernel<T, V>, stream_blocks(total), 256, 0, (cudaStream_t)stream, 
                                                (const T*)x->ptr, x->ld, (T*)y->ptr, y->ld, x->H, x->W, ix, (uint32_t)total))));
  return phs_check_launch("upsample2_fwd");
}

int phs_upsample2_bwd(const phs_tensor* dy, const phs_tensor* dx, int accumulate, void* stream) {
  PHS_REQUIRE(dx && dy && dx->ptr && dy->ptr, "phs_upsample2_bwd: null argument");
  PHS_REQUIRE(dx->dtype == dy->dtype && dx->N == dy->N && dx->C == dy->C && dy->H == 2 * dx->H && dy->W == 2 * dx->W,
              "phs_upsample2_bwd: shape mismatch");
  int v = min_vec(pick_vec(dx), pick_vec(dy));
  int64_t total = (int64_t)dx->N * dx->H * dx->W * (dx->C / v);
  PHS_REQUIRE(total < (1ll << 31), "phs_upsample2_bwd: tensor too large");
  const idx4_t ix = idx4_make(dx->C / v, dx->W, dx->H);
  PHS_DISPATCH_DTYPE(dx->dtype, T,
                     PHS_DISPATCH_VEC(v, V, (phs_launch(upsample2_bwd_kernel<T, V>, stream_blocks(total), 256, 0, (cudaStream_t)stream, 
                                                (const T*)dy->ptr, dy->ld, (T*)dx->ptr, dx->ld, dx->H, dx->W, ix, (uint32_t)total,
                                                accumulate))));
  return phs_check_launch("upsample2_bwd");
}

// ---- helpers ---------------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void copy_cast_kernel(const TS* __restrict__ s, int lds, TD* __restrict__ d, int ldd, int C, int64_t total,
                                 int64_t src_pix) {
  PHS_PDL_PROLOGUE();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t pix = i / C;
    stf<TD>(d + pix * ldd + c, ldf<TS>(s + (pix % src_pix) * lds + c));   // src_pix < all pixels: batch tiling
  }
}

int phs_copy_cast(const phs_tensor* src, const phs_tensor* dst, void* stream) {
  PHS_REQUIRE(src && dst && src->ptr && dst->ptr, "phs_copy_cast: null argument");
  PHS_REQUIRE(src->N > 0 && dst->N % src->N == 0 && src->H == dst->H && src->W == dst->W && src->C == dst->C,
              "phs_copy_cast: shape mismatch");
  int64_t total = (int64_t)dst->N * src->H * src->W * src->C;
  const int64_t sp = (int64_t)src->N * src->H * src->W;
  cudaStream_t st = (cudaStream_t)stream;
  int b = stream_blocks(total);
  if (src->dtype == PHS_F32 && dst->dtype == PHS_F32)
    phs_launch(copy_cast_kernel<float, float>, b, 256, 0, st, (const float*)src->ptr, src->ld, (float*)dst->ptr, dst->ld, src->C, total, sp);
  else if (src->dtype == PHS_F32)
    phs_launch(copy_cast_kernel<float, bf16>, b, 256, 0, st, (const float*)src->ptr, src->ld, (bf16*)dst->ptr, dst->ld, src->C, total, sp);
  else if (dst->dtype == PHS_F32)
    phs_launch(copy_cast_kernel<bf16, float>, b, 256, 0, st, (const bf16*)src->ptr, src->ld, (float*)dst->ptr, dst->ld, src->C, total, sp);
  else
    phs_launch(copy_cast_kernel<bf16, bf16>, b, 256, 0, st, (const bf16*)src->ptr, src->ld, (bf16*)dst->ptr, dst->ld, src->C, total, sp);
  return phs_check_launch("copy_cast");
}

// fp32 -> (hi, lo) bf16 pair with x = hi + lo to 16 mantissa bits (the operand split of the fp32-accurate tensor-core
// mode: x*w ~ hi*w_hi + lo*w_hi + hi*w_lo, three bf16 tcgen05 passes accumulated in fp32)
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ s, int lds, bf16* __restrict__ hi, int ldh,
                                                         bf16* __restrict__ lo, int ldl, int C, int64_t total) {
  PHS_PDL_PROLOGUE();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t pix = i / C;
    const float v = s[pix * lds + c];
    const bf16 h = __float2bfloat16_rn(v);
    hi[pix * ldh + c] = h;
    lo[pix * ldl + c] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

int phs_split_bf16(const phs_tensor* src, const phs_tensor* hi, const phs_tensor* lo, void* stream) {
  PHS_REQUIRE(src && hi && lo && src->ptr && hi->ptr && lo->ptr, "phs_split_bf16: null argument");
  PHS_REQUIRE(src->dtype == PHS_F32 && hi->dtype == PHS_BF16 && lo->dtype == PHS_BF16, "phs_split_bf16: f32 -> bf16, bf16");
  PHS_REQUIRE(src->N == hi->N && src->H == hi->H && src->W == hi->W && src->C == hi->C && src->N == lo->N &&
                  src->H == lo->H && src->W == lo->W && src->C == lo->C, "phs_split_bf16: shape mismatch");
  const int64_t total = (int64_t)src->N * src->H * src->W * src->C;
  phs_launch(split_bf16_kernel, stream_blocks(total), 256, 0, (cudaStream_t)stream, (const float*)src->ptr, src->ld,
             (bf16*)hi->ptr, hi->ld, (bf16*)lo->ptr, lo->ld, src->C, total);
  return phs_check_launch("split_bf16");
}

template <typename T>
__global__ void posterior_input_kernel(const float* __restrict__ x, const uint8_t* __restrict__ s, int Cx, int nl,
                                       T* __restrict__ out, int ld, int64_t npix) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    T* o = out + p * ld;
    for (int c = 0; c < Cx; ++c) stf<T>(o + c, x[p * Cx + c]);
    int lab = s[p];
    for (int c = 0; c < nl; ++c) stf<T>(o + Cx + c, (c == lab ? 1.f : 0.f) - 0.5f);
  }
}

int phs_posterior_input(const float* x, const uint8_t* s, int N, int H, int W, int Cx, int nlabels,
                        const phs_tensor* out, void* stream) {
  PHS_REQUIRE(x && s && out && out->ptr, "phs_posterior_input: null argument");
  PHS_REQUIRE(out->C == Cx + nlabels && out->N == N && out->H == H && out->W == W, "phs_posterior_input: shape mismatch");
  int64_t npix = (int64_t)N * H * W;
  PHS_DISPATCH_DTYPE(out->dtype, T, (posterior_input_kernel<T><<<stream_blocks(npix), 256, 0, (cudaStream_t)stream>>>(
                                        x, s, Cx, nlabels, (T*)out->ptr, out->ld, npix)));
  return phs_check_launch("posterior_input");
}

// out[p][tap*Cin + ci] = x[p + tap][ci] (zero outside the image), remaining channels of out = 0.  Turns the 3x3
// convolution of a 1..7-channel network input into a 1x1 convolution over <= 64 channels that the tensor-core kernels
// take (forward and filter gradient); one thread per (pixel, 8-channel output vector).
template <typename T>
__global__ void __launch_bounds__(256) im2col3x3_kernel(const T* __restrict__ x, int ldx, int Cin, bf16* __restrict__ out,
                                                        int ldo, int Co, int H, int W, idx4_t ix, fdiv_t fcin, uint32_t total) {
  PHS_PDL_PROLOGUE();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int v, wq, hq, n;
    idx4_decode(i, ix, v, wq, hq, n);
    const int64_t pix = ((int64_t)n * H + hq) * W + wq;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = v * 8 + j;
      const int tap = (int)fdiv((uint32_t)k, fcin), ci = k - tap * Cin;
      float val = 0.f;
      if (tap < 9) {
        const int hh = hq + tap / 3 - 1, ww = wq + tap % 3 - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = ldf<T>(x + (pix + (int64_t)(hh - hq) * W + (ww - wq)) * ldx + ci);
      }
      o[j] = val;
    }
    stv<bf16, 8>(out + pix * ldo + v * 8, o);
  }
}

int phs_im2col3x3(const phs_tensor* x, const phs_tensor* out, void* stream) {
  PHS_REQUIRE(x && out && x->ptr && out->ptr, "phs_im2col3x3: null argument");
  PHS_REQUIRE(out->dtype == PHS_BF16 && out->C % 8 == 0 && out->ld % 8 == 0 && ((uintptr_t)out->ptr & 15) == 0,
              "phs_im2col3x3: output must be bf16 with 16-byte aligned 8-channel vectors");
  PHS_REQUIRE(out->C >= 9 * x->C && x->N == out->N && x->H == out->H && x->W == out->W,
              "phs_im2col3x3: output needs >= 9*Cin channels and the input's N,H,W");
  int64_t total = (int64_t)x->N * x->H * x->W * (out->C / 8);
  PHS_REQUIRE(total < (1ll << 31), "phs_im2col3x3: tensor too large");
  const idx4_t ix = idx4_make(out->C / 8, x->W, x->H);
  PHS_DISPATCH_DTYPE(x->dtype, T, (phs_launch(im2col3x3_kernel<T>, stream_blocks(total), 256, 0, (cudaStream_t)stream, 
                                      (const T*)x->ptr, x->ld, x->C, (bf16*)out->ptr, out->ld, out->C, x->H, x->W, ix,
                                      fdiv_make((uint32_t)x->C), (uint32_t)total)));
  return phs_check_launch("im2col3x3");
}

template <typename T>
__global__ void broadcast_z_kernel(const float* __restrict__ z, T* __restrict__ out, int ld, int HW, int zd, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % zd);
    int64_t pix = i / zd;
    int n = (int)(pix / HW);
    stf<T>(out + pix * ld + c, z[n * zd + c]);
  }
}

int phs_broadcast_z(const float* z, const phs_tensor* out, void* stream) {
  PHS_REQUIRE(z && out && out->ptr, "phs_broadcast_z: null argument");
  int HW = out->H * out->W;
  int64_t total = (int64_t)out->N * HW * out->C;
  PHS_DISPATCH_DTYPE(out->dtype, T, (broadcast_z_kernel<T><<<stream_blocks(total), 256, 0, (cudaStream_t)stream>>>(
                                        z, (T*)out->ptr, out->ld, HW, out->C, total)));
  return phs_check_launch("broadcast_z");
}

template <typename T>
__global__ void broadcast_z_bwd_kernel(const T* __restrict__ g, int ld, int HW, int zd, float* __restrict__ dz, int accumulate) {
  // one block per (n, c)
  int n = blockIdx.x / zd, c = blockIdx.x % zd;
  float a = 0.f;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) a += ldf<T>(g + ((int64_t)n * HW + p) * ld + c);
  __shared__ float sm[256];
  sm[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) dz[blockIdx.x] = (accumulate ? dz[blockIdx.x] : 0.f) + sm[0];
}

int phs_broadcast_z_bwd(const phs_tensor* g, float* dz, int accumulate, void* stream) {
  PHS_REQUIRE(g && g->ptr && dz, "phs_broadcast_z_bwd: null argument");
  int HW = g->H * g->W;
  PHS_DISPATCH_DTYPE(g->dtype, T, (broadcast_z_bwd_kernel<T><<<g->N * g->C, 256, 0, (cudaStream_t)stream>>>(
                                      (const T*)g->ptr, g->ld, HW, g->C, dz, accumulate)));
  return phs_check_launch("broadcast_z_bwd");
}

__global__ void fill_kernel(float* p, int64_t n, float v) {
  PHS_PDL_PROLOGUE();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
int phs_fill_f32(float* p, int64_t n, float v, void* stream) {
  PHS_REQUIRE(p || n == 0, "phs_fill_f32: null");
  if (n == 0) return 0;
  phs_launch(fill_kernel, stream_blocks(n), 256, 0, (cudaStream_t)stream, p, n, v);
  return phs_check_launch("fill");
}
__global__ void axpy_kernel(float* d, const float* s, int64_t n, float alpha) {
  PHS_PDL_PROLOGUE();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] += alpha * s[i];
}
int phs_axpy_f32(float* dst, const float* src, int64_t n, float alpha, void* stream) {
  PHS_REQUIRE((dst && src) || n == 0, "phs_axpy_f32: null");
  if (n == 0) return 0;
  phs_launch(axpy_kernel, stream_blocks(n), 256, 0, (cudaStream_t)stream, dst, src, n, alpha);
  return phs_check_launch("axpy");
}

// out[0] += scale * sum(src^2)  (tf.nn.l2_loss stack of add_weight_decay, phiseg_model.py:290-300)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ s, int64_t n, float scale, float* out) {
  float a = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a += s[i] * s[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i];
    atomicAdd(out, t * scale);
  }
}
int phs_sumsq_f32(const float* src, int64_t n, float scale, float* out, void* stream) {
  PHS_REQUIRE((src && out) || n == 0, "phs_sumsq_f32: null");
  if (n == 0) return 0;
  int64_t b = (n + 255) / 256;
  if (b > 296) b = 296;
  sumsq_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>(src, n, scale, out);
  return phs_check_launch("sumsq");
}

// add_weight_decay (phiseg_model.py:290-300) for ALL filters in one launch: segs[i] = (offset, count) of filter i in the
// flat fp32 parameter buffer; out[0] += wd * sum_i l2_loss(W_i) = 0.5 * wd * sum W^2, and (g != null) g += wd * W, the
// gradient of that term.  grid = (blocks per segment, segments).
__global__ void __launch_bounds__(256) weight_decay_kernel(const float* __restrict__ p, float* __restrict__ g,
                                                           const int64_t* __restrict__ segs, float wd, float* out) {
  const int64_t off = segs[2 * blockIdx.y], n = segs[2 * blockIdx.y + 1];
  const float* w = p + off;
  float a = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = w[i];
    a += v * v;
    if (g) g[off + i] += wd * v;
  }
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0 && out) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sm[i];
    atomicAdd(out, 0.5f * wd * t);
  }
}
int phs_weight_decay(const float* params, float* grads, const int64_t* segs, int nseg, float wd, float* loss_out,
                     void* stream) {
  PHS_REQUIRE((params && segs) || nseg == 0, "phs_weight_decay: null");
  PHS_REQUIRE(nseg >= 0 && nseg <= 65535, "phs_weight_decay: nseg=%d", nseg);
  if (nseg == 0) return 0;
  weight_decay_kernel<<<dim3(8, nseg), 256, 0, (cudaStream_t)stream>>>(params, grads, segs, wd, loss_out);
  return phs_check_launch("weight_decay");
}

// argmax over the last axis of [npix, nl] (np.argmax of the accumulated softmax in predict, phiseg_model.py:351-353):
// first maximal index, like numpy.
__global__ void argmax_kernel(const float* __restrict__ s, int64_t npix, int nl, int64_t* __restrict__ out) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    float mx = s[p * nl];
    int am = 0;
    for (int c = 1; c < nl; ++c) {
      float v = s[p * nl + c];
      if (v > mx) { mx = v; am = c; }
    }
    out[p] = am;
  }
}
int phs_argmax_f32(const float* src, int64_t npix, int nlabels, int64_t* out, void* stream) {
  PHS_REQUIRE((src && out) || npix == 0, "phs_argmax_f32: null");
  PHS_REQUIRE(nlabels >= 1, "phs_argmax_f32: nlabels");
  if (npix == 0) return 0;
  argmax_kernel<<<stream_blocks(npix), 256, 0, (cudaStream_t)stream>>>(src, npix, nlabels, out);
  return phs_check_launch("argmax");
}
