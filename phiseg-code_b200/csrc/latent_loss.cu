// Latent heads (softplus / reparameterisation / KL), the multi-scale residual cross-entropy ELBO term,
// logits aggregation for sampling, and the optimizer kernels.
#include "common.cuh"

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sm[32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = l < (blockDim.x + 31) / 32 ? sm[l] : 0.f;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in warp 0
}

// KL(q||p) element: 0.5*((s0^2 + (m1-m0)^2)/(s1^2+1e-10) + log(s1^2+1e-10) - log(s0^2+1e-10) - 1)
// (phiseg_model.py:210-226), 0 = posterior (q), 1 = prior (p)
__device__ __forceinline__ float kl_elem(float m0, float s0, float m1, float s1) {
  float a = s1 * s1 + 1e-10f, c = s0 * s0 + 1e-10f, d = m1 - m0;
  return 0.5f * ((s0 * s0 + d * d) / a + logf(a) - logf(c) - 1.f);
}

// elementwise (phiseg) form: count = N*hw*zd elements
__global__ void latent_fwd_kernel(const float* __restrict__ mu_q, const float* __restrict__ sp_q,
                                  const float* __restrict__ mu_p, const float* __restrict__ sp_p,
                                  const float* __restrict__ eps, int64_t count, int use_prior_z,
                                  float* __restrict__ sigma_q, float* __restrict__ sigma_p, float* __restrict__ z,
                                  float* kl_out, float kl_scale) {
  PHS_PDL_PROLOGUE();
  float kl = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    float mq = 0.f, sq = 0.f, mp = 0.f, sp = 0.f;
    if (mu_q) { mq = mu_q[i]; sq = softplus_f(sp_q[i]); sigma_q[i] = sq; }
    if (mu_p) { mp = mu_p[i]; sp = softplus_f(sp_p[i]); sigma_p[i] = sp; }
    if (z) z[i] = use_prior_z ? mp + sp * eps[i] : mq + sq * eps[i];
    if (mu_q && mu_p) kl += kl_elem(mq, sq, mp, sp);
  }
  if (kl_out) {
    float r = block_sum(kl);
    if (threadIdx.x == 0) atomicAdd(kl_out, r * kl_scale);
  }
}

// ProbUNet form: mu = mean_hw(mu_map), sigma = mean_hw(softplus(sp_map)); one thread per (n, d)
__global__ void latent_fwd_gap_kernel(const float* __restrict__ mu_q, const float* __restrict__ sp_q,
                                      const float* __restrict__ mu_p, const float* __restrict__ sp_p,
                                      const float* __restrict__ eps, int N, int hw, int zd, int use_prior_z,
                                      float* __restrict__ mu_q_out, float* __restrict__ sigma_q,
                                      float* __restrict__ mu_p_out, float* __restrict__ sigma_p, float* __restrict__ z,
                                      float* kl_out, float kl_scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float kl = 0.f;
  if (i < N * zd) {
    int n = i / zd, d = i % zd;
    float mq = 0.f, sq = 0.f, mp = 0.f, sp = 0.f;
    for (int p = 0; p < hw; ++p) {
      size_t j = ((size_t)n * hw + p) * zd + d;
      if (mu_q) { mq += mu_q[j]; sq += softplus_f(sp_q[j]); }
      if (mu_p) { mp += mu_p[j]; sp += softplus_f(sp_p[j]); }
    }
    mq /= hw; sq /= hw; mp /= hw; sp /= hw;
    if (mu_q) { mu_q_out[i] = mq; sigma_q[i] = sq; }
    if (mu_p) { mu_p_out[i] = mp; sigma_p[i] = sp; }
    if (z) z[i] = use_prior_z ? mp + sp * eps[i] : mq + sq * eps[i];
    if (mu_q && mu_p) kl = kl_elem(mq, sq, mp, sp);
  }
  if (kl_out) {
    float r = block_sum(kl);
    if (threadIdx.x == 0) atomicAdd(kl_out, r * kl_scale);
  }
}

int phs_latent_fwd(const float* mu_q, const float* sp_q, const float* mu_p, const float* sp_p, const float* eps, int N,
                   int hw, int zd, int gap, int use_prior_z, float* mu_q_out, float* sigma_q, float* mu_p_out,
                   float* sigma_p, float* z, float* kl_out, float kl_scale, void* stream) {
  PHS_REQUIRE(mu_q || mu_p, "phs_latent_fwd: need at least one of posterior/prior");
  PHS_REQUIRE(!mu_q || (sp_q && sigma_q), "phs_latent_fwd: posterior args");
  PHS_REQUIRE(!mu_p || (sp_p && sigma_p), "phs_latent_fwd: prior args");
  PHS_REQUIRE(!z || eps, "phs_latent_fwd: eps required for z");
  PHS_REQUIRE(!z || (use_prior_z ? mu_p != 0 : mu_q != 0), "phs_latent_fwd: z source missing");
  cudaStream_t st = (cudaStream_t)stream;
  if (gap) {
    PHS_REQUIRE((!mu_q || mu_q_out) && (!mu_p || mu_p_out), "phs_latent_fwd: gap outputs");
    int total = N * zd;
    latent_fwd_gap_kernel<<<(total + 127) / 128, 128, 0, st>>>(mu_q, sp_q, mu_p, sp_p, eps, N, hw, zd, use_prior_z, mu_q_out,
                                                               sigma_q, mu_p_out, sigma_p, z, kl_out, kl_scale);
  } else {
    int64_t count = (int64_t)N * hw * zd;
    int blocks = (int)((count + 255) / 256 < 296 ? (count + 255) / 256 : 296);
    phs_launch(latent_fwd_kernel, blocks, 256, 0, st, mu_q, sp_q, mu_p, sp_p, eps, count, use_prior_z, sigma_q, sigma_p, z, kl_out,
                                              kl_scale);
  }
  return phs_check_launch("latent_fwd");
}

// gradients.  z = mu_q + sigma_q*eps  =>  dmu_q += dz, dsigma_q += dz*eps.  KL weight w:
//   a = s1^2+1e-10, c = s0^2+1e-10, D = m1-m0
//   dKL/dm0 = -D/a, dKL/dm1 = D/a, dKL/ds0 = s0/a - s0/c, dKL/ds1 = s1/a - s1*(s0^2+D^2)/a^2
// sigma = softplus(pre) => dpre = dsigma*sigmoid(pre); gap: each of the hw positions receives 1/hw of it.
__global__ void latent_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ mu_q,
                                  const float* __restrict__ sp_q, const float* __restrict__ sigma_q,
                                  const float* __restrict__ mu_p, const float* __restrict__ sp_p,
                                  const float* __restrict__ sigma_p, const float* __restrict__ eps, int N, int hw,
                                  int zd, int gap, float w, float* __restrict__ d_mu_q, float* __restrict__ d_sp_q,
                                  float* __restrict__ d_mu_p, float* __restrict__ d_sp_p) {
  PHS_PDL_PROLOGUE();
  int64_t count = (int64_t)N * hw * zd;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < count; j += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = j;  // index into the (possibly pooled) latent
    float inv = 1.f;
    if (gap) {
      int d = (int)(j % zd);
      int64_t n = j / ((int64_t)hw * zd);
      i = n * zd + d;
      inv = 1.f / hw;
    }
    // gap: mu_q/mu_p here are the pooled values (outputs of the forward), sp_* the per-position maps
    float m0 = gap ? mu_q[i] : mu_q[j];
    float m1 = gap ? mu_p[i] : mu_p[j];
    float s0 = sigma_q[i], s1 = sigma_p[i];
    float a = s1 * s1 + 1e-10f, c = s0 * s0 + 1e-10f, D = m1 - m0;
    float g = dz ? dz[i] : 0.f;
    float dm0 = g - w * D / a;
    float ds0 = g * eps[i] + w * (s0 / a - s0 / c);
    float dm1 = w * D / a;
    float ds1 = w * (s1 / a - s1 * (s0 * s0 + D * D) / (a * a));
    d_mu_q[j] = dm0 * inv;
    d_sp_q[j] = ds0 * inv * sigmoid_f(sp_q[j]);
    d_mu_p[j] = dm1 * inv;
    d_sp_p[j] = ds1 * inv * sigmoid_f(sp_p[j]);
  }
}

int phs_latent_bwd(const float* dz, const float* mu_q, const float* sp_q, const float* sigma_q, const float* mu_p,
                   const float* sp_p, const float* sigma_p, const float* eps, int N, int hw, int zd, int gap,
                   float kl_scale, float* d_mu_q, float* d_sp_q, float* d_mu_p, float* d_sp_p, void* stream) {
  PHS_REQUIRE(mu_q && sp_q && sigma_q && mu_p && sp_p && sigma_p && eps && d_mu_q && d_sp_q && d_mu_p && d_sp_p,
              "phs_latent_bwd: null argument");
  int64_t count = (int64_t)N * hw * zd;
  int blocks = (int)((count + 255) / 256 < 296 ? (count + 255) / 256 : 296);
  phs_launch(latent_bwd_kernel, blocks, 256, 0, (cudaStream_t)stream, dz, mu_q, sp_q, sigma_q, mu_p, sp_p, sigma_p, eps, N, hw, zd,
                                                             gap, kl_scale, d_mu_q, d_sp_q, d_mu_p, d_sp_p);
  return phs_check_launch("latent_bwd");
}

// ---------------------------------------------------------------------------------------------------------
// multi-scale residual cross-entropy: one thread per full-resolution pixel, nearest-neighbour reads of the
// native-resolution head outputs, top-down accumulation, softmax-xent per level, gradient scatter.
// ---------------------------------------------------------------------------------------------------------
constexpr int XENT_MAXL = 8;
constexpr int XENT_MAXC = 8;
struct XentPtrs {
  const float* logits[XENT_MAXL];
  float* dlogits[XENT_MAXL];
};

// One thread per 2x2 block of full-resolution pixels, one CTA per 32x32-pixel tile; thread t owns the 2x2 block at the
// Z-order (bit-interleaved) position t of the tile, so the 4^(l-1) threads that share a pixel of level l are a
// contiguous, aligned run of threads: the gradients of the coarser levels (nearest-neighbour up-sampled heads: 4^l
// pixels share one head pixel) are summed in a FIXED order - registers over the 2x2 block (level 1), warp shuffles
// (levels 2-3), shared memory across warps (levels 4-5) - and written by exactly one thread.  No atomics up to level 5
// (the reference uses 5 levels): the gradient, and through it the whole backward pass, is reproducible run to run
// (fp32 atomics here were the one source of the 1e-2 run-to-run gradient differences of round 1).  Levels 6-7 span
// several tiles and keep atomics.
__device__ __forceinline__ int compact_even_bits(int v) {   // bits 0,2,4,6 -> 0,1,2,3
  v &= 0x55;
  v = (v | (v >> 1)) & 0x33;
  v = (v | (v >> 2)) & 0x0f;
  return v;
}

template <int NC>   // class slots compiled in (2, 4 or NC): predicated-off slots still cost issue cycles
__global__ void __launch_bounds__(256)
    xent_multiscale_kernel(XentPtrs P, const uint8_t* __restrict__ labels, int N, int H, int W, int nl, int L,
                           float scale, float* __restrict__ loss_out, int tilesW, int tilesH, uint32_t total) {
  __shared__ float red[2][8][NC];        // [level 4 | level 5][warp][class]
  float lsum[XENT_MAXL];
#pragma unroll
  for (int l = 0; l < XENT_MAXL; ++l) lsum[l] = 0.f;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int bx = compact_even_bits(t), by = compact_even_bits(t >> 1);   // 2x2 block inside the 32x32 tile (16 x 16 blocks)
  for (uint32_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tx = tile % tilesW, ty = (tile / tilesW) % tilesH, n = tile / (tilesW * tilesH);
    const int xb = tx * 16 + bx, yb = ty * 16 + by;
    const bool inside = 2 * xb < W && 2 * yb < H;
    // coarse-level gradient sums of this block: gblk[l][c] = sum over the block's pixels of prefix_{i<=l} g_i
    float gblk[XENT_MAXL][NC];
#pragma unroll
    for (int l = 0; l < XENT_MAXL; ++l)
#pragma unroll
      for (int c = 0; c < NC; ++c) gblk[l][c] = 0.f;
    if (inside) {
#pragma unroll
      for (int sub = 0; sub < 4; ++sub) {
        const int x = 2 * xb + (sub & 1), y = 2 * yb + (sub >> 1);
        const int64_t p = ((int64_t)n * H + y) * W + x;
        const int lab = labels[p];
        float acc[NC], gsum[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = gsum[c] = 0.f;
        // pass 1 (top-down): accumulate logits, per-level loss, per-level softmax gradient g_l;
        // d loss / d logits[j] (at full res) = sum_{i<=j} g_i: walk the levels again bottom-up for the prefix sums
        float gl[XENT_MAXL][NC];
#pragma unroll
        for (int l = XENT_MAXL - 1; l >= 0; --l) {
          if (l < L) {
            const int hl = H >> l, wl = W >> l;
            const float* src = P.logits[l] + (((int64_t)n * hl + (y >> l)) * wl + (x >> l)) * nl;
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < NC; ++c)
              if (c < nl) { acc[c] += src[c]; mx = fmaxf(mx, acc[c]); }
            float se = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c)
              if (c < nl) { gl[l][c] = expf(acc[c] - mx); se += gl[l][c]; }
            const float lse = mx + logf(se);
            const float inv = 1.f / se;
#pragma unroll
            for (int c = 0; c < NC; ++c)
              if (c < nl) {
                gl[l][c] = (gl[l][c] * inv - (c == lab ? 1.f : 0.f)) * scale;
                if (c == lab) lsum[l] += lse - acc[c];
              }
          }
        }
        if (P.dlogits[0]) {
#pragma unroll
          for (int l = 0; l < XENT_MAXL; ++l) {
            if (l < L) {
#pragma unroll
              for (int c = 0; c < NC; ++c)
                if (c < nl) {
                  gsum[c] += gl[l][c];
                  if (l == 0) P.dlogits[0][p * nl + c] = gsum[c];
                  else gblk[l][c] += gsum[c];
                }
            }
          }
        }
      }
    }
    if (P.dlogits[0]) {     // (block-uniform: every thread takes part in the shuffles and barriers below)
#pragma unroll
      for (int l = 1; l < XENT_MAXL; ++l) {
        if (l < L) {
          const int hl = H >> l, wl = W >> l;
          float* dst = P.dlogits[l] + (((int64_t)n * hl + (yb >> (l - 1))) * wl + (xb >> (l - 1))) * nl;
          float v[NC];
#pragma unroll
          for (int c = 0; c < NC; ++c) v[c] = gblk[l][c];
          if (l >= 2) {
            // 4^(l-1) consecutive threads share the level-l pixel: butterfly inside the warp (same order every run)
            const int span = l == 2 ? 4 : (l == 3 ? 16 : 32);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
              if (o < span)
#pragma unroll
                for (int c = 0; c < NC; ++c) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
          }
          if (l <= 3) {
            const int span = l == 1 ? 1 : (l == 2 ? 4 : 16);
            if (inside && (t & (span - 1)) == 0)
#pragma unroll
              for (int c = 0; c < NC; ++c)
                if (c < nl) dst[c] = v[c];
          } else if (l <= 5) {
            // 2 (level 4) or 8 (level 5) warps share the pixel: per-warp sums through shared memory, added in warp order
            if (lane == 0)
#pragma unroll
              for (int c = 0; c < NC; ++c) red[l - 4][warp][c] = v[c];
            __syncthreads();
            const int nw = l == 4 ? 2 : 8;
            if (inside && (t & (32 * nw - 1)) == 0) {
#pragma unroll
              for (int c = 0; c < NC; ++c)
                if (c < nl) {
                  float a = 0.f;
                  for (int w2 = 0; w2 < nw; ++w2) a += red[l - 4][warp + w2][c];
                  dst[c] = a;
                }
            }
            __syncthreads();
          } else {
            // levels 6, 7: the pixel spans several tiles (not used by the reference's 5-level configurations)
            if (inside && lane == 0)
#pragma unroll
              for (int c = 0; c < NC; ++c)
                if (c < nl) atomicAdd(dst + c, v[c]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int l = 0; l < XENT_MAXL; ++l) {
    if (l < L) {
      float r = block_sum(lsum[l]);
      if (threadIdx.x == 0) atomicAdd(loss_out + l, r * scale);
    }
  }
}

int phs_xent_multiscale(const float* const* logits, float* const* dlogits, const uint8_t* labels, int N, int H, int W,
                        int nlabels, int L, float scale, float* loss_out, void* stream) {
  PHS_REQUIRE(logits && labels && loss_out, "phs_xent_multiscale: null argument");
  PHS_REQUIRE(L >= 1 && L <= XENT_MAXL && nlabels >= 1 && nlabels <= XENT_MAXC, "phs_xent_multiscale: L=%d nlabels=%d unsupported", L, nlabels);
  PHS_REQUIRE((H >> (L - 1)) << (L - 1) == H && (W >> (L - 1)) << (L - 1) == W, "phs_xent_multiscale: size not divisible");
  XentPtrs P;
  for (int l = 0; l < XENT_MAXL; ++l) {
    P.logits[l] = l < L ? logits[l] : nullptr;
    P.dlogits[l] = (l < L && dlogits) ? dlogits[l] : nullptr;
    PHS_REQUIRE(l >= L || P.logits[l], "phs_xent_multiscale: logits[%d] null", l);
  }
  PHS_REQUIRE(H % 2 == 0 && W % 2 == 0, "phs_xent_multiscale: odd image size");
  const int tilesW = (W + 31) / 32, tilesH = (H + 31) / 32;
  const int64_t ntile = (int64_t)N * tilesW * tilesH;
  PHS_REQUIRE(ntile < (1ll << 31), "phs_xent_multiscale: tensor too large");
  const int blocks = (int)(ntile < 148 * 8 ? ntile : 148 * 8);
  if (nlabels <= 2)
    xent_multiscale_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(P, labels, N, H, W, nlabels, L, scale, loss_out,
                                                                        tilesW, tilesH, (uint32_t)ntile);
  else if (nlabels <= 4)
    xent_multiscale_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(P, labels, N, H, W, nlabels, L, scale, loss_out,
                                                                        tilesW, tilesH, (uint32_t)ntile);
  else
    xent_multiscale_kernel<XENT_MAXC><<<blocks, 256, 0, (cudaStream_t)stream>>>(P, labels, N, H, W, nlabels, L, scale,
                                                                                loss_out, tilesW, tilesH, (uint32_t)ntile);
  return phs_check_launch("xent_multiscale");
}

__global__ void __launch_bounds__(256)
    aggregate_logits_kernel(XentPtrs P, int N, int H, int W, int nl, int L, int rep, float* __restrict__ s_out,
                            float* __restrict__ sm_out, float* __restrict__ sm_accum, int64_t* __restrict__ argmax_out) {
  // N = images * rep rows (sample-major); one thread walks the rep samples of an image pixel so that their softmax
  // sum reaches sm_accum[image pixel] without atomics and in a fixed order
  const int64_t ipix = (int64_t)(N / rep) * H * W;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < ipix; q += (int64_t)gridDim.x * blockDim.x) {
   float accum[XENT_MAXC];
#pragma unroll
   for (int c = 0; c < XENT_MAXC; ++c) accum[c] = 0.f;
   for (int r = 0; r < rep; ++r) {
    const int64_t p = q + r * ipix;
    int x = (int)(p % W);
    int64_t t = p / W;
    int y = (int)(t % H);
    int n = (int)(t / H);
    float acc[XENT_MAXC];
#pragma unroll
    for (int c = 0; c < XENT_MAXC; ++c) acc[c] = 0.f;
    // phiseg_model.py:304-311: s_accum = list[-1]; then += list[0..L-2] in order
    for (int k = 0; k < L; ++k) {
      int l = k == 0 ? L - 1 : k - 1;
      int hl = H >> l, wl = W >> l;
      const float* src = P.logits[l] + (((int64_t)n * hl + (y >> l)) * wl + (x >> l)) * nl;
#pragma unroll
      for (int c = 0; c < XENT_MAXC; ++c)
        if (c < nl) acc[c] += src[c];
    }
    float mx = -INFINITY;
    int am = 0;
#pragma unroll
    for (int c = 0; c < XENT_MAXC; ++c)
      if (c < nl && acc[c] > mx) { mx = acc[c]; am = c; }
    if (s_out)
      for (int c = 0; c < nl; ++c) s_out[p * nl + c] = acc[c];
    if (sm_out || sm_accum) {
      float e[XENT_MAXC], se = 0.f;
#pragma unroll
      for (int c = 0; c < XENT_MAXC; ++c)
        if (c < nl) { e[c] = expf(acc[c] - mx); se += e[c]; }
      float inv = 1.f / se;
#pragma unroll
      for (int c = 0; c < XENT_MAXC; ++c)
        if (c < nl) {
          if (sm_out) sm_out[p * nl + c] = e[c] * inv;
          accum[c] += e[c] * inv;
        }
    }
    if (argmax_out) argmax_out[p] = am;
   }
   if (sm_accum)
#pragma unroll
     for (int c = 0; c < XENT_MAXC; ++c)
       if (c < nl) sm_accum[q * nl + c] += accum[c];
  }
}

int phs_aggregate_logits(const float* const* logits, int N, int H, int W, int nlabels, int L, int rep, float* s_out,
                         float* softmax_out, float* softmax_accum, int64_t* argmax_out, void* stream) {
  PHS_REQUIRE(logits, "phs_aggregate_logits: null argument");
  PHS_REQUIRE(rep >= 1 && N % rep == 0, "phs_aggregate_logits: N=%d is not a multiple of rep=%d", N, rep);
  PHS_REQUIRE(L >= 1 && L <= XENT_MAXL && nlabels >= 1 && nlabels <= XENT_MAXC, "phs_aggregate_logits: L=%d nlabels=%d unsupported", L, nlabels);
  XentPtrs P;
  for (int l = 0; l < XENT_MAXL; ++l) {
    P.logits[l] = l < L ? logits[l] : nullptr;
    P.dlogits[l] = nullptr;
  }
  int64_t npix = (int64_t)(N / rep) * H * W;
  int blocks = (int)((npix + 255) / 256 < 148 * 8 ? (npix + 255) / 256 : 148 * 8);
  aggregate_logits_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(P, N, H, W, nlabels, L, rep, s_out, softmax_out,
                                                                   softmax_accum, argmax_out);
  return phs_check_launch("aggregate_logits");
}

// ---------------------------------------------------------------------------------------------------------
// optimizer
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, float lr_t,
                                                   const float* __restrict__ lr_dev, float b1, float b2, float eps,
                                                   float gs) {
  if (lr_dev) lr_t = lr_dev[0];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gs;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

int phs_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, const float* lr_t_dev,
                  float beta1, float beta2, float eps, float grad_scale, void* stream) {
  PHS_REQUIRE(p && g && m && v, "phs_adam_step: null argument");
  int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr_t, lr_t_dev, beta1, beta2, eps, grad_scale);
  return phs_check_launch("adam_step");
}

// tf.train.MomentumOptimizer(use_nesterov=True): acc = mom*acc + g; p -= lr*(g + mom*acc)
__global__ void __launch_bounds__(256) momentum_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ acc, int64_t n, float lr,
                                                       const float* __restrict__ lr_dev, float mom, float gs) {
  if (lr_dev) lr = lr_dev[0];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gs;
    float a = mom * acc[i] + gi;
    acc[i] = a;
    p[i] -= lr * (gi + mom * a);
  }
}

int phs_momentum_step(float* p, const float* g, float* acc, int64_t n, float lr, const float* lr_dev, float momentum,
                      float grad_scale, void* stream) {
  PHS_REQUIRE(p && g && acc, "phs_momentum_step: null argument");
  int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  momentum_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, acc, n, lr, lr_dev, momentum, grad_scale);
  return phs_check_launch("momentum_step");
}

// bf16 shadows of the conv filters.  blockIdx.y = conv, the blocks of a column walk its 32x32 (ci, co) tiles of every
// tap: the dgrad layout keeps co contiguous (coalesced straight from the load), the forward layout is the per-tap
// transpose and goes through shared memory so that both global accesses are coalesced.
// part = 0: shadow = bf16(w); part = 1: shadow = bf16(w - bf16(w)), the low half of the split used by the fp32-accurate
// tensor-core mode (w = w_hi + w_lo to 16 mantissa bits)
__device__ __forceinline__ bf16 split_part(float v, int part) {
  const bf16 hi = __float2bfloat16_rn(v);
  return part ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
}

__global__ void __launch_bounds__(256) weight_prep_kernel(const float* __restrict__ master, bf16* __restrict__ shadow,
                                                          const int64_t* __restrict__ table, int part) {
  __shared__ float tile[32][33];
  const int64_t* e = table + (int64_t)blockIdx.y * 7;
  const float* src = master + e[0];
  bf16* fwd = shadow + e[1];
  bf16* dg = e[2] >= 0 ? shadow + e[2] : nullptr;
  const int taps = (int)e[3], cin = (int)e[4], cout = (int)e[5];
  const int64_t kpitch = e[6] > 0 ? e[6] : (int64_t)taps * cin;   // row pitch of the forward layout (zero padded)
  const int tci = (cin + 31) / 32, tco = (cout + 31) / 32;
  const int ntiles = taps * tci * tco;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int co0 = (t % tco) * 32, ci0 = ((t / tco) % tci) * 32, tap = t / (tco * tci);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + ty + 8 * j, co = co0 + tx;
      float v = 0.f;
      if (ci < cin && co < cout) {
        v = src[((int64_t)tap * cin + ci) * cout + co];   // HWIO: [tap][ci][co]
        if (dg) dg[(int64_t)ci * taps * cout + (int64_t)(taps - 1 - tap) * cout + co] = split_part(v, part);
      }
      tile[ty + 8 * j][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + ty + 8 * j, ci = ci0 + tx;
      if (ci < cin && co < cout)
        fwd[(int64_t)co * kpitch + (int64_t)tap * cin + ci] = split_part(tile[tx][ty + 8 * j], part);
    }
    __syncthreads();
  }
}

int phs_weight_prep(const float* master, void* shadow, const int64_t* table, int nconv, void* stream) {
  PHS_REQUIRE(master && shadow && table, "phs_weight_prep: null argument");
  if (nconv <= 0) return 0;
  weight_prep_kernel<<<dim3(48, nconv), 256, 0, (cudaStream_t)stream>>>(master, (bf16*)shadow, table, 0);
  return phs_check_launch("weight_prep");
}

int phs_weight_prep_lo(const float* master, void* shadow_lo, const int64_t* table, int nconv, void* stream) {
  PHS_REQUIRE(master && shadow_lo && table, "phs_weight_prep_lo: null argument");
  if (nconv <= 0) return 0;
  weight_prep_kernel<<<dim3(48, nconv), 256, 0, (cudaStream_t)stream>>>(master, (bf16*)shadow_lo, table, 1);
  return phs_check_launch("weight_prep_lo");
}
