// Validation metrics of the PHiSeg evaluation loop on the device (phiseg_model.py:558-640, utils.py:103-118,270-370):
// generalised energy distance (pairwise per-label IoU between sample masks and annotations), the variance-NCC score
// (normalised cross correlation of pixel-wise cross-entropy maps) and the per-label Dice of the mean prediction.
// The reference computes them with numpy / medpy loops per image on the host (100 images x 16 samples every 500 steps);
// here the samples never leave the device: two reductions produce a few hundred integers / doubles that the host turns
// into the scalar scores.
#include "common.cuh"

namespace {

constexpr int MET_MAXL = 8;

__device__ __forceinline__ int load_label(const void* p, int es, int64_t i) {
  return es == 1 ? (int)((const uint8_t*)p)[i] : (int)((const int64_t*)p)[i];
}

// grid (Kb, Ka): block (j, i) counts, for every label l, the pixels where mask a_i == l and mask b_j == l
// (|intersection|), and - in the blocks of column / row 0 - the label histograms of a_i and b_j.
__global__ void __launch_bounds__(256) pairwise_label_kernel(const void* __restrict__ a, int es_a, const void* __restrict__ b,
                                                             int es_b, int64_t P, int nl, int* __restrict__ inter,
                                                             int* __restrict__ cnt_a, int* __restrict__ cnt_b) {
  const int j = blockIdx.x, i = blockIdx.y, Kb = gridDim.x;
  int it[MET_MAXL], ca[MET_MAXL], cb[MET_MAXL];
#pragma unroll
  for (int l = 0; l < MET_MAXL; ++l) it[l] = ca[l] = cb[l] = 0;
  for (int64_t p = threadIdx.x; p < P; p += blockDim.x) {
    const int la = load_label(a, es_a, (int64_t)i * P + p), lb = load_label(b, es_b, (int64_t)j * P + p);
#pragma unroll
    for (int l = 0; l < MET_MAXL; ++l) {
      ca[l] += la == l;
      cb[l] += lb == l;
      it[l] += (la == l) & (lb == l);
    }
  }
  __shared__ int red[3][MET_MAXL];
  if (threadIdx.x < 3 * MET_MAXL) red[threadIdx.x / MET_MAXL][threadIdx.x % MET_MAXL] = 0;
  __syncthreads();
#pragma unroll
  for (int l = 0; l < MET_MAXL; ++l) {
    int v0 = it[l], v1 = ca[l], v2 = cb[l];
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o);
      v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      v2 += __shfl_xor_sync(0xffffffffu, v2, o);
    }
    if ((threadIdx.x & 31) == 0) {       // integer atomics: order independent
      atomicAdd(&red[0][l], v0);
      atomicAdd(&red[1][l], v1);
      atomicAdd(&red[2][l], v2);
    }
  }
  __syncthreads();
  if (threadIdx.x < nl) {
    const int l = threadIdx.x;
    inter[((int64_t)i * Kb + j) * nl + l] = red[0][l];
    if (j == 0 && cnt_a) cnt_a[i * nl + l] = red[1][l];
    if (i == 0 && cnt_b) cnt_b[j * nl + l] = red[2][l];
  }
}

// variance_ncc_dist, per pixel (utils.py:323-362): mean_seg = mean_i s_i;  E_ss = mean_i xent(mean_seg, s_i);
// E_sy[j] = mean_i xent(onehot(gt_j), s_i), with xent(t, s) = -sum_l t_l log(s_l + 1e-8)
__global__ void __launch_bounds__(256) ncc_maps_kernel(const float* __restrict__ sm, const uint8_t* __restrict__ gt, int N, int M,
                                                       int64_t P, int nl, float* __restrict__ e_ss, float* __restrict__ e_sy) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    float mean[MET_MAXL], slog[MET_MAXL];
#pragma unroll
    for (int l = 0; l < MET_MAXL; ++l) mean[l] = slog[l] = 0.f;
    for (int i = 0; i < N; ++i) {
      const float* s = sm + ((int64_t)i * P + p) * nl;
#pragma unroll
      for (int l = 0; l < MET_MAXL; ++l)
        if (l < nl) {
          const float v = s[l];
          mean[l] += v;
          slog[l] += logf(v + 1e-8f);          // sum_i log(s_i,l + eps): all three maps are linear in it
        }
    }
    const float inv = 1.f / (float)N;
    float ess = 0.f;
#pragma unroll
    for (int l = 0; l < MET_MAXL; ++l)
      if (l < nl) ess -= (mean[l] * inv) * (slog[l] * inv);
    e_ss[p] = ess;
    for (int j = 0; j < M; ++j) {
      const int lab = gt[(int64_t)j * P + p];
      float v = 0.f;
#pragma unroll
      for (int l = 0; l < MET_MAXL; ++l)
        if (l == lab) v = -slog[l] * inv;
      e_sy[(int64_t)j * P + p] = v;
    }
  }
}

// one block per annotation j: sums needed for ncc(E_ss, E_sy[j]) (utils.py:103-118): sum a, sum a^2, sum v, sum v^2, sum a v
__global__ void __launch_bounds__(256) ncc_reduce_kernel(const float* __restrict__ e_ss, const float* __restrict__ e_sy, int64_t P,
                                                         double* __restrict__ out) {
  const int j = blockIdx.x;
  double s[5] = {0, 0, 0, 0, 0};
  for (int64_t p = threadIdx.x; p < P; p += blockDim.x) {
    const double a = e_ss[p], v = e_sy[(int64_t)j * P + p];
    s[0] += a; s[1] += a * a; s[2] += v; s[3] += v * v; s[4] += a * v;
  }
  __shared__ double red[5][8];
  for (int k = 0; k < 5; ++k) {
    double v = s[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];    // fixed order
    out[j * 5 + threadIdx.x] = t;
  }
}

// Uncertainty maps (phiseg_model.py:378-475): the reference stacks num_samples full-resolution outputs on the host and
// takes per-pixel moments with numpy.  Here every sampling pass (S rows per image, row = s*B + b) adds, per image pixel,
//   acc[0 .. nl)              sum_s v_c
//   acc[nl .. nl + nl(nl+1)/2) sum_s v_i v_j  (i <= j, row-major upper triangle)
//   acc[last]                 sum_s xent(logits_s, gt)         (softmax_cross_entropy_with_logits, :304-311)
// into doubles; v = softmax(logits) (kind 0) or the raw summed logits clipped to [lo, hi] (kind 1, :390).
__device__ __forceinline__ int tri_index(int i, int j, int nl) { return i * nl - i * (i - 1) / 2 + (j - i); }

__global__ void __launch_bounds__(256) sample_moments_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ gt,
                                                             int S, int B, int64_t P, int nl, int kind, float lo, float hi,
                                                             double* __restrict__ acc) {
  const int na = nl + nl * (nl + 1) / 2 + 1;
  const int64_t total = (int64_t)B * P;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    double s1[MET_MAXL], s2[MET_MAXL * (MET_MAXL + 1) / 2], xe = 0.0;
#pragma unroll
    for (int l = 0; l < MET_MAXL; ++l) s1[l] = 0.0;
#pragma unroll
    for (int l = 0; l < MET_MAXL * (MET_MAXL + 1) / 2; ++l) s2[l] = 0.0;
    const int lab = gt ? (int)gt[q] : -1;
    for (int s = 0; s < S; ++s) {
      const float* r = logits + ((int64_t)s * total + q) * nl;
      float v[MET_MAXL], mx = -3.0e38f;
#pragma unroll
      for (int l = 0; l < MET_MAXL; ++l)
        if (l < nl) { v[l] = r[l]; mx = fmaxf(mx, v[l]); }
      float den = 0.f, vl = 0.f;
#pragma unroll
      for (int l = 0; l < MET_MAXL; ++l)
        if (l < nl) {
          den += expf(v[l] - mx);
          if (l == lab) vl = v[l];
        }
      if (lab >= 0) xe += (double)(logf(den) + mx - vl);
#pragma unroll
      for (int l = 0; l < MET_MAXL; ++l)
        if (l < nl) v[l] = kind == 0 ? expf(v[l] - mx) / den : fminf(fmaxf(v[l], lo), hi);
#pragma unroll
      for (int i = 0; i < MET_MAXL; ++i)
        if (i < nl) {
          s1[i] += (double)v[i];
#pragma unroll
          for (int j = i; j < MET_MAXL; ++j)       // register triangle indexed for MET_MAXL: static after unrolling
            if (j < nl) s2[tri_index(i, j, MET_MAXL)] += (double)v[i] * (double)v[j];
        }
    }
    double* a = acc + q * na;
#pragma unroll
    for (int i = 0; i < MET_MAXL; ++i)
      if (i < nl) {
        a[i] += s1[i];
#pragma unroll
        for (int j = i; j < MET_MAXL; ++j)
          if (j < nl) a[nl + tri_index(i, j, nl)] += s2[tri_index(i, j, MET_MAXL)];
      }
    a[na - 1] += xe;
  }
}

// maps from the accumulated moments of `count` samples; every output pointer may be NULL
//   mean_arg  argmax_c mean_s v_c                                              (:470)
//   std_mean  mean_c sqrt(population variance of v_c)                           (:467-468, np.std)
//   var_sum   sum over the first nl - drop_last classes of the population variance = trace of the covariance (:393-402)
//   cov_det   determinant of the unbiased sample covariance of all classes      (:423-428, np.cov)
//   err       mean_s xent                                                       (:444-446, :472)
__global__ void __launch_bounds__(256) sample_maps_kernel(const double* __restrict__ acc, int64_t total, int nl, int count,
                                                          int drop_last, int64_t* __restrict__ mean_arg,
                                                          float* __restrict__ std_mean, float* __restrict__ var_sum,
                                                          float* __restrict__ cov_det, float* __restrict__ err) {
  const int na = nl + nl * (nl + 1) / 2 + 1;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const double* a = acc + q * na;
    const double inv = 1.0 / (double)count;
    double m[MET_MAXL];
    int best = 0;
    for (int l = 0; l < nl; ++l) {
      m[l] = a[l] * inv;
      if (m[l] > m[best]) best = l;
    }
    if (mean_arg) mean_arg[q] = best;
    double sd = 0.0, vs = 0.0;
    for (int l = 0; l < nl; ++l) {
      const double var = fmax(a[nl + tri_index(l, l, nl)] * inv - m[l] * m[l], 0.0);
      sd += sqrt(var);
      if (l < nl - drop_last) vs += var;
    }
    if (std_mean) std_mean[q] = (float)(sd / nl);
    if (var_sum) var_sum[q] = (float)vs;
    if (err) err[q] = (float)(a[na - 1] * inv);
    if (cov_det) {
      double c[MET_MAXL][MET_MAXL];
      const double un = 1.0 / (double)(count > 1 ? count - 1 : 1);
      for (int i = 0; i < nl; ++i)
        for (int j = i; j < nl; ++j) c[i][j] = c[j][i] = (a[nl + tri_index(i, j, nl)] - (double)count * m[i] * m[j]) * un;
      double det = 1.0;      // Gaussian elimination with partial pivoting (what LAPACK's getrf does for np.linalg.det)
      for (int k = 0; k < nl; ++k) {
        int piv = k;
        for (int i = k + 1; i < nl; ++i)
          if (fabs(c[i][k]) > fabs(c[piv][k])) piv = i;
        if (c[piv][k] == 0.0) { det = 0.0; break; }
        if (piv != k) {
          for (int j = 0; j < nl; ++j) { const double t = c[k][j]; c[k][j] = c[piv][j]; c[piv][j] = t; }
          det = -det;
        }
        det *= c[k][k];
        for (int i = k + 1; i < nl; ++i) {
          const double f = c[i][k] / c[k][k];
          for (int j = k; j < nl; ++j) c[i][j] -= f * c[k][j];
        }
      }
      cov_det[q] = (float)det;
    }
  }
}

}  // namespace

extern "C" {

int phs_pairwise_label_stats(const void* masks_a, int elem_size_a, int Ka, const void* masks_b, int elem_size_b, int Kb,
                             int64_t npix, int nlabels, int* inter, int* count_a, int* count_b, void* stream) {
  PHS_REQUIRE(masks_a && masks_b && inter, "phs_pairwise_label_stats: null argument");
  PHS_REQUIRE((elem_size_a == 1 || elem_size_a == 8) && (elem_size_b == 1 || elem_size_b == 8),
              "phs_pairwise_label_stats: masks must be uint8 or int64");
  PHS_REQUIRE(nlabels >= 1 && nlabels <= MET_MAXL, "phs_pairwise_label_stats: nlabels=%d unsupported", nlabels);
  PHS_REQUIRE(Ka >= 1 && Kb >= 1 && Ka <= 65535 && npix >= 1, "phs_pairwise_label_stats: bad sizes");
  pairwise_label_kernel<<<dim3(Kb, Ka), 256, 0, (cudaStream_t)stream>>>(masks_a, elem_size_a, masks_b, elem_size_b, npix, nlabels,
                                                                       inter, count_a, count_b);
  return phs_check_launch("pairwise_label_stats");
}

int phs_ncc_maps(const float* softmax, const uint8_t* gt, int N, int M, int64_t npix, int nlabels, float* e_ss, float* e_sy,
                 double* sums, void* stream) {
  PHS_REQUIRE(softmax && gt && e_ss && e_sy && sums, "phs_ncc_maps: null argument");
  PHS_REQUIRE(nlabels >= 1 && nlabels <= MET_MAXL && N >= 1 && M >= 1 && npix >= 1, "phs_ncc_maps: bad sizes");
  const int blocks = (int)((npix + 255) / 256 < 148 * 4 ? (npix + 255) / 256 : 148 * 4);
  ncc_maps_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(softmax, gt, N, M, npix, nlabels, e_ss, e_sy);
  ncc_reduce_kernel<<<M, 256, 0, (cudaStream_t)stream>>>(e_ss, e_sy, npix, sums);
  return phs_check_launch("ncc_maps");
}

int phs_sample_moments(const float* logits, const uint8_t* gt, int S, int B, int64_t npix, int nlabels, int kind, float lo,
                       float hi, double* acc, void* stream) {
  PHS_REQUIRE(logits && acc, "phs_sample_moments: null argument");
  PHS_REQUIRE(nlabels >= 1 && nlabels <= MET_MAXL && S >= 1 && B >= 1 && npix >= 1, "phs_sample_moments: bad sizes");
  PHS_REQUIRE(kind == 0 || kind == 1, "phs_sample_moments: kind must be 0 (softmax) or 1 (clipped logits)");
  const int64_t total = (int64_t)B * npix;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  sample_moments_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(logits, gt, S, B, npix, nlabels, kind, lo, hi, acc);
  return phs_check_launch("sample_moments");
}

int phs_sample_maps(const double* acc, int64_t total_pix, int nlabels, int count, int drop_last, int64_t* mean_arg,
                    float* std_mean, float* var_sum, float* cov_det, float* err, void* stream) {
  PHS_REQUIRE(acc, "phs_sample_maps: null argument");
  PHS_REQUIRE(nlabels >= 1 && nlabels <= MET_MAXL && count >= 1 && total_pix >= 1 && drop_last >= 0 && drop_last < nlabels,
              "phs_sample_maps: bad sizes");
  const int blocks = (int)((total_pix + 255) / 256 < 148 * 8 ? (total_pix + 255) / 256 : 148 * 8);
  sample_maps_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(acc, total_pix, nlabels, count, drop_last, mean_arg, std_mean,
                                                              var_sum, cov_det, err);
  return phs_check_launch("sample_maps");
}

}  // extern "C"
