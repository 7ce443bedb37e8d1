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

}  // extern "C"
