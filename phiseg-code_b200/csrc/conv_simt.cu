// CUDA-core (fp32 accumulate) implicit-GEMM convolution: forward, input gradient, filter gradient.
// Used (a) as the fp32 "parity mode" of the engine, (b) for the layers whose shapes the tensor-core kernels do
// not take (Cin in {1,2,3,5}, Cout in {2,4,6}: network inputs, latent heads, y_lvl heads), and (c) as the on-device
// cross-check of the tcgen05 kernels.  Reference semantics: tf.nn.conv2d SAME stride 1 (tfwrapper/layers.py:123).
#include "common.cuh"

namespace {

struct ConvGeom {
  int N, H, W;
  int Cin, Cout;  // channels of the kernel's *input* and *output* tensors (swapped roles under dgrad)
  int ks, dgrad;
  int ldx, ldy;
  int64_t M;  // N*H*W
  int K;      // ks*ks*Cin
};

// B operand element: forward W[tap][ci][co] (HWIO); dgrad: in=dy (co), out=dx (ci): W[flip tap][ci=out][co=in]
__device__ __forceinline__ float weight_at(const float* __restrict__ w, const ConvGeom& g, int k, int oc) {
  int taps = g.ks * g.ks;
  int tap = k / g.Cin, ic = k - tap * g.Cin;
  if (!g.dgrad) return w[((size_t)tap * g.Cin + ic) * g.Cout + oc];
  return w[((size_t)(taps - 1 - tap) * g.Cout + oc) * g.Cin + ic];
}

template <typename TI, typename TO, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256)
    conv_simt_kernel(const TI* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                     TO* __restrict__ y, ConvGeom g, int accumulate) {
  constexpr int BK = 16;
  static_assert((BM / TM) * (BN / TN) == 256, "thread tiling");
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int pad = g.ks / 2;

  // A-tile loader assignment: element e = tid + i*256 -> (k = e % BK, m = e / BK)
  constexpr int A_PER = BM * BK / 256;
  int a_k = tid % BK;
  int a_h[A_PER], a_w[A_PER];
  int64_t a_base[A_PER];
  bool a_ok[A_PER];
#pragma unroll
  for (int i = 0; i < A_PER; ++i) {
    int64_t m = m0 + (tid + i * 256) / BK;
    a_ok[i] = m < g.M;
    int64_t mm = a_ok[i] ? m : 0;
    int wq = (int)(mm % g.W);
    int64_t t = mm / g.W;
    a_h[i] = (int)(t % g.H);
    a_w[i] = wq;
    a_base[i] = (t / g.H) * (int64_t)g.H * g.W;  // pixel index of (n,0,0)
  }
  constexpr int B_PER = (BK * BN + 255) / 256;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += BK) {
    {
      int k = k0 + a_k;
      int tap = k / g.Cin, ic = k - tap * g.Cin;
      int dh = tap / g.ks - pad, dw = tap % g.ks - pad;
#pragma unroll
      for (int i = 0; i < A_PER; ++i) {
        float v = 0.f;
        int hh = a_h[i] + dh, ww = a_w[i] + dw;
        if (a_ok[i] && k < g.K && hh >= 0 && hh < g.H && ww >= 0 && ww < g.W)
          v = ldf<TI>(x + (a_base[i] + (int64_t)hh * g.W + ww) * g.ldx + ic);
        As[a_k][(tid + i * 256) / BK] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int e = tid + i * 256;
      if (e < BK * BN) {
        int kk = e / BN, nn = e % BN;
        int k = k0 + kk, oc = n0 + nn;
        Bs[kk][nn] = (k < g.K && oc < g.Cout) ? weight_at(w, g, k, oc) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int oc = n0 + tx * TN + j;
      if (oc >= g.Cout) continue;
      float v = acc[i][j] + (bias ? bias[oc] : 0.f);
      TO* q = y + m * g.ldy + oc;
      if (accumulate) v += ldf<TO>(q);
      stf<TO>(q, v);
    }
  }
}

// filter gradient: dW[k][co] += sum_m A[m][k] * dY[m][co]; tile 64(k) x 64(co), split over m (blockIdx.z)
template <typename TX, typename TD>
__global__ void __launch_bounds__(256)
    wgrad_simt_kernel(const TX* __restrict__ x, const TD* __restrict__ dy, float* __restrict__ dw, ConvGeom g,
                      int64_t m_per_split) {
  constexpr int BKO = 64, BN = 64, BMR = 16;
  __shared__ float As[BMR][BKO + 4];
  __shared__ float Bs[BMR][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int k0 = blockIdx.x * BKO, n0 = blockIdx.y * BN;
  const int pad = g.ks / 2;
  const int64_t ms = (int64_t)blockIdx.z * m_per_split;
  const int64_t me = min(g.M, ms + m_per_split);
  // loader: column (k or co) = tid % 64, rows tid/64 + 4*i
  const int col = tid % 64, r0 = tid / 64;
  const int k = k0 + col;
  const bool k_ok = k < g.K;
  int tap = k_ok ? k / g.Cin : 0;
  int ic = k - tap * g.Cin;
  int dh = tap / g.ks - pad, dwd = tap % g.ks - pad;
  const int oc = n0 + col;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t mb = ms; mb < me; mb += BMR) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int r = r0 + 4 * i;
      int64_t m = mb + r;
      float va = 0.f, vb = 0.f;
      if (m < me) {
        int wq = (int)(m % g.W);
        int64_t t = m / g.W;
        int hq = (int)(t % g.H);
        int64_t nb = (t / g.H) * (int64_t)g.H * g.W;
        int hh = hq + dh, ww = wq + dwd;
        if (k_ok && hh >= 0 && hh < g.H && ww >= 0 && ww < g.W) va = ldf<TX>(x + (nb + (int64_t)hh * g.W + ww) * g.ldx + ic);
        if (oc < g.Cout) vb = ldf<TD>(dy + m * g.ldy + oc);
      }
      As[r][col] = va;
      Bs[r][col] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < BMR; ++r) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[r][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[r][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int kk = k0 + ty * 4 + i;
    if (kk >= g.K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int o = n0 + tx * 4 + j;
      if (o < g.Cout) atomicAdd(dw + (size_t)kk * g.Cout + o, acc[i][j]);
    }
  }
}

template <typename TD>
__global__ void __launch_bounds__(256) bias_grad_kernel(const TD* __restrict__ dy, int ld, int C, int64_t M, float* __restrict__ db) {
  // small C only: thread t handles channel t % C over a strided set of pixels
  int c = threadIdx.x % C;
  int lanes = blockDim.x / C;
  int lane = threadIdx.x / C;
  float a = 0.f;
  if (lane < lanes)
    for (int64_t m = (int64_t)blockIdx.x * lanes + lane; m < M; m += (int64_t)gridDim.x * lanes) a += ldf<TD>(dy + m * ld + c);
  __shared__ float sm[256];
  sm[threadIdx.x] = lane < lanes ? a : 0.f;
  __syncthreads();
  if (threadIdx.x < C) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += sm[l * C + threadIdx.x];
    atomicAdd(db + threadIdx.x, s);
  }
}

}  // namespace

int small_conv_try(const phs_tensor* x, const float* w, const float* bias, const phs_tensor* y, int ksize, int dgrad,
                   int accumulate, cudaStream_t st);
int small_wgrad_try(const phs_tensor* x, const phs_tensor* dy, float* dw, int ksize, cudaStream_t st);

int conv2d_simt(const phs_tensor* x, const float* w, const float* bias, const phs_tensor* y, int ksize, int dgrad,
                int accumulate, cudaStream_t st) {
  {
    // tiny channel count on one side: dedicated streaming kernels (small_conv.cu)
    int r = small_conv_try(x, w, bias, y, ksize, dgrad, accumulate, st);
    if (r != 0) return r == 1 ? 0 : r;
  }
  ConvGeom g;
  g.N = x->N; g.H = x->H; g.W = x->W;
  g.Cin = x->C; g.Cout = y->C;
  g.ks = ksize; g.dgrad = dgrad;
  g.ldx = x->ld; g.ldy = y->ld;
  g.M = (int64_t)x->N * x->H * x->W;
  g.K = ksize * ksize * x->C;
  const bool narrow = y->C <= 16;
  const int BN = narrow ? 16 : 64;
  dim3 grid((unsigned)ceil_div64(g.M, 64), (unsigned)((y->C + BN - 1) / BN));
#define LAUNCH(TI, TO)                                                                                               \
  do {                                                                                                               \
    if (narrow)                                                                                                      \
      conv_simt_kernel<TI, TO, 64, 16, 4, 1><<<grid, 256, 0, st>>>((const TI*)x->ptr, w, bias, (TO*)y->ptr, g, accumulate); \
    else                                                                                                             \
      conv_simt_kernel<TI, TO, 64, 64, 4, 4><<<grid, 256, 0, st>>>((const TI*)x->ptr, w, bias, (TO*)y->ptr, g, accumulate); \
  } while (0)
  if (x->dtype == PHS_F32 && y->dtype == PHS_F32) LAUNCH(float, float);
  else if (x->dtype == PHS_F32) LAUNCH(float, bf16);
  else if (y->dtype == PHS_F32) LAUNCH(bf16, float);
  else LAUNCH(bf16, bf16);
#undef LAUNCH
  return phs_check_launch("conv2d_simt");
}

int conv2d_wgrad_simt(const phs_tensor* x, const phs_tensor* dy, float* dw, float* db, int ksize, int accumulate,
                      cudaStream_t st) {
  ConvGeom g;
  g.N = x->N; g.H = x->H; g.W = x->W;
  g.Cin = x->C; g.Cout = dy->C;
  g.ks = ksize; g.dgrad = 0;
  g.ldx = x->ld; g.ldy = dy->ld;
  g.M = (int64_t)x->N * x->H * x->W;
  g.K = ksize * ksize * x->C;
  if (!accumulate) {
    cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)g.K * g.Cout, st);
    if (db) cudaMemsetAsync(db, 0, sizeof(float) * g.Cout, st);
  }
  int rc = small_wgrad_try(x, dy, dw, ksize, st);
  if (rc != 0 && rc != 1) return rc;
  const bool handled = rc == 1;
  int gx = (g.K + 63) / 64, gy = (g.Cout + 63) / 64;
  int64_t splits = (148 * 4 + gx * gy - 1) / (gx * gy);
  int64_t max_splits = ceil_div64(g.M, 256);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t mps = ceil_div64(ceil_div64(g.M, splits), 16) * 16;
  splits = ceil_div64(g.M, mps);
  dim3 grid(gx, gy, (unsigned)splits);
  if (handled) {
  } else if (x->dtype == PHS_F32 && dy->dtype == PHS_F32)
    wgrad_simt_kernel<float, float><<<grid, 256, 0, st>>>((const float*)x->ptr, (const float*)dy->ptr, dw, g, mps);
  else if (x->dtype == PHS_F32)
    wgrad_simt_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)x->ptr, (const bf16*)dy->ptr, dw, g, mps);
  else if (dy->dtype == PHS_F32)
    wgrad_simt_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)x->ptr, (const float*)dy->ptr, dw, g, mps);
  else
    wgrad_simt_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)x->ptr, (const bf16*)dy->ptr, dw, g, mps);
  rc = phs_check_launch("conv2d_wgrad_simt");
  if (rc) return rc;
  if (db) {
    PHS_REQUIRE(dy->C <= 256, "bias gradient: C=%d too large for the head kernel", dy->C);
    int blocks = (int)(ceil_div64(g.M, 256) < 296 ? ceil_div64(g.M, 256) : 296);
    if (dy->dtype == PHS_F32)
      bias_grad_kernel<float><<<blocks, 256, 0, st>>>((const float*)dy->ptr, dy->ld, dy->C, g.M, db);
    else
      bias_grad_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)dy->ptr, dy->ld, dy->C, g.M, db);
    rc = phs_check_launch("bias_grad");
  }
  return rc;
}
