// Input pipeline on the device (data/batch_provider.py:43-67,131-272, utils.py:18-37).
//
// The reference keeps the data set in an HDF5 file, gathers a batch on the host, picks one annotator per image and
// augments image by image with OpenCV: cv2.warpAffine (random rotation, bilinear, constant border 0) followed by
// cv2.resize of a random square crop back to the full size (bilinear), label masks as one-hot planes + argmax, then
// optional flips.  Here the data set is RESIDENT in HBM (LIDC: ~1.2 GB of 180), the host only draws the random
// parameters (in the reference's np.random call order) and ONE launch gathers, picks the annotator, augments and
// converts a whole batch: one thread per output pixel evaluates flip -> crop-resize taps -> rotation taps on the fly,
// so no intermediate image exists.
//
// The resampling arithmetic restates OpenCV's: warpAffine's fixed-point source coordinates (AB_BITS = 10, 1/32-pixel
// interpolation table, round-half-even), resize's float coefficients with its different clamping rules along x and y,
// the left-to-right summation order, one-hot planes accumulated in double and a first-maximum argmax.  Products and
// sums are written with explicit round-to-nearest intrinsics so that the compiler cannot contract them into FMAs.
#include "common.cuh"

namespace {

constexpr int AUG_ROTATE = 1, AUG_SCALE = 2, AUG_FLIPLR = 4, AUG_FLIPUD = 8;

template <typename W> __device__ __forceinline__ W mul_rn(W a, W b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename W> __device__ __forceinline__ W add_rn(W a, W b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

// warpAffine: source position of destination pixel (y, x) under the inverted matrix m, in 1/32-pixel fixed point
struct RotTaps {
  int ix, iy;
  float w[4];   // (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy fx : exact multiples of 2^-10
};

__device__ __forceinline__ RotTaps rot_taps(const double* m, int y, int x) {
  const double AB_SCALE = 1024.0;
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(m[0], (double)x), AB_SCALE));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(m[3], (double)x), AB_SCALE));
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)y), m[2]), AB_SCALE)) + 16;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)y), m[5]), AB_SCALE)) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  RotTaps t;
  t.ix = X >> 5;
  t.iy = Y >> 5;
  const float fx = (float)(X & 31) * (1.f / 32.f), fy = (float)(Y & 31) * (1.f / 32.f);
  const float gx = 1.f - fx, gy = 1.f - fy;
  t.w[0] = gy * gx; t.w[1] = gy * fx; t.w[2] = fy * gx; t.w[3] = fy * fx;
  return t;
}

// image after the (optional) rotation at integer position (y, x) of the full frame
template <typename T>
__device__ __forceinline__ T rotated_pixel(const T* __restrict__ img, int H, int W, const phs_aug_params& p, int y, int x) {
  if (!(p.flags & AUG_ROTATE)) return img[(size_t)y * W + x];
  const RotTaps t = rot_taps(p.minv, y, x);
  T acc = (T)0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int yy = t.iy + (k >> 1), xx = t.ix + (k & 1);
    const T v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[(size_t)yy * W + xx] : (T)0;   // BORDER_CONSTANT, 0
    const T term = mul_rn<T>(v, (T)t.w[k]);
    acc = k == 0 ? term : add_rn<T>(acc, term);
  }
  return acc;
}

// label after the (optional) rotation: one-hot planes through the same taps, first maximum (np.argmax)
__device__ __forceinline__ int rotated_label(const uint8_t* __restrict__ lab, int H, int W, int A, int nl,
                                             const phs_aug_params& p, int y, int x) {
  if (!(p.flags & AUG_ROTATE)) return lab[((size_t)y * W + x) * A];
  const RotTaps t = rot_taps(p.minv, y, x);
  int l4[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int yy = t.iy + (k >> 1), xx = t.ix + (k & 1);
    l4[k] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (int)lab[((size_t)yy * W + xx) * A] : -1;
  }
  int best = 0;
  double best_v = 0.0;
  for (int c = 0; c < nl; ++c) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double term = l4[k] == c ? (double)t.w[k] : 0.0;
      acc = k == 0 ? term : __dadd_rn(acc, term);
    }
    if (c == 0 || acc > best_v) { best = c; best_v = acc; }
  }
  return best;
}

// cv2.resize INTER_LINEAR taps of destination index d for a source of `src` samples stretched to `dst`
struct ResizeTap { int i0, i1; float a0, a1; };

__device__ __forceinline__ ResizeTap resize_tap(int d, int src, int dst, bool along_x) {
  const double scale = 1.0 / ((double)dst / (double)src);
  float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  ResizeTap t;
  if (along_x) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
    t.i0 = s;
    t.i1 = min(s + 1, src - 1);
  } else {       // rows are clamped, their coefficients are not touched
    t.i0 = min(max(s, 0), src - 1);
    t.i1 = min(max(s + 1, 0), src - 1);
  }
  t.a0 = 1.f - f;
  t.a1 = f;
  return t;
}

template <typename T>
__global__ void __launch_bounds__(256) augment_kernel(const T* __restrict__ images, const uint8_t* __restrict__ labels, int H,
                                                      int W, int A, int nl, const phs_aug_params* __restrict__ params,
                                                      float* __restrict__ x_out, uint8_t* __restrict__ s_out) {
  const int b = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= H * W) return;
  const phs_aug_params p = params[b];
  const T* img = images + (size_t)p.src * H * W;
  const uint8_t* lab = labels ? labels + (size_t)p.src * H * W * A + p.annot : nullptr;
  int y = q / W, x = q % W;
  if (p.flags & AUG_FLIPLR) x = W - 1 - x;
  if (p.flags & AUG_FLIPUD) y = H - 1 - y;
  T v;
  int l = 0;
  if (p.flags & AUG_SCALE) {
    // resize_image(img[py:py+r, px:px+r], (H, W)): horizontal pass in the source rows, then the vertical blend
    const ResizeTap tx = resize_tap(x, p.crop, W, true), ty = resize_tap(y, p.crop, H, false);
    const int r0 = p.py + ty.i0, r1 = p.py + ty.i1, c0 = p.px + tx.i0, c1 = p.px + tx.i1;
    const T h0 = add_rn<T>(mul_rn<T>(rotated_pixel<T>(img, H, W, p, r0, c0), (T)tx.a0),
                           mul_rn<T>(rotated_pixel<T>(img, H, W, p, r0, c1), (T)tx.a1));
    const T h1 = add_rn<T>(mul_rn<T>(rotated_pixel<T>(img, H, W, p, r1, c0), (T)tx.a0),
                           mul_rn<T>(rotated_pixel<T>(img, H, W, p, r1, c1), (T)tx.a1));
    v = add_rn<T>(mul_rn<T>(h0, (T)ty.a0), mul_rn<T>(h1, (T)ty.a1));
    if (lab) {
      const int l00 = rotated_label(lab, H, W, A, nl, p, r0, c0), l01 = rotated_label(lab, H, W, A, nl, p, r0, c1);
      const int l10 = rotated_label(lab, H, W, A, nl, p, r1, c0), l11 = rotated_label(lab, H, W, A, nl, p, r1, c1);
      double best_v = 0.0;
      for (int c = 0; c < nl; ++c) {
        const double g0 = __dadd_rn(l00 == c ? (double)tx.a0 : 0.0, l01 == c ? (double)tx.a1 : 0.0);
        const double g1 = __dadd_rn(l10 == c ? (double)tx.a0 : 0.0, l11 == c ? (double)tx.a1 : 0.0);
        const double acc = __dadd_rn(__dmul_rn(g0, (double)ty.a0), __dmul_rn(g1, (double)ty.a1));
        if (c == 0 || acc > best_v) { l = c; best_v = acc; }
      }
    }
  } else {
    v = rotated_pixel<T>(img, H, W, p, y, x);
    if (lab) l = rotated_label(lab, H, W, A, nl, p, y, x);
  }
  x_out[(size_t)b * H * W + q] = (float)v;
  if (s_out) s_out[(size_t)b * H * W + q] = (uint8_t)l;
}

}  // namespace

extern "C" {

int phs_augment_batch(const void* images, int image_dtype, const uint8_t* labels, int H, int W, int annotators, int nlabels,
                      const phs_aug_params* params, int B, float* x_out, uint8_t* s_out, void* stream) {
  PHS_REQUIRE(images && params && x_out, "phs_augment_batch: null argument");
  PHS_REQUIRE((labels == nullptr) == (s_out == nullptr), "phs_augment_batch: labels and s_out go together");
  PHS_REQUIRE(image_dtype == PHS_F32 || image_dtype == PHS_F64, "phs_augment_batch: images must be float32 or float64");
  PHS_REQUIRE(H >= 2 && W >= 2 && H <= 4096 && W <= 4096 && B >= 1 && B <= 65535 && annotators >= 1,
              "phs_augment_batch: bad sizes");
  // more than 4 labels: the reference switches to nearest-neighbour label interpolation (batch_provider.py:203-207)
  PHS_REQUIRE(nlabels >= 1 && nlabels <= 4, "phs_augment_batch: one-hot label interpolation covers nlabels <= 4 (got %d)", nlabels);
  const dim3 grid((H * W + 255) / 256, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (image_dtype == PHS_F64)
    augment_kernel<double><<<grid, 256, 0, st>>>((const double*)images, labels, H, W, annotators, nlabels, params, x_out, s_out);
  else
    augment_kernel<float><<<grid, 256, 0, st>>>((const float*)images, labels, H, W, annotators, nlabels, params, x_out, s_out);
  return phs_check_launch("augment_batch");
}

}  // extern "C"
