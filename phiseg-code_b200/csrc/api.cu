// C-ABI plumbing of libphiseg_sm100.so: error reporting, version/arch queries and the convolution dispatchers
// (CUDA-core fp32 kernels vs tcgen05 tensor-core kernels).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void phs_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Programmatic dependent launch is OPT-IN (PHS_PDL=1) since the end of round 2: it bought 0.29 ms per step, but two training
// runs stalled inside a step after the last kernel reworks and the cause could not be pinned down on the device any more
// (DESIGN.md section 2).  Off, the kernels of a lane are strictly serialised and griddepcontrol.* are no-ops - the
// execution model the round-1 driver runs (1 to 8 GPUs) went through.
int phs_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PHS_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

// PHS_PDL_TC=1: tensor-memory kernels are launched early too (off by default, see PHS_PDL_WAIT in common.cuh)
int phs_pdl_tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PHS_PDL_TC");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

int phs_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    phs_set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// internal entry points (conv_simt.cu, conv_tc.cu)
int conv2d_simt(const phs_tensor* x, const float* w, const float* bias, const phs_tensor* y, int ksize, int dgrad,
                int accumulate, cudaStream_t st);
int conv2d_wgrad_simt(const phs_tensor* x, const phs_tensor* dy, float* dw, float* db, int ksize, int accumulate,
                      cudaStream_t st);
int conv2d_tc(const phs_tensor* x, const void* w, const float* bias, const phs_tensor* y, int ksize, int dgrad,
              int accumulate, double* stats, cudaStream_t st);
int conv2d_wgrad_tc(const phs_tensor* x, const phs_tensor* dy, float* dw, float* db, int ksize, int accumulate,
                    cudaStream_t st);

extern "C" {

int phs_version(void) { return 100; }

/* CRC32C (Castagnoli) of a host buffer, continuing from `crc` (0 to start): the checksum TensorFlow checkpoint bundles and
 * the LevelDB table format carry (tfwrapper/checkpoint.py).  Plain host code: slicing-by-1 table, ~400 MB/s. */
unsigned int phs_crc32c(const void* data, size_t n, unsigned int crc) {
  static unsigned int table[256];
  static bool init = false;
  if (!init) {
    for (unsigned int i = 0; i < 256; ++i) {
      unsigned int c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  unsigned int c = crc ^ 0xFFFFFFFFu;
  const unsigned char* p = (const unsigned char*)data;
  for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
int phs_arch(void) { return 100; }
const char* phs_last_error(void) { return g_err; }

int phs_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

static int check_conv_args(const char* fn, const phs_tensor* x, const void* w, const phs_tensor* y, int ksize) {
  PHS_REQUIRE(x && y && w && x->ptr && y->ptr, "%s: null argument", fn);
  PHS_REQUIRE(ksize == 1 || ksize == 3, "%s: ksize=%d (only 1 and 3: the reference uses no other, layers.py:96)", fn, ksize);
  PHS_REQUIRE(x->N == y->N && x->H == y->H && x->W == y->W, "%s: SAME stride-1 needs equal N,H,W (%d,%d,%d) vs (%d,%d,%d)",
              fn, x->N, x->H, x->W, y->N, y->H, y->W);
  PHS_REQUIRE(x->N > 0 && x->H > 0 && x->W > 0 && x->C > 0 && y->C > 0, "%s: empty tensor", fn);
  PHS_REQUIRE(x->ld >= x->C && y->ld >= y->C, "%s: pitch smaller than channel count", fn);
  PHS_REQUIRE((x->dtype == PHS_F32 || x->dtype == PHS_BF16) && (y->dtype == PHS_F32 || y->dtype == PHS_BF16),
              "%s: bad dtype", fn);
  return 0;
}

int phs_conv2d(const phs_tensor* x, const void* w, const float* bias, const phs_tensor* y, int ksize, int dgrad,
               int accumulate, int impl, void* stream) {
  int rc = check_conv_args("phs_conv2d", x, w, y, ksize);
  if (rc) return rc;
  if (impl == PHS_IMPL_SIMT) return conv2d_simt(x, (const float*)w, bias, y, ksize, dgrad, accumulate, (cudaStream_t)stream);
  if (impl == PHS_IMPL_TC) return conv2d_tc(x, w, bias, y, ksize, dgrad, accumulate, nullptr, (cudaStream_t)stream);
  PHS_REQUIRE(false, "phs_conv2d: unknown impl %d", impl);
}

int phs_conv2d_stats(const phs_tensor* x, const void* w, const float* bias, const phs_tensor* y, int ksize,
                     double* stats, void* stream) {
  int rc = check_conv_args("phs_conv2d_stats", x, w, y, ksize);
  if (rc) return rc;
  PHS_REQUIRE(stats, "phs_conv2d_stats: null stats");
  return conv2d_tc(x, w, bias, y, ksize, 0, 0, stats, (cudaStream_t)stream);
}

int phs_conv2d_stats_acc(const phs_tensor* x, const void* w, const float* bias, const phs_tensor* y, int ksize,
                         double* stats, void* stream) {
  int rc = check_conv_args("phs_conv2d_stats_acc", x, w, y, ksize);
  if (rc) return rc;
  PHS_REQUIRE(stats, "phs_conv2d_stats_acc: null stats");
  return conv2d_tc(x, w, bias, y, ksize, 0, 2, stats, (cudaStream_t)stream);
}

int phs_conv2d_wgrad(const phs_tensor* x, const phs_tensor* dy, float* dw, float* db, int ksize, int accumulate,
                     int impl, void* stream) {
  int rc = check_conv_args("phs_conv2d_wgrad", x, dw, dy, ksize);
  if (rc) return rc;
  if (impl == PHS_IMPL_SIMT) return conv2d_wgrad_simt(x, dy, dw, db, ksize, accumulate, (cudaStream_t)stream);
  if (impl == PHS_IMPL_TC) return conv2d_wgrad_tc(x, dy, dw, db, ksize, accumulate, (cudaStream_t)stream);
  PHS_REQUIRE(false, "phs_conv2d_wgrad: unknown impl %d", impl);
}

}  // extern "C"
