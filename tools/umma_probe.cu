// Hardware probe (not part of the library): how does tcgen05.mma address a swizzled shared-memory operand whose start
// address is NOT aligned to the swizzle pattern and whose 8-row groups are NOT a multiple of the pattern apart?
// This decides whether one "halo" tile in shared memory can serve all 9 taps of a 3x3 convolution through shifted
// descriptors.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <vector>
#include "../phiseg-code_b200/csrc/tc_ptx.cuh"

using namespace tc;
typedef __nv_bfloat16 bf16;

struct Probe {
  int kind;         // 0: K-major A, 1: MN-major A
  int row_bytes;    // 128 (SW128) or 64 (SW64)
  int shift_rows;   // start address offset in rows
  int pitch_rows;   // K-major: rows between 8-row groups (SBO = pitch*row_bytes); MN-major: SBO rows (normally 8)
  int base_mode;    // 0: base_offset 0, 1: base_offset = (start >> 7) & 7
  int value_mode;   // 0: element value = row index, 1: element value = column index
  int M;            // 128 or 64
};

// smem image: rows of row_bytes, value by (row, col), stored with the TMA swizzle on ABSOLUTE address bits
__device__ void fill_rows(uint8_t* base, int rows, int row_bytes, int value_mode, int col0) {
  const int cols = row_bytes / 2;
  for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
    int r = i / cols, c = i % cols;
    uint32_t lin = r * row_bytes + c * 2;
    uint32_t a = smem_u32(base) + lin;
    uint32_t mask = row_bytes == 128 ? 7u : 3u;
    uint32_t phys = a ^ (((a >> 7) & mask) << 4);
    float v = value_mode == 0 ? (float)r : (float)(c + col0);
    *reinterpret_cast<bf16*>(base + (phys - smem_u32(base))) = __float2bfloat16(v);
  }
}

__global__ void probe_kernel(Probe p, float* out /* [128 lanes][64 cols] */) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  uint8_t* base = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
  uint8_t* a_img = base;                 // up to 64 KB
  uint8_t* b_img = base + 64 * 1024;     // identity
  const int N = p.kind == 0 ? 64 : 16;
  // zero everything
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
  __syncthreads();
  if (p.kind == 0) {
    fill_rows(a_img, 400, p.row_bytes, p.value_mode, 0);
    // B[n][k] = delta(n,k), K-major rows of row_bytes (N = K = row_bytes/2 = 64, or 32 for SW64 -> N = 64 needs K = 64:
    // use two K chunks for SW64? keep it simple: for SW64 K = 32, N = 64, rows n >= 32 are zero)
    const int cols = p.row_bytes / 2;
    for (int n = threadIdx.x; n < 64; n += blockDim.x) {
      if (n < cols) {
        uint32_t lin = n * p.row_bytes + n * 2;
        uint32_t a = smem_u32(b_img) + lin;
        uint32_t mask = p.row_bytes == 128 ? 7u : 3u;
        uint32_t phys = a ^ (((a >> 7) & mask) << 4);
        *reinterpret_cast<bf16*>(b_img + (phys - smem_u32(b_img))) = __float2bfloat16(1.f);
      }
    }
  } else {
    // MN-major A: two slabs of (rows = K) x (64 or 32 M-columns); slab 1 holds M columns cols..2*cols-1
    const int cols = p.row_bytes / 2;
    const int slab_bytes = 16 * 1024;
    fill_rows(a_img, 128, p.row_bytes, p.value_mode, 0);
    fill_rows(a_img + slab_bytes, 128, p.row_bytes, p.value_mode, cols);
    // B MN-major: [K rows = 16][N = 16 cols]... store as rows of 32 B?  Use K-major B instead: B[n][k] = delta(n,k),
    // N = 16, K = 16, no swizzle needed beyond row 0..15 of a SW32 tile: keep SW128 rows (only first 32 B used)
    for (int n = threadIdx.x; n < 16; n += blockDim.x) {
      uint32_t lin = n * 128 + n * 2;
      uint32_t a = smem_u32(b_img) + lin;
      uint32_t phys = a ^ (((a >> 7) & 7u) << 4);
      *reinterpret_cast<bf16*>(b_img + (phys - smem_u32(b_img))) = __float2bfloat16(1.f);
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_s), 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  // zero the accumulator region first via an MMA with zero operands?  Simply rely on accumulate=0 of the first MMA;
  // for M=64 the other lanes keep garbage: pre-store zeros with tcgen05.st
  {
    uint32_t lane_base = tmem + ((uint32_t)((threadIdx.x / 32) * 32) << 16);
    for (int c = 0; c < 64; ++c)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_base + c), "r"(0x7fc00000u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint64_t lay = p.row_bytes == 128 ? LAYOUT_SW128 : LAYOUT_SW64;
    if (p.kind == 0) {
      const uint32_t a0 = smem_u32(a_img) + p.shift_rows * p.row_bytes;
      const uint32_t bo = p.base_mode ? ((a0 >> 7) & 7) : 0;
      const uint32_t idesc = idesc_bf16(p.M, 64, 0, 0);
      const int ksteps = p.row_bytes / 32;
      for (int k = 0; k < ksteps; ++k) {
        uint64_t da = smem_desc(a0 + k * 32, 16, p.pitch_rows * p.row_bytes, lay, bo);
        uint64_t db = smem_desc(smem_u32(b_img) + k * 32, 16, 8 * p.row_bytes, lay, 0);
        umma_bf16(tmem, da, db, idesc, k != 0);
      }
    } else {
      const uint32_t a0 = smem_u32(a_img) + p.shift_rows * p.row_bytes;
      const uint32_t bo = p.base_mode ? ((a0 >> 7) & 7) : 0;
      const uint32_t idesc = idesc_bf16(p.M, 16, 1, 0);
      uint64_t da = smem_desc(a0, 16 * 1024, p.pitch_rows * p.row_bytes, lay, bo);
      uint64_t db = smem_desc(smem_u32(b_img), 16, 1024, LAYOUT_SW128, 0);
      umma_bf16(tmem, da, db, idesc, 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  {
    const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint32_t r[16];
    for (int c0 = 0; c0 < 64; c0 += 16) {
      tmem_ld16(tmem + ((uint32_t)(w * 32) << 16) + c0, r);
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) out[(w * 32 + lane) * 64 + c0 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 64 * 4);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  std::vector<float> h(128 * 64);
  std::vector<Probe> probes;
  for (int rb : {128, 64})
    for (int shift : {0, 1, 2, 3, 10, 11})
      for (int pitch : {8, 10, 16})
        for (int bm : {0, 1}) {
          if (shift % 8 == 0 && bm == 1) continue;
          probes.push_back(Probe{0, rb, shift, pitch, bm, 0, 128});
          probes.push_back(Probe{0, rb, shift, pitch, bm, 1, 128});
        }
  for (int rb : {128, 64})
    for (int shift : {0, 1, 2, 9})
      for (int bm : {0, 1}) {
        if (shift == 0 && bm == 1) continue;
        probes.push_back(Probe{1, rb, shift, 8, bm, 0, 128});
        probes.push_back(Probe{1, rb, shift, 8, bm, 1, 128});
      }
  // TMEM layout of an M=64 MMA (K-major, aligned)
  probes.push_back(Probe{0, 128, 0, 8, 0, 0, 64});
  for (const Probe& p : probes) {
    cudaMemset(d_out, 0xff, 128 * 64 * 4);
    probe_kernel<<<1, 128, 100 * 1024>>>(p, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("kind %d rb %d shift %d pitch %d base %d val %d M %d: CUDA error %s\n", p.kind, p.row_bytes, p.shift_rows,
             p.pitch_rows, p.base_mode, p.value_mode, p.M, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(h.data(), d_out, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    int bad = 0, first_bad = -1;
    const int cols = p.row_bytes / 2;
    if (p.M == 64) {
      printf("M=64 TMEM layout: lane -> value(row index) at col 0:");
      for (int l = 0; l < 128; ++l) {
        float v = h[l * 64];
        if (v == v) printf(" %d:%g", l, v);
      }
      printf("\n");
      continue;
    }
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < (p.kind == 0 ? cols : 16); ++n) {
        float expect;
        if (p.kind == 0) {
          int R = (m / 8) * p.pitch_rows + (m % 8) + p.shift_rows;
          expect = p.value_mode == 0 ? (float)R : (float)n;
        } else {
          int R = n + p.shift_rows;   // K row
          expect = p.value_mode == 0 ? (float)R : (float)m;
          if (m >= 2 * cols) continue;
        }
        if (h[m * 64 + n] != expect) {
          if (first_bad < 0) first_bad = m * 64 + n;
          ++bad;
        }
      }
    printf("kind %d rb %3d shift %2d pitch %2d base %d val %d: %s", p.kind, p.row_bytes, p.shift_rows, p.pitch_rows,
           p.base_mode, p.value_mode, bad ? "MISMATCH" : "ok");
    if (bad) {
      int m = first_bad / 64;
      printf(" (%d bad; row %d got:", bad, m);
      for (int n = 0; n < 8; ++n) printf(" %g", h[m * 64 + n]);
      printf(" | row %d:", m + 8 < 128 ? m + 8 : m);
      for (int n = 0; n < 4; ++n) printf(" %g", h[(m + 8 < 128 ? m + 8 : m) * 64 + n]);
      printf(")");
    }
    printf("\n");
  }
  return 0;
}
