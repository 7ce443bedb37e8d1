// Host-side check of the multiply-high division constants (csrc/common.cuh: fdiv_make) that the streaming kernels use
// for their index arithmetic: q = (umulhi(n, m) + n) >> s must equal n / d for every n < 2^31.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../phiseg-code_b200/csrc/common.cuh"

static uint32_t fdiv_host(uint32_t n, const fdiv_t& f) { return (uint32_t)(((((uint64_t)n * f.m) >> 32) + n) >> f.s); }

int main() {
  uint64_t checked = 0;
  uint32_t seed = 12345u;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed; };
  for (uint32_t d = 1; d <= 4096; ++d) {
    const fdiv_t f = fdiv_make(d);
    const uint32_t probes[] = {0u, 1u, d - 1, d, d + 1, 2 * d - 1, 2 * d, 0x7fffffffu, 0x7ffffffeu, 0x40000000u};
    for (uint32_t n : probes) {
      if (n > 0x7fffffffu) continue;
      if (fdiv_host(n, f) != n / d) { printf("FAIL d=%u n=%u\n", d, n); return 1; }
      ++checked;
    }
    for (int k = 0; k < 64; ++k) {
      const uint32_t n = rnd() & 0x7fffffffu;
      if (fdiv_host(n, f) != n / d) { printf("FAIL d=%u n=%u\n", d, n); return 1; }
      ++checked;
    }
  }
  for (int k = 0; k < 200000; ++k) {
    const uint32_t d = (rnd() % 0x00ffffffu) + 1, n = rnd() & 0x7fffffffu;
    const fdiv_t f = fdiv_make(d);
    if (fdiv_host(n, f) != n / d) { printf("FAIL d=%u n=%u\n", d, n); return 1; }
    ++checked;
  }
  // idx4: flat index -> (vector, w, h, n)
  const idx4_t ix = idx4_make(24, 40, 24);
  for (uint32_t i = 0; i < 24u * 40u * 24u * 3u; i += 7) {
    const uint32_t pix = fdiv_host(i, ix.nvec), cv = i - pix * 24u;
    const uint32_t row = fdiv_host(pix, ix.W), w = pix - row * 40u;
    const uint32_t img = fdiv_host(row, ix.H), h = row - img * 24u;
    if (((img * 24u + h) * 40u + w) * 24u + cv != i) { printf("FAIL idx4 i=%u\n", i); return 1; }
    ++checked;
  }
  printf("OK %llu\n", (unsigned long long)checked);
  return 0;
}
