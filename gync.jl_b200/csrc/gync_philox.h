// Philox4x32-10 counter-based RNG and the draw addressing used by the Metropolis kernel.
// Host+device (the host build is only for the known-answer tests).
#pragma once
#include <math.h>
#include <stdint.h>
#ifndef GYNC_HD
#  ifdef __CUDACC__
#    define GYNC_HD __host__ __device__ __forceinline__
#  else
#    define GYNC_HD inline
#  endif
#endif

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) — counter-based, so a (seed, chain, iteration, draw)
// tuple addresses every random number and runs are resumable from `iter_offset`.
struct Philox {
  static GYNC_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  static GYNC_HD void gen(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, hi0, lo0);
      mulhilo(0xCD9E8D57u, c2, hi1, lo1);
      const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
  // 53-bit uniform in (0,1) from two 32-bit words
  static GYNC_HD double u01(uint32_t hi, uint32_t lo) {
    const uint64_t v = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
  }
};

// draw block b of (chain, iteration): two uniforms
GYNC_HD static void gync_draw_u2(uint64_t seed, uint64_t chain, uint64_t iter, uint32_t block, double& u0, double& u1) {
  uint32_t r[4];
  Philox::gen(seed, block, (uint32_t)iter, (uint32_t)(iter >> 32), (uint32_t)chain, r);
  u0 = Philox::u01(r[0], r[1]);
  u1 = Philox::u01(r[2], r[3]);
}

// two standard normals (Box-Muller)
GYNC_HD static void gync_draw_n2(uint64_t seed, uint64_t chain, uint64_t iter, uint32_t block, double& z0, double& z1) {
  double u0, u1;
  gync_draw_u2(seed, chain, iter, block, u0, u1);
  const double r = sqrt(-2.0 * log(u0));
  const double a = 6.283185307179586476925286766559 * u1;
  z0 = r * cos(a);
  z1 = r * sin(a);
}

