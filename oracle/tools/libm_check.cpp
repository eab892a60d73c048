// libm_check.cpp -- ORACLE / TEST INFRASTRUCTURE.
// Exhaustively compares v-img_b200/csrc/vimg_libm.cuh (compiled for the host) with the
// host's glibc sinf / cosf / sincosf over every float with |x| < 120 (the range of glibc's one-step reduction), both signs.
//   g++ -O2 -fopenmp -ffp-contract=off oracle/tools/libm_check.cpp -o oracle/_ref/libm_check -lm
// Exit code 0 = bit-identical everywhere.  `libm_check quick` samples every 4099th float.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../v-img_b200/csrc/vimg_libm.cuh"

static float from_bits(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static uint32_t bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

int main(int argc, char** argv) {
  const uint32_t step = (argc > 1 && !strcmp(argv[1], "quick")) ? 4099u : 1u;
  const uint32_t hi = bits(120.0f);
  unsigned long long bad_sin = 0, bad_cos = 0, bad_sc = 0, n = 0;
#pragma omp parallel for reduction(+ : bad_sin, bad_cos, bad_sc, n) schedule(static)
  for (long long i = 0; i < (long long)hi; i += step) {
    for (int sgn = 0; sgn < 2; ++sgn) {
      float x = from_bits((uint32_t)i | (sgn ? 0x80000000u : 0u));
      float s = sinf(x), c = cosf(x), s2, c2;
      sincosf(x, &s2, &c2);
      float ms = vimg_libm::glibc_sinf(x), mc = vimg_libm::glibc_cosf(x);
      bad_sin += bits(ms) != bits(s);
      bad_cos += bits(mc) != bits(c);
      bad_sc += (bits(s2) != bits(s)) || (bits(c2) != bits(c));
      n++;
    }
  }
  printf("checked %llu floats: sinf mismatches %llu, cosf mismatches %llu, sincosf!=sinf/cosf %llu\n", n,
         bad_sin, bad_cos, bad_sc);
  return (bad_sin || bad_cos || bad_sc) ? 1 : 0;
}
