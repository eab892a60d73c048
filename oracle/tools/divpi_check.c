/* ORACLE TOOL (test infrastructure).  Checks, for EVERY float x in [0, 2], that the three-operation
 * constant division used on the device (vimg_materials.cuh, div_pi):
 *     q0 = x * RN(1/pi);  r = fma(-q0, pi, x);  q = fma(r, RN(1/pi), q0)
 * equals the IEEE double quotient (double)x / pi that the reference computes
 * (src/material/lambertian.cpp: `/ std::numbers::pi` on a float promoted to double).
 * build: gcc -O2 -fopenmp -ffp-contract=off oracle/tools/divpi_check.c -o oracle/_ref/divpi_check -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
int main(void) {
  const double pi = 3.141592653589793238462643383279502884, inv_pi = 0.318309886183790671537767526745028724;
  uint32_t hi_bits;
  float two = 2.0f;
  memcpy(&hi_bits, &two, 4);
  unsigned long long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
  for (uint64_t b = 0; b <= hi_bits; ++b) {
    uint32_t u = (uint32_t)b;
    float xf;
    memcpy(&xf, &u, 4);
    const double x = (double)xf;
    const double q0 = x * inv_pi;
    const double r = fma(-q0, pi, x);
    const double q = fma(r, inv_pi, q0);
    const double ref = x / pi;
    if (memcmp(&q, &ref, 8) != 0) ++bad;
  }
  printf("div_pi: %u floats in [0, 2], %llu mismatches\n", hi_bits + 1u, bad);
  return bad != 0;
}
