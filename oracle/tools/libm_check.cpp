// libm_check.cpp -- ORACLE / TEST INFRASTRUCTURE.
// Exhaustively compares v-img_b200/csrc/vimg_libm.cuh (compiled for the host) with the
// host's glibc sinf / cosf / sincosf over every float with |x| < 120 (the range of glibc's one-step reduction), both signs;
// log2f and logf over EVERY float bit pattern (NaN results compared as "both NaN"); powf over every float of
// [0, 2] for the exponents the reference uses (1/2.4f, 2.4f, 2.2f: color_utils.h, agx.cpp) and over
// 2^28 pseudo-random (x, y) pairs with x in (0, 2^8), y in (-8, 8).
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

  auto same = [](float a, float b) { return bits(a) == bits(b) || (a != a && b != b); };
  unsigned long long bad_log2 = 0, bad_log = 0, n_log2 = 0;
#pragma omp parallel for reduction(+ : bad_log2, bad_log, n_log2) schedule(static)
  for (long long i = 0; i <= 0xffffffffll; i += step) {
    const float x = from_bits((uint32_t)i);
    bad_log2 += !same(vimg_libm::glibc_log2f(x), log2f(x));
    bad_log += !same(vimg_libm::glibc_logf(x), logf(x));
    n_log2++;
  }
  printf("checked %llu floats: log2f mismatches %llu, logf mismatches %llu\n", n_log2, bad_log2, bad_log);

  const float exps[3] = {1.0f / 2.4f, 2.4f, 2.2f};
  unsigned long long bad_pow = 0, n_pow = 0;
  const uint32_t hi_pow = bits(2.0f);
#pragma omp parallel for reduction(+ : bad_pow, n_pow) schedule(static)
  for (long long i = 0; i <= (long long)hi_pow; i += step) {
    const float x = from_bits((uint32_t)i);
    for (int e = 0; e < 3; ++e) {
      bad_pow += !same(vimg_libm::glibc_powf(x, exps[e]), powf(x, exps[e]));
      n_pow++;
    }
  }
  unsigned long long bad_pow_r = 0, n_pow_r = 0;
  const long long n_rand = step == 1 ? (1ll << 28) : (1ll << 18);
#pragma omp parallel for reduction(+ : bad_pow_r, n_pow_r) schedule(static)
  for (long long i = 0; i < n_rand; ++i) {
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull + 0x1234567ull;  // splitmix64
    h = (h ^ (h >> 30)) * 0xBF58476D1CE4E5B9ull;
    h = (h ^ (h >> 27)) * 0x94D049BB133111EBull;
    h ^= h >> 31;
    const float x = from_bits((uint32_t)(h % bits(256.0f)) + 1u);
    const float y = ((float)((h >> 32) & 0xffffffu) / 16777216.0f - 0.5f) * 16.0f;
    bad_pow_r += !same(vimg_libm::glibc_powf(x, y), powf(x, y));
    n_pow_r++;
  }
  printf("checked %llu (x, exponent) pairs: powf mismatches %llu; %llu random pairs: mismatches %llu\n", n_pow,
         bad_pow, n_pow_r, bad_pow_r);

  // acosf, atanf: every float bit pattern; atan2f: every y with x in a set of awkward values, and
  // 2^30 pseudo-random (y, x) bit patterns
  unsigned long long bad_acos = 0, bad_atan = 0, bad_atan2 = 0, n1 = 0, n2 = 0;
  const float xs[10] = {1.0f, -1.0f, 0.5f, -3.0f, 1e-30f, -1e30f, 0.0f, -0.0f, INFINITY, 0.7071068f};
#pragma omp parallel for reduction(+ : bad_acos, bad_atan, bad_atan2, n1, n2) schedule(static)
  for (long long i = 0; i <= 0xffffffffll; i += step) {
    const float x = from_bits((uint32_t)i);
    bad_acos += !same(vimg_libm::glibc_acosf(x), acosf(x));
    bad_atan += !same(vimg_libm::glibc_atanf(x), atanf(x));
    n1++;
    for (int j = 0; j < 10; ++j) {
      bad_atan2 += !same(vimg_libm::glibc_atan2f(x, xs[j]), atan2f(x, xs[j]));
      n2++;
    }
  }
  const long long n_rand2 = step == 1 ? (1ll << 30) : (1ll << 20);
#pragma omp parallel for reduction(+ : bad_atan2, n2) schedule(static)
  for (long long i = 0; i < n_rand2; ++i) {
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull + 0x7654321ull;
    h = (h ^ (h >> 30)) * 0xBF58476D1CE4E5B9ull;
    h = (h ^ (h >> 27)) * 0x94D049BB133111EBull;
    h ^= h >> 31;
    float y = from_bits((uint32_t)h), x = from_bits((uint32_t)(h >> 32));
    if (i & 1) {  // the renderer's case: components of a unit vector
      y = (float)((uint32_t)h) / 4294967296.0f * 2.0f - 1.0f;
      x = (float)((uint32_t)(h >> 32)) / 4294967296.0f * 2.0f - 1.0f;
    }
    bad_atan2 += !same(vimg_libm::glibc_atan2f(y, x), atan2f(y, x));
    n2++;
  }
  printf("checked %llu floats: acosf mismatches %llu, atanf mismatches %llu; %llu (y, x) pairs: atan2f mismatches %llu\n",
         n1, bad_acos, bad_atan, n2, bad_atan2);

  // double log2: 2^30 pseudo-random doubles (uniform bit patterns of positive doubles incl. subnormals, and
  // the neighbourhood of 1 that takes the separate polynomial), every float of [2^-20, 2^20] widened
  auto dbits = [](double v) { uint64_t u; memcpy(&u, &v, 8); return u; };
  auto dfrom = [](uint64_t u) { double v; memcpy(&v, &u, 8); return v; };
  unsigned long long bad_log2d = 0, n_log2d = 0;
  const long long n_rand3 = step == 1 ? (1ll << 30) : (1ll << 20);
#pragma omp parallel for reduction(+ : bad_log2d, n_log2d) schedule(static)
  for (long long i = 0; i < n_rand3; ++i) {
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull + 0xabcdefull;
    h = (h ^ (h >> 30)) * 0xBF58476D1CE4E5B9ull;
    h = (h ^ (h >> 27)) * 0x94D049BB133111EBull;
    h ^= h >> 31;
    double x;
    switch (i & 3) {
      case 0: x = dfrom(h % 0x7ff0000000000001ull); break;                             // any non-negative double
      case 1: x = dfrom(0x3fe0000000000000ull + (h % 0x0020000000000000ull)); break;  // [0.5, 2)
      case 2: x = 1.0 + ((double)(int64_t)(h >> 11) / 9007199254740992.0 - 0.5) * 0.125; break;  // around 1
      default: x = (double)from_bits((uint32_t)(h % 0x7f800000u)); break;            // floats widened
    }
    const double a = vimg_libm::glibc_log2(x), b = log2(x);
    bad_log2d += !(dbits(a) == dbits(b) || (a != a && b != b));
    n_log2d++;
  }
  const uint32_t f_lo = bits(0x1p-20f), f_hi = bits(0x1p20f);
#pragma omp parallel for reduction(+ : bad_log2d, n_log2d) schedule(static)
  for (long long i = f_lo; i <= (long long)f_hi; i += step) {
    const double x = (double)from_bits((uint32_t)i);
    bad_log2d += dbits(vimg_libm::glibc_log2(x)) != dbits(log2(x));
    n_log2d++;
  }
  {
    const double sp[] = {0.0, -0.0, 1.0, -1.0, INFINITY, -INFINITY, NAN, 4096.0, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308};
    for (double x : sp) {
      const double a = vimg_libm::glibc_log2(x), b = log2(x);
      bad_log2d += !(dbits(a) == dbits(b) || (a != a && b != b));
      n_log2d++;
    }
  }
  printf("checked %llu doubles: log2 mismatches %llu\n", n_log2d, bad_log2d);
  return (bad_sin || bad_cos || bad_sc || bad_log2 || bad_log || bad_pow || bad_pow_r || bad_acos || bad_atan || bad_atan2 || bad_log2d) ? 1 : 0;
}
