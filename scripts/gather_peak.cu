// gather_peak.cu -- what the memory system delivers for the access pattern of the streamed shade / END kernels, with
// no arithmetic in the way: per item one 32-byte queue record read in order, one 128-byte-aligned slot line gathered
// by the slot number in it (112 bytes read), 96 bytes of that line written back, one 48-byte record appended in order.
// = 288 algorithmic bytes per item, the figure bench.py's roofline uses for the shade kernels (304 with the shadow ray).
// Prints GB/s of those bytes for several slot counts (= address ranges) and items per launch, next to a plain copy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_peak gather_peak.cu && ./gather_peak
//   ./gather_peak --one SLOTS ITEMS : that one case (random order), one JSON line -- what bench.py's roofline.pattern_peak runs
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <string>
#include <cstdlib>
#include <cuda_runtime.h>

struct __align__(128) Line { float4 c[8]; };

template <int ITEMS_IN_FLIGHT>
__global__ void __launch_bounds__(256) gather_kernel(const float4* __restrict__ qa, const float4* __restrict__ qb,
                                                     Line* __restrict__ lines, float4* __restrict__ ta,
                                                     float4* __restrict__ tb, float4* __restrict__ tc, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x + threadIdx.x; base < n; base += stride * ITEMS_IN_FLIGHT) {
    float4 a[ITEMS_IN_FLIGHT], b[ITEMS_IN_FLIGHT], l[ITEMS_IN_FLIGHT][7];
#pragma unroll
    for (int k = 0; k < ITEMS_IN_FLIGHT; ++k) {
      const uint32_t it = base + k * stride;
      if (it < n) {
        a[k] = __ldcs(qa + it);
        b[k] = __ldcs(qb + it);
      }
    }
#pragma unroll
    for (int k = 0; k < ITEMS_IN_FLIGHT; ++k) {
      const uint32_t it = base + k * stride;
      if (it < n) {
        const float4* src = lines[__float_as_uint(a[k].x)].c;
#pragma unroll
        for (int j = 0; j < 7; ++j) l[k][j] = __ldcg(src + j);
      }
    }
#pragma unroll
    for (int k = 0; k < ITEMS_IN_FLIGHT; ++k) {
      const uint32_t it = base + k * stride;
      if (it < n) {
        float4* dst = lines[__float_as_uint(a[k].x)].c;
        float4 s = b[k];
#pragma unroll
        for (int j = 0; j < 7; ++j) { s.x += l[k][j].x; s.y += l[k][j].y; s.z += l[k][j].z; s.w += l[k][j].w; }
        __stcg(dst + 0, s); __stcg(dst + 1, l[k][0]);            // sector 0
        __stcg(dst + 4, l[k][1]); __stcg(dst + 5, l[k][2]);      // sector 2
        __stcg(dst + 6, l[k][3]); __stcg(dst + 7, l[k][4]);      // sector 3
        __stcs(ta + it, s); __stcs(tb + it, l[k][5]); __stcs(tc + it, l[k][6]);
      }
    }
  }
}

__global__ void copy_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv) {
  const bool one = argc == 4 && std::string(argv[1]) == "--one";
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const uint32_t max_slots = 8u << 20;
  Line* lines; float4 *qa, *qb, *ta, *tb, *tc, *flush;
  CK(cudaMalloc(&lines, (size_t)max_slots * sizeof(Line)));
  CK(cudaMemset(lines, 0, (size_t)max_slots * sizeof(Line)));
  for (float4** p : {&qa, &qb, &ta, &tb, &tc}) CK(cudaMalloc(p, (size_t)max_slots * 16));
  const size_t flush_n = (512u << 20) / 16;
  CK(cudaMalloc(&flush, flush_n * 16 * 2));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double copy_gbs = 0;
  {  // plain copy, 512 MB in + 512 MB out
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
      cudaEventRecord(e0); copy_kernel<<<sms * 8, 256>>>(flush, flush + flush_n, flush_n); cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
    }
    copy_gbs = 2.0 * flush_n * 16 / best / 1e6;
    if (!one) printf("copy 512 MB -> 512 MB: %.1f GB/s\n", copy_gbs);
  }
  if (one) {
    const uint32_t slots = (uint32_t)std::min<unsigned long long>(std::strtoull(argv[2], nullptr, 10), max_slots);
    const uint32_t n = (uint32_t)std::min<unsigned long long>(std::strtoull(argv[3], nullptr, 10), slots);
    std::mt19937 rng1(1);
    std::vector<uint32_t> perm(slots);
    for (uint32_t i = 0; i < slots; ++i) perm[i] = i;
    std::shuffle(perm.begin(), perm.end(), rng1);
    std::vector<float4> h(n);
    for (uint32_t i = 0; i < n; ++i) { h[i] = make_float4(0, 0, 0, 0); ((uint32_t*)&h[i])[0] = perm[i]; }
    CK(cudaMemcpy(qa, h.data(), (size_t)n * 16, cudaMemcpyHostToDevice));
    float best = 1e9f;
    for (int r = 0; r < 8; ++r) {
      copy_kernel<<<sms * 8, 256>>>(flush, flush + flush_n, flush_n);  // evict the lines from L2
      cudaEventRecord(e0);
      gather_kernel<2><<<sms * 8, 256>>>(qa, qb, lines, ta, tb, tc, n);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r > 0) best = std::min(best, ms);
    }
    printf("{\"slots\": %u, \"items\": %u, \"bytes_per_item\": 288, \"us\": %.2f, \"pattern_gbs\": %.1f, \"copy_gbs\": %.1f}\n", slots, n,
           best * 1e3, 288.0 * n / best / 1e6, copy_gbs);
    return 0;
  }
  std::mt19937 rng(1);
  printf("%10s %10s %8s %14s %4s %10s %12s\n", "slots", "items", "range", "order", "mlp", "us", "GB/s(288B)");
  // order of the slot numbers in the queue: random (what appends in arrival order converge to), random inside buckets
  // of 16 K / 1 K / 64 slots that follow each other in order, sorted
  const uint32_t buckets[] = {0u, 16384u, 1024u, 64u, 1u};
  const char* names[] = {"random", "bucket 16K", "bucket 1K", "bucket 64", "sorted"};
  for (uint32_t slots : {1u << 20, 1u << 21, 1u << 22}) {
    for (double frac : {0.35, 1.0}) {
      const uint32_t n = (uint32_t)(slots * frac);
      for (int o = 0; o < 5; ++o) {
        // the n active slots: a random subset; then ordered as asked
        std::vector<uint32_t> perm(slots);
        for (uint32_t i = 0; i < slots; ++i) perm[i] = i;
        std::shuffle(perm.begin(), perm.end(), rng);
        perm.resize(n);
        if (buckets[o]) {
          const uint32_t bsz = buckets[o];
          std::stable_sort(perm.begin(), perm.end(), [bsz](uint32_t x, uint32_t y) { return x / bsz < y / bsz; });
        }
        std::vector<float4> h(n);
        for (uint32_t i = 0; i < n; ++i) { h[i] = make_float4(0, 0, 0, 0); ((uint32_t*)&h[i])[0] = perm[i]; }
        CK(cudaMemcpy(qa, h.data(), (size_t)n * 16, cudaMemcpyHostToDevice));
        for (int mlp : {1, 2}) {
          float best = 1e9f;
          for (int r = 0; r < 6; ++r) {
            copy_kernel<<<sms * 8, 256>>>(flush, flush + flush_n, flush_n);  // evict the lines from L2
            cudaEventRecord(e0);
            if (mlp == 1) gather_kernel<1><<<sms * 8, 256>>>(qa, qb, lines, ta, tb, tc, n);
            else gather_kernel<2><<<sms * 8, 256>>>(qa, qb, lines, ta, tb, tc, n);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0) best = std::min(best, ms);
          }
          printf("%10u %10u %5u MB %14s %4d %10.1f %12.1f\n", slots, n, (unsigned)((size_t)slots * 128 >> 20), names[o], mlp, best * 1e3, 288.0 * n / best / 1e6);
        }
      }
    }
  }
  return 0;
}
