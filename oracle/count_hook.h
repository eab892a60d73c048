// Force-included by the `count` flavour of the oracle build (TEST INFRASTRUCTURE).
// The Makefile derives an instrumented bvh.h at build time (sed, output under
// oracle/_ref/, never committed) that bumps these counters inside BVH::hit<T>
// (include/bvh.h:83-225): [0] closest-hit calls, [1] any-hit calls, [2] inner
// nodes popped, [3] leaves popped, [4] primitive tests.
#pragma once
#ifdef __cplusplus
#include <atomic>
#include <cstdint>
extern std::atomic<uint64_t> vref_counters_g[6];
#define VREF_BUMP(i) vref_counters_g[(i)].fetch_add(1, std::memory_order_relaxed)
#endif
