// Force-included by the oracle build (TEST INFRASTRUCTURE).
// libstdc++ 13 has no std::sqrtf; the reference calls it
// (src/geometry/triangle.cpp:229, sphere.cpp:87,112, dielectric.cpp:20,63).
#pragma once
#ifdef __cplusplus
#include <cmath>
namespace std {
inline float sqrtf(float x) { return ::sqrtf(x); }
}
#endif
