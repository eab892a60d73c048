// Shim (TEST INFRASTRUCTURE): the only fastgltf symbol the reference's hot-path
// headers need is the Wrap enum (include/texture/texture_common.h:4,9-20).
// Values are the glTF sampler constants fastgltf v0.9.0 uses.
#pragma once
#include <cstdint>
namespace fastgltf {
enum class Wrap : std::uint16_t { ClampToEdge = 33071, MirroredRepeat = 33648, Repeat = 10497 };
}
