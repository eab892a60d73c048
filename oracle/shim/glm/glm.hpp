// Minimal glm-compatible shim (TEST INFRASTRUCTURE, part of oracle/).
//
// The reference pins glm 1.0.1 through CPM (CMakeLists.txt:67-71); glm is not
// vendored in /root/reference and there is no network here, so the oracle
// build supplies this header set instead.  It restates glm 1.0.1's *scalar*
// code paths (no SIMD, default precision) for exactly the API surface the
// reference's hot path and JSON loader touch.  Operation ORDER is part of the
// parity contract, so every function below states the expression it evaluates:
//   dot(vec2)  = x*x' + y*y'
//   dot(vec3)  = (x*x' + y*y') + z*z'
//   dot(vec4)  = (x*x' + y*y') + (z*z' + w*w')
//   normalize  = v * (1 / sqrt(dot(v, v)))
//   length     = sqrt(dot(v, v)),  length2 = dot(v, v)
//   cross      = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y)
//   min(a, b)  = (b < a) ? b : a,  max(a, b) = (a < b) ? b : a   (per component)
//   clamp      = min(max(x, lo), hi)
//   mix(x,y,a) = x * (1 - a) + y * a
//   mat4 * vec4 = (m0*v0 + m1*v1) + (m2*v2 + m3*v3)
//   v / s      = per-component division (no reciprocal)
// Fidelity to the real glm cannot be verified offline; bit-exact parity is
// therefore pinned to "reference sources + this shim".
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace glm {

template <int L, typename T> struct vec;

// ------------------------------------------------------------------ vec2
template <typename T> struct vec<2, T> {
  union { T x, r, s; };
  union { T y, g, t; };
  constexpr vec() : x(0), y(0) {}
  constexpr vec(const vec&) = default;
  constexpr vec& operator=(const vec&) = default;
  template <typename U> constexpr explicit vec(U v) : x(static_cast<T>(v)), y(static_cast<T>(v)) {}
  template <typename A, typename B> constexpr vec(A a, B b)
      : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
  template <typename U> constexpr vec(const vec<2, U>& v)
      : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
  template <typename U> constexpr vec(const vec<3, U>& v);
  template <typename U> constexpr vec(const vec<4, U>& v);
  constexpr T& operator[](int i) { return i == 0 ? x : y; }
  constexpr const T& operator[](int i) const { return i == 0 ? x : y; }
};

// ------------------------------------------------------------------ vec3
template <typename T> struct vec<3, T> {
  union { T x, r, s; };
  union { T y, g, t; };
  union { T z, b, p; };
  constexpr vec() : x(0), y(0), z(0) {}
  constexpr vec(const vec&) = default;
  constexpr vec& operator=(const vec&) = default;
  template <typename U> constexpr explicit vec(U v)
      : x(static_cast<T>(v)), y(static_cast<T>(v)), z(static_cast<T>(v)) {}
  template <typename A, typename B, typename C> constexpr vec(A a, B b, C c)
      : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
  template <typename U, typename C> constexpr vec(const vec<2, U>& v, C c)
      : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(c)) {}
  template <typename U> constexpr vec(const vec<3, U>& v)
      : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
  template <typename U> constexpr vec(const vec<4, U>& v);
  constexpr T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  constexpr const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

// ------------------------------------------------------------------ vec4
template <typename T> struct vec<4, T> {
  union { T x, r, s; };
  union { T y, g, t; };
  union { T z, b, p; };
  union { T w, a, q; };
  constexpr vec() : x(0), y(0), z(0), w(0) {}
  constexpr vec(const vec&) = default;
  constexpr vec& operator=(const vec&) = default;
  template <typename U> constexpr explicit vec(U v)
      : x(static_cast<T>(v)), y(static_cast<T>(v)), z(static_cast<T>(v)), w(static_cast<T>(v)) {}
  template <typename A, typename B, typename C, typename D> constexpr vec(A a, B b, C c, D d)
      : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)), w(static_cast<T>(d)) {}
  template <typename U, typename D> constexpr vec(const vec<3, U>& v, D d)
      : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)),
        w(static_cast<T>(d)) {}
  template <typename U> constexpr vec(const vec<4, U>& v)
      : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)),
        w(static_cast<T>(v.w)) {}
  constexpr T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
  constexpr const T& operator[](int i) const {
    return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w));
  }
};

template <typename T> template <typename U> constexpr vec<2, T>::vec(const vec<3, U>& v)
    : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
template <typename T> template <typename U> constexpr vec<2, T>::vec(const vec<4, U>& v)
    : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
template <typename T> template <typename U> constexpr vec<3, T>::vec(const vec<4, U>& v)
    : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}

typedef vec<2, float> vec2;
typedef vec<3, float> vec3;
typedef vec<4, float> vec4;
typedef vec<2, int> ivec2;
typedef vec<3, int> ivec3;
typedef vec<2, uint16_t> u16vec2;

// ------------------------------------------------ component-wise arithmetic
#define GLM_SHIM_BINOP(OP)                                                                   \
  template <typename T> constexpr vec<2, T> operator OP(const vec<2, T>& a, const vec<2, T>& b) { \
    return vec<2, T>(a.x OP b.x, a.y OP b.y);                                                \
  }                                                                                          \
  template <typename T> constexpr vec<3, T> operator OP(const vec<3, T>& a, const vec<3, T>& b) { \
    return vec<3, T>(a.x OP b.x, a.y OP b.y, a.z OP b.z);                                    \
  }                                                                                          \
  template <typename T> constexpr vec<4, T> operator OP(const vec<4, T>& a, const vec<4, T>& b) { \
    return vec<4, T>(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w);                        \
  }                                                                                          \
  constexpr vec2 operator OP(const vec2& a, float s) { return vec2(a.x OP s, a.y OP s); }    \
  constexpr vec3 operator OP(const vec3& a, float s) {                                       \
    return vec3(a.x OP s, a.y OP s, a.z OP s);                                               \
  }                                                                                          \
  constexpr vec4 operator OP(const vec4& a, float s) {                                       \
    return vec4(a.x OP s, a.y OP s, a.z OP s, a.w OP s);                                     \
  }                                                                                          \
  constexpr vec2 operator OP(float s, const vec2& a) { return vec2(s OP a.x, s OP a.y); }    \
  constexpr vec3 operator OP(float s, const vec3& a) {                                       \
    return vec3(s OP a.x, s OP a.y, s OP a.z);                                               \
  }                                                                                          \
  constexpr vec4 operator OP(float s, const vec4& a) {                                       \
    return vec4(s OP a.x, s OP a.y, s OP a.z, s OP a.w);                                     \
  }
GLM_SHIM_BINOP(+)
GLM_SHIM_BINOP(-)
GLM_SHIM_BINOP(*)
GLM_SHIM_BINOP(/)
#undef GLM_SHIM_BINOP

// compound assignment: glm casts a scalar of any type U to T first
// (type_vec3.inl: `this->x /= static_cast<T>(scalar)`).
#define GLM_SHIM_ASSIGNOP(OP)                                                                \
  template <int L, typename T> constexpr vec<L, T>& operator OP##=(vec<L, T>& a,             \
                                                                   const vec<L, T>& b) {     \
    for (int i = 0; i < L; ++i) a[i] OP## = b[i];                                            \
    return a;                                                                                \
  }                                                                                          \
  template <int L, typename T, typename U>                                                   \
  constexpr vec<L, T>& operator OP##=(vec<L, T>& a, U s) {                                   \
    const T st = static_cast<T>(s);                                                          \
    for (int i = 0; i < L; ++i) a[i] OP## = st;                                              \
    return a;                                                                                \
  }
GLM_SHIM_ASSIGNOP(+)
GLM_SHIM_ASSIGNOP(-)
GLM_SHIM_ASSIGNOP(*)
GLM_SHIM_ASSIGNOP(/)
#undef GLM_SHIM_ASSIGNOP

template <typename T> constexpr vec<2, T> operator-(const vec<2, T>& a) {
  return vec<2, T>(-a.x, -a.y);
}
template <typename T> constexpr vec<3, T> operator-(const vec<3, T>& a) {
  return vec<3, T>(-a.x, -a.y, -a.z);
}
template <typename T> constexpr vec<4, T> operator-(const vec<4, T>& a) {
  return vec<4, T>(-a.x, -a.y, -a.z, -a.w);
}
template <typename T> constexpr vec<2, T> operator+(const vec<2, T>& a) { return a; }
template <typename T> constexpr vec<3, T> operator+(const vec<3, T>& a) { return a; }

template <int L, typename T> constexpr bool operator==(const vec<L, T>& a, const vec<L, T>& b) {
  for (int i = 0; i < L; ++i)
    if (!(a[i] == b[i])) return false;
  return true;
}
template <int L, typename T> constexpr bool operator!=(const vec<L, T>& a, const vec<L, T>& b) {
  return !(a == b);
}

// ------------------------------------------------------- geometric functions
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) {
  const float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
  return tx + ty + tz;
}
inline float dot(const vec4& a, const vec4& b) {
  const float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z, tw = a.w * b.w;
  return (tx + ty) + (tz + tw);
}
inline float length(const vec2& v) { return std::sqrt(dot(v, v)); }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline float length2(const vec2& v) { return dot(v, v); }
inline float length2(const vec3& v) { return dot(v, v); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline vec2 normalize(const vec2& v) { return v * inversesqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
inline vec4 normalize(const vec4& v) { return v * inversesqrt(dot(v, v)); }
inline vec3 cross(const vec3& a, const vec3& b) {
  return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}

// ---------------------------------------------------------- common functions
inline float min(float a, float b) { return (b < a) ? b : a; }
inline float max(float a, float b) { return (a < b) ? b : a; }
inline vec2 min(const vec2& a, const vec2& b) { return vec2(min(a.x, b.x), min(a.y, b.y)); }
inline vec2 max(const vec2& a, const vec2& b) { return vec2(max(a.x, b.x), max(a.y, b.y)); }
inline vec3 min(const vec3& a, const vec3& b) {
  return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z));
}
inline vec3 max(const vec3& a, const vec3& b) {
  return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z));
}
inline vec3 min(const vec3& a, float b) { return vec3(min(a.x, b), min(a.y, b), min(a.z, b)); }
inline vec3 max(const vec3& a, float b) { return vec3(max(a.x, b), max(a.y, b), max(a.z, b)); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline vec3 clamp(const vec3& v, float lo, float hi) {
  return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi));
}
inline vec3 clamp(const vec3& v, const vec3& lo, const vec3& hi) {
  return vec3(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y), clamp(v.z, lo.z, hi.z));
}
inline vec2 abs(const vec2& v) { return vec2(std::abs(v.x), std::abs(v.y)); }
inline vec3 abs(const vec3& v) { return vec3(std::abs(v.x), std::abs(v.y), std::abs(v.z)); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline vec2 mix(const vec2& x, const vec2& y, const vec2& a) {
  return x * (vec2(1.0f) - a) + y * a;
}
inline vec3 mix(const vec3& x, const vec3& y, const vec3& a) {
  return x * (vec3(1.0f) - a) + y * a;
}
inline vec3 mix(const vec3& x, const vec3& y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 pow(const vec3& b, const vec3& e) {
  return vec3(std::pow(b.x, e.x), std::pow(b.y, e.y), std::pow(b.z, e.z));
}
inline vec3 log2(const vec3& v) { return vec3(std::log2(v.x), std::log2(v.y), std::log2(v.z)); }
inline vec3 exp2(const vec3& v) { return vec3(std::exp2(v.x), std::exp2(v.y), std::exp2(v.z)); }
inline vec3 sqrt(const vec3& v) { return vec3(std::sqrt(v.x), std::sqrt(v.y), std::sqrt(v.z)); }

// ------------------------------------------------------------------ matrices
struct mat3 {
  vec3 c[3];
  constexpr mat3() : mat3(1.0f) {}
  constexpr explicit mat3(float d) : c{vec3(d, 0, 0), vec3(0, d, 0), vec3(0, 0, d)} {}
  constexpr mat3(const vec3& a, const vec3& b, const vec3& cc) : c{a, b, cc} {}
  // nine scalars of any arithmetic type (agx.cpp passes double literals), column by column
  template <typename A0, typename A1, typename A2, typename B0, typename B1, typename B2, typename C0,
            typename C1, typename C2>
  constexpr mat3(A0 a0, A1 a1, A2 a2, B0 b0, B1 b1, B2 b2, C0 c0, C1 c1, C2 c2)
      : c{vec3(a0, a1, a2), vec3(b0, b1, b2), vec3(c0, c1, c2)} {}
  constexpr vec3& operator[](int i) { return c[i]; }
  constexpr const vec3& operator[](int i) const { return c[i]; }
};
// type_mat3x3.inl: m * v = m0*v0 + m1*v1 + m2*v2 written per component
inline vec3 operator*(const mat3& m, const vec3& v) {
  return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
              m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
              m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}

struct mat4 {
  vec4 c[4];
  mat4() : mat4(1.0f) {}
  explicit mat4(float d) {
    c[0] = vec4(d, 0, 0, 0);
    c[1] = vec4(0, d, 0, 0);
    c[2] = vec4(0, 0, d, 0);
    c[3] = vec4(0, 0, 0, d);
  }
  mat4(const vec4& a, const vec4& b, const vec4& cc, const vec4& d) {
    c[0] = a;
    c[1] = b;
    c[2] = cc;
    c[3] = d;
  }
  explicit mat4(const mat3& m) {
    c[0] = vec4(m[0], 0.0f);
    c[1] = vec4(m[1], 0.0f);
    c[2] = vec4(m[2], 0.0f);
    c[3] = vec4(0, 0, 0, 1);
  }
  vec4& operator[](int i) { return c[i]; }
  const vec4& operator[](int i) const { return c[i]; }
};
typedef mat4 mat4x4;
typedef mat3 mat3x3;

// type_mat4x4.inl operator*(mat4, vec4): Add0 = m0*v0 + m1*v1, Add1 = m2*v2 + m3*v3
inline vec4 operator*(const mat4& m, const vec4& v) {
  const vec4 Mul0 = m[0] * vec4(v[0]);
  const vec4 Mul1 = m[1] * vec4(v[1]);
  const vec4 Add0 = Mul0 + Mul1;
  const vec4 Mul2 = m[2] * vec4(v[2]);
  const vec4 Mul3 = m[3] * vec4(v[3]);
  const vec4 Add1 = Mul2 + Mul3;
  return Add0 + Add1;
}
// host-side only (scene loaders): column i = a0*b[i][0] + a1*b[i][1] + a2*b[i][2] + a3*b[i][3]
inline mat4 operator*(const mat4& a, const mat4& b) {
  mat4 r;
  for (int i = 0; i < 4; ++i)
    r[i] = a[0] * b[i][0] + a[1] * b[i][1] + a[2] * b[i][2] + a[3] * b[i][3];
  return r;
}

// ------------------------------------------------- gtx/transform, quaternion
inline mat4 scale(const vec3& v) {
  const mat4 m(1.0f);
  return mat4(m[0] * v[0], m[1] * v[1], m[2] * v[2], m[3]);
}
inline mat4 translate(const vec3& v) {
  mat4 m(1.0f);
  m[3] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3];
  return m;
}
struct quat {
  float x, y, z, w;  // reference's from_json writes q.x..q.w (json_scene.cpp:30-35)
  quat() : x(0), y(0), z(0), w(1) {}
};
// gtc/quaternion.inl mat3_cast
inline mat4 toMat4(const quat& q) {
  mat3 R(1.0f);
  const float qxx(q.x * q.x), qyy(q.y * q.y), qzz(q.z * q.z);
  const float qxz(q.x * q.z), qxy(q.x * q.y), qyz(q.y * q.z);
  const float qwx(q.w * q.x), qwy(q.w * q.y), qwz(q.w * q.z);
  R[0][0] = 1.0f - 2.0f * (qyy + qzz);
  R[0][1] = 2.0f * (qxy + qwz);
  R[0][2] = 2.0f * (qxz - qwy);
  R[1][0] = 2.0f * (qxy - qwz);
  R[1][1] = 1.0f - 2.0f * (qxx + qzz);
  R[1][2] = 2.0f * (qyz + qwx);
  R[2][0] = 2.0f * (qxz + qwy);
  R[2][1] = 2.0f * (qyz - qwx);
  R[2][2] = 1.0f - 2.0f * (qxx + qyy);
  return mat4(R);
}

}  // namespace glm
