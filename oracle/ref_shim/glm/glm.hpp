// TEST INFRASTRUCTURE (oracle). Minimal stand-in for the slice of glm that the
// reference's hot path touches; lets /root/reference/source/world.h compile
// headless.  glm itself is not vendored by the reference and carries no version
// pin (reference Makefile:5 only adds an include path), so this file restates
// glm's *published scalar definitions* for exactly the call sites listed in
// SURVEY.md 8c:
//   length(v)      = sqrt(dot(v,v))
//   inversesqrt(x) = 1 / sqrt(x)
//   normalize(v)   = v * inversesqrt(dot(v,v))
//   dot(a,b)       = a.x*b.x + a.y*b.y (+ a.z*b.z)   (left-to-right sum)
//   cross(a,b)     = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y)
//   converting ctors are static_cast per component (float->int truncates)
// Nothing here is shipped in the product library.
#pragma once
#include <cmath>

namespace glm {

template <typename T>
struct tvec2 {
  T x, y;
  constexpr tvec2() : x(0), y(0) {}
  template <typename A> constexpr tvec2(A s) : x(static_cast<T>(s)), y(static_cast<T>(s)) {}
  template <typename A, typename B> constexpr tvec2(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
  template <typename U> constexpr tvec2(const tvec2<U>& o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}
  T& operator[](int i) { return i == 0 ? x : y; }
  const T& operator[](int i) const { return i == 0 ? x : y; }
  template <typename U> tvec2& operator+=(const tvec2<U>& o) { x += static_cast<T>(o.x); y += static_cast<T>(o.y); return *this; }
  template <typename U> tvec2& operator-=(const tvec2<U>& o) { x -= static_cast<T>(o.x); y -= static_cast<T>(o.y); return *this; }
  template <typename U> tvec2& operator/=(const tvec2<U>& o) { x /= static_cast<T>(o.x); y /= static_cast<T>(o.y); return *this; }
  tvec2& operator*=(T s) { x *= s; y *= s; return *this; }
  tvec2& operator/=(T s) { x /= s; y /= s; return *this; }
};

typedef tvec2<float> vec2;
typedef tvec2<int> ivec2;

// same-type component-wise arithmetic
template <typename T> constexpr tvec2<T> operator+(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x + b.x, a.y + b.y); }
template <typename T> constexpr tvec2<T> operator-(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x - b.x, a.y - b.y); }
template <typename T> constexpr tvec2<T> operator*(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x * b.x, a.y * b.y); }
template <typename T> constexpr tvec2<T> operator/(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x / b.x, a.y / b.y); }
template <typename T> constexpr tvec2<T> operator*(T s, const tvec2<T>& v) { return tvec2<T>(s * v.x, s * v.y); }
template <typename T> constexpr tvec2<T> operator*(const tvec2<T>& v, T s) { return tvec2<T>(v.x * s, v.y * s); }
template <typename T> constexpr tvec2<T> operator/(const tvec2<T>& v, T s) { return tvec2<T>(v.x / s, v.y / s); }
template <typename T> constexpr bool operator==(const tvec2<T>& a, const tvec2<T>& b) { return a.x == b.x && a.y == b.y; }

// the mixed forms the reference relies on (glm resolves these through implicit
// conversion of the integer vector to the float one)
inline vec2 operator+(const ivec2& a, const vec2& b) { return vec2(a) + b; }
inline vec2 operator+(const vec2& a, const ivec2& b) { return a + vec2(b); }
inline vec2 operator-(const vec2& a, const ivec2& b) { return a - vec2(b); }
inline vec2 operator*(int s, const vec2& v) { return static_cast<float>(s) * v; }
inline vec2 operator*(double s, const vec2& v) { return static_cast<float>(s) * v; }
inline vec2 operator*(float s, const ivec2& v) { return s * vec2(v); }
inline vec2 operator/(const vec2& v, int s) { return v / static_cast<float>(s); }
inline vec2 operator/(const vec2& v, double s) { return v / static_cast<float>(s); }

struct vec3 {
  float x, y, z;
  constexpr vec3() : x(0), y(0), z(0) {}
  template <typename A> constexpr vec3(A s) : x(static_cast<float>(s)), y(static_cast<float>(s)), z(static_cast<float>(s)) {}
  template <typename A, typename B, typename C>
  constexpr vec3(A a, B b, C c) : x(static_cast<float>(a)), y(static_cast<float>(b)), z(static_cast<float>(c)) {}
  vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
};
constexpr vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
constexpr vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
constexpr vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
constexpr vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
constexpr vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }

inline float dot(const vec2& a, const vec2& b) { vec2 t = a * b; return t.x + t.y; }
inline float dot(const vec3& a, const vec3& b) { vec3 t = a * b; return t.x + t.y + t.z; }
inline vec3 cross(const vec3& a, const vec3& b) {
  return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline float inversesqrt(float v) { return 1.0f / std::sqrt(v); }
inline float length(const vec2& v) { return std::sqrt(dot(v, v)); }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline vec2 normalize(const vec2& v) { return v * inversesqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }

}  // namespace glm
