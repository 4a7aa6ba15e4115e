// Host-side float32 helpers for the PreRender mirror (vh_* layer).
//
// The host layer has to hand the device the SAME trees, triangle order and camera matrix the
// reference's PreRender would (the traversal order, NodesT and tie-breaks depend on them), so the
// scalar helpers keep the reference's float32 evaluation order and x86 min/max operand semantics:
//   math/dim_amd64.s:8-21 (Max/Min return the 2nd operand on NaN/equal), math/boundingbox.go,
//   math/vec3_amd64.s:11-43 (normalize = RSQRTSS + one Newton step), math/matrix4.go, math/quat.go,
//   math/animdecomp.go (polar decomposition used by Camera.PreRender).
// Compiled with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#if defined(__SSE__)
#include <xmmintrin.h>
#endif

namespace vh {

static const float kInf = std::numeric_limits<float>::infinity();
static const float kPi32 = 3.14159265358f;  // math/const.go:8

inline float fmax_x86(float x, float y) { return x > y ? x : y; }
inline float fmin_x86(float x, float y) { return x < y ? x : y; }
inline float fabs32(float x) { uint32_t u; std::memcpy(&u, &x, 4); u &= 0x7fffffffu; std::memcpy(&x, &u, 4); return x; }

struct V3 {
  float x, y, z;
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 scale(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3 lerp(V3 a, V3 b, float t) { return {(1.0f - t) * a.x + t * b.x, (1.0f - t) * a.y + t * b.y, (1.0f - t) * a.z + t * b.z}; }

// math/vec3_amd64.s:11-43
inline V3 normalize(V3 a) {
  float x0 = a.x * a.x, x1 = a.y * a.y, x2 = a.z * a.z;
  x1 = x1 + x0;
  x1 = x1 + x2;
#if defined(__SSE__)
  float g = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x1)));
#else
  float g = 1.0f / sqrtf(x1);
#endif
  float h = 0.5f * x1;
  float gg = g * g;
  gg = gg * h;
  float c = 1.5f - gg;
  float r = c * g;
  return {a.x * r, a.y * r, a.z * r};
}

// Axis-aligned box kept as two float triples; Grow* follow math/boundingbox.go:94-123 argument order.
struct Box {
  float lo[3], hi[3];
  void reset() { for (int i = 0; i < 3; i++) { lo[i] = kInf; hi[i] = -kInf; } }
  void grow_point(float X, float Y, float Z) {
    lo[0] = fmin_x86(X, lo[0]); hi[0] = fmax_x86(X, hi[0]);
    lo[1] = fmin_x86(Y, lo[1]); hi[1] = fmax_x86(Y, hi[1]);
    lo[2] = fmin_x86(Z, lo[2]); hi[2] = fmax_x86(Z, hi[2]);
  }
  void grow_box(const Box& p) {
    for (int k = 0; k < 3; k++) { lo[k] = fmin_x86(lo[k], p.lo[k]); hi[k] = fmax_x86(hi[k], p.hi[k]); }
  }
  float dim(int a) const { return hi[a] - lo[a]; }
  float area() const { return dim(0) * dim(1) * 2.0f + dim(1) * dim(2) * 2 + dim(0) * dim(2) * 2; }
  int max_dim() const {
    if (dim(0) < dim(1)) return dim(1) < dim(2) ? 2 : 1;
    return dim(0) < dim(2) ? 2 : 0;
  }
  V3 centroid() const { return {(hi[0] + lo[0]) * 0.5f, (hi[1] + lo[1]) * 0.5f, (hi[2] + lo[2]) * 0.5f}; }
};
inline Box box_lerp(const Box& a, const Box& b, float t) {
  Box o;
  for (int k = 0; k < 3; k++) { o.lo[k] = (1 - t) * a.lo[k] + t * b.lo[k]; o.hi[k] = (1 - t) * a.hi[k] + t * b.hi[k]; }
  return o;
}

// ---- 4x4 column-major matrices (math/matrix4.go:9-18) --------------------------------------
struct M4 {
  float m[16];
  float at(int i, int j) const { return m[j * 4 + i]; }
  void set(int i, int j, float v) { m[j * 4 + i] = v; }
};
inline M4 m4_identity() { M4 r{}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
inline M4 m4_mul(const M4& a, const M4& b) {
  M4 c{};
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) c.m[j * 4 + i] += a.m[k * 4 + i] * b.m[j * 4 + k];
  return c;
}
inline M4 m4_transpose(const M4& a) { M4 c; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c.m[i * 4 + j] = a.m[j * 4 + i]; return c; }
inline M4 m4_scale(float s, const M4& a) { M4 c; for (int i = 0; i < 16; i++) c.m[i] = s * a.m[i]; return c; }
inline M4 m4_add(const M4& a, const M4& b) { M4 c; for (int i = 0; i < 16; i++) c.m[i] = a.m[i] + b.m[i]; return c; }
inline M4 m4_lerp(const M4& a, const M4& b, float t) { M4 c; for (int i = 0; i < 16; i++) c.m[i] = (1.0f - t) * a.m[i] + t * b.m[i]; return c; }
void m4_cofactors(const float* m, float* inv);  // nodes.cpp
bool m4_inverse(const M4& a, M4* out);
float m4_det(const M4& a);

}  // namespace vh
