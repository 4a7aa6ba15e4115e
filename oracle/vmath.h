// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// CPU restatement of the reference's float32 math helpers (jamiec7919/vermeer, Go/amd64).
// parity unpinned: the reference's own tests hold no golden vectors for this path
// (SURVEY.md §4); this restatement is anchored structurally (see tests/).
//
// Follows:
//   math/vec3.go:15-165            (Vec3 ops, evaluated left-to-right in float32)
//   math/vec3_amd64.s:11-56        (Vec3Normalize = RSQRTSS + one Newton step; Vec3Length = SQRTSS)
//   math/dim_amd64.s:8-21          (Max/Min with x86 MAXSS/MINSS operand semantics)
//   math/ferror.go:16-67           (EpsilonFloat32, MachineEpsilon32, Gamma, NextFloatUp/Down)
//   math/unsafe.go:15-41           (SignMask, Xorf)
//   math/boundingbox.go:9-127      (BoundingBox)
//   math/matrix4.go, quat.go, animdecomp.go (camera matrix via polar decomposition)
//   math/sincos.go:16-70           (float32 trig = float64 libm then round)
//
// Build with: -O2 -msse4.1 -ffp-contract=off (Go/amd64 never fuses mul+add; every float32
// operation rounds to float32).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <xmmintrin.h>

namespace orc {

static const float kInfPos = std::numeric_limits<float>::infinity();
static const float kInfNeg = -std::numeric_limits<float>::infinity();

// math/const.go:8  Pi float32 = 3.14159265358
static const float kPi = 3.14159265358f;
// math/ferror.go:16-19
static const float kEpsilonFloat32 = 1.19209290E-07f;
static const float kMachineEpsilon32 = (float)(1.19209290E-07 * 0.5);

// math/dim_amd64.s:8-13: MAXSS X1,X0 -> X0 = (X0 > X1) ? X0 : X1 (second operand on NaN / equal)
static inline float Max(float x, float y) { return x > y ? x : y; }
// math/dim_amd64.s:16-21
static inline float Min(float x, float y) { return x < y ? x : y; }

static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// math/abs.go:14-16
static inline float Abs(float x) { return u2f(f2u(x) & 0x7fffffffu); }
// math/unsafe.go:15-21, 34-37
static inline uint32_t SignMask(float v) { return f2u(v) & 0x80000000u; }
static inline float Xorf(float v, uint32_t i) { return u2f(f2u(v) ^ i); }

static inline float Sqrt(float x) { return sqrtf(x); }    // SQRTSS, correctly rounded
static inline float Floor(float x) { return floorf(x); }  // ROUNDSS
static inline float Ceil(float x) { return ceilf(x); }

// math/sincos.go: float32 -> float64 stdlib -> float32
static inline float Sin(float x) { return (float)std::sin((double)x); }
static inline float Cos(float x) { return (float)std::cos((double)x); }
static inline float Tan(float x) { return (float)std::tan((double)x); }
static inline float Acos(float x) { return (float)std::acos((double)x); }
static inline float Atan(float x) { return (float)std::atan((double)x); }
static inline float Atan2(float y, float x) { return (float)std::atan2((double)y, (double)x); }
// math/pow.go:11-19
static inline float Log2(float x) { return (float)std::log2((double)x); }
static inline float Exp(float x) { return (float)std::exp((double)x); }

// math/inf.go:24-26
static inline bool IsInf(float v) { return v > std::numeric_limits<float>::max() || v < -std::numeric_limits<float>::max(); }

// math/ferror.go:22-24
static inline float Gamma(int32_t n) {
  return ((float)n * kMachineEpsilon32) / (1 - (float)n * kMachineEpsilon32);
}

// math/ferror.go:27-45
static inline float NextFloatUp(float v) {
  if (IsInf(v) && v > 0) return v;
  if (v == -0.0f) v = 0.0f;
  uint32_t ui = f2u(v);
  if (v >= 0.0f) ui++; else ui--;
  return u2f(ui);
}
// math/ferror.go:48-66
static inline float NextFloatDown(float v) {
  if (IsInf(v) && v < 0) return v;
  if (v == -0.0f) v = 0.0f;
  uint32_t ui = f2u(v);
  if (v >= 0.0f) ui--; else ui++;
  return u2f(ui);
}

struct Vec3 {
  float v[3];
  float& operator[](int i) { return v[i]; }
  const float& operator[](int i) const { return v[i]; }
};
static inline Vec3 V3(float x, float y, float z) { Vec3 r; r.v[0] = x; r.v[1] = y; r.v[2] = z; return r; }

static inline Vec3 Vec3Mad(Vec3 a, Vec3 b, float s) { return V3(a[0] + (b[0] * s), a[1] + (b[1] * s), a[2] + (b[2] * s)); }
static inline Vec3 Vec3Add(Vec3 a, Vec3 b) { return V3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
static inline Vec3 Vec3Add3(Vec3 a, Vec3 b, Vec3 c) { return V3(a[0] + b[0] + c[0], a[1] + b[1] + c[1], a[2] + b[2] + c[2]); }
static inline Vec3 Vec3Sub(Vec3 a, Vec3 b) { return V3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
static inline Vec3 Vec3Lerp(Vec3 a, Vec3 b, float t) {
  return V3((1.0f - t) * a[0] + t * b[0], (1.0f - t) * a[1] + t * b[1], (1.0f - t) * a[2] + t * b[2]);
}
static inline Vec3 Vec3Scale(float s, Vec3 a) { return V3(a[0] * s, a[1] * s, a[2] * s); }
static inline Vec3 Vec3Cross(Vec3 a, Vec3 b) {
  return V3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
static inline float Vec3Dot(Vec3 a, Vec3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline float Vec3DotAbs(Vec3 a, Vec3 b) { return Abs(a[0] * b[0] + a[1] * b[1] + a[2] * b[2]); }
static inline float Vec3Length2(Vec3 a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }
static inline Vec3 Vec3Neg(Vec3 a) { return V3(-a[0], -a[1], -a[2]); }

// math/vec3_amd64.s:46-56: (a1*a1 + a0*a0) + a2*a2 then SQRTSS
static inline float Vec3Length(Vec3 a) {
  float x0 = a[0] * a[0], x1 = a[1] * a[1], x2 = a[2] * a[2];
  x1 = x1 + x0;
  x1 = x1 + x2;
  return sqrtf(x1);
}

// math/vec3_amd64.s:11-43. RSQRTSS is hardware-approximate (vendor specific); the Newton step is
//   g' = (1.5 - (g*g)*(0.5*len2)) * g   in exactly that operand order.
static inline Vec3 Vec3Normalize(Vec3 a) {
  float x0 = a[0] * a[0], x1 = a[1] * a[1], x2 = a[2] * a[2];
  x1 = x1 + x0;
  x1 = x1 + x2;
  float g = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x1)));
  float h = 0.5f * x1;
  float gg = g * g;
  gg = gg * h;
  float c = 1.5f - gg;
  float r = c * g;
  return V3(a[0] * r, a[1] * r, a[2] * r);
}

static inline Vec3 Vec3BasisProject(Vec3 U, Vec3 V, Vec3 W, Vec3 S) { return V3(Vec3Dot(U, S), Vec3Dot(V, S), Vec3Dot(W, S)); }
static inline Vec3 Vec3BasisExpand(Vec3 U, Vec3 V, Vec3 W, Vec3 S) {
  return V3(U[0] * S[0] + V[0] * S[1] + W[0] * S[2], U[1] * S[0] + V[1] * S[1] + W[1] * S[2],
            U[2] * S[0] + V[2] * S[1] + W[2] * S[2]);
}

// math/boundingbox.go
struct BoundingBox {
  float b[2][3];
  void Reset() { for (int i = 0; i < 3; i++) { b[0][i] = kInfPos; b[1][i] = kInfNeg; } }
  float Dim(int axis) const { return b[1][axis] - b[0][axis]; }
  float SurfaceArea() const { return Dim(0) * Dim(1) * 2.0f + Dim(1) * Dim(2) * 2 + Dim(0) * Dim(2) * 2; }
  Vec3 Centroid() const { return V3((b[1][0] + b[0][0]) * 0.5f, (b[1][1] + b[0][1]) * 0.5f, (b[1][2] + b[0][2]) * 0.5f); }
  int MaxDim() const {  // boundingbox.go:59-76
    if (Dim(0) < Dim(1)) { return (Dim(1) < Dim(2)) ? 2 : 1; }
    return (Dim(0) < Dim(2)) ? 2 : 0;
  }
  void Grow(float X, float Y, float Z) {  // boundingbox.go:114-123 (note argument order of Min/Max)
    b[0][0] = Min(X, b[0][0]); b[1][0] = Max(X, b[1][0]);
    b[0][1] = Min(Y, b[0][1]); b[1][1] = Max(Y, b[1][1]);
    b[0][2] = Min(Z, b[0][2]); b[1][2] = Max(Z, b[1][2]);
  }
  void GrowVec3(Vec3 P) { Grow(P[0], P[1], P[2]); }
  void GrowBox(const BoundingBox& p) {  // boundingbox.go:106-111
    for (int k = 0; k < 3; k++) { b[0][k] = Min(b[0][k], p.b[0][k]); b[1][k] = Max(b[1][k], p.b[1][k]); }
  }
};
static inline BoundingBox InfBox() { BoundingBox x; for (int i = 0; i < 3; i++) { x.b[0][i] = kInfPos; x.b[1][i] = kInfPos; } return x; }
static inline BoundingBox BoundingBoxLerp(const BoundingBox& b0, const BoundingBox& b1, float t) {
  BoundingBox o;
  for (int i = 0; i < 2; i++) for (int k = 0; k < 3; k++) o.b[i][k] = (1 - t) * b0.b[i][k] + t * b1.b[i][k];
  return o;
}

// math/matrix4.go — column major, m[(j*4)+i] = row i, col j
struct Matrix4 {
  float m[16];
  float Elt(int i, int j) const { return m[(j * 4) + i]; }
  void Set(int i, int j, float v) { m[(j * 4) + i] = v; }
};
static inline Matrix4 Matrix4Identity() { Matrix4 r{}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
static inline Matrix4 Matrix4Null() { Matrix4 r{}; return r; }
static inline Matrix4 Matrix4Add(const Matrix4& a, const Matrix4& b) { Matrix4 c; for (int i = 0; i < 16; i++) c.m[i] = a.m[i] + b.m[i]; return c; }
static inline Matrix4 Matrix4Mul(const Matrix4& a, const Matrix4& b) {  // matrix4.go:89-99
  Matrix4 c{};
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) c.m[(j * 4) + i] += a.m[(k * 4) + i] * b.m[(j * 4) + k];
  return c;
}
static inline Matrix4 Matrix4Transpose(const Matrix4& a) { Matrix4 c; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c.m[(i * 4) + j] = a.m[(j * 4) + i]; return c; }
static inline Matrix4 Matrix4Scale(float s, const Matrix4& a) { Matrix4 x; for (int i = 0; i < 16; i++) x.m[i] = s * a.m[i]; return x; }
static inline Matrix4 Matrix4Lerp(const Matrix4& a, const Matrix4& b, float t) { Matrix4 x; for (int i = 0; i < 16; i++) x.m[i] = (1.0f - t) * a.m[i] + t * b.m[i]; return x; }
static inline Matrix4 Matrix4Translate(float X, float Y, float Z) { Matrix4 c = Matrix4Identity(); c.m[12] = X; c.m[13] = Y; c.m[14] = Z; return c; }
static inline Matrix4 Matrix4Basis(Vec3 u, Vec3 v, Vec3 w) {
  Matrix4 o{};
  o.m[0] = u[0]; o.m[1] = u[1]; o.m[2] = u[2];
  o.m[4] = v[0]; o.m[5] = v[1]; o.m[6] = v[2];
  o.m[8] = w[0]; o.m[9] = w[1]; o.m[10] = w[2];
  o.m[15] = 1.0f;
  return o;
}
static inline Vec3 Matrix4MulPoint(const Matrix4& a, Vec3 b) {
  return V3(a.m[0] * b[0] + a.m[4] * b[1] + a.m[8] * b[2] + a.m[12], a.m[1] * b[0] + a.m[5] * b[1] + a.m[9] * b[2] + a.m[13],
            a.m[2] * b[0] + a.m[6] * b[1] + a.m[10] * b[2] + a.m[14]);
}
static inline Vec3 Matrix4MulVec(const Matrix4& a, Vec3 b) {
  return V3(a.m[0] * b[0] + a.m[4] * b[1] + a.m[8] * b[2], a.m[1] * b[0] + a.m[5] * b[1] + a.m[9] * b[2],
            a.m[2] * b[0] + a.m[6] * b[1] + a.m[10] * b[2]);
}

// matrix4.go:113-147 / 150-171 — cofactor expansion, terms in source order.
static inline void cofactors(const float* m, float* inv) {
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
}
static inline bool Matrix4Inverse(const Matrix4& a, Matrix4* c) {
  float inv[16];
  const float* m = a.m;
  cofactors(m, inv);
  float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  if (det == 0.0f) { *c = Matrix4Null(); return false; }
  det = 1.0f / det;
  for (int i = 0; i < 16; i++) c->m[i] = inv[i] * det;
  return true;
}
static inline float Matrix4Det(const Matrix4& a) {
  float inv[16];
  const float* m = a.m;
  cofactors(m, inv);
  return m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
}
// matrix4.go:331-339 (note: signed difference, not absolute)
static inline bool Matrix4Eq(const Matrix4& A, const Matrix4& B, float eps) {
  for (int i = 0; i < 16; i++) if (A.m[i] - B.m[i] > eps) return false;
  return true;
}
// matrix4.go:345-363
static inline bool Matrix4PolarFactor(const Matrix4& m, Matrix4* out) {
  Matrix4 Q = m;
  for (int i = 0; i < 10; i++) {
    Matrix4 Qinv;
    if (!Matrix4Inverse(Q, &Qinv)) { *out = Matrix4Null(); return false; }
    Matrix4 Qnew = Matrix4Scale(.5f, Matrix4Add(Q, Matrix4Transpose(Qinv)));
    if (Matrix4Eq(Qnew, Q, 0.000001f)) { *out = Qnew; return true; }
    Q = Qnew;
  }
  *out = Matrix4Null();
  return false;
}

// math/quat.go
struct Quat { float X, Y, Z, W; };
static inline Quat Matrix4ToQuat(const Matrix4& m) {  // quat.go:124-154
  Quat q;
  float tr = m.Elt(0, 0) + m.Elt(1, 1) + m.Elt(2, 2);
  if (tr > 0.0f) {
    float S = Sqrt(tr + 1.0f) * 2;
    q.W = 0.25f * S;
    q.X = (m.Elt(2, 1) - m.Elt(1, 2)) / S;
    q.Y = (m.Elt(0, 2) - m.Elt(2, 0)) / S;
    q.Z = (m.Elt(1, 0) - m.Elt(0, 1)) / S;
  } else if ((m.Elt(0, 0) > m.Elt(1, 1)) && (m.Elt(0, 0) > m.Elt(2, 2))) {
    float S = Sqrt(1.0f + m.Elt(0, 0) - m.Elt(1, 1) - m.Elt(2, 2)) * 2;
    q.W = (m.Elt(2, 1) - m.Elt(1, 2)) / S;
    q.X = 0.25f * S;
    q.Y = (m.Elt(0, 1) + m.Elt(1, 0)) / S;
    q.Z = (m.Elt(0, 2) + m.Elt(2, 0)) / S;
  } else if (m.Elt(1, 1) > m.Elt(2, 2)) {
    float S = Sqrt(1.0f + m.Elt(1, 1) - m.Elt(0, 0) - m.Elt(2, 2)) * 2;
    q.W = (m.Elt(0, 2) - m.Elt(2, 0)) / S;
    q.X = (m.Elt(0, 1) + m.Elt(1, 0)) / S;
    q.Y = 0.25f * S;
    q.Z = (m.Elt(1, 2) + m.Elt(2, 1)) / S;
  } else {
    float S = Sqrt(1.0f + m.Elt(2, 2) - m.Elt(0, 0) - m.Elt(1, 1)) * 2;
    q.W = (m.Elt(1, 0) - m.Elt(0, 1)) / S;
    q.X = (m.Elt(0, 2) + m.Elt(2, 0)) / S;
    q.Y = (m.Elt(1, 2) + m.Elt(2, 1)) / S;
    q.Z = 0.25f * S;
  }
  return q;
}
static inline Matrix4 QuatToMatrix4(Quat q) {  // quat.go:73-96
  Matrix4 m{};
  float n = 1.0f / Sqrt(q.X * q.X + q.Y * q.Y + q.Z * q.Z + q.W * q.W);
  float qX = q.X * n, qY = q.Y * n, qZ = q.Z * n, qW = q.W * n;
  m.Set(0, 0, 1 - 2 * qY * qY - 2 * qZ * qZ);
  m.Set(0, 1, 2 * qX * qY - 2 * qW * qZ);
  m.Set(0, 2, 2 * qX * qZ + 2 * qW * qY);
  m.Set(1, 0, 2 * qX * qY + 2 * qW * qZ);
  m.Set(1, 1, 1 - 2 * qX * qX - 2 * qZ * qZ);
  m.Set(1, 2, 2 * qY * qZ - 2 * qW * qX);
  m.Set(2, 0, 2 * qX * qZ - 2 * qW * qY);
  m.Set(2, 1, 2 * qY * qZ + 2 * qW * qX);
  m.Set(2, 2, 1 - 2 * qX * qX - 2 * qY * qY);
  m.Set(3, 3, 1.0f);
  return m;
}
static inline Quat QuatSlerp(Quat qa, Quat qb, float t) {  // quat.go:30-66
  Quat qm;
  float cosHalfTheta = qa.W * qb.W + qa.X * qb.X + qa.Y * qb.Y + qa.Z * qb.Z;
  if (Abs(cosHalfTheta) >= 1.0f) return qa;
  float halfTheta = Acos(cosHalfTheta);
  float sinHalfTheta = Sqrt(1.0f - cosHalfTheta * cosHalfTheta);
  if (Abs(sinHalfTheta) < 0.001f) {
    qm.W = (qa.W * 0.5f + qb.W * 0.5f); qm.X = (qa.X * 0.5f + qb.X * 0.5f);
    qm.Y = (qa.Y * 0.5f + qb.Y * 0.5f); qm.Z = (qa.Z * 0.5f + qb.Z * 0.5f);
    return qm;
  }
  float ratioA = Sin((1 - t) * halfTheta) / sinHalfTheta;
  float ratioB = Sin(t * halfTheta) / sinHalfTheta;
  qm.W = (qa.W * ratioA + qb.W * ratioB); qm.X = (qa.X * ratioA + qb.X * ratioB);
  qm.Y = (qa.Y * ratioA + qb.Y * ratioB); qm.Z = (qa.Z * ratioA + qb.Z * ratioB);
  return qm;
}

// math/animdecomp.go
struct TransformDecomp { Vec3 T; Quat R; Matrix4 S; };
// math/octanormal.go:16 sign()
static inline float signf_(float v) { return v >= 0.0f ? 1.0f : -1.0f; }
// math/clamp.go:8-10
static inline float Clamp(float x, float mn, float mx) { return Max(mn, Min(x, mx)); }
static inline TransformDecomp TransformDecompMatrix4(Matrix4 m) {  // animdecomp.go:21-63
  TransformDecomp d;
  float sign = signf_(Matrix4Det(m));
  for (int i = 0; i < 3; i++) d.T[i] = m.Elt(i, 3);
  m.Set(0, 3, 0); m.Set(1, 3, 0); m.Set(2, 3, 0);
  if (sign < 0.0f) m = Matrix4Mul(Matrix4Scale(-1, Matrix4Identity()), m);
  Matrix4 Q;
  if (!Matrix4PolarFactor(m, &Q)) { d.R = Quat{0, 0, 0, 1}; d.S = Matrix4Identity(); return d; }
  Matrix4 S = Matrix4Mul(Matrix4Transpose(Q), m);
  d.R = Matrix4ToQuat(Q);
  if (sign < 0.0f) { d.S = Matrix4Mul(Matrix4Scale(-1, Matrix4Identity()), S); d.S.m[15] = 1; }
  else d.S = S;
  return d;
}
static inline Matrix4 TransformDecompToMatrix4(const TransformDecomp& d) {  // animdecomp.go:66-71
  return Matrix4Mul(Matrix4Translate(d.T[0], d.T[1], d.T[2]), Matrix4Mul(QuatToMatrix4(d.R), d.S));
}
static inline TransformDecomp TransformDecompLerp(const TransformDecomp& a, const TransformDecomp& b, float t) {
  TransformDecomp o;
  o.T = Vec3Lerp(a.T, b.T, t);
  o.R = QuatSlerp(a.R, b.R, t);
  o.S = Matrix4Lerp(a.S, b.S, t);
  return o;
}

}  // namespace orc
