// Object-transform arithmetic shared by the host flattening (context.cu) and the device kernels: the per-ray part of
// GeomInstance.Trace (builtin/geom/instance/instance.go:16-33,73-96) in the reference's float32 operation order.
//   math.Matrix4Mul / Transpose / Inverse      math/matrix4.go:89-99,102-147
//   math.QuatToMatrix4 / QuatSlerp              math/quat.go:30-66,73-96
//   math.TransformDecompLerp / ToMatrix4        math/animdecomp.go:66-83
//   math.Matrix4MulPoint / MulVec               math/matrix4.go (column-major storage: element (i,j) at [j*4+i])
// Device code is compiled with -fmad=false and host code with -ffp-contract=off, so both evaluate these expressions like
// Go/amd64 does; Acos/Sin go through float64 like math/sincos.go.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define VG_HD __host__ __device__ inline
#else
#define VG_HD inline
#endif

namespace vg {

struct Mat4 {
  float m[16];
};
// m.TransformDecomp (math/animdecomp.go:12-17): T Vec3, R Quat{X,Y,Z,W}, S Matrix4 — 23 contiguous floats
struct XfSRT {
  float T[3];
  float R[4];
  float S[16];
};

VG_HD Mat4 m4_mul(const Mat4& a, const Mat4& b) {
  Mat4 c;
  for (int i = 0; i < 16; i++) c.m[i] = 0.0f;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 4; k++) c.m[(j * 4) + i] += a.m[(k * 4) + i] * b.m[(j * 4) + k];
  return c;
}
VG_HD Mat4 m4_transpose(const Mat4& a) {
  Mat4 c;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) c.m[(i * 4) + j] = a.m[(j * 4) + i];
  return c;
}
VG_HD void m4_cofactors(const float* m, float* inv) {  // math/matrix4.go:113-132, term order kept
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
VG_HD Mat4 m4_inverse(const Mat4& a) {  // matrix4.go:134-147: the null matrix when singular
  Mat4 c;
  float inv[16];
  m4_cofactors(a.m, inv);
  float det = a.m[0] * inv[0] + a.m[1] * inv[4] + a.m[2] * inv[8] + a.m[3] * inv[12];
  if (det == 0.0f) {
    for (int i = 0; i < 16; i++) c.m[i] = 0.0f;
    return c;
  }
  det = 1.0f / det;
  for (int i = 0; i < 16; i++) c.m[i] = inv[i] * det;
  return c;
}
VG_HD Mat4 quat_to_m4(float X, float Y, float Z, float W) {  // quat.go:73-96
  Mat4 m;
  for (int i = 0; i < 16; i++) m.m[i] = 0.0f;
  const float n = 1.0f / sqrtf(X * X + Y * Y + Z * Z + W * W);
  const float x = X * n, y = Y * n, z = Z * n, w = W * n;
  m.m[0 * 4 + 0] = 1 - 2 * y * y - 2 * z * z; m.m[1 * 4 + 0] = 2 * x * y - 2 * w * z; m.m[2 * 4 + 0] = 2 * x * z + 2 * w * y;
  m.m[0 * 4 + 1] = 2 * x * y + 2 * w * z; m.m[1 * 4 + 1] = 1 - 2 * x * x - 2 * z * z; m.m[2 * 4 + 1] = 2 * y * z - 2 * w * x;
  m.m[0 * 4 + 2] = 2 * x * z - 2 * w * y; m.m[1 * 4 + 2] = 2 * y * z + 2 * w * x; m.m[2 * 4 + 2] = 1 - 2 * x * x - 2 * y * y;
  m.m[15] = 1.0f;
  return m;
}
VG_HD Mat4 srt_to_m4(const XfSRT& d) {  // animdecomp.go:66-71: Translate(T) * (QuatToMatrix4(R) * S)
  Mat4 t;
  for (int i = 0; i < 16; i++) t.m[i] = 0.0f;
  t.m[0] = t.m[5] = t.m[10] = t.m[15] = 1.0f;
  t.m[12] = d.T[0]; t.m[13] = d.T[1]; t.m[14] = d.T[2];
  Mat4 s;
  for (int i = 0; i < 16; i++) s.m[i] = d.S[i];
  return m4_mul(t, m4_mul(quat_to_m4(d.R[0], d.R[1], d.R[2], d.R[3]), s));
}
VG_HD XfSRT srt_lerp(const XfSRT& a, const XfSRT& b, float t) {  // animdecomp.go:74-83
  XfSRT o;
  for (int i = 0; i < 3; i++) o.T[i] = (1.0f - t) * a.T[i] + t * b.T[i];  // Vec3Lerp (vec3.go)
  // QuatSlerp (quat.go:30-66); R = {X,Y,Z,W}
  const float c = a.R[3] * b.R[3] + a.R[0] * b.R[0] + a.R[1] * b.R[1] + a.R[2] * b.R[2];
  if (fabsf(c) >= 1.0f) {
    for (int i = 0; i < 4; i++) o.R[i] = a.R[i];
  } else {
    const float half = (float)acos((double)c);
    const float s = sqrtf(1.0f - c * c);
    if (fabsf(s) < 0.001f) {
      for (int i = 0; i < 4; i++) o.R[i] = a.R[i] * 0.5f + b.R[i] * 0.5f;
    } else {
      const float ra = (float)sin((double)((1 - t) * half)) / s;
      const float rb = (float)sin((double)(t * half)) / s;
      for (int i = 0; i < 4; i++) o.R[i] = a.R[i] * ra + b.R[i] * rb;
    }
  }
  for (int i = 0; i < 16; i++) o.S[i] = (1.0f - t) * a.S[i] + t * b.S[i];  // Matrix4Lerp
  return o;
}
// TransformSRTArray.TimeKey (instance.go:16-33) + ToMatrix4 + Inverse for one ray time
VG_HD void xf_matrices(const XfSRT* keys, int nkeys, float time, Mat4* M, Mat4* Minv) {
  if (nkeys > 1) {
    const float k = time * (float)(nkeys - 1);
    const float fk = floorf(k);
    const float frac = k - fk;
    const int key = (int)fk, key2 = (int)ceilf(k);
    *M = srt_to_m4(srt_lerp(keys[key], keys[key2], frac));
  } else {
    *M = srt_to_m4(keys[0]);
  }
  *Minv = m4_inverse(*M);
}
VG_HD void m4_mul_point(const Mat4& a, float x, float y, float z, float* o) {
  o[0] = a.m[0] * x + a.m[4] * y + a.m[8] * z + a.m[12];
  o[1] = a.m[1] * x + a.m[5] * y + a.m[9] * z + a.m[13];
  o[2] = a.m[2] * x + a.m[6] * y + a.m[10] * z + a.m[14];
}
VG_HD void m4_mul_vec(const Mat4& a, float x, float y, float z, float* o) {
  o[0] = a.m[0] * x + a.m[4] * y + a.m[8] * z;
  o[1] = a.m[1] * x + a.m[5] * y + a.m[9] * z;
  o[2] = a.m[2] * x + a.m[6] * y + a.m[10] * z;
}

}  // namespace vg
