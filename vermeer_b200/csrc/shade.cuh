// Device-side shading for the wavefront path integrator: QMC sequences, hero-wavelength colour,
// Oren-Nayar / mirror BSDFs, dielectric Fresnel, triangle-light sampling and the camera.
//
// Reference functions mirrored (evaluation order kept; compiled with -fmad=false):
//   ldseq.VanDerCorput/Sobol/RasterXY          math/ldseq/ldseq.go:50-96, raster.go:10-57 (m = 12 rows)
//   colour.Spectrum / Smits'99 / CIE / sRGB    colour/spectrum.go:49-112, spectrum_smits9.go:16-84,
//                                              cie1931_2deg.go:62-95, space_srgb.go:8-14
//   sample.CosineHemisphere / UniformDisk2D    math/sample/sample.go:18-27,105-129
//   bsdf.OrenNayar                             builtin/shader/bsdf/orennayar.go:24-73
//   bsdf.Specular, fresnel.Dielectric          builtin/shader/bsdf/specular.go:13-101, fresnel/dielectric.go:34-47
//   light.Tri sampling                         builtin/light/triangle.go:79-343,376-535, disk.go:38-51
//   camera.ComputeRay                          builtin/camera/camera.go:221-323
//   ShaderContext.OffsetP                      core/shader.go:139-162
// float32 trig goes through float64 exactly like math/sincos.go:16-70 (CUDA's double libm instead of Go's).
// Vec3Normalize (RSQRTSS + Newton, math/vec3_amd64.s:11-43) is hardware-approximate on the CPU and is
// replaced by a correctly rounded reciprocal square root: everything downstream of a normalize is
// tolerance-only (SURVEY.md note N).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// The CIE / Smits tables are indexed by a per-lane wavelength bin: in __constant__ memory a warp's lookup is serialised over
// its distinct addresses (ncu: the largest stall lines of k_shade), in global memory it is one L1-resident LDG.
#ifdef VG_TABLES_CONSTANT
#define VG_TABLE_QUAL __constant__ const
#else
#define VG_TABLE_QUAL __device__ const
#endif
#include "colour_tables.h"

namespace vg {

struct f3 {
  float x, y, z;
};
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 add3(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 sub3(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 scale3(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 neg3(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float len2_3(f3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ f3 mad3(f3 a, f3 b, float s) { return mk3(a.x + (b.x * s), a.y + (b.y * s), a.z + (b.z * s)); }
// math/vec3_amd64.s:46-56 (a1^2 + a0^2) + a2^2
__device__ __forceinline__ float length3(f3 a) {
  float x0 = a.x * a.x, x1 = a.y * a.y, x2 = a.z * a.z;
  x1 = x1 + x0;
  x1 = x1 + x2;
  return sqrtf(x1);
}
__device__ __forceinline__ f3 normalize3(f3 a) {
  float x0 = a.x * a.x, x1 = a.y * a.y, x2 = a.z * a.z;
  x1 = x1 + x0;
  x1 = x1 + x2;
  const float r = 1.0f / sqrtf(x1);  // two correctly rounded float ops: within 1 ulp, like the CPU's RSQRTSS + Newton step
  return mk3(a.x * r, a.y * r, a.z * r);
}
// FAST shading variant: one MUFU.RSQ (<= 2 ulp) instead of sqrt + divide
template <bool FAST>
__device__ __forceinline__ f3 normalize3t(f3 a) {
  if (!FAST) return normalize3(a);
  float x0 = a.x * a.x, x1 = a.y * a.y, x2 = a.z * a.z;
  x1 = x1 + x0;
  x1 = x1 + x2;
  const float r = rsqrtf(x1);
  return mk3(a.x * r, a.y * r, a.z * r);
}
// FAST shading variant of the scalar helpers: approximate division (MUFU.RCP + multiply, <= 2 ulp) and square root
// (MUFU.SQRT, <= 1 ulp; 0 -> 0, negative -> NaN like sqrtf) instead of the correctly rounded multi-instruction sequences
// -prec-div / -prec-sqrt produce. The traversal needs those flags (its arithmetic is bit-exact); the shading quantities that go
// through here are all downstream of a normalize, i.e. tolerance-only (SURVEY.md note N), and the precise path keeps IEEE.
template <bool FAST>
__device__ __forceinline__ float divt(float a, float b) { return FAST ? __fdividef(a, b) : a / b; }
template <bool FAST>
__device__ __forceinline__ float sqrtt(float x) {
  if (!FAST) return sqrtf(x);
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
template <bool FAST>
__device__ __forceinline__ float length3t(f3 a) {
  float x0 = a.x * a.x, x1 = a.y * a.y, x2 = a.z * a.z;
  x1 = x1 + x0;
  x1 = x1 + x2;
  return sqrtt<FAST>(x1);
}
__device__ __forceinline__ f3 basis_project(f3 U, f3 V, f3 W, f3 S) { return mk3(dot3(U, S), dot3(V, S), dot3(W, S)); }
__device__ __forceinline__ f3 basis_expand(f3 U, f3 V, f3 W, f3 S) {
  return mk3(U.x * S.x + V.x * S.y + W.x * S.z, U.y * S.x + V.y * S.y + W.y * S.z, U.z * S.x + V.z * S.y + W.z * S.z);
}
__device__ __forceinline__ float maxf_x86(float x, float y) { return x > y ? x : y; }  // math/dim_amd64.s
__device__ __forceinline__ float minf_x86(float x, float y) { return x < y ? x : y; }

// float32 trig. PRECISE: through float64 exactly like math/sincos.go:16-70 (CUDA's double libm instead of Go's).
// FAST (default): CUDA's single-precision libm (<= 2 ulp), about 3x cheaper on the shading kernel and far smaller
// code; both variants sit inside the image tolerance (everything here is downstream of a normalize anyway).
template <bool FAST>
struct Trig {
  static __device__ __forceinline__ float sin(float x) { return FAST ? ::sinf(x) : (float)::sin((double)x); }
  static __device__ __forceinline__ float cos(float x) { return FAST ? ::cosf(x) : (float)::cos((double)x); }
  static __device__ __forceinline__ float tan(float x) { return FAST ? ::tanf(x) : (float)::tan((double)x); }
  static __device__ __forceinline__ float acos(float x) { return FAST ? ::acosf(x) : (float)::acos((double)x); }
  static __device__ __forceinline__ float atan(float x) { return FAST ? ::atanf(x) : (float)::atan((double)x); }
  static __device__ __forceinline__ float atan2(float y, float x) { return FAST ? ::atan2f(y, x) : (float)::atan2((double)y, (double)x); }
};
__device__ __forceinline__ float sin32(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float cos32(float x) { return (float)cos((double)x); }

#define VG_PI32 3.14159265358f
#define VG_PI64 3.14159265358979323846

// ---- QMC (integer exact) ----------------------------------------------------------------------------
__device__ __forceinline__ uint64_t vdc_u(uint64_t i, uint64_t scramble) {
  // 64-bit bit reversal == the swap network of ldseq.go:55-77
  return (scramble ^ __brevll(i)) >> (64 - 52);
}
__device__ __forceinline__ double vdc(uint64_t i, uint64_t scramble) { return (double)vdc_u(i, scramble) / (double)(1ull << 52); }
__device__ __forceinline__ uint64_t sobol_u(uint64_t i, uint64_t scramble) {
  uint64_t r = scramble >> (64 - 52);
  for (uint64_t v = 1ull << (52 - 1); i != 0; i >>= 1) {
    if (i & 1) r ^= v;
    v ^= v >> 1;
  }
  return r;
}
__device__ __forceinline__ double sobol(uint64_t i, uint64_t scramble) { return (double)sobol_u(i, scramble) / (double)(1ull << 52); }

__constant__ const uint32_t kVdcSobolM12[28] = {0x808, 0xc0c, 0xa0a, 0xf0f, 0x888, 0xccc, 0xaaa, 0xfff, 0x800, 0xc00, 0xa00, 0xf00, 0x880, 0xcc0,
                                                0xaa0, 0xff0, 0x808, 0xc0c, 0xa0a, 0xf0f, 0x888, 0xccc, 0xaaa, 0xfff, 0x800, 0xc00, 0xa00, 0xf00};
__constant__ const uint32_t kVdcSobolInvM12[24] = {0xf0f000, 0x505000, 0x303000, 0x101000, 0xff0000, 0x550000, 0x330000, 0x110000,
                                                   0xf0000,  0x50000,  0x30000,  0x10000,  0x888800, 0x444400, 0x222200, 0x111100,
                                                   0x800080, 0x400040, 0x200020, 0x100010, 0x80008,  0x40004,  0x20002,  0x10001};
// raster.go:10-57 with m = 12 and scrambleX = scrambleY = 0 (core/render.go:89)
__device__ __forceinline__ void raster_xy12(uint32_t frame, uint32_t px, uint32_t py, double* rx, double* ry) {
  uint64_t index = (uint64_t)frame << 24;
  uint32_t delta = 0;
  for (uint32_t c = 0, f = frame; f != 0; f >>= 1, c++)
    if (f & 1) delta ^= kVdcSobolM12[c];
  uint32_t b = ((px << 12) | py) ^ delta;
  for (uint32_t c = 0; b != 0; b >>= 1, c++)
    if (b & 1) index ^= (uint64_t)kVdcSobolInvM12[c];
  *rx = (double)vdc_u(index, 0) / (double)(1ull << 40);
  *ry = (double)sobol_u(index, 0) / (double)(1ull << 40);
}

// Byte-sliced form of the same enumeration: the XOR over the set bits of `b` (24 bits) and of `index` (52 bits) is looked up
// 8 bits at a time in tables built on the host from the same matrices (RenderState::qmc). Integer exact, so identical.
struct QmcTables {
  uint64_t sob[7][256];   // XOR of the Sobol' direction numbers v_k (ldseq.go:83-96) over the set bits of byte k
  uint32_t rinv[3][256];  // XOR of vdCSobolInvMatrices[12] rows (raster.go:24-36) over the set bits of byte k
  // Spectrum.FromRGB of white (the Oren-Nayar reflectance's spectrum, orennayar.go:66-69) per Smits bin: [bin + 1], bin = -1..9.
  // RGBToSpectrumSmits99 depends on the colour and the wavelength's bin only (spectrum_smits9.go:16-25,48-84), so a constant colour
  // has 11 possible values; k_spectrum_tables evaluates them once with the same device function the per-vertex path used.
  float white_spec[12];
};
__device__ __forceinline__ void raster_xy12_tab(const QmcTables* __restrict__ T, uint32_t frame, uint32_t px, uint32_t py, double* rx, double* ry) {
  uint64_t index = (uint64_t)frame << 24;
  uint32_t delta = 0;
  for (uint32_t c = 0, f = frame; f != 0; f >>= 1, c++)
    if (f & 1) delta ^= kVdcSobolM12[c];
  const uint32_t b = ((px << 12) | py) ^ delta;
  index ^= (uint64_t)(__ldg(&T->rinv[0][b & 255u]) ^ __ldg(&T->rinv[1][(b >> 8) & 255u]) ^ __ldg(&T->rinv[2][(b >> 16) & 255u]));
  *rx = (double)vdc_u(index, 0) / (double)(1ull << 40);
  uint64_t r = 0;
  int k = 0;
  for (uint64_t i = index; i != 0; i >>= 8, k++) r ^= __ldg(&T->sob[k][i & 255u]);
  *ry = (double)r / (double)(1ull << 40);
}

// ---- colour -------------------------------------------------------------------------------------------
struct Spec4 {
  float c[4];
};
__device__ __forceinline__ float wavelength(float lambda, int j) {  // spectrum.go:101-112
  float v = (lambda - 450.0f + ((float)j / 4.0f) * 300.0f);
  if (v >= 300.0f) v -= 300.0f;
  v += 450.0f;
  return v;
}
// The 4 hero wavelengths of a path depend on its Lambda only (spectrum.go:101-112), so the Smits bin
// (spectrum_smits9.go:16-25: -1 outside [380,720)) and the three nearest-bin CIE observer values (cie1931_2deg.go:62-95:
// -1 outside [360,830)) are looked up once per path vertex instead of once per FromRGB / ToRGB call.
struct Hero {
  int sbin[4];
  float cx[4], cy[4], cz[4];
};
__device__ __forceinline__ Hero hero_setup(float lambda) {
  Hero h;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float wl = wavelength(lambda, k);
    h.sbin[k] = (wl < 380.0f || wl >= 720.0f) ? -1 : (int)(((wl - 380.0f) / (720.0f - 380.0f)) * 10.0f);
    if (wl < 360.0f || wl >= 830.0f) {
      h.cx[k] = h.cy[k] = h.cz[k] = -1.0f;
    } else {
      const int bin = (int)(((wl - 360.0f) / (830.0f - 360.0f)) * 95.0f);
      h.cx[k] = kCieX[bin]; h.cy[k] = kCieY[bin]; h.cz[k] = kCieZ[bin];
    }
  }
  return h;
}
__device__ __forceinline__ float smits_eval(const float* s, int bin) { return bin < 0 ? 0.0f : s[bin]; }
__device__ inline float rgb_to_spectrum(float r, float g, float b, int bin) {  // spectrum_smits9.go:48-84
  float c = 0.0f;
  if (r <= g && r <= b) {
    c += r * smits_eval(kSmitsWhite, bin);
    if (g <= b) { c += (g - r) * smits_eval(kSmitsCyan, bin); c += (b - g) * smits_eval(kSmitsBlue, bin); }
    else { c += (b - r) * smits_eval(kSmitsCyan, bin); c += (g - b) * smits_eval(kSmitsGreen, bin); }
  } else if (g <= r && g <= b) {
    c += g * smits_eval(kSmitsWhite, bin);
    if (r <= b) { c += (r - g) * smits_eval(kSmitsMagenta, bin); c += (b - r) * smits_eval(kSmitsBlue, bin); }
    else { c += (b - g) * smits_eval(kSmitsMagenta, bin); c += (r - b) * smits_eval(kSmitsRed, bin); }
  } else {
    c += b * smits_eval(kSmitsWhite, bin);
    if (r <= g) { c += (r - b) * smits_eval(kSmitsYellow, bin); c += (g - r) * smits_eval(kSmitsGreen, bin); }
    else { c += (g - b) * smits_eval(kSmitsYellow, bin); c += (r - g) * smits_eval(kSmitsRed, bin); }
  }
  return c;
}
__device__ __forceinline__ Spec4 spec_from_rgb(f3 rgb, const Hero& h) {  // spectrum.go:49-54
  Spec4 s;
#pragma unroll
  for (int k = 0; k < 4; k++) s.c[k] = rgb_to_spectrum(rgb.x, rgb.y, rgb.z, h.sbin[k]);
  return s;
}
// the same for a colour whose 11 per-bin values were tabulated (light emissions, white): four loads instead of four Smits evaluations
__device__ __forceinline__ Spec4 spec_from_table(const float* __restrict__ tab, const Hero& h) {
  Spec4 s;
#pragma unroll
  for (int k = 0; k < 4; k++) s.c[k] = __ldg(tab + h.sbin[k] + 1);
  return s;
}
__device__ __forceinline__ f3 spec_to_rgb(const Spec4& s, const Hero& h) {  // spectrum.go:57-72
  float x = 0, y = 0, z = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    x += s.c[i] * h.cx[i];
    y += s.c[i] * h.cy[i];
    z += s.c[i] * h.cz[i];
  }
  return mk3(x * 3.2404542f + y * -1.5371385f + z * -0.4985314f, x * -0.9692660f + y * 1.8760108f + z * 0.0415560f,
             x * 0.0556434f + y * -0.2040259f + z * 1.0572252f);
}

// ---- sampling warps ---------------------------------------------------------------------------------
template <bool FAST>
__device__ __forceinline__ f3 cosine_hemisphere(double u0, double u1) {  // sample.go:18-27
  if (FAST) {
    const float r = sqrtt<true>((float)(1 - u0));
    float sn, cs;
    sincospif((float)(2 * u1), &sn, &cs);
    return mk3(r * cs, r * sn, sqrtt<true>((float)u0));
  }
  const double r = sqrt(1 - u0);
  const double theta = 2 * VG_PI64 * u1;
  return mk3((float)(r * cos(theta)), (float)(r * sin(theta)), (float)sqrt(u0));
}
__device__ inline void uniform_disk2d(float radius, float r0, float r1, float* xo, float* yo) {  // sample.go:105-129
  const float x = -1 + 2 * r0, y = -1 + 2 * r1;
  float r = 0, theta = 0;
  if (x > -y && x > y) { r = x; theta = (VG_PI32 / 4) * y / x; }
  else if (x > -y && x < y) { r = y; theta = (VG_PI32 / 4) * (2 - x / y); }
  else if (x < y && x < -y) { r = -x; theta = (VG_PI32 / 4) * (4 + y / x); }
  else if (x > y && x < -y) { r = -y; theta = (VG_PI32 / 4) * (6 - x / y); }
  *xo = radius * r * cos32(theta);
  *yo = radius * r * sin32(theta);
}

// ---- BSDFs ------------------------------------------------------------------------------------------
struct Frame {
  f3 U, V, N;
};
// orennayar.go:34-39
__device__ __forceinline__ double oren_pdf(const Frame& f, f3 wo) {
  const f3 o = basis_project(f.U, f.V, f.N, wo);
  return (double)maxf_x86(0.0f, o.z) / VG_PI64;
}
// float32(bsdf.PDF(wo)) as EvaluateLightSamples uses it (core/shader.go:284,323). FAST: one float multiply by 1/pi instead
// of the float64 divide (the two differ by at most 1 ulp of float32).
template <bool FAST>
__device__ __forceinline__ float oren_pdf32(const Frame& f, f3 wo) {
  if (!FAST) return (float)oren_pdf(f, wo);
  return maxf_x86(0.0f, dot3(f.N, wo)) * 0.318309886183790671538f;
}
// orennayar.go:42-73 split into its per-vertex part (depends on omegaI, lambda and the material only) and its
// per-sample part; the arithmetic and its order are unchanged.
struct OrenVertex {
  float A, B;          // from Roughness^2 (orennayar.go:25,47-49)
  float phiI, thetaI;  // atan2(omegaI.y, omegaI.x), acos(omegaI.z)           (precise variant)
  float cosI, sinI, cphiI, sphiI;  // the same two angles as cosine/sine pairs  (FAST variant, no inverse trig)
  Spec4 white;         // FromRGB({1,1,1}) at this path's hero wavelength
};
// (cos, sin) of atan2(y, x) without the angle; atan2(0, 0) = 0 -> (1, 0)
__device__ __forceinline__ void unit2(float x, float y, float* c, float* s) {
  const float l2 = x * x + y * y;
  if (l2 > 0.0f) {
    const float r = rsqrtf(l2);
    *c = x * r;
    *s = y * r;
  } else {
    *c = 1.0f;
    *s = 0.0f;
  }
}
template <bool FAST>
__device__ __forceinline__ OrenVertex oren_vertex(f3 omegaI, float roughness2, const Hero& hero, const float* __restrict__ white_tab = nullptr) {
  OrenVertex v;
  const float sigma = roughness2;
  v.A = 1 - (0.5f * (sigma * sigma) / ((sigma * sigma) + 0.57f));
  v.B = 0.45f * (sigma * sigma) / ((sigma * sigma) + 0.09f);
  if (FAST) {
    v.cosI = omegaI.z;
    v.sinI = sqrtt<true>(1.0f - omegaI.z * omegaI.z);  // NaN for |z| > 1, like acos
    unit2(omegaI.x, omegaI.y, &v.cphiI, &v.sphiI);
    v.phiI = v.thetaI = 0.0f;
  } else {
    v.phiI = Trig<FAST>::atan2(omegaI.y, omegaI.x);
    v.thetaI = Trig<FAST>::acos(omegaI.z);
    v.cosI = v.sinI = v.cphiI = v.sphiI = 0.0f;
  }
  v.white = white_tab ? spec_from_table(white_tab, hero) : spec_from_rgb(mk3(1, 1, 1), hero);
  return v;
}
template <bool FAST>
__device__ __forceinline__ Spec4 oren_eval(const Frame& f, const OrenVertex& ov, f3 wo) {
  const f3 o = basis_project(f.U, f.V, f.N, wo);
  float Cc, gamma;
  if (FAST) {
    // the same expression without inverse trig: theta = acos(cos) is decreasing, so max/min of the angles are picked by
    // comparing cosines; sin(alpha) tan(beta) and cos(phiO - phiI) follow from the cosine/sine pairs. NaN cases keep the
    // reference's x86 max/min semantics (a NaN thetaI is dropped, a NaN thetaO propagates).
    const float cosO = o.z, sinO = sqrtt<true>(1.0f - o.z * o.z);
    const bool ok = (ov.sinI == ov.sinI) && (sinO == sinO);
    const bool aI = ok && ov.cosI < cosO;  // alpha = thetaI
    const bool bI = ok && ov.cosI > cosO;  // beta = thetaI
    const float sinA = aI ? ov.sinI : sinO;
    const float tanB = bI ? __fdividef(ov.sinI, ov.cosI) : __fdividef(sinO, cosO);
    Cc = sinA * tanB;
    float cphiO, sphiO;
    unit2(o.x, o.y, &cphiO, &sphiO);
    gamma = cphiO * ov.cphiI + sphiO * ov.sphiI;
  } else {
    const float phiO = Trig<FAST>::atan2(o.y, o.x);
    const float thetaO = Trig<FAST>::acos(o.z);
    const float alpha = maxf_x86(ov.thetaI, thetaO);
    const float beta = minf_x86(ov.thetaI, thetaO);
    Cc = Trig<FAST>::sin(alpha) * Trig<FAST>::tan(beta);
    gamma = Trig<FAST>::cos(phiO - ov.phiI);
  }
  const float sc = o.z * (ov.A + (ov.B * maxf_x86(0.0f, gamma) * Cc));
  Spec4 rho = ov.white;
  const float k = FAST ? sc * 0.318309886183790671538f : sc / (float)VG_PI64;
#pragma unroll
  for (int i = 0; i < 4; i++) rho.c[i] *= k;
  return rho;
}
// fresnel/dielectric.go:34-47
__device__ __forceinline__ float dielectric_kr(float eta, float c) {
  float g = (eta * eta) - 1 + (c * c);
  if (g < 0.0f) return 1.0f;
  g = sqrtf(g);
  const float a = (g - c) / (g + c);
  const float b = (c * (g + c) - 1) / (c * (g - c) + 1);
  return 0.5f * (a * a) * (1 + (b * b));
}
__device__ __forceinline__ f3 reflect_z(f3 w) {  // specular.go:13-19 with N = (0,0,1)
  const f3 n = mk3(0, 0, 1);
  return sub3(scale3(2.0f * dot3(n, w), n), w);
}

// ---- OffsetP (core/shader.go:139-162) ----------------------------------------------------------------
__device__ __forceinline__ float next_up(float v) {  // math/ferror.go:27-45
  if (isinf(v) && v > 0) return v;
  if (v == -0.0f) v = 0.0f;
  uint32_t ui = __float_as_uint(v);
  if (v >= 0.0f) ui++; else ui--;
  return __uint_as_float(ui);
}
__device__ __forceinline__ float next_down(float v) {  // math/ferror.go:48-66
  if (isinf(v) && v < 0) return v;
  if (v == -0.0f) v = 0.0f;
  uint32_t ui = __float_as_uint(v);
  if (v >= 0.0f) ui--; else ui++;
  return __uint_as_float(ui);
}
__device__ __forceinline__ float offset1(float p, float o) {
  float r = p + o;
  if (o > 0) r = next_up(r);
  else if (o < -0.0f) r = next_down(r);
  return r;
}
__device__ __forceinline__ f3 offset_p(f3 P, f3 Poffset, int dir) {
  const f3 o = dir < 0 ? neg3(Poffset) : Poffset;
  return mk3(offset1(P.x, o.x), offset1(P.y, o.y), offset1(P.z, o.z));
}

// ---- triangle light ---------------------------------------------------------------------------------
// builtin/light/triangle.go:79-134 (Moller-Trumbore, no culling)
template <bool FAST = false>
__device__ inline bool ray_triangle(f3 Ro, f3 Rd, f3 P0, f3 P1, f3 P2, f3* pout) {
  const f3 e1 = sub3(P1, P0), e2 = sub3(P2, P0);
  const f3 P = cross3(Rd, e2);
  const float det = dot3(e1, P);
  if (det > -1e-6f && det < 1e-6f) return false;
  const float inv_det = divt<FAST>(1.0f, det);
  const f3 T = sub3(Ro, P0);
  const float u = dot3(T, P) * inv_det;
  if (u < 0 || u > 1) return false;
  const f3 Q = cross3(T, e1);
  const float v = dot3(Rd, Q) * inv_det;
  if (v < 0 || u + v > 1) return false;
  const float t = dot3(e2, Q) * inv_det;
  if (t > 1e-6f) {
    const f3 a = scale3(1 - u - v, P0), b = scale3(u, P1), c = scale3(v, P2);
    *pout = mk3(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z);
    return true;
  }
  return false;
}
// triangle.go:376-417, 430-459, 474-489: the part of Arvo's spherical-triangle sampling that depends only on the
// shading point and the light (unit vectors to the vertices, side c, angle alpha, solid angle) — computed once per
// (vertex, light) and shared by SampleArea and ValidSample, which both recompute it in the reference.
struct SphTri {
  f3 pa, pb, pc;
  float alpha, c, area;        // precise variant: the angle at A, the side AB, the solid angle
  float cosAlpha, sinAlpha, cosC;  // FAST variant: the same two angles as cosine/sine values
};
template <bool FAST>
__device__ inline SphTri spherical_setup(f3 p0, f3 p1, f3 p2, f3 p) {
  SphTri t;
  t.pa = normalize3t<FAST>(sub3(p0, p));
  t.pb = normalize3t<FAST>(sub3(p1, p));
  t.pc = normalize3t<FAST>(sub3(p2, p));
  if (FAST) {
    // Same quantities without the six inverse-trig calls: the vertex angle alpha is the dihedral angle between the planes
    // (pa,pb) and (pa,pc); the solid angle by Van Oosterom & Strackee, tan(O/2) = |pa.(pb x pc)| / (1 + pa.pb + pb.pc + pc.pa).
    const f3 n1 = cross3(t.pa, t.pb), n2 = cross3(t.pa, t.pc);
    const float trip = fabsf(dot3(t.pa, cross3(t.pb, t.pc)));
    const float rn = rsqrtf(len2_3(n1) * len2_3(n2));
    t.cosAlpha = dot3(n1, n2) * rn;
    t.sinAlpha = trip * rn;
    t.cosC = dot3(t.pa, t.pb);
    const float denom = 1.0f + t.cosC + dot3(t.pb, t.pc) + dot3(t.pc, t.pa);
    t.area = 2.0f * atan2f(trip, denom);
    t.alpha = t.c = 0.0f;
    return t;
  }
  t.cosAlpha = t.sinAlpha = t.cosC = 0.0f;
  const float as = Trig<FAST>::acos(dot3(t.pb, t.pc)), bs = Trig<FAST>::acos(dot3(t.pc, t.pa)), cs = Trig<FAST>::acos(dot3(t.pa, t.pb));
  const float ssu = (as + bs + cs) / 2;
  const float sa = Trig<FAST>::sin(ssu - as), sb = Trig<FAST>::sin(ssu - bs), scs = Trig<FAST>::sin(ssu - cs), ss = Trig<FAST>::sin(ssu);
  const float tanA2 = sqrtf(sb * scs / (ss * sa));
  const float tanB2 = sqrtf(sa * scs / (ss * sb));
  const float tanC2 = sqrtf(sa * sb / (ss * scs));
  const float alpha = 2 * Trig<FAST>::atan(tanA2), beta = 2 * Trig<FAST>::atan(tanB2), gamma = 2 * Trig<FAST>::atan(tanC2);
  t.alpha = alpha;
  t.c = cs;
  t.area = alpha + beta + gamma - VG_PI32;
  return t;
}
// triangle.go:490-535 (Arvo). Returns the unit direction; pdf = 1/solid angle.
template <bool FAST>
__device__ inline f3 sample_spherical_triangle(const SphTri& st, double r0, double r1) {
  const f3 pa = st.pa, pb = st.pb, pc = st.pc;
  const float alpha = st.alpha;
  const float areaHat = (float)r0 * st.area;
  float s, t, sinAlpha, cosAlpha, cosC;
  if (FAST) {
    float sh, ch;
    sincosf(areaHat, &sh, &ch);
    sinAlpha = st.sinAlpha; cosAlpha = st.cosAlpha; cosC = st.cosC;
    s = sh * cosAlpha - ch * sinAlpha;  // sin(areaHat - alpha)
    t = ch * cosAlpha + sh * sinAlpha;  // cos(areaHat - alpha)
  } else {
    s = Trig<FAST>::sin(areaHat - alpha); t = Trig<FAST>::cos(areaHat - alpha);
    sinAlpha = Trig<FAST>::sin(alpha); cosAlpha = Trig<FAST>::cos(alpha);
    cosC = Trig<FAST>::cos(st.c);
  }
  const float u = t - cosAlpha;
  const float v = s + sinAlpha * cosC;
  float q = divt<FAST>((v * t - u * s) * cosAlpha - v, (v * s + u * t) * sinAlpha);
  q = maxf_x86(-1.0f, minf_x86(q, 1.0f));
  float w = dot3(pc, pa);
  f3 v31 = normalize3t<FAST>(mk3(pc.x - w * pa.x, pc.y - w * pa.y, pc.z - w * pa.z));
  const float sq = sqrtt<FAST>(1 - q * q);
  const f3 v4 = mk3(q * pa.x + sq * v31.x, q * pa.y + sq * v31.y, q * pa.z + sq * v31.z);
  const float z = 1 - (float)r1 * (1 - dot3(v4, pb));
  w = dot3(v4, pb);
  const f3 v42 = normalize3t<FAST>(mk3(v4.x - w * pb.x, v4.y - w * pb.y, v4.z - w * pb.z));
  return add3(scale3(z, pb), scale3(sqrtt<FAST>(1 - z * z), v42));
}
// disk.go:38-51
template <bool FAST = false>
__device__ __forceinline__ float ray_plane(f3 Ro, f3 Rd, f3 P, f3 N) {
  const float denom = dot3(N, Rd);
  if (fabsf(denom) > 1e-6f) return divt<FAST>(dot3(sub3(P, Ro), N), denom);
  return 0.0f;
}

}  // namespace vg
