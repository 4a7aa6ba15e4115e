// ORACLE — TEST INFRASTRUCTURE ONLY. See shading.h for the reference files this restates.
#include "shading.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <stdexcept>
#include <thread>

#include "ldseq.h"

namespace orc {

static inline float sqr(float x) { return x * x; }

// ---------------------------------------------------------------------------------------------
// math/sample/sample.go:18-27
static Vec3 CosineHemisphere(double u0, double u1) {
  double r = std::sqrt(1 - u0);
  double theta = 2 * M_PI * u1;
  double x = r * std::cos(theta);
  double y = r * std::sin(theta);
  return V3((float)x, (float)y, (float)std::sqrt(u0));
}

// math/sample/sample.go:105-129
static void UniformDisk2D(float radius, float r0, float r1, float* xo, float* yo) {
  float x = -1 + 2 * r0;
  float y = -1 + 2 * r1;
  float r = 0, theta = 0;
  if (x > -y && x > y) {
    r = x;
    theta = (kPi / 4) * y / x;
  } else if (x > -y && x < y) {
    r = y;
    theta = (kPi / 4) * (2 - x / y);
  } else if (x < y && x < -y) {
    r = -x;
    theta = (kPi / 4) * (4 + y / x);
  } else if (x > y && x < -y) {
    r = -y;
    theta = (kPi / 4) * (6 - x / y);
  }
  *xo = radius * r * Cos(theta);
  *yo = radius * r * Sin(theta);
}

// builtin/shader/bsdf/orennayar.go:16-73
struct OrenNayar : BSDF {
  float Lambda;
  Vec3 OmegaI;
  float Roughness;
  Vec3 U, V, N;
  OrenNayar(float lambda, Vec3 omegaI, float roughness, Vec3 U_, Vec3 V_, Vec3 N_)
      : Lambda(lambda), OmegaI(Vec3BasisProject(U_, V_, N_, omegaI)), Roughness(roughness * roughness), U(U_), V(V_), N(N_) {}
  Vec3 Sample(double r0, double r1) override { return Vec3BasisExpand(U, V, N, CosineHemisphere(r0, r1)); }
  double PDF(Vec3 _omegaO) override {
    Vec3 omegaO = Vec3BasisProject(U, V, N, _omegaO);
    double ODotN = (double)Max(0, omegaO[2]);
    return ODotN / M_PI;
  }
  Spectrum Eval(Vec3 _omegaO) override {
    Vec3 omegaO = Vec3BasisProject(U, V, N, _omegaO);
    float sigma = Roughness;
    float A = 1 - (0.5f * (sigma * sigma) / ((sigma * sigma) + 0.57f));
    float B = 0.45f * (sigma * sigma) / ((sigma * sigma) + 0.09f);
    float phiI = Atan2(OmegaI[1], OmegaI[0]);
    float phiO = Atan2(omegaO[1], omegaO[0]);
    float thetaI = Acos(OmegaI[2]);
    float thetaO = Acos(omegaO[2]);
    float alpha = Max(thetaI, thetaO);
    float beta = Min(thetaI, thetaO);
    float C = Sin(alpha) * Tan(beta);
    float gamma = Cos(phiO - phiI);
    float scale = omegaO[2] * (A + (B * Max(0, gamma) * C));
    Spectrum rho;
    rho.Lambda = Lambda;
    rho.FromRGB(MakeRGB(1, 1, 1));
    rho.Scale(scale / (float)M_PI);
    return rho;
  }
};

// core.Fresnel (core/shader.go:16-19)
struct Fresnel {
  virtual ~Fresnel() {}
  virtual RGB Kr(float cosTheta) const = 0;
};

// builtin/shader/fresnel/dielectric.go:34-47
struct Dielectric : Fresnel {
  float eta;
  explicit Dielectric(float e) : eta(e) {}
  RGB Kr(float cosTheta) const override {
    float c = cosTheta;
    float g = (eta * eta) - 1 + (c * c);
    if (g < 0.0f) return MakeRGB(1, 1, 1);
    g = Sqrt(g);
    float fr = 0.5f * sqr((g - c) / (g + c)) * (1 + sqr((c * (g + c) - 1) / (c * (g - c) + 1)));
    return MakeRGB(fr, fr, fr);
  }
};

// builtin/shader/fresnel/conductor.go:19-89 ('Artist Friendly Metallic Fresnel', Gulbrandsen)
struct Conductor : Fresnel {
  RGB r, g;
  Conductor(RGB r_, RGB g_) : r(r_), g(g_) {}
  static float nMin(float r) { return (1 - r) / (1 + r); }
  static float nMax(float r) { return (1 + Sqrt(r)) / (1 - Sqrt(r)); }
  static float getN(float r, float g) { return nMin(r) * g + (1 - g) * nMax(r); }
  static float getK2(float r, float n) {
    float nr = (n + 1) * (n + 1) * r - (n - 1) * (n - 1);
    return nr / (1 - r);
  }
  static float fresnel(float r, float g, float cosTheta) {
    float nr = Clamp(r, 0, 0.99f);
    float n = getN(nr, g);
    float k2 = getK2(nr, n);
    float rsNum = n * n + k2 - 2 * n * cosTheta + cosTheta * cosTheta;
    float rsDen = n * n + k2 + 2 * n * cosTheta + cosTheta * cosTheta;
    float rs = rsNum / rsDen;
    float rpNum = (n * n + k2) * cosTheta * cosTheta - 2 * n * cosTheta + 1;
    float rpDen = (n * n + k2) * cosTheta * cosTheta + 2 * n * cosTheta + 1;
    float rp = rpNum / rpDen;
    return 0.5f * (rs + rp);
  }
  RGB Kr(float cosTheta) const override {
    RGB out;
    for (int k = 0; k < 3; k++) out[k] = fresnel(r[k], g[k], cosTheta);
    return out;
  }
};

// builtin/shader/bsdf/specular.go:13-101
static Vec3 reflectV(Vec3 omegaR, Vec3 N) { return Vec3Sub(Vec3Scale(2.0f * Vec3Dot(N, omegaR), N), omegaR); }
struct Specular : BSDF {
  float Lambda;
  Vec3 OmegaR;
  const Fresnel* fresnel;
  Vec3 U, V, N;
  Specular(float lambda, Vec3 omegaI, const Fresnel* f, Vec3 U_, Vec3 V_, Vec3 N_)
      : Lambda(lambda), OmegaR(Vec3BasisProject(U_, V_, N_, omegaI)), fresnel(f), U(U_), V(V_), N(N_) {}
  Vec3 Sample(double, double) override {
    Vec3 omegaO = reflectV(OmegaR, V3(0, 0, 1));
    omegaO = Vec3Normalize(omegaO);
    return Vec3BasisExpand(U, V, N, omegaO);
  }
  double PDF(Vec3 _omegaO) override {
    Vec3 omegaO = Vec3BasisProject(U, V, N, _omegaO);
    Vec3 omegaORefl = reflectV(OmegaR, V3(0, 0, 1));
    if (Vec3Dot(omegaO, omegaORefl) < 0.9999f) return 0;
    return 1;
  }
  Spectrum Eval(Vec3 _omegaO) override {
    Spectrum rho;
    Vec3 omegaO = Vec3BasisProject(U, V, N, _omegaO);
    Vec3 omegaORefl = reflectV(OmegaR, V3(0, 0, 1));
    if (Vec3Dot(omegaO, omegaORefl) < 0.9999f) return rho;
    RGB fr = fresnel->Kr(OmegaR[2]);
    rho.Lambda = Lambda;
    rho.FromRGB(fr);
    rho.Scale(Vec3DotAbs(omegaO, V3(0, 0, 1)));
    return rho;
  }
};

// builtin/shader/bsdf/microfacetggx.go:16-146
static inline float chi(float x) { return x > 0.0f ? 1.0f : 0.0f; }
static inline float signGGX(float v) { return v < 0 ? -1.0f : 1.0f; }
static float ggxSmithG1(Vec3 omega, Vec3 omegaM, float alpha) {
  float ODotN = Vec3Dot(omega, omegaM);
  float denom = 1 + Sqrt(1 + (alpha * alpha) * ((1.0f / (omega[2] * omega[2])) - 1));
  return chi(ODotN / omega[2]) * 2 / denom;
}
static float ggxD(Vec3 omegaM, float alpha) {
  float numer = alpha * alpha * chi(omegaM[2]);
  if (omegaM[2] == 1.0f) return 1.0f / (kPi * alpha * alpha);
  float denom = kPi * sqr(omegaM[2] * omegaM[2]) * sqr(alpha * alpha + ((1.0f / (omegaM[2] * omegaM[2])) - 1));
  return numer / denom;
}
struct MicrofacetGGX : BSDF {
  float Lambda;
  Vec3 OmegaR;
  float Roughness;  // roughness^2 (NewMicrofacetGGX, :87); Sample/PDF/Eval square it again (alpha = roughness^4)
  const Fresnel* fresnel;
  Vec3 U, V, N;
  MicrofacetGGX(float lambda, Vec3 omegaI, const Fresnel* f, float roughness, Vec3 U_, Vec3 V_, Vec3 N_)
      : Lambda(lambda), OmegaR(Vec3BasisProject(U_, V_, N_, omegaI)), Roughness(roughness * roughness), fresnel(f), U(U_), V(V_), N(N_) {}
  Vec3 Sample(double r0, double r1) override {
    float alpha = sqr(Roughness);
    double thetaM = std::atan2((double)alpha * std::sqrt(r0), std::sqrt(1 - r0));
    double phiM = 2.0 * M_PI * r1;
    Vec3 omegaM = V3(Sin((float)thetaM) * Cos((float)phiM), Sin((float)thetaM) * Sin((float)phiM), Cos((float)thetaM));
    Vec3 omegaO = Vec3Sub(Vec3Scale(2.0f * Vec3DotAbs(omegaM, OmegaR), omegaM), OmegaR);
    return Vec3BasisExpand(U, V, N, Vec3Normalize(omegaO));
  }
  double PDF(Vec3 _omegaO) override {
    Vec3 omegaO = Vec3BasisProject(U, V, N, _omegaO);
    float alpha = sqr(Roughness);
    Vec3 omegaM = Vec3Scale(signGGX(OmegaR[2]), Vec3Normalize(Vec3Add(OmegaR, omegaO)));
    double pdf = (double)(ggxD(omegaM, alpha) * omegaM[2]);
    if (std::isnan(pdf)) return 0;
    return pdf;
  }
  Spectrum Eval(Vec3 _omegaO) override {
    Spectrum rho;
    Vec3 omegaI = Vec3BasisProject(U, V, N, _omegaO);
    float alpha = sqr(Roughness);
    Vec3 h = Vec3Scale(signGGX(OmegaR[2]), Vec3Normalize(Vec3Add(OmegaR, omegaI)));
    RGB fr = fresnel->Kr(Vec3DotAbs(OmegaR, h));
    float numer = ggxSmithG1(OmegaR, h, alpha) * ggxSmithG1(omegaI, h, alpha) * ggxD(h, alpha);
    float denom = 4 * Abs(OmegaR[2]) * Abs(omegaI[2]);
    rho.Lambda = Lambda;
    rho.FromRGB(fr);
    rho.Scale(Abs(omegaI[2]) * numer / denom);
    for (int k = 0; k < 4; k++)
      if (rho.C[k] < 0 || std::isnan((double)rho.C[k])) rho.C[k] = 0;
    return rho;
  }
};

// ---------------------------------------------------------------------------------------------
// builtin/shader/std.go:77-296
// builtin/shader/debug.go:42-49
void DebugShader::Eval(ShaderContext* sg) { sg->OutRGB = Colour; }
RGB DebugShader::EvalEmission(ShaderContext*, Vec3) { return RGB{}; }

// param.RGBUniform / Float32Uniform (core/param/param.go): a constant map, or builtin/maps/texture.go:22-46 when the .vnf
// value was a file name
static TexCoord tex_coord(const ShaderContext* sg) {
  TexCoord t;
  t.U = sg->U; t.V = sg->V;
  t.Dduvdx[0] = sg->Dduvdx[0]; t.Dduvdx[1] = sg->Dduvdx[1];
  t.Dduvdy[0] = sg->Dduvdy[0]; t.Dduvdy[1] = sg->Dduvdy[1];
  t.PixelDelta[0] = sg->PixelDelta[0]; t.PixelDelta[1] = sg->PixelDelta[1];
  return t;
}
RGB ShaderStd::rgb(int slot, const RGB& constant, const ShaderContext* sg) const {
  if (!tex[slot].tex) return constant;
  if (!sg) throw std::runtime_error("Shader " + Name + ": a texture map on a light's emission is out of scope (the light passes its own lsg)");
  float c[3];
  tex[slot].Sample(tex_coord(sg), c);
  return MakeRGB(c[0], c[1], c[2]);
}
float ShaderStd::f32(int slot, float constant, const ShaderContext* sg) const {
  if (!tex[slot].tex) return constant;
  if (!sg) throw std::runtime_error("Shader " + Name + ": a texture map on a light's emission is out of scope (the light passes its own lsg)");
  float c[3];
  tex[slot].Sample(tex_coord(sg), c);
  return c[tex[slot].Chan];
}

void ShaderStd::Eval(ShaderContext* sg) {
  if (sg->Level > 3) return;

  Vec3 V = Vec3Cross(sg->N, sg->DdPdu);
  if (Vec3Length2(V) < 0.1f) V = Vec3Cross(sg->N, sg->DdPdv);
  V = Vec3Normalize(V);
  Vec3 U = Vec3Normalize(Vec3Cross(sg->N, V));

  float diffRoughness = 0.5f;
  if (hasDiffuseRoughness) diffRoughness = f32(kDiffuseRoughness, DiffuseRoughness, sg);
  OrenNayar diffBrdf(sg->Lambda, Vec3Neg(sg->Rd), diffRoughness, U, V, sg->N);

  RGB diffContrib;
  RGB diffColour;
  if (hasDiffuseColour) diffColour = rgb(kDiffuseColour, DiffuseColour, sg);

  float diffWeight = 0, spec1Weight = 0;
  if (hasDiffuseStrength) diffWeight = f32(kDiffuseStrength, DiffuseStrength, sg);
  if (hasSpec1Strength) spec1Weight = f32(kSpec1Strength, Spec1Strength, sg);
  float totalWeight = diffWeight + spec1Weight;
  diffWeight /= totalWeight;
  spec1Weight /= totalWeight;
  if (totalWeight == 0.0f) throw std::runtime_error("Shader " + Name + " has no weight");

  if (diffWeight > 0.0f) {
    sg->LightsPrepare();
    while (sg->NextLight()) {
      if (sg->Lp->DiffuseShadeMult() > 0.0f) {
        RGB col = sg->EvaluateLightSamples(&diffBrdf);
        col.Mul(diffColour);
        diffContrib.Add(col);
      }
    }
    diffContrib.Scale(diffWeight);
  }

  float ior = 1.7f;
  if (hasIOR) ior = f32(kIOR, IOR, sg);

  // std.go:172-192
  Dielectric dielectric(ior);
  RGB refl = MakeRGB(0.5f, 0.5f, 0.5f), edge = MakeRGB(0.5f, 0.5f, 0.5f);
  if (hasSpec1FresnelRefl) refl = rgb(kSpec1FresnelRefl, Spec1FresnelRefl, sg);
  if (hasSpec1FresnelEdge) refl = rgb(kSpec1FresnelEdge, Spec1FresnelEdge, sg);  // sic: std.go:187-189 assigns the edge tint to `refl`
  Conductor conductor(refl, edge);
  const Fresnel* fresnel = spec1FresnelModel == 1 ? static_cast<const Fresnel*>(&conductor) : static_cast<const Fresnel*>(&dielectric);

  RGB spec1Contrib;
  if (spec1Weight > 0.0f) {
    float spec1Roughness = 0.5f;
    if (hasSpec1Roughness) spec1Roughness = f32(kSpec1Roughness, Spec1Roughness, sg);
    RGB spec1Colour;
    if (hasSpec1Colour) spec1Colour = rgb(kSpec1Colour, Spec1Colour, sg);
    Specular specBRDF(sg->Lambda, Vec3Neg(sg->Rd), fresnel, U, V, sg->N);
    MicrofacetGGX ggxBRDF(sg->Lambda, Vec3Neg(sg->Rd), fresnel, spec1Roughness, U, V, sg->N);
    BSDF& spec1BRDF = spec1Roughness == 0.0f ? static_cast<BSDF&>(specBRDF) : static_cast<BSDF&>(ggxBRDF);

    TraceSample samp;
    Ray ray;
    ray.Task = sg->task;
    int spec1Samples = 0;
    if (spec1Roughness == 0.0f) spec1Samples = 1;
    for (int i = 0; i < spec1Samples; i++) {
      uint64_t idx = (uint64_t)(sg->I * spec1Samples + i);
      double r0 = VanDerCorput(idx, sg->Scramble[0]);
      double r1 = Sobol(idx, sg->Scramble[1]);
      Vec3 spec1OmegaO = spec1BRDF.Sample(r0, r1);
      double pdf = spec1BRDF.PDF(spec1OmegaO);
      if (Vec3Dot(spec1OmegaO, sg->Ng) <= 0.0f) continue;
      ray.Init(RayTypeReflected, sg->OffsetP(1), spec1OmegaO, kInfPos, (uint8_t)(sg->Level + 1), sg);
      bool traced;
      if (sg->Level + 1 > 3 && !sg->task->trace_last_level) {
        traced = false;  // optional: skip the level-4 ray whose shader returns black (std.go:95)
      } else {
        traced = Trace(&ray, &samp);
      }
      if (traced) {
        Spectrum rho = spec1BRDF.Eval(spec1OmegaO);
        rho.Scale(1.0f / (float)pdf);
        RGB col = rho.ToRGB();
        col.Mul(spec1Colour);
        col.Mul(samp.Colour);
        for (int k = 0; k < 3; k++)
          if (col[k] < 0 || std::isnan((double)col[k])) col[k] = 0;
        spec1Contrib.Add(col);
      }
    }
    if (spec1Samples > 0) spec1Contrib.Scale(spec1Weight / (float)spec1Samples);

    if (spec1Roughness > 0.0f) {  // std.go:269-284: glossy lobe = direct light only, NOT scaled by spec1Weight
      sg->LightsPrepare();
      while (sg->NextLight()) {
        RGB col = sg->EvaluateLightSamples(&spec1BRDF);
        col.Mul(spec1Colour);
        spec1Contrib.Add(col);
      }
    }
  }

  RGB contrib;
  RGB emissContrib = EvalEmission(sg, Vec3Neg(sg->Rd));
  contrib.Add(emissContrib);
  contrib.Add(diffContrib);
  contrib.Add(spec1Contrib);
  sg->OutRGB = contrib;
}

// builtin/shader/std.go:299-316
RGB ShaderStd::EvalEmission(ShaderContext* sg, Vec3) {
  RGB emissColour;
  float emissStrength = 0;
  if (hasEmissionColour) emissColour = rgb(kEmissionColour, EmissionColour, sg);
  if (hasEmissionStrength) emissStrength = f32(kEmissionStrength, EmissionStrength, sg);
  else return RGB{};
  emissColour.Scale(emissStrength);
  return emissColour;
}

// ---------------------------------------------------------------------------------------------
// builtin/light/triangle.go:79-134
static bool rayTriangleIntersect(Vec3 Ro, Vec3 Rd, Vec3 P0, Vec3 P1, Vec3 P2, Vec3* pout, float* tout) {
  Vec3 e1 = Vec3Sub(P1, P0);
  Vec3 e2 = Vec3Sub(P2, P0);
  Vec3 P = Vec3Cross(Rd, e2);
  float det = Vec3Dot(e1, P);
  if (det > -1e-6f && det < 1e-6f) return false;
  float inv_det = 1 / det;
  Vec3 T = Vec3Sub(Ro, P0);
  float u = Vec3Dot(T, P) * inv_det;
  if (u < 0 || u > 1) return false;
  Vec3 Q = Vec3Cross(T, e1);
  float v = Vec3Dot(Rd, Q) * inv_det;
  if (v < 0 || u + v > 1) return false;
  float t = Vec3Dot(e2, Q) * inv_det;
  if (t > 1e-6f) {
    *pout = Vec3Add3(Vec3Scale(1 - u - v, P0), Vec3Scale(u, P1), Vec3Scale(v, P2));
    *tout = t;
    return true;
  }
  return false;
}

static float triangleArea(Vec3 P0, Vec3 P1, Vec3 P2) { return 0.5f * Vec3Length(Vec3Cross(Vec3Sub(P1, P0), Vec3Sub(P2, P0))); }

// builtin/light/disk.go:38-51
static float rayPlaneIntersect(Vec3 Ro, Vec3 Rd, Vec3 P, Vec3 N) {
  float denom = Vec3Dot(N, Rd);
  if (Abs(denom) > 1e-6f) {
    Vec3 p0l0 = Vec3Sub(P, Ro);
    return Vec3Dot(p0l0, N) / denom;
  }
  return 0;
}

// builtin/light/triangle.go:376-417
static void triangleSidesToAngles(float as, float bs, float cs, float* a, float* b, float* c) {
  float ssu = (as + bs + cs) / 2;
  float sinssuas = Sin(ssu - as);
  float sinssubs = Sin(ssu - bs);
  float sinssucs = Sin(ssu - cs);
  float sinssu = Sin(ssu);
  float tanA2 = Sqrt(sinssubs * sinssucs / (sinssu * sinssuas));
  float tanB2 = Sqrt(sinssuas * sinssucs / (sinssu * sinssubs));
  float tanC2 = Sqrt(sinssuas * sinssubs / (sinssu * sinssucs));
  *a = 2 * Atan(tanA2);
  *b = 2 * Atan(tanB2);
  *c = 2 * Atan(tanC2);
}
// builtin/light/triangle.go:430-459
static void triangleVerticesToSides(Vec3 pa, Vec3 pb, Vec3 pc, float* a, float* b, float* c) {
  float adot = Vec3Dot(pb, pc);
  float bdot = Vec3Dot(pc, pa);
  float cdot = Vec3Dot(pa, pb);
  *a = Acos(adot);
  *b = Acos(bdot);
  *c = Acos(cdot);
}

// builtin/light/triangle.go:474-535
static Vec3 sampleSphericalTriangle(Vec3 p0, Vec3 p1, Vec3 p2, Vec3 p, double r0, double r1, double* pdf) {
  Vec3 pa = Vec3Normalize(Vec3Sub(p0, p));
  Vec3 pb = Vec3Normalize(Vec3Sub(p1, p));
  Vec3 pc = Vec3Normalize(Vec3Sub(p2, p));
  float a, b, c;
  triangleVerticesToSides(pa, pb, pc, &a, &b, &c);
  float alpha, beta, gamma;
  triangleSidesToAngles(a, b, c, &alpha, &beta, &gamma);
  float area = alpha + beta + gamma - kPi;
  float areaHat = (float)r0 * area;
  float s = Sin(areaHat - alpha), t = Cos(areaHat - alpha);  // m.Sincos
  float sinAlpha = Sin(alpha), cosAlpha = Cos(alpha);
  float u = t - cosAlpha;
  float v = s + sinAlpha * Cos(c);
  float q = ((v * t - u * s) * cosAlpha - v) / ((v * s + u * t) * sinAlpha);
  q = Max(-1, Min(q, 1));
  float w = Vec3Dot(pc, pa);
  Vec3 v31;
  for (int k = 0; k < 3; k++) v31[k] = pc[k] - w * pa[k];
  v31 = Vec3Normalize(v31);
  Vec3 v4;
  for (int k = 0; k < 3; k++) v4[k] = q * pa[k] + Sqrt(1 - q * q) * v31[k];
  float z = 1 - (float)r1 * (1 - Vec3Dot(v4, pb));
  w = Vec3Dot(v4, pb);
  Vec3 v42;
  for (int k = 0; k < 3; k++) v42[k] = v4[k] - w * pb[k];
  v42 = Vec3Normalize(v42);
  Vec3 x = Vec3Add(Vec3Scale(z, pb), Vec3Scale(Sqrt(1 - z * z), v42));
  *pdf = 1 / (double)area;
  return x;
}

// builtin/light/triangle.go:537-565 (UVs omitted)
PolyMesh* Tri::createMesh() {
  PolyMesh* msh = new PolyMesh();
  msh->Name = Name + ":<mesh>";
  msh->shader.push_back(shader);
  msh->Verts.MotionKeys = 1;
  msh->Normals.MotionKeys = 1;
  Vec3 N = Vec3Normalize(Vec3Cross(Vec3Sub(P1, P0), Vec3Sub(P2, P0)));
  msh->Verts.Elems = {P0, P1, P2};
  msh->Verts.ElemsPerKey = 3;
  msh->Normals.Elems = {N, N, N};
  msh->Normals.ElemsPerKey = 3;
  return msh;
}

// builtin/light/triangle.go:136-230
bool Tri::ValidSample(ShaderContext* sg, BSDFSample* sample) {
  Vec3 P;
  float tdummy;
  if (!rayTriangleIntersect(sg->P, sample->D, P0, P1, P2, &P, &tdummy)) return false;

  if (Vec3Dot(sg->Ng, Vec3Sub(P0, P)) < 0 || Vec3Dot(sg->Ng, Vec3Sub(P1, P)) < 0 || Vec3Dot(sg->Ng, Vec3Sub(P2, P)) < 0) {
    double pdf = (double)(1 / triangleArea(P0, P1, P2));
    Vec3 D = Vec3Sub(P, sg->P);
    sample->Ldist = Vec3Length(D);
    sample->Ld = Vec3Normalize(D);
    Vec3 N = Vec3Normalize(Vec3Cross(Vec3Sub(P1, P0), Vec3Sub(P2, P0)));
    if (Vec3Dot(sample->Ld, N) > 0 || Vec3Dot(sample->Ld, sg->Ng) < 0) return false;
    sample->Liu.Lambda = sg->Lambda;
    RGB E = shader->EvalEmission(nullptr, Vec3Neg(sample->Ld));
    sample->Liu.FromRGB(E);
    sample->PdfLight = (float)pdf * sqr(sample->Ldist) / Vec3DotAbs(sample->Ld, N);
  } else {
    Vec3 pa = Vec3Normalize(Vec3Sub(P0, sg->P));
    Vec3 pb = Vec3Normalize(Vec3Sub(P1, sg->P));
    Vec3 pc = Vec3Normalize(Vec3Sub(P2, sg->P));
    float a, b, c;
    triangleVerticesToSides(pa, pb, pc, &a, &b, &c);
    float alpha, beta, gamma;
    triangleSidesToAngles(a, b, c, &alpha, &beta, &gamma);
    float area = alpha + beta + gamma - kPi;
    double pdf = (double)(1 / area);
    Vec3 D = Vec3Sub(P, sg->P);
    sample->Ldist = Vec3Length(D);
    sample->Ld = Vec3Normalize(D);
    Vec3 N = Vec3Normalize(Vec3Cross(Vec3Sub(P1, P0), Vec3Sub(P2, P0)));
    if (Vec3Dot(sample->Ld, N) > 0 || Vec3Dot(sample->Ld, sg->Ng) < 0) return false;
    sample->Liu.Lambda = sg->Lambda;
    RGB E = shader->EvalEmission(nullptr, Vec3Neg(sample->Ld));
    sample->Liu.FromRGB(E);
    sample->PdfLight = (float)pdf;
  }
  return true;
}

// builtin/light/triangle.go:232-282
void Tri::sampleByArea(ShaderContext* sg, int n) {
  double pdf = (double)(1 / triangleArea(P0, P1, P2));
  Vec3 N = Vec3Normalize(Vec3Cross(Vec3Sub(P1, P0), Vec3Sub(P2, P0)));
  for (int i = 0; i < n; i++) {
    uint64_t idx = (uint64_t)(sg->I * n + i);
    double r0 = VanDerCorput(idx, sg->Scramble[0]);
    double r1 = Sobol(idx, sg->Scramble[1]);
    Vec3 P = Vec3Add3(P0, Vec3Scale((float)(r1 * std::sqrt(1 - r0)), Vec3Sub(P1, P0)), Vec3Scale((float)(1 - std::sqrt(1 - r0)), Vec3Sub(P2, P0)));
    Vec3 D = Vec3Sub(P, sg->P);
    LightSample ls{};
    ls.Ldist = Vec3Length(D);
    ls.Ld = Vec3Normalize(D);
    if (Vec3Dot(ls.Ld, N) > 0 || Vec3Dot(ls.Ld, sg->Ng) < 0) continue;
    ls.Liu.Lambda = sg->Lambda;
    RGB E = shader->EvalEmission(nullptr, Vec3Neg(ls.Ld));
    ls.Liu.FromRGB(E);
    ls.Pdf = (float)pdf * sqr(ls.Ldist) / Vec3DotAbs(ls.Ld, N);
    ls.P = P;
    sg->Lsamples.push_back(ls);
  }
}

// builtin/light/triangle.go:285-343
void Tri::SampleArea(ShaderContext* sg, int n) {
  if (Vec3Dot(sg->Ng, Vec3Sub(P0, sg->P)) < 0 || Vec3Dot(sg->Ng, Vec3Sub(P1, sg->P)) < 0 || Vec3Dot(sg->Ng, Vec3Sub(P2, sg->P)) < 0) {
    sampleByArea(sg, n);
    return;
  }
  for (int i = 0; i < n; i++) {
    uint64_t idx = (uint64_t)(sg->I * n + i);
    double r0 = VanDerCorput(idx, sg->Scramble[0]);
    double r1 = Sobol(idx, sg->Scramble[1]);
    double pdf;
    Vec3 x = sampleSphericalTriangle(P0, P1, P2, sg->P, r0, r1, &pdf);
    Vec3 N = Vec3Normalize(Vec3Cross(Vec3Sub(P1, P0), Vec3Sub(P2, P0)));
    float t = rayPlaneIntersect(sg->P, x, P0, N);
    Vec3 P = Vec3Mad(sg->P, x, t);
    Vec3 D = Vec3Sub(P, sg->P);
    LightSample ls{};
    ls.Ldist = Vec3Length(D);
    ls.Ld = Vec3Normalize(D);
    if (Vec3Dot(ls.Ld, N) > 0 || Vec3Dot(ls.Ld, sg->Ng) < 0) continue;
    ls.Liu.Lambda = sg->Lambda;
    RGB E = shader->EvalEmission(nullptr, Vec3Neg(ls.Ld));
    ls.Liu.FromRGB(E);
    ls.Pdf = (float)pdf;
    ls.P = P;
    sg->Lsamples.push_back(ls);
  }
}

// ---------------------------------------------------------------------------------------------
// builtin/light/disk.go:38-68
static bool rayPlaneIntersectOk(Vec3 Ro, Vec3 Rd, Vec3 P, Vec3 N, float* t) {
  float denom = Vec3Dot(N, Rd);
  if (Abs(denom) > 1e-6f) {
    Vec3 p0l0 = Vec3Sub(P, Ro);
    *t = Vec3Dot(p0l0, N) / denom;
    return *t >= 0;
  }
  *t = 0;
  return false;
}
static bool rayDiskIntersect(Vec3 Ro, Vec3 Rd, Vec3 P, Vec3 N, float radius, float* t) {
  if (rayPlaneIntersectOk(Ro, Rd, P, N, t)) {
    Vec3 p = Vec3Mad(Ro, Rd, *t);
    Vec3 v = Vec3Sub(p, P);
    float d2 = Vec3Dot(v, v);
    return Sqrt(d2) <= radius;
  }
  *t = 0;
  return false;
}

// builtin/light/disk.go:90-97
void Disk::PreRender() {
  N = Vec3Normalize(Vec3Sub(LookAt, P));
  T = Vec3Normalize(Vec3Cross(N, Up));
  B = Vec3Cross(N, T);
}

// builtin/light/disk.go:224-262 (UVs omitted): a fan of `Segments` unindexed triangles
PolyMesh* Disk::createMesh() {
  PolyMesh* msh = new PolyMesh();
  msh->Name = Name + ":<mesh>";
  msh->shader.push_back(shader);
  int nv = Segments;
  float dv = 2 * kPi / (float)nv;
  float ang = 0;
  msh->Verts.MotionKeys = 1;
  msh->Normals.MotionKeys = 1;
  for (int i = 0; i < nv; i++) {
    msh->Verts.Elems.push_back(P);
    msh->Verts.Elems.push_back(Vec3Add(P, Vec3Add(Vec3Scale(Radius * Cos(ang), B), Vec3Scale(Radius * Sin(ang), T))));
    msh->Verts.Elems.push_back(Vec3Add(P, Vec3Add(Vec3Scale(Radius * Cos(ang + dv), B), Vec3Scale(Radius * Sin(ang + dv), T))));
    msh->Verts.ElemsPerKey += 3;
    msh->Normals.Elems.push_back(N);
    msh->Normals.Elems.push_back(N);
    msh->Normals.Elems.push_back(N);
    msh->Normals.ElemsPerKey += 3;
    ang += dv;
  }
  return msh;
}

// builtin/light/disk.go:123-169
bool Disk::ValidSample(ShaderContext* sg, BSDFSample* sample) {
  double pdf = (double)(1.0f / (kPi * Radius * Radius));
  float t;
  if (!rayDiskIntersect(sg->P, sample->D, P, N, Radius, &t)) return false;
  Vec3 p = Vec3Mad(sg->P, sample->D, t);
  Vec3 V = Vec3Sub(p, sg->P);
  if (Vec3Dot(V, sg->Ng) <= 0.0f || Vec3Dot(V, N) >= 0.0f) return false;
  sample->Ldist = Vec3Length(V);
  sample->Ld = Vec3Normalize(V);
  sample->Liu.Lambda = sg->Lambda;
  RGB E = shader->EvalEmission(nullptr, Vec3Neg(sample->Ld));
  sample->Liu.FromRGB(E);
  sample->PdfLight = (float)pdf * (sample->Ldist * sample->Ldist) / (Abs(Vec3Dot(sample->Ld, N)));
  return true;
}

// builtin/light/disk.go:172-217
void Disk::SampleArea(ShaderContext* sg, int n) {
  for (int i = 0; i < n; i++) {
    uint64_t idx = (uint64_t)(sg->I * n + i);
    double r0 = VanDerCorput(idx, sg->Scramble[0]);
    double r1 = Sobol(idx, sg->Scramble[1]);
    float u = Radius * Sqrt((float)r0) * Cos(2 * kPi * (float)r1);
    float v = Radius * Sqrt((float)r0) * Sin(2 * kPi * (float)r1);
    double pdf = (double)(1.0f / (kPi * Radius * Radius));
    Vec3 Pl = Vec3Add3(P, Vec3Scale(u, B), Vec3Scale(v, T));
    Vec3 V = Vec3Sub(Pl, sg->P);
    LightSample ls{};
    if (Vec3Dot(V, sg->Ng) > 0.0f && Vec3Dot(V, N) < 0.0f) {
      ls.Ldist = Vec3Length(V);
      ls.Ld = Vec3Normalize(V);
      ls.Liu.Lambda = sg->Lambda;
      RGB E = shader->EvalEmission(nullptr, Vec3Neg(ls.Ld));
      ls.Liu.FromRGB(E);
      ls.Pdf = (float)pdf * (ls.Ldist * ls.Ldist) / Abs(Vec3Dot(ls.Ld, N));
      sg->Lsamples.push_back(ls);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// builtin/geom/sphere/trace.go:53-109 == builtin/light/sphere.go:69-127
static bool solveQuadratic(float a, float b, float c, float* x0, float* x1) {
  float discr = b * b - 4 * a * c;
  if (discr < 0) return false;
  if (discr == 0) {
    *x1 = -0.5f * b / a;
    *x0 = *x1;
  } else {
    float q;
    if (b > 0) q = -0.5f * (b + Sqrt(discr));
    else q = -0.5f * (b - Sqrt(discr));
    *x0 = q / a;
    *x1 = c / q;
  }
  if (*x0 > *x1) std::swap(*x0, *x1);
  return true;
}
static bool raySphereIntersect(Vec3 Ro, Vec3 Rd, Vec3 P, float radius, float* t) {
  Vec3 L = Vec3Sub(Ro, P);
  float a = Vec3Dot(Rd, Rd);
  float b = 2 * Vec3Dot(Rd, L);
  float c = Vec3Dot(L, L) - radius * radius;
  float t0 = 0, t1 = 0;
  if (!solveQuadratic(a, b, c, &t0, &t1)) return false;
  if (t0 > t1) std::swap(t0, t1);
  if (t0 < 0) {
    t0 = t1;
    if (t0 < 0) return false;
  }
  *t = t0;
  return true;
}

// builtin/geom/sphere/trace.go:13-51 (U,V surface parameters omitted: every in-scope shader parameter is a Constant map)
bool SphereGeom::Trace(Ray* ray, ShaderContext* sg) {
  float t;
  if (raySphereIntersect(ray->P, ray->D, P, Radius, &t)) {
    if (t < ray->Tclosest) {
      ray->Tclosest = t;
      Vec3 Ph = Vec3Mad(ray->P, ray->D, t);
      Vec3 N = Vec3Normalize(Vec3Sub(Ph, P));
      sg->Poffset = Vec3Scale(0.001f, N);
      sg->Ng = N;
      sg->N = N;
      sg->Bu = 0;
      sg->Bv = 0;
      sg->Bw = 0;
      sg->P = Ph;
      sg->Po = sg->P;
      sg->DdPdu = V3(1, 0, 0);
      sg->DdPdv = V3(0, 0, 1);
      sg->shader = shader;
      sg->ElemID = 0;  // the reference leaves ElemID untouched (stale); the oracle and the device both report 0
      return true;
    }
  }
  return false;
}
// builtin/geom/sphere/sphere.go:66-76
BoundingBox SphereGeom::Bounds(float) const {
  BoundingBox b;
  for (int k = 0; k < 3; k++) {
    b.b[0][k] = P[k] - Radius;
    b.b[1][k] = P[k] + Radius;
  }
  return b;
}

// builtin/light/sphere.go:132-183
bool SphereLight::ValidSample(ShaderContext* sg, BSDFSample* sample) {
  float t;
  if (!raySphereIntersect(sg->P, sample->D, P, Radius, &t)) return false;
  Vec3 x = Vec3Mad(sg->P, sample->D, t);
  Vec3 D = Vec3Sub(x, sg->P);
  sample->Ldist = Vec3Length(D);
  sample->Ld = Vec3Normalize(D);
  sample->Liu.Lambda = sg->Lambda;
  RGB E = shader->EvalEmission(nullptr, Vec3Neg(sample->D));
  sample->Liu.FromRGB(E);
  Vec3 V = Vec3Sub(P, sg->P);
  float l = Vec3Length(V);
  sample->PdfLight = 1 / (2 * kPi * (1 - Sqrt(1 - sqr(Radius / l))));
  return true;
}

// builtin/light/sphere.go:186-270 (sg.Sample is never set on this path: 0)
void SphereLight::SampleArea(ShaderContext* sg, int n) {
  Vec3 V = Vec3Sub(P, sg->P);
  float l = Vec3Length(V);
  Vec3 w = Vec3Normalize(V);
  Vec3 v = Vec3Normalize(Vec3Cross(w, sg->Ng));
  Vec3 u = Vec3Cross(w, v);
  for (int i = 0; i < n; i++) {
    uint64_t idx = (uint64_t)(sg->I * n + 0 + i);
    double r0 = VanDerCorput(idx, sg->Scramble[0]);
    double r1 = Sobol(idx, sg->Scramble[1]);
    float theta = Acos(1 - (float)r0 + (float)r0 * Sqrt(1 - sqr(Radius / l)));
    float phi = 2 * kPi * (float)r1;
    Vec3 a = V3(Cos(phi) * Sin(theta), Sin(phi) * Sin(theta), Cos(theta));
    Vec3 omega = Vec3BasisExpand(u, v, w, a);
    if (Vec3Dot(omega, sg->Ng) < 0) continue;
    float t;
    if (!raySphereIntersect(sg->P, omega, P, Radius, &t)) continue;
    Vec3 x = Vec3Mad(sg->P, omega, t);
    Vec3 D = Vec3Sub(x, sg->P);
    LightSample ls{};
    ls.Ldist = Vec3Length(D);
    ls.Ld = Vec3Normalize(D);
    ls.Liu.Lambda = sg->Lambda;
    RGB E = shader->EvalEmission(nullptr, Vec3Neg(omega));
    ls.Liu.FromRGB(E);
    ls.Pdf = 1 / (2 * kPi * (1 - Sqrt(1 - sqr(Radius / l))));
    sg->Lsamples.push_back(ls);
  }
}

// ---------------------------------------------------------------------------------------------
static inline float degToRad(float deg) { return deg * kPi / 180.0f; }

// builtin/camera/camera.go:80-98,109-193 (single key: the else branch :154-187 with i=0)
void Camera::PreRender(float frameAspect) {
  if (Aspect == 0.0f) Aspect = frameAspect;
  TanThetaFocal = Tan(degToRad(Fov / 2)) * Focal;
  if (FromKeys.empty()) FromKeys.push_back(From);
  if (ToKeys.empty()) ToKeys.push_back(To);
  if (RollKeys.empty()) RollKeys.push_back(Roll);
  LocalToWorld.clear();
  decomp.clear();
  if (Type == "LookAt") {
    // camera.go:109-193 (calcLookatMatrices): one matrix per key of whichever of From / To has more keys; the other one and
    // Roll are interpolated at time = i / keys (NOT i / (keys-1): kept)
    const int nF = (int)FromKeys.size(), nT = (int)ToKeys.size(), nR = (int)RollKeys.size();
    const bool byTarget = nT > nF;
    const int n = byTarget ? nT : nF;
    for (int i = 0; i < n; i++) {
      float time = (float)i / (float)n;
      Vec3 eye, tgt;
      {
        const std::vector<Vec3>& other = byTarget ? FromKeys : ToKeys;
        float k = time * (float)((int)other.size() - 1);
        float t = k - Floor(k);
        int key = (int)Floor(k), key2 = (int)Ceil(k);
        Vec3 P = Vec3Lerp(other[key], other[key2], t);
        if (byTarget) { eye = P; tgt = ToKeys[i]; } else { eye = FromKeys[i]; tgt = P; }
      }
      Vec3 W = Vec3Normalize(Vec3Sub(eye, tgt));  // W points away from the target
      Vec3 u = Vec3Normalize(Vec3Cross(Up, W));
      Vec3 v = Vec3Normalize(Vec3Cross(u, W));
      float roll = 0;
      {
        float k = time * (float)(nR - 1);
        float t = k - Floor(k);
        int key = (int)Floor(k), key2 = (int)Ceil(k);
        roll = (1 - t) * RollKeys[key] + t * RollKeys[key2];
      }
      Vec3 U = Vec3Add(Vec3Scale(Cos(roll), u), Vec3Scale(Sin(roll), v));
      Vec3 V = Vec3Add(Vec3Scale(-Sin(roll), u), Vec3Scale(Cos(roll), v));
      LocalToWorld.push_back(Matrix4Mul(Matrix4Translate(eye[0], eye[1], eye[2]), Matrix4Basis(U, V, W)));
    }
  } else {
    // camera.go:205-216 (matrixCalc)
    for (const Matrix4& w2l : WorldToLocal) {
      Matrix4 inv;
      Matrix4Inverse(w2l, &inv);
      LocalToWorld.push_back(inv);
    }
  }
  for (const Matrix4& m : LocalToWorld) decomp.push_back(TransformDecompMatrix4(m));
  M = MatrixAt(0.0f);
}

// camera.go:225-236
Matrix4 Camera::MatrixAt(float time) const {
  if (decomp.empty()) return Matrix4Identity();
  float k = time * (float)((int)decomp.size() - 1);
  float t = k - Floor(k);
  int key = (int)Floor(k), key2 = (int)Ceil(k);
  return TransformDecompToMatrix4(TransformDecompLerp(decomp[key], decomp[key2], t));
}

// builtin/camera/camera.go:221-323
void Camera::ComputeRay(float Sx, float Sy, double lensU, double lensV, const ShaderContext* sc, Ray* ray) const {
  float camu = Sx * TanThetaFocal;
  float camv = Sy * (TanThetaFocal / Aspect);
  Vec3 U = V3(1, 0, 0), V = V3(0, 1, 0), W = V3(0, 0, 1);
  Vec3 s = Vec3Sub(Vec3Add(Vec3Scale(camu, U), Vec3Scale(camv, V)), Vec3Scale(Focal, W));
  Vec3 D, P, d;
  const Matrix4 M = decomp.size() > 1 ? MatrixAt(sc->Time) : this->M;
  if (Radius > 0.0f) {
    float x, y;
    UniformDisk2D(Radius, (float)lensU, (float)lensV, &x, &y);
    Vec3 e = Vec3Add(Vec3Scale(x, U), Vec3Scale(y, V));
    d = Matrix4MulVec(M, Vec3Sub(s, e));
    D = Vec3Normalize(d);
    P = Matrix4MulPoint(M, e);
  } else {
    d = Matrix4MulVec(M, s);
    D = Vec3Normalize(d);
    P = Matrix4MulPoint(M, V3(0, 0, 0));
  }
  // camera.go:300-306 (the finite-difference dx, dy of :268-295 are computed and then discarded by the reference)
  const Vec3 right = V3(M.m[0], M.m[1], M.m[2]);
  const Vec3 up = V3(M.m[4], M.m[5], M.m[6]);
  ray->DdPdx = V3(0, 0, 0);
  ray->DdPdy = V3(0, 0, 0);
  ray->DdDdx = Vec3Scale(1 / (Vec3Dot(d, d) * Sqrt(Vec3Dot(d, d))), Vec3Sub(Vec3Scale(Vec3Dot(d, d), right), Vec3Scale(Vec3Dot(d, right), d)));
  ray->DdDdy = Vec3Scale(1 / (Vec3Dot(d, d) * Sqrt(Vec3Dot(d, d))), Vec3Sub(Vec3Scale(Vec3Dot(d, d), up), Vec3Scale(Vec3Dot(d, up), d)));
  ray->Init(RayTypeCamera, P, D, kInfPos, 0, sc);
}

// camera.go:316-317: sc.Image.PixelDelta, the same two numbers on every call
void Camera::PixelDelta(int w, int h, float out[2]) const {
  out[0] = 2 * TanThetaFocal / (float)w;
  out[1] = 2 * TanThetaFocal / (Aspect * (float)h);
}

// ---------------------------------------------------------------------------------------------
// builtin/filter/filter.go:88-173
void FilterSampler::Create(int n_, double w_, double (*f)(double, double, void*), void* ctx) {
  n = n_;
  w = w_;
  std::vector<double> filter((size_t)n * n);
  double du = w / (double)(n - 1);
  double u = -w / 2;
  double F = 0;
  for (int j = 0; j < n; j++) {
    double dv = w / (double)(n - 1);
    double v = -w / 2;
    for (int i = 0; i < n; i++) {
      double fuv = f(u, v, ctx);
      filter[j + (i * n)] = fuv;
      F += fuv;
      v += dv;
    }
    u += du;
  }
  std::vector<std::vector<double>> pdf(n, std::vector<double>(n));
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) pdf[j][i] = filter[j + (i * n)] / F;
  std::vector<double> pV(n, 0.0);
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) pV[j] += pdf[j][i];
  std::vector<std::vector<double>> pVU(n, std::vector<double>(n));
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) pVU[j][i] = pdf[j][i] / pV[j];
  double p = 0;
  cdfV.assign(n, 0.0);
  for (int i = 0; i < n; i++) {
    p += pV[i];
    cdfV[i] = p;
  }
  cdfVU.assign(n, std::vector<double>(n));
  for (int j = 0; j < n; j++) {
    double q = 0;
    for (int i = 0; i < n; i++) {
      q += pVU[j][i];
      cdfVU[j][i] = q;
    }
  }
}

// builtin/filter/filter.go:40-86 (including the `du / w` of the first bin, kept as is)
void FilterSampler::WarpSample(double r0, double r1, double* uo, double* vo) const {
  double u = 0, v = 0;
  int uI = -1;
  for (int i = 0; i < (int)cdfV.size(); i++) {
    uI = i;
    if (r0 < cdfV[i]) {
      if (i == 0) {
        double du = r0 / cdfV[i];
        u = (-w / 2) + (du / w);
      } else {
        double du = (r0 - cdfV[i - 1]) / (cdfV[i] - cdfV[i - 1]);
        u = (-w / 2) + w * ((double)i + du) / (double)(n - 1);
      }
      break;
    }
  }
  const std::vector<double>& c = cdfVU[uI];
  for (int i = 0; i < (int)c.size(); i++) {
    if (r1 < c[i]) {
      if (i == 0) {
        double dv = r1 / c[i];
        v = (-w / 2) + (dv / w);
      } else {
        double dv = (r1 - c[i - 1]) / (c[i] - c[i - 1]);
        v = (-w / 2) + w * ((double)i + dv) / (double)(n - 1);
      }
      *uo = u;
      *vo = v;
      return;
    }
  }
  *uo = 0;
  *vo = 0;
}

// builtin/filter/airy.go:34-72
double BesselJ1(double x) {
  double ax = std::fabs(x);
  if (ax < 8.0) {
    double y = x * x;
    double ans1 = x * (72362614232.0 + y * (-7895059235.0 + y * (242396853.1 + y * (-2972611.439 + y * (15704.48260 + y * (-30.16036606))))));
    double ans2 = 144725228442.0 + y * (2300535178.0 + y * (18583304.74 + y * (99447.43394 + y * (376.9991397 + y * 1.0))));
    return ans1 / ans2;
  }
  double z = 8.0 / ax;
  double y = z * z;
  double xx = ax - 2.356194491;
  double ans1 = 1.0 + y * (0.183105e-2 + y * (-0.3516396496e-4 + y * (0.2457520174e-5 + y * (-0.240337019e-6))));
  double ans2 = 0.04687499995 + y * (-0.2002690873e-3 + y * (0.8449199096e-5 + y * (-0.88228987e-6 + y * 0.105787412e-6)));
  double ans = std::sqrt(0.636619772 / ax) * (std::cos(xx) * ans1 - z * std::sin(xx) * ans2);
  if (x < 0.0) ans = -ans;
  return ans;
}

static double sqrd(double x) { return x * x; }
// builtin/filter/airy.go:74-101
static double airyFn(double x, double y, void* ctx) {
  PixelFilter* f = (PixelFilter*)ctx;
  double q = std::sqrt(x * x + y * y);
  if (q > (double)(f->Width / 2)) return 0;
  double lambda = 550.0;
  double N = 5.6;
  double v = (20000 / (double)f->Width) * (M_PI * q) / (lambda * N);
  return (double)f->Peak * sqrd(2 * BesselJ1(v) / v);
}
// builtin/filter/gauss.go:31-50
static double gaussFn(double x, double y, void* ctx) {
  PixelFilter* f = (PixelFilter*)ctx;
  double q = std::sqrt(x * x + y * y);
  if (q > (double)(f->Width / 2)) return 0;
  double sigma = (double)(1.0 / std::sqrt((double)f->Width));
  return (1 / (2 * M_PI * sqrd(sigma))) * std::exp((double)(-(x * x + y * y) / 2 * sqrd(sigma)));
}
void PixelFilter::PreRender() { sampler.Create(Res, (double)Width, kind == 1 ? airyFn : gaussFn, this); }

// ---------------------------------------------------------------------------------------------
ShaderStd* Renderer::findShader(const std::string& name) {
  for (auto& s : shaders) if (s->Name == name) return s.get();
  return nullptr;
}

// core/core.go:36-61: nodes PreRender in creation order (lights append their meshes, which are
// pre-rendered in the next round), then scene.PreRender.
void Renderer::PreRender() {
  if (prerendered) return;
  framebuffer.assign((size_t)XRes * YRes * 3, 0.0f);
  camera.PreRender((float)XRes / (float)YRes);
  if (filter) filter->PreRender();
  int gid = 0;
  for (auto& m : meshes) {
    m->PreRender();
    m->id = gid++;
    scene.geoms.push_back(m.get());
  }
  for (auto& in : instances) {
    in->PreRender();
    in->id = gid++;
    scene.geoms.push_back(in.get());
  }
  // lights PreRender in creation order; each appends its geom (core.AddNode), which is pre-rendered in the next round
  const size_t firstLightGeom = meshes.size();
  std::vector<Geom*> lightGeoms;
  for (Light* l : lightOrder) {
    if (Tri* t = dynamic_cast<Tri*>(l)) {
      PolyMesh* g = t->createMesh();
      meshes.emplace_back(g);
      t->geom = g;
      lightGeoms.push_back(g);
    } else if (Disk* d = dynamic_cast<Disk*>(l)) {
      d->PreRender();
      PolyMesh* g = d->createMesh();
      meshes.emplace_back(g);
      d->geom = g;
      lightGeoms.push_back(g);
    } else if (SphereLight* sl = dynamic_cast<SphereLight*>(l)) {
      SphereGeom* g = new SphereGeom();
      g->Name = sl->Name + ":<sphere>";
      g->P = sl->P;
      g->Radius = sl->Radius;
      g->shader = sl->shader;
      sphereGeoms.emplace_back(g);
      sl->geom = g;
      lightGeoms.push_back(g);
    }
    scene.lights.push_back(l);
  }
  for (size_t i = firstLightGeom; i < meshes.size(); i++) meshes[i]->PreRender();
  for (Geom* g : lightGeoms) {
    g->id = gid++;
    scene.geoms.push_back(g);
  }
  scene.initAccel();
  prerendered = true;
}

// core/render.go:89-124
void Renderer::GenerateCameraRay(int iter, int x, int y, ShaderContext* sc, Ray* ray) const {
  int pixIdx = x + y * XRes;
  double rasterX, rasterY;
  RasterXY12((uint32_t)iter, (uint32_t)x, (uint32_t)y, 0, 0, &rasterX, &rasterY);
  const pixelscramble& scr = framescramble[pixIdx];
  double time = VanDerCorput((uint64_t)iter, scr.time);
  double lambda = (720 - 450) * VanDerCorput((uint64_t)iter, scr.lambda) + 450;
  double lensU = VanDerCorput((uint64_t)iter, scr.lensU);
  double lensV = Sobol((uint64_t)iter, scr.lensV);
  if (filter) {  // core/render.go:99-107
    double pixu = rasterX - std::floor(rasterX);
    double pixv = rasterY - std::floor(rasterY);
    double u, v;
    filter->sampler.WarpSample(pixu, pixv, &u, &v);
    rasterX = std::floor(rasterX) + 0.5 + u;
    rasterY = std::floor(rasterY) + 0.5 + v;
  }
  int w = XRes, h = YRes;
  float Sx = (float)(-1.0 + 2.0 * (rasterX / (double)w));
  float Sy = -(float)(-1.0 + 2.0 * (rasterY / (double)h));
  sc->Lambda = (float)lambda;
  sc->Time = (float)time;
  camera.ComputeRay(Sx, Sy, lensU, lensV, sc, ray);
  ray->I = iter;
  ray->Scramble[0] = scr.scramble[0];
  ray->Scramble[1] = scr.scramble[1];
}

// core/render.go:66-137,140-218. Worker threads pull 32x32 tiles; the per-pixel running mean
// (fb*iter + c)/(iter+1) with 1-based iter is kept (quirk d).
RenderStats Renderer::Render(int iterBegin, int iterEnd, int nthreads) {
  PreRender();
  if ((int)framescramble.size() != XRes * YRes) throw std::runtime_error("framescramble not set");
  RenderStats stats;
  // nthreads < 0: the reference's own worker policy — min(MaxGoRoutines, 10) goroutines (core/render.go:190; MaxGoRoutines is
  // given as -nthreads) and the two global atomic ray counters (core/stats.go:26-33)
  const bool faithful = nthreads < 0;
  if (faithful) nthreads = std::min(-nthreads, 10);
  std::atomic<uint64_t> sharedRays(0), sharedShadow(0);
  auto t0 = std::chrono::steady_clock::now();
  const int tilesX = (XRes + 31) / 32, tilesY = (YRes + 31) / 32;
  for (int iter0 = iterBegin; iter0 < iterEnd; iter0++) {
    const int iter = iter0 + 1;
    std::atomic<int> next(0);
    std::vector<RenderTask> tasks(nthreads);
    auto worker = [&](int ti) {
      RenderTask* task = &tasks[ti];
      task->scene = &scene;
      task->trace_last_level = trace_last_level;
      if (faithful) {
        task->sharedRayCount = &sharedRays;
        task->sharedShadowRayCount = &sharedShadow;
      }
      camera.PixelDelta(XRes, YRes, task->PixelDelta);
      Ray ray;
      ray.Task = task;
      ShaderContext sc;
      sc.task = task;
      for (;;) {
        int tile = next.fetch_add(1);
        if (tile >= tilesX * tilesY) break;
        int tx = (tile % tilesX) * 32, ty = (tile / tilesX) * 32;
        for (int j = 0; j < 32; j++)
          for (int i = 0; i < 32; i++) {
            int x = i + tx, y = j + ty;
            if (x >= XRes || y >= YRes) continue;
            GenerateCameraRay(iter, x, y, &sc, &ray);
            TraceSample samp;
            Trace(&ray, &samp);
            float* px = &framebuffer[(size_t)(x + y * XRes) * 3];
            for (int k = 0; k < 3; k++) px[k] = (px[k] * (float)iter + samp.Colour[k]) / (float)(iter + 1);
          }
      }
    };
    if (nthreads <= 1) {
      worker(0);
    } else {
      std::vector<std::thread> th;
      for (int i = 0; i < nthreads; i++) th.emplace_back(worker, i);
      for (auto& t : th) t.join();
    }
    for (auto& t : tasks) {
      stats.rayCount += t.rayCount;
      stats.shadowRayCount += t.shadowRayCount;
    }
  }
  stats.rayCount += sharedRays.load();
  stats.shadowRayCount += sharedShadow.load();
  stats.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return stats;
}

}  // namespace orc
