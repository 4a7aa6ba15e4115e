// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// CPU restatement of the reference's default shader, BSDFs, Fresnel, triangle light, camera and
// frame loop. parity unpinned (no reference tests; Go stdlib trig is replaced by glibc libm in
// double, rounded to float exactly where the reference rounds).
// Follows:
//   builtin/shader/std.go:77-316              (ShaderStd.Eval / EvalEmission)
//   builtin/shader/bsdf/orennayar.go:16-73    (OrenNayar)
//   builtin/shader/bsdf/specular.go:13-101    (Specular mirror)
//   builtin/shader/bsdf/microfacetggx.go:16-146 (MicrofacetGGX glossy lobe)
//   builtin/shader/fresnel/dielectric.go:16-47, conductor.go:19-89, models.go:13-17 (Dielectric, Conductor)
//   math/sample/sample.go:18-30,105-129       (CosineHemisphere, UniformDisk2D)
//   builtin/light/triangle.go:71-343,376-535  (Tri light: area + spherical-triangle sampling)
//   builtin/light/disk.go:21-270              (Disk light; rayPlaneIntersect / rayDiskIntersect)
//   builtin/light/sphere.go:17-304            (Sphere light)
//   builtin/geom/sphere/sphere.go:15-86, trace.go:13-109 (Sphere geom, only used by the sphere light)
//   builtin/camera/camera.go:80-98,109-193,221-323 (Camera)
//   core/render.go:18-23,66-137,166-205       (pixelscramble, render, Render)
// Out of scope here: Quad light (both sampling entry points panic in the reference, quad.go:88,94).
#pragma once
#include <string>
#include <vector>

#include "core.h"
#include "scene.h"
#include "texture.h"

namespace orc {

struct ShaderStd : Shader {
  std::string Name;
  bool hasEmissionColour = false, hasEmissionStrength = false;
  RGB EmissionColour;
  float EmissionStrength = 0;
  bool hasDiffuseColour = false, hasDiffuseStrength = false, hasDiffuseRoughness = false;
  RGB DiffuseColour;
  float DiffuseStrength = 0, DiffuseRoughness = 0;
  bool hasSpec1Colour = false, hasSpec1Strength = false, hasSpec1Roughness = false;
  RGB Spec1Colour;
  float Spec1Strength = 0, Spec1Roughness = 0;
  bool hasIOR = false;
  float IOR = 0;
  int spec1FresnelModel = 0;  // fresnel.DielectricModel = 0 (the zero value when Spec1FresnelModel is unset), ConductorModel = 1
  bool hasSpec1FresnelRefl = false, hasSpec1FresnelEdge = false;
  RGB Spec1FresnelRefl, Spec1FresnelEdge;
  // parameter slots in node order (std.go:28-46); a slot with a texture bound is a maps.Texture / TextureTrilinear
  enum { kEmissionColour = 0, kEmissionStrength, kDiffuseColour, kDiffuseStrength, kDiffuseRoughness, kSpec1Colour, kSpec1Strength,
         kSpec1Roughness, kIOR, kSpec1FresnelModel, kSpec1FresnelRefl, kSpec1FresnelEdge, kNumSlots };
  TextureMap tex[kNumSlots];
  RGB rgb(int slot, const RGB& constant, const ShaderContext* sg) const;
  float f32(int slot, float constant, const ShaderContext* sg) const;

  void Eval(ShaderContext* sg) override;
  RGB EvalEmission(ShaderContext* sg, Vec3 omegaO) override;
};

// builtin/shader/debug.go:15-49 (node "DebugShader"): OutRGB = Colour, no emission. Derived from ShaderStd only so that the
// renderer's shader list keeps one element type; nothing of ShaderStd is used.
struct DebugShader : ShaderStd {
  RGB Colour;
  void Eval(ShaderContext* sg) override;
  RGB EvalEmission(ShaderContext* sg, Vec3 omegaO) override;
};

struct Tri : Light {
  std::string Name;
  Vec3 P0, P1, P2;
  int Samples = 1;
  Shader* shader = nullptr;
  PolyMesh* geom = nullptr;

  PolyMesh* createMesh();
  void SampleArea(ShaderContext* sg, int n) override;
  void sampleByArea(ShaderContext* sg, int n);
  float DiffuseShadeMult() override { return 1; }
  int NumSamples(ShaderContext*) override { return 1 << (unsigned)Samples; }
  bool ValidSample(ShaderContext* sg, BSDFSample* sample) override;
  Geom* GetGeom() override { return geom; }
};

// builtin/light/disk.go:21-34
struct Disk : Light {
  std::string Name;
  Vec3 P, Up, LookAt;
  Vec3 T, B, N;
  float Radius = 0;
  int Segments = 20, Samples = 1;
  Shader* shader = nullptr;
  PolyMesh* geom = nullptr;

  void PreRender();  // disk.go:90-97: N, T, B
  PolyMesh* createMesh();
  void SampleArea(ShaderContext* sg, int n) override;
  float DiffuseShadeMult() override { return 1; }
  int NumSamples(ShaderContext*) override { return 1 << (unsigned)Samples; }
  bool ValidSample(ShaderContext* sg, BSDFSample* sample) override;
  Geom* GetGeom() override { return geom; }
};

// builtin/geom/sphere/sphere.go:15-86, trace.go:13-109
struct SphereGeom : Geom {
  std::string Name;
  Vec3 P;
  float Radius = 1;
  Shader* shader = nullptr;
  bool Trace(Ray*, ShaderContext*) override;
  int MotionKeys() const override { return 1; }
  BoundingBox Bounds(float time) const override;
};

// builtin/light/sphere.go:17-304
struct SphereLight : Light {
  std::string Name;
  Vec3 P;
  float Radius = 1;
  int Samples = 1;
  Shader* shader = nullptr;
  SphereGeom* geom = nullptr;

  void SampleArea(ShaderContext* sg, int n) override;
  float DiffuseShadeMult() override { return 1; }
  int NumSamples(ShaderContext*) override { return 1 << (unsigned)Samples; }
  bool ValidSample(ShaderContext* sg, BSDFSample* sample) override;
  Geom* GetGeom() override { return geom; }
};

struct Camera {
  // Type "LookAt": From/To/Roll with any number of motion keys (param.PointArray / Float32Array); Type "Matrix":
  // WorldToLocal matrices (camera.go:48-73,109-216)
  std::string Type = "LookAt";
  std::vector<Vec3> FromKeys, ToKeys;
  std::vector<float> RollKeys;  // empty = Roll.Elems == nil
  std::vector<Matrix4> WorldToLocal;
  Vec3 Up;
  float Aspect = 0, Fov = 90, Focal = 12, Radius = 0;
  float TanThetaFocal = 0;
  std::vector<Matrix4> LocalToWorld;
  std::vector<TransformDecomp> decomp;
  // convenience for single-key LookAt cameras
  Vec3 From, To;
  float Roll = 0;
  Matrix4 M;  // LocalToWorld at Time 0 after decompose/recompose (what every ray of a single-key camera uses)
  void PreRender(float frameAspect);
  Matrix4 MatrixAt(float time) const;  // camera.go:225-236
  void ComputeRay(float Sx, float Sy, double lensU, double lensV, const ShaderContext* sc, Ray* ray) const;
  void PixelDelta(int w, int h, float out[2]) const;
};

// builtin/filter/filter.go:29-173 — filter-importance-sampling tables (marginal + conditional CDFs)
struct FilterSampler {
  std::vector<double> cdfV;
  std::vector<std::vector<double>> cdfVU;
  int n = 0;
  double w = 0;
  void Create(int n, double w, double (*f)(double, double, void*), void* ctx);
  void WarpSample(double r0, double r1, double* u, double* v) const;
};
double BesselJ1(double x);  // builtin/filter/airy.go:34-72
struct PixelFilter {         // AiryFilter (airy.go:13-101) / GaussianFilter (gauss.go:13-50)
  int kind = 0;              // 1 = Airy, 2 = Gaussian
  float Width = 0, Peak = 0;
  int Res = 0;
  FilterSampler sampler;
  void PreRender();
};

// core/render.go:18-23
struct pixelscramble {
  uint64_t lensU, lensV, time, lambda, scramble[2];
};

struct RenderStats {
  uint64_t rayCount = 0, shadowRayCount = 0;
  double seconds = 0;
};

struct Renderer {
  Scene scene;
  Camera camera;
  int XRes = 1024, YRes = 1024;
  std::vector<float> framebuffer;
  std::vector<pixelscramble> framescramble;
  std::vector<std::unique_ptr<PolyMesh>> meshes;
  std::vector<std::unique_ptr<Instance>> instances;  // GeomInstance nodes, created after the meshes they refer to
  std::vector<std::unique_ptr<ShaderStd>> shaders;
  std::vector<std::unique_ptr<Texture>> textures;  // texture.TexStore (texture/texture.go:45-47), filled before PreRender
  std::vector<std::unique_ptr<Tri>> tris;
  std::vector<std::unique_ptr<Disk>> disks;
  std::vector<std::unique_ptr<SphereLight>> sphereLights;
  std::vector<std::unique_ptr<SphereGeom>> sphereGeoms;
  std::vector<Light*> lightOrder;  // every light in node-creation order (core.AddNode -> scene.AddLight, core/core.go:77-93)
  bool prerendered = false;
  std::unique_ptr<PixelFilter> filter;  // core.filter (core/core.go:15), set by AddNode for a PixelFilter node
  bool trace_last_level = true;  // false: skip the level-4 mirror ray whose shader returns black (std.go:95)

  ShaderStd* findShader(const std::string& name);
  void PreRender();
  // iterations [iterBegin, iterEnd) are the reference's 0-based `iter`; render() receives iter+1.
  RenderStats Render(int iterBegin, int iterEnd, int nthreads);
  // one camera sample; used by Render and by tests that want the primary ray
  void GenerateCameraRay(int iter1, int x, int y, ShaderContext* sc, Ray* ray) const;
};

}  // namespace orc
