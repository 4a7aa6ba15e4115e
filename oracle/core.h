// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// CPU restatement of the reference's ray / shading-context records and interfaces.
// parity unpinned (the reference has no tests on this path).
// Follows:
//   core/ray.go:13-20,27-53,56-93,103-148,152-166   (Ray, Init, Setup, RenderTask)
//   core/shader.go:53-102,129-162                    (ShaderContext, ApplyTransform, OffsetP)
//   core/geom.go:8-18, core/light.go:13-42, core/shader.go:16-49 (Geom/Light/BSDF/Shader interfaces)
//   qbvh/qbvh.go:18-20, qbvh/mqbvh.go:13-15          (Primitive / MotionPrimitive)
// Ray differentials (core/ray.go:40-44,72-87,95-99) are carried: they feed the texture footprint of texture.h.
#pragma once
#include <atomic>
#include <cstdint>
#include <vector>

#include "colour.h"
#include "vmath.h"

namespace orc {

enum : uint32_t {
  RayTypeCamera = 1u << 0,
  RayTypeShadow = 1u << 1,
  RayTypeReflected = 1u << 2,
  RayTypeRefracted = 1u << 3,
  RayTypeGlossy = 1u << 4,
};

// core/ray.go:20
static const float ShadowRayEpsilon = 0.0001f;

struct Ray;
struct ShaderContext;
struct Scene;

// core/ray.go:152-166 (pools omitted: they do not affect results)
struct alignas(16) RenderTask {
  struct {
    alignas(16) float T[4];
    alignas(16) int32_t Hits[4];
    alignas(16) float Boxes[24];
    int32_t StackTop;
    struct { float T; int32_t Node; } Stack[90];
  } Traversal;
  Scene* scene = nullptr;        // the reference uses a package global (core/core.go:11)
  uint64_t rayCount = 0;         // core/stats.go:26-33 (kept per task, summed at the end)
  uint64_t shadowRayCount = 0;
  // reference-faithful statistics: the reference bumps two process-global counters with atomic.AddUint64 on EVERY ray
  // (core/stats.go:26-33, core/trace.go:28-32). When set, TraceProbe does the same on these shared counters instead of
  // the per-task ones: the contended cache line is a real cost of the reference on many cores (bench.py's "faithful" CPU leg).
  std::atomic<uint64_t>* sharedRayCount = nullptr;
  std::atomic<uint64_t>* sharedShadowRayCount = nullptr;
  bool trace_last_level = true;  // trace the level-4 mirror ray like the reference does (std.go:243)
  float PixelDelta[2] = {0, 0};  // core.Image.PixelDelta: a package global in the reference (render.go:47, camera.go:316-317)
  RenderTask() { Traversal.StackTop = 0; }
};

// core/ray.go:27-53
struct Ray {
  Vec3 P, D, Dinv;
  float Tclosest;
  float S[3];
  int32_t Kx, Ky, Kz;
  float Time, Lambda;
  uint8_t Level;
  uint32_t Type;
  int64_t I;
  uint64_t Scramble[2];
  int64_t NodesT, LeafsT;
  int64_t TrisT;  // oracle-only counter: sum of LeafCount over visited triangle leaves (bytes model)
  RenderTask* Task;
  Vec3 DdPdx{}, DdPdy{}, DdDdx{}, DdDdy{};  // core/ray.go:40-44
  void DifferentialTransfer(ShaderContext* sc) const;  // core/ray.go:95-104

  void Setup();
  void Init(uint32_t ty, Vec3 P, Vec3 D, float maxdist, uint8_t level, const ShaderContext* sc);
};

struct Geom {
  virtual ~Geom() {}
  virtual bool Trace(Ray*, ShaderContext*) = 0;
  virtual int MotionKeys() const = 0;
  virtual BoundingBox Bounds(float time) const = 0;
  int id = -1;  // oracle-only: creation order, to report geomID
};

struct Primitive {
  virtual bool TraceElems(Ray* ray, ShaderContext* sg, int base, int count) = 0;
};
struct MotionPrimitive {
  virtual bool TraceMotionElems(float time, int key, int key2, Ray* ray, ShaderContext* sg, int base, int count) = 0;
};

struct BSDFSample {
  Vec3 D;
  double Pdf;
  float PdfLight;
  Spectrum Liu;
  Vec3 Ld;
  float Ldist;
};
struct LightSample {
  Vec3 P;
  float Pdf;
  Spectrum Liu;
  Vec3 Ld;
  float Ldist;
};

struct BSDF {
  virtual Vec3 Sample(double r0, double r1) = 0;
  virtual Spectrum Eval(Vec3 omegaO) = 0;
  virtual double PDF(Vec3 omegaO) = 0;
};

struct Shader {
  virtual ~Shader() {}
  virtual void Eval(ShaderContext* sc) = 0;
  virtual RGB EvalEmission(ShaderContext* sc, Vec3 omegaO) = 0;
};

struct Light {
  virtual ~Light() {}
  virtual void SampleArea(ShaderContext* sg, int n) = 0;
  virtual float DiffuseShadeMult() = 0;
  virtual int NumSamples(ShaderContext* sg) = 0;
  virtual bool ValidSample(ShaderContext* sg, BSDFSample* sample) = 0;
  virtual Geom* GetGeom() = 0;
};

// core/shader.go:53-102 (fields used by the in-scope path)
struct ShaderContext {
  uint8_t Level = 0;
  int64_t I = 0;
  int NSamples = 0;
  uint64_t Scramble[2] = {0, 0};
  float Lambda = 0, Time = 0;
  Vec3 Ro{}, Rd{};
  uint32_t ElemID = 0;
  Geom* geom = nullptr;
  Shader* shader = nullptr;
  Vec3 Po{}, P{}, Poffset{};
  Vec3 N{}, Ng{};
  Vec3 DdPdu{}, DdPdv{};
  float U = 0, V = 0;                           // surface parameters (texture coordinates), core/shader.go:78-79
  Vec3 DdPdx{}, DdPdy{}, DdDdx{}, DdDdy{}, DdNdx{}, DdNdy{};  // core/shader.go:81-86
  float Dduvdx[2] = {0, 0}, Dduvdy[2] = {0, 0};
  float PixelDelta[2] = {0, 0};                 // core.Image.PixelDelta (render.go:32-35): a process-wide constant of the camera
  float Bu = 0, Bv = 0, Bw = 0;  // Bw is oracle-only (W is a local in trace.go:113)
  std::vector<Light*> Lights;
  int Lidx = 0;
  std::vector<LightSample> Lsamples;
  Light* Lp = nullptr;
  RGB OutRGB{};
  Matrix4 Transform = Matrix4Identity(), InvTransform = Matrix4Identity();  // core/trace.go:57-58
  bool transformSet = false;  // oracle-only: an Instance overwrote Transform (identity otherwise)
  RenderTask* task = nullptr;

  void ApplyTransform();
  Vec3 OffsetP(int dir) const;
  void LightsPrepare();
  bool NextLight();
  RGB EvaluateLightSamples(BSDF* bsdf);
};

// core/trace.go:13-21
struct TraceSample {
  RGB Colour{};
  Vec3 Point{};
  uint32_t ElemID = 0;
  Geom* geom = nullptr;
};

bool TraceProbe(Ray* ray, ShaderContext* sg);
bool Trace(Ray* ray, TraceSample* samp);

}  // namespace orc
