// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// CPU restatement of the reference's triangle-mesh geom and default two-level scene.
// parity unpinned (no reference tests); anchored by BVH == brute force with the same triangle routine.
// Follows:
//   builtin/geom/polymesh/polymesh.go:17-118   (PolyMesh, PreRender, MotionKeys)
//   builtin/geom/polymesh/init.go:12-134       (polygon fan triangulation -> idxp)
//   builtin/geom/polymesh/buildqbvh.go:14-212  (initAccel, initMotionBoxes(Rec))
//   builtin/geom/polymesh/trace.go:15-104,108-194,276-360,504-515,520-719 (Trace, TraceElems, TraceMotionElems)
//   builtin/geom/polymesh/bounds.go:25-53      (Bounds)
//   builtin/scene/scene.go:15-268              (Scene: Trace, TraceElems, LightsPrepare, initAccel, initMotionBoxes)
//   builtin/geom/instance/instance.go:16-160   (GeomInstance: SRT-interpolated transform, ray re-Setup, user-given bounds)
// UVs and ray differentials (trace.go:350-502) are computed for static meshes, like the reference (TraceMotionElems sets
// U,V but leaves every differential 0, trace.go:677-684). The mesh's own Transform (quirk q) is restated for documentation only.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "core.h"
#include "qbvh.h"

namespace orc {

struct PointArray {
  int MotionKeys = 0;
  int ElemsPerKey = 0;
  std::vector<Vec3> Elems;
};

struct PolyMesh : Geom, Primitive, MotionPrimitive {
  std::string Name;
  float RayBias = 0;
  PointArray Verts;
  std::vector<int32_t> PolyCount, FaceIdx;
  std::vector<int32_t> ShaderIdx;
  PointArray Normals;
  std::vector<int32_t> NormalIdx;
  bool hasPolyCount = false, hasFaceIdx = false, hasNormalIdx = false, hasUVIdx = false;
  struct Vec2 { float v[2]; };
  std::vector<Vec2> UV;           // param.Vec2Array, one key (polymesh.go:36)
  std::vector<int32_t> UVIdx;     // polymesh.go:37
  std::vector<uint32_t> uvtriidx;  // triangulated UV indexes (polymesh.go:45)

  int facecount = 0;
  std::vector<uint32_t> idxp;
  std::vector<uint32_t> normalidx;
  std::vector<uint8_t> shaderidx;
  struct {
    MotionQBVH mqbvh;
    std::vector<Node> qbvh;
    std::vector<int32_t> idx;
  } accel;
  std::vector<Shader*> shader;
  BoundingBox bounds;
  std::vector<BoundingBox> motionBounds;
  bool ref_compat_motion = true;  // quirk (b): motion leaves test face i, not accel.idx[i]

  void init();
  void initAccel();
  void initMotionBoxes();
  BoundingBox initMotionBoxesRec(int key, int32_t node);
  void PreRender() { init(); facecount = (int)idxp.size() / 3; initAccel(); }

  // PolyMesh.Transform (polymesh.go:32-33, init.go:14-18): restated ONLY to document what the reference does with it (quirk q);
  // the GPU path refuses non-identity mesh transforms and tests/test_oracle_mesh_transform.py shows why.
  std::vector<Matrix4> Transform;
  std::vector<TransformDecomp> transformSRT;
  bool Trace(Ray*, ShaderContext*) override;
  int MotionKeys() const override { return accel.qbvh.empty() ? (int)accel.mqbvh.Boxes.size() : 1; }
  BoundingBox Bounds(float time) const override;
  bool TraceElems(Ray* ray, ShaderContext* sg, int base, int count) override;
  bool TraceMotionElems(float time, int key, int key2, Ray* ray, ShaderContext* sg, int base, int count) override;
  // brute force over all triangles with the same routine (anchor for the BVH)
  bool TraceBrute(Ray* ray, ShaderContext* sg);
};

// builtin/geom/instance/instance.go:36-160. "Instance duplicates an existing geom but with a new transform."
struct Instance : Geom {
  std::string Name;
  Geom* geom = nullptr;                    // ins.geom (the target keeps being a scene geom of its own)
  std::vector<Vec3> BMin, BMax;            // param.PointArray elements
  std::vector<Matrix4> Transform;          // param.MatrixArray elements
  std::vector<TransformDecomp> transformSRT;
  std::vector<BoundingBox> bounds;

  void PreRender();                        // instance.go:117-146
  TransformDecomp TimeKey(float time) const;  // instance.go:16-33
  bool Trace(Ray*, ShaderContext*) override;
  int MotionKeys() const override { return geom->MotionKeys(); }
  BoundingBox Bounds(float) const override { return bounds[0]; }
};

struct Scene : Primitive, MotionPrimitive {
  std::vector<Node> qbvh;
  MotionQBVH mqbvh;
  std::vector<Geom*> geoms;
  std::vector<Light*> lights;
  BoundingBox bounds;

  bool Trace(Ray* ray, ShaderContext* sg);
  bool TraceElems(Ray* ray, ShaderContext* sc, int base, int count) override;
  bool TraceMotionElems(float time, int key, int key2, Ray* ray, ShaderContext* sc, int base, int count) override;
  void LightsPrepare(ShaderContext* sg);
  void initAccel();
  void initMotionBoxes(int keys);
  BoundingBox initMotionBoxesRec(int key, int32_t node);
};

}  // namespace orc
