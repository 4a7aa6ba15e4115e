// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// CPU restatement of the reference's 4-wide BVH: node format, binned-SAH builder, motion
// (topology-only) builder, stack traversal and the SSE 4-box slab test.
// parity unpinned (the reference has no tests here); anchored by: SSE box test == the reference's own
// scalar twin intersectBoxesSlow2 on finite inputs, and BVH closest hit == brute force (tests/).
// Follows:
//   qbvh/qbvh.go:24-113              (Node, LeafCount/LeafBase, SetLeaf, SetEmptyLeaf)
//   qbvh/build.go:22-307             (calcMinCost, calcBox, binarySplit, buildAccelRec, BuildAccel)
//   qbvh/mqbvh.go:20-98, motionbuild.go:12-122 (MotionNode, BuildAccelMotion)
//   qbvh/intersect.go:52-87,91-246   (intersectBoxesSlow2, Trace)
//   qbvh/intersect_amd64.s:13-100    (intersectBoxes)
//   qbvh/motionintersect.go:22-127   (TraceMotion)
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "core.h"

namespace orc {

static const int MaxLeafCount = 16;

// qbvh/qbvh.go:31-36 — 128 bytes
struct alignas(16) Node {
  float Boxes[4 * 3 * 2];
  uint32_t Axis0, Axis1, Axis2;
  int32_t Children[4];
  int32_t Parent;

  void SetBounds(int idx, const BoundingBox& bounds) {
    for (int i = 0; i < 2; i++) for (int k = 0; k < 3; k++) Boxes[idx + (i * 12) + (k * 4)] = bounds.b[i][k];
  }
  BoundingBox Bounds(int idx) const {
    BoundingBox bb;
    for (int i = 0; i < 2; i++) for (int k = 0; k < 3; k++) bb.b[i][k] = Boxes[idx + (i * 12) + (k * 4)];
    return bb;
  }
  void SetEmptyLeaf(int idx) { Children[idx] = -1; SetBounds(idx, InfBox()); }
  void SetLeaf(int idx, uint32_t first, uint32_t count) {
    if (count == 0) { SetEmptyLeaf(idx); return; }
    uint32_t v = (1u << 31) | ((first << 4) & 0xfffffff0u) | ((count - 1) & 0xf);
    Children[idx] = (int32_t)v;
  }
};
static_assert(sizeof(Node) == 128, "qbvh.Node is 128 bytes");

// qbvh/qbvh.go:61-64. NOTE the 23-bit decode (mask before shift) — reference quirk (c).
static inline int LeafCount(int32_t l) { return (int)(((l)&0xf) + 1); }
static inline int LeafBase(int32_t l) { return (int)(((l)&0x7ffffff) >> 4); }

// qbvh/mqbvh.go:20-37
struct MotionNodeBoxes { float v[24];
  void SetBounds(int idx, const BoundingBox& bounds) {
    for (int i = 0; i < 2; i++) for (int k = 0; k < 3; k++) v[idx + (i * 12) + (k * 4)] = bounds.b[i][k];
  }
};
struct MotionNode {
  int32_t Axis0, Axis1, Axis2;
  int32_t Children[4];
  int32_t Parent;
  uint32_t pad[2];
  void SetLeaf(int idx, uint32_t first, uint32_t count) {
    if (count == 0) { Children[idx] = -1; return; }
    uint32_t v = (1u << 31) | ((first << 4) & 0xfffffff0u) | ((count - 1) & 0xf);
    Children[idx] = (int32_t)v;
  }
};
struct MotionQBVH {
  std::vector<std::vector<MotionNodeBoxes>> Boxes;  // [key][node]
  std::vector<MotionNode> Nodes;
};

std::vector<Node> BuildAccel(BoundingBox* boxes, Vec3* centroids, int32_t* indxs, int n, int leafMax, BoundingBox* bounds);
std::vector<MotionNode> BuildAccelMotion(BoundingBox* boxes, Vec3* centroids, int32_t* indxs, int n, int leafMax);

void intersectBoxes(const Ray* ray, const float* boxes, int32_t* hits, float* t);
void intersectBoxesSlow2(const Ray* ray, const float* boxes, int32_t* hits, float* t);
void intersectBoxesSlow(const Ray* ray, const float* boxes, int32_t* hits, float* t);  // intersect.go:17-50

bool QTrace(const std::vector<Node>& qbvh, Primitive* prim, Ray* ray, ShaderContext* sg);
bool QTraceMotion(const MotionQBVH& qbvh, float time, int key, int key2, MotionPrimitive* prim, Ray* ray, ShaderContext* sg);

}  // namespace orc
