// Device-side scene layout (see DESIGN.md "Data layout in HBM").
//
// Everything the traversal kernels touch lives in five flat arrays:
//   nodes      128-B static QBVH nodes of every mesh and of the scene level, concatenated; child links
//              re-encoded to GLOBAL indices (layout of one node = qbvh.Node, qbvh/qbvh.go:31-36: six float4 of
//              min/max x,y,z for the 4 children + axes + children), so one node is one 128-B line = 8 LDG.128.
//   mtopo      32-B motion-node topology records (qbvh/mqbvh.go:24-29 re-packed), global index = n_static + i.
//   mboxes     96-B motion box sets, [mesh block][key][node] (qbvh/mqbvh.go:20, Boxes[key][node]).
//   tris       48-B pre-gathered static triangles in leaf order: 3 x float4 {x,y,z,w}; w carries
//              geom id / prim id / bias term, so a leaf of <=16 triangles is one contiguous burst instead of the
//              reference's idxp -> Verts double indirection (polymesh/trace.go:119-125).
//   mtris      the same record per motion key: key k of slot s is mtris[3*(s + k*key_stride)].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "xform_math.h"

namespace vg {

// Child link encoding on the device (host flattening re-encodes the reference's links):
//   c >= 0                      interior node, global index (static if c < n_static, else motion node c - n_static)
//   c == -1                     empty child (qbvh.go:67-79)
//   bit31=1 bit30=0             triangle leaf: bit29 = motion triangles, base = (c>>4)&0x1FFFFFF, count = (c&15)+1
//   bit31=1 bit30=1 (c != -1)   geom leaf (scene level, leafMax=1): bit29 = 0: root node global index = c & 0x1FFFFFFF
//                               bit29 = 1: analytic sphere geom, record at triangle slot c & 0x1FFFFFF of `tris`
//                               ({centre.xyz, geom id}, {radius, 0, 0, prim 0}, {0,0,0,0})
static const uint32_t kLeafBit = 0x80000000u;
static const uint32_t kGeomBit = 0x40000000u;
static const uint32_t kMotionTriBit = 0x20000000u;
static const uint32_t kSphereBit = 0x20000000u;      // with kGeomBit
static const uint32_t kGeomRootMask = 0x1FFFFFFFu;
//                               bit28 = 1 (bit29 = 0): instance geom, payload = index into `xforms` (enter: transform the ray)
//                               bit28 = 1 and bit29 = 1: the stack sentinel pushed on entering an instance (leave: restore)
static const uint32_t kXformBit = 0x10000000u;       // with kGeomBit; mesh roots of such scenes must stay below 2^28
static const uint32_t kXformMask = 0x0FFFFFFFu;

struct DevXform {
  int32_t root;      // global node index of the target mesh's root
  int32_t geom;      // geom id of the instance (what a hit reports, scene.go:65)
  int32_t nkeys;     // transform keys; 1 = static: M / Minv pre-multiplied on the host in xf_static[2*i], [2*i+1]
  int32_t key_base;  // first XfSRT of this instance in xf_keys
  int32_t inner;     // instance of an instance: the xform applied next on the way to the mesh (instance.go:95), or -1
  int32_t pad_[3];
};
static const uint32_t kLeafBaseMask = 0x1FFFFFFu;
// float4s per static triangle record: 3 = the packed 48-B record; 4 = padded to 64 B so that a record is two aligned 256-bit loads
// (LDG.E.256, sm_100) instead of three 128-bit ones, at +33 % triangle memory
#ifndef VG_TRI_STRIDE
#define VG_TRI_STRIDE 3
#endif
static const int kTriStride = VG_TRI_STRIDE;  // 25 bits: 33.5 M triangle slots per kind

struct __align__(16) DevNode {  // 128 B
  float4 lo_x, lo_y, lo_z, hi_x, hi_y, hi_z;
  uint4 m0;  // axis0, axis1, axis2, child0
  uint4 m1;  // child1, child2, child3, parent(unused)
};

struct __align__(16) DevMotionNode {  // 32 B
  int32_t child[4];
  uint32_t axes_keys;       // axis0 | axis1<<2 | axis2<<4 | keys<<8
  uint32_t box_base;        // index (units of 6 float4) of this node's key-0 box set
  uint32_t box_key_stride;  // units of 6 float4 between keys
  uint32_t tri_key_stride;  // triangle slots between keys (for the leaves below this node)
};

struct DevGeom {  // per geom (creation order)
  int32_t tri_base;     // global slot of the mesh's first triangle (static or motion space)
  int32_t prim_base;    // into prim_material[]
  int32_t normal_base;  // slot base into tri_normals[] (3 float4 per slot), or -1
  int32_t keys;         // 1 = static mesh, > 1 = motion mesh, 0 = analytic sphere (tri_base = its record)
  int32_t tri_key_stride;
  int32_t n_tris;
  int32_t uv_base;      // slot base into tri_uv[] (3 float2 per slot), or -1: the mesh has no UVs
  int32_t xform;        // GeomInstance: index into DevScene::xforms of this geom's transform chain; -1 otherwise
};

struct DevScene {
  const DevNode* nodes;
  const DevMotionNode* mtopo;
  const float4* mboxes;
  const float4* tris;
  const float4* mtris;
  const DevGeom* geoms;
  const uint8_t* prim_material;  // global material id per (geom, prim)
  const float4* tri_normals;     // optional per-slot vertex normals (static meshes only)
  const float2* tri_uv;          // optional per-slot texture coordinates (static meshes only), 3 per slot
  int32_t n_static;              // number of static nodes (motion node global index = n_static + i)
  int32_t root;                  // global index of the scene-level root node
  int32_t n_geoms;
  const DevXform* xforms;
  const XfSRT* xf_keys;
  const Mat4* xf_static;
  int32_t n_mtris;               // motion triangle slots (selects the kernels whose cooperative leaf phase handles them)
  int32_t n_xforms;              // instance geoms (selects the kernels that carry the transform enter/leave code)
  int32_t n_spheres;             // analytic sphere geoms in the scene-level tree (selects the kernels that carry their leaf test)
};

}  // namespace vg
