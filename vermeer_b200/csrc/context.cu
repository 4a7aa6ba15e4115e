// Device layer of the C ABI (include/vermeer_gpu.h, vg_*): context, scene flattening, batch TraceProbe.
#include "context.h"

#include <algorithm>
#include <cstring>

#include "kernels.h"

using namespace vg;

namespace {
std::string g_create_err;

#define VG_LOCK(ctx)             \
  if (!(ctx)) return VG_ERR_INVALID; \
  std::lock_guard<std::mutex> lock_((ctx)->mu)

#define VG_CUDA(ctx, call)                                   \
  do {                                                       \
    cudaError_t e_ = (call);                                 \
    if (e_ != cudaSuccess) return (ctx)->cuda_fail(e_, #call); \
  } while (0)

// The reference's leaf decode, including the 23-bit LeafBase (qbvh/qbvh.go:61-64, quirk c).
inline float i2f(int32_t i) { float f; std::memcpy(&f, &i, 4); return f; }
inline int refLeafCount(int32_t l) { return (int)((l & 0xf) + 1); }
inline int refLeafBase(int32_t l) { return (int)((l & 0x7ffffff) >> 4); }

template <class T>
cudaError_t upload(DevBuf<T>& buf, const std::vector<T>& v, cudaStream_t s) {
  cudaError_t e = buf.reserve(v.size());
  if (e != cudaSuccess) return e;
  if (v.empty()) return cudaSuccess;
  return cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}
}  // namespace

extern "C" {

int vg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char* vg_last_error(vg_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int vg_create(vg_ctx** out, int device_ordinal) {
  if (!out) return VG_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    g_create_err = std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e);
    return VG_ERR_NO_DEVICE;
  }
  if (device_ordinal < 0 || device_ordinal >= n) {
    g_create_err = "device ordinal out of range";
    return VG_ERR_INVALID;
  }
  e = cudaSetDevice(device_ordinal);
  if (e != cudaSuccess) {
    g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return VG_ERR_CUDA;
  }
  vg_ctx* c = new vg_ctx();
  c->device = device_ordinal;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_ordinal) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess ||
      (e = c->d_counters.reserve(16)) != cudaSuccess) {
    g_create_err = std::string("context setup: ") + cudaGetErrorString(e);
    delete c;
    return VG_ERR_CUDA;
  }
  cudaMemsetAsync(c->d_counters.p, 0, 16 * sizeof(unsigned long long), c->stream);
  *out = c;
  return VG_OK;
}

void vg_destroy(vg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  render_destroy(ctx);
  ctx->d_nodes.release(); ctx->d_mtopo.release(); ctx->d_mboxes.release(); ctx->d_tris.release();
  ctx->d_mtris.release(); ctx->d_normals.release(); ctx->d_geoms.release(); ctx->d_prim_material.release();
  ctx->d_xforms.release(); ctx->d_xf_keys.release(); ctx->d_xf_static.release();
  ctx->d_rays.release(); ctx->d_hits.release(); ctx->d_counters.release(); ctx->d_cam_keys.release();
  for (int i = 0; i < 3; i++) {
    if (ctx->pipe_stream[i]) cudaStreamDestroy(ctx->pipe_stream[i]);
    if (ctx->pipe_done[i]) cudaEventDestroy(ctx->pipe_done[i]);
  }
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int vg_scene_begin(vg_ctx* ctx, int n_geoms) {
  VG_LOCK(ctx);
  if (n_geoms < 0) return ctx->fail(VG_ERR_INVALID, "n_geoms < 0");
  ctx->meshes.assign(n_geoms, MeshStage());
  ctx->scene = SceneStage();
  ctx->committed = false;
  render_invalidate(ctx);
  return VG_OK;
}

int vg_mesh_upload(vg_ctx* ctx, int geom_id, const VgNode* nodes, int n_nodes, const uint32_t* idxp, int n_tris, const float* verts,
                   int n_verts, const uint8_t* shaderidx, const int32_t* material_ids, int n_materials, const float* normals,
                   int n_normals, const uint32_t* normalidx, float raybias) {
  VG_LOCK(ctx);
  if (geom_id < 0 || geom_id >= (int)ctx->meshes.size()) return ctx->fail(VG_ERR_INVALID, "geom_id out of range");
  if (!nodes || n_nodes <= 0 || !idxp || n_tris < 0 || !verts || n_verts <= 0) return ctx->fail(VG_ERR_INVALID, "vg_mesh_upload: null/empty input");
  MeshStage& m = ctx->meshes[geom_id];
  m = MeshStage();
  m.present = true;
  m.nodes.assign(nodes, nodes + n_nodes);
  m.idxp.assign(idxp, idxp + (size_t)n_tris * 3);
  m.n_tris = n_tris;
  m.verts.assign(verts, verts + (size_t)n_verts * 3);
  m.n_verts = n_verts;
  for (uint32_t i : m.idxp)
    if (i >= (uint32_t)n_verts) return ctx->fail(VG_ERR_INVALID, "vg_mesh_upload: vertex index out of range");
  if (shaderidx) m.shaderidx.assign(shaderidx, shaderidx + n_tris);
  if (material_ids) m.material_ids.assign(material_ids, material_ids + n_materials);
  if (normals && normalidx) {
    m.normals.assign(normals, normals + (size_t)n_normals * 3);
    m.normalidx.assign(normalidx, normalidx + (size_t)n_tris * 3);
    for (uint32_t i : m.normalidx)
      if (i >= (uint32_t)n_normals) return ctx->fail(VG_ERR_INVALID, "vg_mesh_upload: normal index out of range");
  }
  m.raybias = raybias;
  ctx->committed = false;
  return VG_OK;
}

int vg_mesh_upload_motion(vg_ctx* ctx, int geom_id, const VgMotionNode* topo, int n_nodes, const float* boxes, int keys, const uint32_t* idxp,
                          const int32_t* accel_idx, int n_tris, const float* verts, int n_verts, const uint8_t* shaderidx,
                          const int32_t* material_ids, int n_materials, float raybias, int ref_compat) {
  VG_LOCK(ctx);
  if (geom_id < 0 || geom_id >= (int)ctx->meshes.size()) return ctx->fail(VG_ERR_INVALID, "geom_id out of range");
  if (!topo || n_nodes <= 0 || !boxes || keys < 2 || keys > 255 || !idxp || !accel_idx || !verts || n_verts <= 0)
    return ctx->fail(VG_ERR_INVALID, "vg_mesh_upload_motion: null/empty input or keys outside [2,255]");
  MeshStage& m = ctx->meshes[geom_id];
  m = MeshStage();
  m.present = true;
  m.motion = true;
  m.keys = keys;
  m.topo.assign(topo, topo + n_nodes);
  m.boxes.assign(boxes, boxes + (size_t)keys * n_nodes * 24);
  m.idxp.assign(idxp, idxp + (size_t)n_tris * 3);
  m.accel_idx.assign(accel_idx, accel_idx + n_tris);
  m.n_tris = n_tris;
  m.verts.assign(verts, verts + (size_t)keys * n_verts * 3);
  m.n_verts = n_verts;
  for (uint32_t i : m.idxp)
    if (i >= (uint32_t)n_verts) return ctx->fail(VG_ERR_INVALID, "vg_mesh_upload_motion: vertex index out of range");
  for (int32_t f : m.accel_idx)
    if (f < 0 || f >= n_tris) return ctx->fail(VG_ERR_INVALID, "vg_mesh_upload_motion: accel_idx out of range");
  if (shaderidx) m.shaderidx.assign(shaderidx, shaderidx + n_tris);
  if (material_ids) m.material_ids.assign(material_ids, material_ids + n_materials);
  m.raybias = raybias;
  m.ref_compat = ref_compat;
  ctx->committed = false;
  return VG_OK;
}

int vg_sphere_upload(vg_ctx* ctx, int geom_id, const float* centre, float radius, int32_t material_id) {
  VG_LOCK(ctx);
  if (geom_id < 0 || geom_id >= (int)ctx->meshes.size()) return ctx->fail(VG_ERR_INVALID, "geom_id out of range");
  if (!centre) return ctx->fail(VG_ERR_INVALID, "vg_sphere_upload: null centre");
  MeshStage& m = ctx->meshes[geom_id];
  m = MeshStage();
  m.present = true;
  m.sphere = true;
  m.centre[0] = centre[0]; m.centre[1] = centre[1]; m.centre[2] = centre[2];
  m.radius = radius;
  m.n_tris = 1;  // one record in the triangle array, one prim (ElemID 0)
  m.material_ids.assign(1, material_id);
  ctx->committed = false;
  return VG_OK;
}

int vg_instance_upload(vg_ctx* ctx, int geom_id, int target_geom_id, const VgTransformSRT* srt, int keys) {
  VG_LOCK(ctx);
  if (geom_id < 0 || geom_id >= (int)ctx->meshes.size()) return ctx->fail(VG_ERR_INVALID, "geom_id out of range");
  if (target_geom_id < 0 || target_geom_id >= (int)ctx->meshes.size() || target_geom_id == geom_id)
    return ctx->fail(VG_ERR_INVALID, "vg_instance_upload: target geom out of range");
  if (!srt || keys < 1 || keys > 255) return ctx->fail(VG_ERR_INVALID, "vg_instance_upload: need 1..255 transform keys");
  MeshStage& m = ctx->meshes[geom_id];
  m = MeshStage();
  m.present = true;
  m.instance = true;
  m.target = target_geom_id;
  m.srt.assign(srt, srt + keys);
  m.n_tris = 0;
  ctx->committed = false;
  return VG_OK;
}

int vg_scene_upload(vg_ctx* ctx, const VgNode* nodes, int n_nodes, const int32_t* geom_of_slot, int n_slots) {
  VG_LOCK(ctx);
  if (!nodes || n_nodes <= 0 || !geom_of_slot || n_slots < 0) return ctx->fail(VG_ERR_INVALID, "vg_scene_upload: null/empty input");
  ctx->scene = SceneStage();
  ctx->scene.present = true;
  ctx->scene.nodes.assign(nodes, nodes + n_nodes);
  ctx->scene.geom_of_slot.assign(geom_of_slot, geom_of_slot + n_slots);
  ctx->committed = false;
  return VG_OK;
}

int vg_scene_upload_motion(vg_ctx* ctx, const VgMotionNode* topo, int n_nodes, const float* boxes, int keys, const int32_t* geom_of_slot, int n_slots) {
  VG_LOCK(ctx);
  if (!topo || n_nodes <= 0 || !boxes || keys < 2 || keys > 255 || !geom_of_slot) return ctx->fail(VG_ERR_INVALID, "vg_scene_upload_motion: bad input");
  ctx->scene = SceneStage();
  ctx->scene.present = true;
  ctx->scene.motion = true;
  ctx->scene.keys = keys;
  ctx->scene.topo.assign(topo, topo + n_nodes);
  ctx->scene.boxes.assign(boxes, boxes + (size_t)keys * n_nodes * 24);
  ctx->scene.geom_of_slot.assign(geom_of_slot, geom_of_slot + n_slots);
  ctx->committed = false;
  return VG_OK;
}

int vg_scene_commit(vg_ctx* ctx) {
  VG_LOCK(ctx);
  if (!ctx->scene.present) return ctx->fail(VG_ERR_INVALID, "vg_scene_commit: no scene-level tree uploaded");
  const int G = (int)ctx->meshes.size();
  for (int g = 0; g < G; g++)
    if (!ctx->meshes[g].present) return ctx->fail(VG_ERR_INVALID, "vg_scene_commit: geom " + std::to_string(g) + " was not uploaded");

  for (int g = 0; g < G; g++) {
    const MeshStage& m = ctx->meshes[g];
    if (!m.instance) continue;
    const MeshStage& tg = ctx->meshes[m.target];
    if (tg.instance || tg.sphere) return ctx->fail(VG_ERR_UNSUPPORTED, "GeomInstance of an Instance or a Sphere is outside this path (PolyMesh targets only)");
  }

  // ---- index spaces ----
  std::vector<int64_t> node_base(G), tri_base(G), prim_base(G), normal_base(G, -1), mbox_base(G, 0);
  int64_t n_static = 0, n_motion = 0, n_tris = 0, n_mtris = 0, n_prims = 0, n_mboxes = 0, n_normal_slots = 0;
  for (int g = 0; g < G; g++) {
    const MeshStage& m = ctx->meshes[g];
    if (!m.motion) {
      node_base[g] = n_static;
      n_static += (int64_t)m.nodes.size();
      tri_base[g] = n_tris;
      n_tris += m.n_tris;
    } else {
      node_base[g] = n_motion;  // relative to n_static, fixed up below
      n_motion += (int64_t)m.topo.size();
      tri_base[g] = n_mtris;
      n_mtris += (int64_t)m.n_tris * m.keys;
      mbox_base[g] = n_mboxes;
      n_mboxes += (int64_t)m.topo.size() * m.keys;
    }
    prim_base[g] = n_prims;
    n_prims += m.n_tris;
  }
  // normals only for static meshes that carry them: slot space parallel to `tris`
  bool any_normals = false;
  for (int g = 0; g < G; g++) any_normals |= (!ctx->meshes[g].motion && !ctx->meshes[g].normalidx.empty());
  if (any_normals) n_normal_slots = n_tris;
  const SceneStage& S = ctx->scene;
  int64_t scene_node_base, scene_mbox_base = 0;
  if (!S.motion) {
    scene_node_base = n_static;
    n_static += (int64_t)S.nodes.size();
  } else {
    scene_node_base = n_motion;
    n_motion += (int64_t)S.topo.size();
    scene_mbox_base = n_mboxes;
    n_mboxes += (int64_t)S.topo.size() * S.keys;
  }
  if (n_tris > (int64_t)kLeafBaseMask || n_mtris > (int64_t)kLeafBaseMask)
    return ctx->fail(VG_ERR_UNSUPPORTED, "scene exceeds 2^25 triangle slots");
  if (n_static + n_motion >= (int64_t)kGeomRootMask) return ctx->fail(VG_ERR_UNSUPPORTED, "too many nodes");
  auto motion_global = [&](int64_t rel) { return (int32_t)(n_static + rel); };
  auto mesh_root = [&](int g) -> int32_t {
    return ctx->meshes[g].motion ? motion_global(node_base[g]) : (int32_t)node_base[g];
  };

  std::vector<DevNode> nodes((size_t)n_static);
  std::vector<DevMotionNode> mtopo((size_t)n_motion);
  std::vector<float4> mboxes((size_t)n_mboxes * 6);
  std::vector<float4> tris((size_t)n_tris * 3), mtris((size_t)n_mtris * 3), normals((size_t)n_normal_slots * 3);
  std::vector<DevGeom> geoms((size_t)G);
  std::vector<uint8_t> prim_material((size_t)n_prims, 255);

  auto put_static_node = [&](DevNode& d, const VgNode& s, const int32_t c[4]) {
    std::memcpy(&d, &s, 96);
    d.m0 = make_uint4(s.axis0, s.axis1, s.axis2, (uint32_t)c[0]);
    d.m1 = make_uint4((uint32_t)c[1], (uint32_t)c[2], (uint32_t)c[3], 0u);
  };

  for (int g = 0; g < G; g++) {
    const MeshStage& m = ctx->meshes[g];
    DevGeom& dg = geoms[g];
    dg.tri_base = (int32_t)tri_base[g];
    dg.prim_base = (int32_t)prim_base[g];
    dg.normal_base = -1;
    dg.keys = m.motion ? m.keys : 1;
    dg.tri_key_stride = m.n_tris;
    dg.n_tris = m.n_tris;
    dg.pad0 = dg.pad1 = 0;
    if (m.material_ids.size() > 255) return ctx->fail(VG_ERR_UNSUPPORTED, "more than 255 shaders on one mesh");

    if (m.instance) continue;  // filled from the target below
    if (m.sphere) {
      dg.keys = 0;
      float4* t = &tris[(size_t)tri_base[g] * 3];
      t[0] = make_float4(m.centre[0], m.centre[1], m.centre[2], i2f((int32_t)g));
      t[1] = make_float4(m.radius, 0.f, 0.f, i2f(0));
      t[2] = make_float4(0.f, 0.f, 0.f, 0.f);
      prim_material[(size_t)prim_base[g]] = (uint8_t)m.material_ids[0];
    } else if (!m.motion) {
      for (size_t i = 0; i < m.nodes.size(); i++) {
        const VgNode& s = m.nodes[i];
        int32_t c[4];
        for (int k = 0; k < 4; k++) {
          const int32_t ch = s.children[k];
          if (ch >= 0) {
            if (ch >= (int32_t)m.nodes.size()) return ctx->fail(VG_ERR_INVALID, "child index out of range");
            c[k] = (int32_t)(node_base[g] + ch);
          } else if (ch == -1) {
            c[k] = -1;
          } else {
            const int first = refLeafBase(ch), count = refLeafCount(ch);
            if (first + count > m.n_tris) return ctx->fail(VG_ERR_INVALID, "leaf range outside the mesh (reference LeafBase decodes 23 bits: meshes must have < 2^23 triangles)");
            c[k] = (int32_t)(kLeafBit | ((uint32_t)(tri_base[g] + first) << 4) | (uint32_t)(count - 1));
          }
        }
        put_static_node(nodes[(size_t)node_base[g] + i], s, c);
      }
      // polymesh/trace.go:182: bias term (EpsilonFloat32 + RayBias), one float32 add
      const float bias = 1.19209290E-07f + m.raybias;
      for (int i = 0; i < m.n_tris; i++) {
        float4* t = &tris[(size_t)(tri_base[g] + i) * 3];
        for (int j = 0; j < 3; j++) {
          const float* v = &m.verts[(size_t)m.idxp[(size_t)i * 3 + j] * 3];
          t[j] = make_float4(v[0], v[1], v[2], 0.f);
        }
        t[0].w = i2f((int32_t)g);
        t[1].w = i2f((int32_t)i);
        t[2].w = bias;
        if (!m.material_ids.empty()) {
          const int si = m.shaderidx.empty() ? 0 : m.shaderidx[i];
          if (si < (int)m.material_ids.size()) prim_material[(size_t)prim_base[g] + i] = (uint8_t)m.material_ids[si];
        }
      }
      if (!m.normalidx.empty()) {
        dg.normal_base = (int32_t)tri_base[g];
        for (int i = 0; i < m.n_tris; i++)
          for (int j = 0; j < 3; j++) {
            const float* v = &m.normals[(size_t)m.normalidx[(size_t)i * 3 + j] * 3];
            normals[(size_t)(tri_base[g] + i) * 3 + j] = make_float4(v[0], v[1], v[2], 0.f);
          }
      }
    } else {
      const int nn = (int)m.topo.size();
      for (int i = 0; i < nn; i++) {
        const VgMotionNode& s = m.topo[i];
        DevMotionNode& d = mtopo[(size_t)node_base[g] + i];
        for (int k = 0; k < 4; k++) {
          const int32_t ch = s.children[k];
          if (ch >= 0) {
            if (ch >= nn) return ctx->fail(VG_ERR_INVALID, "child index out of range");
            d.child[k] = motion_global(node_base[g] + ch);
          } else if (ch == -1) {
            d.child[k] = -1;
          } else {
            const int first = refLeafBase(ch), count = refLeafCount(ch);
            if (first + count > m.n_tris) return ctx->fail(VG_ERR_INVALID, "leaf range outside the mesh");
            d.child[k] = (int32_t)(kLeafBit | kMotionTriBit | ((uint32_t)(tri_base[g] + first) << 4) | (uint32_t)(count - 1));
          }
        }
        d.axes_keys = (uint32_t)(s.axis0 & 3) | ((uint32_t)(s.axis1 & 3) << 2) | ((uint32_t)(s.axis2 & 3) << 4) | ((uint32_t)m.keys << 8);
        d.box_base = (uint32_t)(mbox_base[g] + i);
        d.box_key_stride = (uint32_t)nn;
        d.tri_key_stride = (uint32_t)m.n_tris;
        for (int k = 0; k < m.keys; k++)
          std::memcpy(&mboxes[(size_t)(mbox_base[g] + (int64_t)k * nn + i) * 6], &m.boxes[((size_t)k * nn + i) * 24], 96);
      }
      for (int i = 0; i < m.n_tris; i++) {
        // quirk (b): the reference tests face i at leaf slot i although the leaf box bounds face accel_idx[i]
        const int f = m.ref_compat ? i : m.accel_idx[i];
        for (int k = 0; k < m.keys; k++) {
          float4* t = &mtris[(size_t)(tri_base[g] + (int64_t)k * m.n_tris + i) * 3];
          for (int j = 0; j < 3; j++) {
            const float* v = &m.verts[((size_t)k * m.n_verts + m.idxp[(size_t)f * 3 + j]) * 3];
            t[j] = make_float4(v[0], v[1], v[2], 0.f);
          }
          t[0].w = i2f((int32_t)g);
          t[1].w = i2f((int32_t)f);
          t[2].w = m.raybias;  // trace.go:612: `<= RayBias*|det|`
        }
        if (!m.material_ids.empty()) {
          const int si = m.shaderidx.empty() ? 0 : m.shaderidx[f];
          if (si < (int)m.material_ids.size()) prim_material[(size_t)prim_base[g] + f] = (uint8_t)m.material_ids[si];
        }
      }
    }
  }

  // ---- instances: the hit record of an instance is the target mesh's, under the instance's geom id ----
  std::vector<DevXform> xforms;
  std::vector<XfSRT> xf_keys;
  std::vector<Mat4> xf_static;
  std::vector<int> xform_of_geom((size_t)G, -1);
  for (int g = 0; g < G; g++) {
    const MeshStage& m = ctx->meshes[g];
    if (!m.instance) continue;
    geoms[g] = geoms[m.target];
    DevXform x;
    x.root = mesh_root(m.target);
    x.geom = g;
    x.nkeys = (int32_t)m.srt.size();
    x.key_base = (int32_t)xf_keys.size();
    for (const VgTransformSRT& k : m.srt) {
      XfSRT d;
      std::memcpy(&d, &k, sizeof(d));
      xf_keys.push_back(d);
    }
    Mat4 M, Minv;
    xf_matrices(&xf_keys[(size_t)x.key_base], 1, 0.0f, &M, &Minv);  // what every ray computes when there is one key
    xf_static.push_back(M);
    xf_static.push_back(Minv);
    xform_of_geom[g] = (int)xforms.size();
    xforms.push_back(x);
  }
  if (!xforms.empty() && n_static + n_motion >= (int64_t)kXformMask) return ctx->fail(VG_ERR_UNSUPPORTED, "too many nodes for a scene with instances");

  // ---- scene level: leaves (leafMax = 1) become links to the mesh roots ----
  auto scene_child = [&](int32_t ch, int nn, int64_t base, bool motion, int32_t* out) -> bool {
    if (ch >= 0) {
      if (ch >= nn) return false;
      *out = motion ? motion_global(base + ch) : (int32_t)(base + ch);
    } else if (ch == -1) {
      *out = -1;
    } else {
      const int first = refLeafBase(ch), count = refLeafCount(ch);
      if (count != 1 || first >= (int)S.geom_of_slot.size()) return false;
      const int g = S.geom_of_slot[first];
      if (g < 0 || g >= G) return false;
      if (ctx->meshes[g].instance) *out = (int32_t)(kLeafBit | kGeomBit | kXformBit | (uint32_t)xform_of_geom[g]);
      else if (ctx->meshes[g].sphere) *out = (int32_t)(kLeafBit | kGeomBit | kSphereBit | (uint32_t)tri_base[g]);
      else *out = (int32_t)(kLeafBit | kGeomBit | (uint32_t)mesh_root(g));
    }
    return true;
  };
  if (!S.motion) {
    const int nn = (int)S.nodes.size();
    for (int i = 0; i < nn; i++) {
      int32_t c[4];
      for (int k = 0; k < 4; k++)
        if (!scene_child(S.nodes[i].children[k], nn, scene_node_base, false, &c[k])) return ctx->fail(VG_ERR_INVALID, "bad scene-level child link (leaves must hold exactly one geom)");
      put_static_node(nodes[(size_t)scene_node_base + i], S.nodes[i], c);
    }
  } else {
    const int nn = (int)S.topo.size();
    for (int i = 0; i < nn; i++) {
      DevMotionNode& d = mtopo[(size_t)scene_node_base + i];
      for (int k = 0; k < 4; k++)
        if (!scene_child(S.topo[i].children[k], nn, scene_node_base, true, &d.child[k])) return ctx->fail(VG_ERR_INVALID, "bad scene-level child link (leaves must hold exactly one geom)");
      d.axes_keys = (uint32_t)(S.topo[i].axis0 & 3) | ((uint32_t)(S.topo[i].axis1 & 3) << 2) | ((uint32_t)(S.topo[i].axis2 & 3) << 4) | ((uint32_t)S.keys << 8);
      d.box_base = (uint32_t)(scene_mbox_base + i);
      d.box_key_stride = (uint32_t)nn;
      d.tri_key_stride = 0;
      for (int k = 0; k < S.keys; k++)
        std::memcpy(&mboxes[(size_t)(scene_mbox_base + (int64_t)k * nn + i) * 6], &S.boxes[((size_t)k * nn + i) * 24], 96);
    }
  }

  // ---- to HBM ----
  VG_CUDA(ctx, cudaSetDevice(ctx->device));
  VG_CUDA(ctx, upload(ctx->d_nodes, nodes, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_mtopo, mtopo, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_mboxes, mboxes, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_tris, tris, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_mtris, mtris, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_normals, normals, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_geoms, geoms, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_prim_material, prim_material, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_xforms, xforms, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_xf_keys, xf_keys, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_xf_static, xf_static, ctx->stream));
  VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->scene_bytes = nodes.size() * sizeof(DevNode) + mtopo.size() * sizeof(DevMotionNode) + (mboxes.size() + tris.size() + mtris.size() + normals.size()) * sizeof(float4);

  DevScene& d = ctx->dev;
  d.nodes = ctx->d_nodes.p;
  d.mtopo = ctx->d_mtopo.p;
  d.mboxes = ctx->d_mboxes.p;
  d.tris = ctx->d_tris.p;
  d.mtris = ctx->d_mtris.p;
  d.geoms = ctx->d_geoms.p;
  d.prim_material = ctx->d_prim_material.p;
  d.tri_normals = any_normals ? ctx->d_normals.p : nullptr;
  d.n_static = (int32_t)n_static;
  d.root = S.motion ? motion_global(scene_node_base) : (int32_t)scene_node_base;
  d.xforms = ctx->d_xforms.p;
  d.xf_keys = ctx->d_xf_keys.p;
  d.xf_static = ctx->d_xf_static.p;
  d.n_xforms = (int32_t)xforms.size();
  d.n_mtris = (int32_t)n_mtris;
  d.n_geoms = G;
  d.n_spheres = 0;
  for (int g = 0; g < G; g++) d.n_spheres += ctx->meshes[g].sphere ? 1 : 0;
  ctx->committed = true;
  render_invalidate(ctx);
  return VG_OK;
}

int vg_set_materials(vg_ctx* ctx, const VgMaterial* mats, int n) {
  VG_LOCK(ctx);
  if (n < 0 || (n > 0 && !mats)) return ctx->fail(VG_ERR_INVALID, "vg_set_materials: bad input");
  if (n > 254) return ctx->fail(VG_ERR_UNSUPPORTED, "more than 254 materials");
  ctx->materials.assign(mats, mats + n);
  render_invalidate(ctx);
  return VG_OK;
}

int vg_set_lights(vg_ctx* ctx, const VgTriLight* lights, int n) {
  VG_LOCK(ctx);
  if (n < 0 || (n > 0 && !lights)) return ctx->fail(VG_ERR_INVALID, "vg_set_lights: bad input");
  ctx->lights.clear();
  for (int i = 0; i < n; i++) {
    VgLight l{};
    l.type = VG_LIGHT_TRI;
    l.samples = lights[i].samples;
    l.material = lights[i].material;
    l.geom = lights[i].geom;
    std::memcpy(l.p0, lights[i].p0, 12);
    std::memcpy(l.p1, lights[i].p1, 12);
    std::memcpy(l.p2, lights[i].p2, 12);
    ctx->lights.push_back(l);
  }
  render_invalidate(ctx);
  return VG_OK;
}

int vg_set_area_lights(vg_ctx* ctx, const VgLight* lights, int n) {
  VG_LOCK(ctx);
  if (n < 0 || (n > 0 && !lights)) return ctx->fail(VG_ERR_INVALID, "vg_set_area_lights: bad input");
  for (int i = 0; i < n; i++)
    if (lights[i].type < VG_LIGHT_TRI || lights[i].type > VG_LIGHT_SPHERE) return ctx->fail(VG_ERR_INVALID, "vg_set_area_lights: unknown light type");
  ctx->lights.assign(lights, lights + n);
  render_invalidate(ctx);
  return VG_OK;
}

int vg_set_camera(vg_ctx* ctx, const VgCamera* cam) {
  VG_LOCK(ctx);
  if (!cam) return ctx->fail(VG_ERR_INVALID, "vg_set_camera: null");
  ctx->camera = *cam;
  ctx->have_camera = true;
  ctx->cam_nkeys = 0;
  return VG_OK;
}

int vg_set_camera_motion(vg_ctx* ctx, const VgCamera* cam, const VgTransformSRT* decomp, int keys) {
  VG_LOCK(ctx);
  if (!cam) return ctx->fail(VG_ERR_INVALID, "vg_set_camera_motion: null");
  if (keys < 0 || keys > 255 || (keys > 0 && !decomp)) return ctx->fail(VG_ERR_INVALID, "vg_set_camera_motion: bad keys");
  static_assert(sizeof(VgTransformSRT) == sizeof(vg::XfSRT), "VgTransformSRT is m.TransformDecomp's 23 floats");
  ctx->camera = *cam;
  ctx->have_camera = true;
  ctx->cam_nkeys = keys > 1 ? keys : 0;
  if (keys > 1) {
    VG_CUDA(ctx, cudaSetDevice(ctx->device));
    VG_CUDA(ctx, ctx->d_cam_keys.reserve((size_t)keys));
    VG_CUDA(ctx, cudaMemcpyAsync(ctx->d_cam_keys.p, decomp, (size_t)keys * sizeof(vg::XfSRT), cudaMemcpyHostToDevice, ctx->stream));
    VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return VG_OK;
}

int vg_set_frame(vg_ctx* ctx, int xres, int yres) {
  VG_LOCK(ctx);
  // RasterXY(12, ...) stratifies a 2^12 x 2^12 grid (math/ldseq/raster.go:49)
  if (xres <= 0 || yres <= 0 || xres > 4096 || yres > 4096) return ctx->fail(VG_ERR_INVALID, "frame must be within 4096x4096 (RasterXY m=12)");
  ctx->xres = xres;
  ctx->yres = yres;
  render_invalidate(ctx);
  return VG_OK;
}

int vg_set_partition(vg_ctx* ctx, int rank, int world) {
  VG_LOCK(ctx);
  if (world <= 0 || rank < 0 || rank >= world) return ctx->fail(VG_ERR_INVALID, "vg_set_partition: bad rank/world");
  ctx->rank = rank;
  ctx->world = world;
  render_invalidate(ctx);
  return VG_OK;
}

int vg_set_scramble(vg_ctx* ctx, const uint64_t* table, int64_t npix) {
  VG_LOCK(ctx);
  if (!table || npix <= 0) return ctx->fail(VG_ERR_INVALID, "vg_set_scramble: bad input");
  VG_CUDA(ctx, cudaSetDevice(ctx->device));
  return render_set_scramble(ctx, table, npix);
}

int vg_set_filter(vg_ctx* ctx, int n, double w, const double* cdfV, const double* cdfVU) {
  VG_LOCK(ctx);
  if (n == 0) {
    ctx->filter_n = 0;
    ctx->filter_cdf.clear();
    render_invalidate(ctx);
    return VG_OK;
  }
  if (n < 2 || n > 1024 || !(w > 0) || !cdfV || !cdfVU) return ctx->fail(VG_ERR_INVALID, "vg_set_filter: bad input");
  ctx->filter_n = n;
  ctx->filter_w = w;
  ctx->filter_cdf.assign(cdfV, cdfV + n);
  ctx->filter_cdf.insert(ctx->filter_cdf.end(), cdfVU, cdfVU + (size_t)n * n);
  render_invalidate(ctx);
  return VG_OK;
}

int vg_set_option(vg_ctx* ctx, const char* name, int value) {
  VG_LOCK(ctx);
  if (!name) return ctx->fail(VG_ERR_INVALID, "null option name");
  if (!std::strcmp(name, "trace_last_level")) ctx->opt_trace_last_level = value != 0;
  else if (!std::strcmp(name, "precise_trig")) ctx->opt_precise_trig = value != 0;
  else if (!std::strcmp(name, "tma_stage")) ctx->opt_traversal = value != 0 ? 1 : 2;
  else if (!std::strcmp(name, "primary_per_lane")) ctx->opt_primary_per_lane = value != 0;
  else if (!std::strcmp(name, "shadow_unordered")) ctx->opt_shadow_unordered = value != 0;
  else if (!std::strcmp(name, "generic_shade")) {
    ctx->opt_generic_shade = value != 0;
    render_invalidate(ctx);
  }
  else if (!std::strcmp(name, "pixel_block")) {
    ctx->opt_pixel_block = value != 0;
    render_invalidate(ctx);
  }
  else if (!std::strcmp(name, "traversal")) {
    if (value < 0 || value > 2) return ctx->fail(VG_ERR_INVALID, "traversal must be 0 (per-lane), 1 (TMA-staged queue) or 2 (cooperative leaves)");
    ctx->opt_traversal = value;
  }
  else if (!std::strcmp(name, "iters_per_batch")) {
    if (value < 1 || value > 64) return ctx->fail(VG_ERR_INVALID, "iters_per_batch must be in [1,64]");
    ctx->opt_iters_per_batch = value;
    render_invalidate(ctx);
  } else
    return ctx->fail(VG_ERR_INVALID, std::string("unknown option ") + name);
  return VG_OK;
}

static int trace_device_locked(vg_ctx* ctx, const VgRay* d_rays, int64_t n, VgHit* d_hits, uint32_t flags) {
  if (!ctx->committed) return ctx->fail(VG_ERR_INVALID, "scene not committed");
  if (n == 0) return VG_OK;
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) blocks_per_sm = trace_batch_blocks_per_sm();
  long long want = (n + kTraceBlock - 1) / kTraceBlock;
  long long grid = (long long)ctx->sm_count * blocks_per_sm;
  if (grid > want) grid = want;
  VG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  VG_CUDA(ctx, launch_trace_batch(ctx->dev, d_rays, d_hits, n, (flags & VG_TRACE_ANY_HIT) != 0, ctx->opt_traversal, ctx->d_counters.p, ctx->d_counters.p + 1, (int)grid, ctx->stream));
  VG_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  ctx->stats.trace_ms = ms;
  ctx->stats.rays += (uint64_t)n;
  if (flags & VG_TRACE_ANY_HIT) ctx->stats.shadow_rays += (uint64_t)n;
  ctx->stats.kernel_launches += 1;
  return VG_OK;
}

int vg_trace_batch_device(vg_ctx* ctx, const VgRay* d_rays, int64_t n, VgHit* d_hits, uint32_t flags) {
  VG_LOCK(ctx);
  if (n < 0 || (n > 0 && (!d_rays || !d_hits))) return ctx->fail(VG_ERR_INVALID, "vg_trace_batch_device: bad input");
  VG_CUDA(ctx, cudaSetDevice(ctx->device));
  return trace_device_locked(ctx, d_rays, n, d_hits, flags);
}

static bool pinned_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

int vg_trace_batch(vg_ctx* ctx, const VgRay* rays, int64_t n, VgHit* hits, uint32_t flags) {
  VG_LOCK(ctx);
  if (n < 0 || (n > 0 && (!rays || !hits))) return ctx->fail(VG_ERR_INVALID, "vg_trace_batch: bad input");
  if (n == 0) return VG_OK;
  VG_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->committed) return ctx->fail(VG_ERR_INVALID, "scene not committed");
  VG_CUDA(ctx, ctx->d_rays.reserve((size_t)n));
  VG_CUDA(ctx, ctx->d_hits.reserve((size_t)n));
  const int64_t chunk = 1 << 19;  // 16 MB of rays
  if (n >= 2 * chunk && pinned_host(rays) && pinned_host(hits)) {
    // Page-locked caller buffers: the batch goes through in chunks on three streams, so the H2D copy of one chunk, the
    // traversal of the previous one and the D2H copy of the one before overlap (PCIe is full duplex; each direction carries
    // 32 B per ray, which at ~50 GB/s is slower than the traversal itself).
    for (int s = 0; s < 3; s++) {
      if (!ctx->pipe_stream[s]) VG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->pipe_stream[s], cudaStreamNonBlocking));
      if (!ctx->pipe_done[s]) VG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pipe_done[s], cudaEventDisableTiming));
    }
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) blocks_per_sm = trace_batch_blocks_per_sm();
    VG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    for (int s = 0; s < 3; s++) VG_CUDA(ctx, cudaStreamWaitEvent(ctx->pipe_stream[s], ctx->ev0, 0));
    int c = 0;
    for (int64_t off = 0; off < n; off += chunk, c++) {
      const int64_t m = std::min(chunk, n - off);
      cudaStream_t st = ctx->pipe_stream[c % 3];
      VG_CUDA(ctx, cudaMemcpyAsync(ctx->d_rays.p + off, rays + off, (size_t)m * sizeof(VgRay), cudaMemcpyHostToDevice, st));
      long long grid = std::min<long long>((long long)ctx->sm_count * blocks_per_sm, (m + kTraceBlock - 1) / kTraceBlock);
      VG_CUDA(ctx, launch_trace_batch(ctx->dev, ctx->d_rays.p + off, ctx->d_hits.p + off, m, (flags & VG_TRACE_ANY_HIT) != 0, ctx->opt_traversal,
                                      ctx->d_counters.p + 4 + (c % 3), ctx->d_counters.p + 1, (int)grid, st));
      VG_CUDA(ctx, cudaMemcpyAsync(hits + off, ctx->d_hits.p + off, (size_t)m * sizeof(VgHit), cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < 3; s++) {
      VG_CUDA(ctx, cudaEventRecord(ctx->pipe_done[s], ctx->pipe_stream[s]));
      VG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipe_done[s], 0));
    }
    VG_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->stats.trace_ms = ms;  // copies included: the whole pipelined call
    ctx->stats.rays += (uint64_t)n;
    if (flags & VG_TRACE_ANY_HIT) ctx->stats.shadow_rays += (uint64_t)n;
    ctx->stats.kernel_launches += (uint64_t)c;
    return VG_OK;
  }
  VG_CUDA(ctx, cudaMemcpyAsync(ctx->d_rays.p, rays, (size_t)n * sizeof(VgRay), cudaMemcpyHostToDevice, ctx->stream));
  int rc = trace_device_locked(ctx, ctx->d_rays.p, n, ctx->d_hits.p, flags);
  if (rc != VG_OK) return rc;
  VG_CUDA(ctx, cudaMemcpyAsync(hits, ctx->d_hits.p, (size_t)n * sizeof(VgHit), cudaMemcpyDeviceToHost, ctx->stream));
  VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VG_OK;
}

int vg_render(vg_ctx* ctx, int iter_begin, int iter_end, float* fb_out) {
  VG_LOCK(ctx);
  VG_CUDA(ctx, cudaSetDevice(ctx->device));
  return render_run(ctx, iter_begin, iter_end, fb_out);
}

int vg_clear_framebuffer(vg_ctx* ctx) {
  VG_LOCK(ctx);
  VG_CUDA(ctx, cudaSetDevice(ctx->device));
  return render_clear(ctx);
}

int vg_framebuffer_device(vg_ctx* ctx, float** d_fb) {
  VG_LOCK(ctx);
  if (!d_fb) return ctx->fail(VG_ERR_INVALID, "null out pointer");
  VG_CUDA(ctx, cudaSetDevice(ctx->device));
  return render_fb_device(ctx, d_fb);
}

int vg_get_stats(vg_ctx* ctx, VgStats* out) {
  VG_LOCK(ctx);
  if (!out) return ctx->fail(VG_ERR_INVALID, "null out pointer");
  unsigned long long c[3] = {0, 0, 0};
  cudaSetDevice(ctx->device);
  cudaMemcpy(c, ctx->d_counters.p, sizeof(c), cudaMemcpyDeviceToHost);
  *out = ctx->stats;
  out->nodes_t += c[1];  // batch-trace kernels accumulate on the device; the render adds its totals on the host
  out->tris_t += c[2];
  return VG_OK;
}

int vg_reset_stats(vg_ctx* ctx) {
  VG_LOCK(ctx);
  cudaSetDevice(ctx->device);
  cudaMemsetAsync(ctx->d_counters.p, 0, 16 * sizeof(unsigned long long), ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  ctx->stats = VgStats{};
  return VG_OK;
}

}  // extern "C"
