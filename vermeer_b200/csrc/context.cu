// Device layer of the C ABI (include/vermeer_gpu.h, vg_*): context, scene flattening, batch TraceProbe.
#include "context.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

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

// Host array WITHOUT value initialisation: the flattened scene is written exactly once, geom by geom on several threads, so
// zero-filling 0.5 GB first (what std::vector does) only adds a single-threaded pass of page faults (C3: 0.31 of 0.47 s).
template <class T>
struct HostArr {
  std::unique_ptr<T[]> p;
  size_t n = 0;
  explicit HostArr(size_t count) : p(count ? new T[count] : nullptr), n(count) {}
  T* data() { return p.get(); }
  const T* data() const { return p.get(); }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};
// Pageable host memory -> device through page-locked slots: a few threads each copy their chunks into their own two 4-MB
// slots and issue the DMA of a chunk as soon as it is staged (cudaMemcpyAsync from pageable memory stages through ONE driver
// buffer at ~4 GB/s; C3's 0.55 GB of scene records took 140 ms that way).
struct StagedCopier {
  static const int kThreads = 4, kSlots = 2;
  static const size_t kChunk = (size_t)4 << 20;
  char* pinned = nullptr;
  cudaEvent_t ev[kThreads][kSlots] = {};
  cudaError_t init() {
    if (pinned) return cudaSuccess;
    cudaError_t e = cudaMallocHost((void**)&pinned, kChunk * kThreads * kSlots);
    if (e != cudaSuccess) { pinned = nullptr; return e; }
    for (int t = 0; t < kThreads; t++)
      for (int k = 0; k < kSlots; k++)
        if ((e = cudaEventCreateWithFlags(&ev[t][k], cudaEventDisableTiming)) != cudaSuccess) return e;
    return cudaSuccess;
  }
  cudaError_t copy(void* dst, const void* src, size_t bytes, cudaStream_t stream, int device) {
    cudaError_t e = init();
    if (e != cudaSuccess) return e;
    const size_t nchunks = (bytes + kChunk - 1) / kChunk;
    cudaError_t errs[kThreads];
    for (int t = 0; t < kThreads; t++) errs[t] = cudaSuccess;
    auto work = [&](int t) {
      if (t > 0) cudaSetDevice(device);
      int use = 0;
      for (size_t c = (size_t)t; c < nchunks; c += kThreads, use++) {
        const int k = use % kSlots;
        if (use >= kSlots) cudaEventSynchronize(ev[t][k]);  // the DMA that last read this slot
        char* slot = pinned + ((size_t)t * kSlots + k) * kChunk;
        const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
        std::memcpy(slot, (const char*)src + off, len);
        cudaError_t e2 = cudaMemcpyAsync((char*)dst + off, slot, len, cudaMemcpyHostToDevice, stream);
        if (e2 == cudaSuccess) e2 = cudaEventRecord(ev[t][k], stream);
        if (e2 != cudaSuccess) errs[t] = e2;
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < kThreads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    for (int t = 0; t < kThreads; t++)
      if (errs[t] != cudaSuccess) return errs[t];
    return cudaStreamSynchronize(stream);  // the slots are reused by the next call
  }
  void release() {
    if (!pinned) return;
    cudaFreeHost(pinned);
    pinned = nullptr;
    for (int t = 0; t < kThreads; t++)
      for (int k = 0; k < kSlots; k++)
        if (ev[t][k]) cudaEventDestroy(ev[t][k]);
  }
};
StagedCopier g_copier;  // used under a context's mutex; contexts of one process share the slots, so serialise on this too
std::mutex g_copier_mu;

template <class T>
cudaError_t upload(DevBuf<T>& buf, const HostArr<T>& v, cudaStream_t s, int device = -1) {
  cudaError_t e = buf.reserve(v.size());
  if (e != cudaSuccess) return e;
  if (v.empty()) return cudaSuccess;
  const size_t bytes = v.size() * sizeof(T);
  if (device >= 0 && bytes >= ((size_t)16 << 20)) {
    std::lock_guard<std::mutex> lock(g_copier_mu);
    return g_copier.copy(buf.p, v.data(), bytes, s, device);
  }
  return cudaMemcpyAsync(buf.p, v.data(), bytes, cudaMemcpyHostToDevice, s);
}

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
  comm_destroy(ctx);
  render_destroy(ctx);
  build_scratch_destroy(ctx);
  ctx->d_nodes.release(); ctx->d_mtopo.release(); ctx->d_mboxes.release(); ctx->d_tris.release();
  ctx->d_mtris.release(); ctx->d_normals.release(); ctx->d_geoms.release(); ctx->d_prim_material.release();
  ctx->d_tri_uv.release(); ctx->d_texels.release(); ctx->d_tex_levels.release(); ctx->d_textures.release();
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
  // VG_TIMING=1: stage times of the commit on stderr (scene preparation is outside the benchmarked path, but not free)
  const bool timing = std::getenv("VG_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[vg_scene_commit] %-22s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  if (!ctx->scene.present) return ctx->fail(VG_ERR_INVALID, "vg_scene_commit: no scene-level tree uploaded");
  const int G = (int)ctx->meshes.size();
  for (int g = 0; g < G; g++)
    if (!ctx->meshes[g].present) return ctx->fail(VG_ERR_INVALID, "vg_scene_commit: geom " + std::to_string(g) + " was not uploaded");

  for (int g = 0; g < G; g++) {
    const MeshStage& m = ctx->meshes[g];
    if (!m.instance) continue;
    const MeshStage& tg = ctx->meshes[m.target];
    if (tg.sphere) return ctx->fail(VG_ERR_UNSUPPORTED, "GeomInstance of a Sphere is outside this path (PolyMesh and GeomInstance targets only)");
    {  // an instance of an instance of ...: the chain must end at a mesh
      int t = m.target, hops = 0;
      while (t >= 0 && t < G && ctx->meshes[t].present && ctx->meshes[t].instance && hops <= G) { t = ctx->meshes[t].target; hops++; }
      if (t < 0 || t >= G || !ctx->meshes[t].present || ctx->meshes[t].instance || ctx->meshes[t].sphere)
        return ctx->fail(VG_ERR_INVALID, "GeomInstance: the chain of instance targets does not end at a PolyMesh");
    }
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

  HostArr<DevNode> nodes((size_t)n_static);
  HostArr<DevMotionNode> mtopo((size_t)n_motion);
  HostArr<float4> mboxes((size_t)n_mboxes * 6);
  HostArr<float4> tris((size_t)n_tris * kTriStride), mtris((size_t)n_mtris * 3), normals((size_t)n_normal_slots * 3);
  bool any_uv = false;
  for (int g = 0; g < G; g++) any_uv |= (!ctx->meshes[g].motion && !ctx->meshes[g].uvtriidx.empty());
  HostArr<float2> tri_uv(any_uv ? (size_t)n_tris * 3 : 0);
  std::vector<DevGeom> geoms((size_t)G);
  HostArr<uint8_t> prim_material((size_t)n_prims);  // each geom fills its own range (255 = no shader) before it assigns

  lap("allocate host arrays");
  auto put_static_node = [&](DevNode& d, const VgNode& s, const int32_t c[4]) {
    std::memcpy(&d, &s, 96);
    d.m0 = make_uint4(s.axis0, s.axis1, s.axis2, (uint32_t)c[0]);
    d.m1 = make_uint4((uint32_t)c[1], (uint32_t)c[2], (uint32_t)c[3], 0u);
  };

  // one geom's nodes, triangle records, normals, UVs and material ids; geoms write disjoint ranges of the output arrays, so the
  // loop runs on several host threads (C3: 1024 meshes, 10 M triangles, 0.48 GB of records)
  struct FlatErr { int code; const char* msg; };
  auto flatten_geom = [&](int g) -> FlatErr {
    const MeshStage& m = ctx->meshes[g];
    DevGeom& dg = geoms[g];
    dg.tri_base = (int32_t)tri_base[g];
    dg.prim_base = (int32_t)prim_base[g];
    dg.normal_base = -1;
    dg.keys = m.motion ? m.keys : 1;
    dg.tri_key_stride = m.n_tris;
    dg.n_tris = m.n_tris;
    dg.uv_base = -1;
    dg.xform = -1;
    if (m.material_ids.size() > 255) return FlatErr{VG_ERR_UNSUPPORTED, "more than 255 shaders on one mesh"};
    if (!m.instance) std::memset(prim_material.data() + prim_base[g], 255, (size_t)(m.sphere ? 1 : m.n_tris));

    if (m.instance) return FlatErr{VG_OK, nullptr};  // filled from the target below
    if (m.sphere) {
      dg.keys = 0;
      float4* t = &tris[(size_t)tri_base[g] * kTriStride];
      t[0] = make_float4(m.centre[0], m.centre[1], m.centre[2], i2f((int32_t)g));
      t[1] = make_float4(m.radius, 0.f, 0.f, i2f(0));
      t[2] = make_float4(0.f, 0.f, 0.f, 0.f);
      prim_material[(size_t)prim_base[g]] = (uint8_t)m.material_ids[0];
    } else if (!m.motion) {
      // Device node order. The reference numbers nodes in depth-first preorder (child 0 next to its parent, children 1-3 a
      // subtree away). Option node_order=1 renumbers each mesh breadth-first, so the four children of a node are neighbours
      // and the top of the tree is one compact block; links are re-encoded anyway, results and NodesT do not depend on it.
      std::vector<int32_t> place;  // reference index -> position inside this mesh's block (identity when empty)
      if (ctx->opt_node_order == 1 && m.nodes.size() > 1) {
        place.assign(m.nodes.size(), -1);
        std::vector<int32_t> order;
        order.reserve(m.nodes.size());
        order.push_back(0);
        place[0] = 0;
        for (size_t q = 0; q < order.size(); q++)
          for (int k = 0; k < 4; k++) {
            const int32_t ch = m.nodes[(size_t)order[q]].children[k];
            if (ch >= 0 && ch < (int32_t)m.nodes.size() && place[(size_t)ch] < 0) {
              place[(size_t)ch] = (int32_t)order.size();
              order.push_back(ch);
            }
          }
        if (order.size() != m.nodes.size()) place.clear();  // unreachable nodes: keep the reference order
      }
      auto pos = [&](size_t i) { return place.empty() ? (int64_t)i : (int64_t)place[i]; };
      for (size_t i = 0; i < m.nodes.size(); i++) {
        const VgNode& s = m.nodes[i];
        int32_t c[4];
        for (int k = 0; k < 4; k++) {
          const int32_t ch = s.children[k];
          if (ch >= 0) {
            if (ch >= (int32_t)m.nodes.size()) return FlatErr{VG_ERR_INVALID, "child index out of range"};
            c[k] = (int32_t)(node_base[g] + pos((size_t)ch));
          } else if (ch == -1) {
            c[k] = -1;
          } else {
            const int first = refLeafBase(ch), count = refLeafCount(ch);
            if (first + count > m.n_tris) return FlatErr{VG_ERR_INVALID, "leaf range outside the mesh (reference LeafBase decodes 23 bits: meshes must have < 2^23 triangles)"};
            c[k] = (int32_t)(kLeafBit | ((uint32_t)(tri_base[g] + first) << 4) | (uint32_t)(count - 1));
          }
        }
        put_static_node(nodes[(size_t)node_base[g] + (size_t)pos(i)], s, c);
      }
      // polymesh/trace.go:182: bias term (EpsilonFloat32 + RayBias), one float32 add
      const float bias = 1.19209290E-07f + m.raybias;
      for (int i = 0; i < m.n_tris; i++) {
        float4* t = &tris[(size_t)(tri_base[g] + i) * kTriStride];
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
      if (!m.uvtriidx.empty()) {
        dg.uv_base = (int32_t)tri_base[g];
        for (int i = 0; i < m.n_tris; i++)
          for (int j = 0; j < 3; j++) {
            const float* v = &m.uv[(size_t)m.uvtriidx[(size_t)i * 3 + j] * 2];
            tri_uv[(size_t)(tri_base[g] + i) * 3 + j] = make_float2(v[0], v[1]);
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
            if (ch >= nn) return FlatErr{VG_ERR_INVALID, "child index out of range"};
            d.child[k] = motion_global(node_base[g] + ch);
          } else if (ch == -1) {
            d.child[k] = -1;
          } else {
            const int first = refLeafBase(ch), count = refLeafCount(ch);
            if (first + count > m.n_tris) return FlatErr{VG_ERR_INVALID, "leaf range outside the mesh"};
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
    return FlatErr{VG_OK, nullptr};
  };
  {
    std::vector<FlatErr> ferr((size_t)G, FlatErr{VG_OK, nullptr});
    const unsigned hw = std::thread::hardware_concurrency();
    const int nthreads = (int)std::min<int64_t>(std::min<unsigned>(hw ? hw : 1, 16), (n_tris + n_mtris) / 65536 + 1);
    if (nthreads >= 2 && G >= 2) {
      std::atomic<int> next(0);
      auto work = [&] {
        for (;;) {
          const int g = next.fetch_add(1);
          if (g >= G) return;
          ferr[(size_t)g] = flatten_geom(g);
        }
      };
      std::vector<std::thread> th;
      for (int t = 1; t < nthreads; t++) th.emplace_back(work);
      work();
      for (auto& t : th) t.join();
    } else {
      for (int g = 0; g < G; g++) ferr[(size_t)g] = flatten_geom(g);
    }
    for (int g = 0; g < G; g++)
      if (ferr[(size_t)g].code != VG_OK) return ctx->fail(ferr[(size_t)g].code, ferr[(size_t)g].msg);
  }

  lap("flatten geoms");
  // ---- instances: the hit record of an instance is the target mesh's, under the instance's geom id ----
  std::vector<DevXform> xforms;
  std::vector<XfSRT> xf_keys;
  std::vector<Mat4> xf_static;
  std::vector<int> xform_of_geom((size_t)G, -1);
  for (int g = 0; g < G; g++) {
    const MeshStage& m = ctx->meshes[g];
    if (!m.instance) continue;
    int final_target = m.target;  // the mesh at the end of the chain
    while (ctx->meshes[final_target].instance) final_target = ctx->meshes[final_target].target;
    geoms[g] = geoms[final_target];
    DevXform x;
    x.root = mesh_root(final_target);
    x.geom = g;
    x.inner = -1;
    x.pad_[0] = x.pad_[1] = x.pad_[2] = 0;
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
    geoms[g].xform = (int32_t)xforms.size();
    xforms.push_back(x);
  }
  for (int g = 0; g < G; g++) {  // chains: the xform applied after this one is the target instance's
    const MeshStage& m = ctx->meshes[g];
    if (m.instance && ctx->meshes[m.target].instance) xforms[(size_t)xform_of_geom[g]].inner = xform_of_geom[m.target];
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
  lap("scene level");
  VG_CUDA(ctx, cudaSetDevice(ctx->device));
  VG_CUDA(ctx, upload(ctx->d_nodes, nodes, ctx->stream, ctx->device));
  VG_CUDA(ctx, upload(ctx->d_mtopo, mtopo, ctx->stream, ctx->device));
  VG_CUDA(ctx, upload(ctx->d_mboxes, mboxes, ctx->stream, ctx->device));
  VG_CUDA(ctx, upload(ctx->d_tris, tris, ctx->stream, ctx->device));
  VG_CUDA(ctx, upload(ctx->d_mtris, mtris, ctx->stream, ctx->device));
  VG_CUDA(ctx, upload(ctx->d_normals, normals, ctx->stream, ctx->device));
  VG_CUDA(ctx, upload(ctx->d_tri_uv, tri_uv, ctx->stream, ctx->device));
  VG_CUDA(ctx, upload(ctx->d_geoms, geoms, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_prim_material, prim_material, ctx->stream, ctx->device));
  VG_CUDA(ctx, upload(ctx->d_xforms, xforms, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_xf_keys, xf_keys, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_xf_static, xf_static, ctx->stream));
  VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  lap("host -> device");
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
  d.tri_uv = any_uv ? ctx->d_tri_uv.p : nullptr;
  d.n_static = (int32_t)n_static;
  d.root = S.motion ? motion_global(scene_node_base) : (int32_t)scene_node_base;
  d.xforms = ctx->d_xforms.p;
  d.xf_keys = ctx->d_xf_keys.p;
  d.xf_static = ctx->d_xf_static.p;
  d.n_xforms = (int32_t)xforms.size();
  d.n_mtris = (int32_t)n_mtris;
  ctx->n_tri_slots = n_tris;
  d.n_geoms = G;
  d.n_spheres = 0;
  for (int g = 0; g < G; g++) d.n_spheres += ctx->meshes[g].sphere ? 1 : 0;
  // Option l2_persist_nodes: pin the static node array in the persisting part of L2 (access-policy window on the stream the
  // render kernels run on), so that triangle records streaming through cannot evict the tree of a scene larger than L2.
  if (ctx->opt_l2_persist_nodes && n_static > 0) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
      const size_t bytes = (size_t)n_static * sizeof(DevNode);
      const size_t carve = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, bytes);
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
      cudaStreamAttrValue attr;
      std::memset(&attr, 0, sizeof(attr));
      attr.accessPolicyWindow.base_ptr = (void*)ctx->d_nodes.p;
      attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
      attr.accessPolicyWindow.hitRatio = bytes <= carve ? 1.0f : (float)((double)carve / (double)bytes);
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaGetLastError();
    }
  }
  ctx->committed = true;
  render_invalidate(ctx);
  return VG_OK;
}

int vg_set_materials(vg_ctx* ctx, const VgMaterial* mats, int n) {
  VG_LOCK(ctx);
  if (n < 0 || (n > 0 && !mats)) return ctx->fail(VG_ERR_INVALID, "vg_set_materials: bad input");
  if (n > 254) return ctx->fail(VG_ERR_UNSUPPORTED, "more than 254 materials");
  ctx->materials.assign(mats, mats + n);
  ctx->mat_tex.assign((size_t)n, MatTex());
  render_invalidate(ctx);
  return VG_OK;
}

// ---- texture store ----------------------------------------------------------------------------------------------------
int vg_texture_upload(vg_ctx* ctx, const uint8_t* rgb8, int w, int h, int* tex_id) {
  VG_LOCK(ctx);
  if (!rgb8 || w <= 0 || h <= 0 || !tex_id) return ctx->fail(VG_ERR_INVALID, "vg_texture_upload: null/empty input");
  if (w > 32768 || h > 32768) return ctx->fail(VG_ERR_UNSUPPORTED, "vg_texture_upload: image larger than 32768 texels on a side");
  // maxlevel := int(Ceil(Log2(Max(w, h)))) (mipmap.go:123): exact for integers
  int maxlevel = 0;
  while ((1 << maxlevel) < std::max(w, h)) maxlevel++;
  if (maxlevel <= 0) return ctx->fail(VG_ERR_INVALID, "vg_texture_upload: a 1x1 image has no mip level (the reference panics, mipmap.go:127)");
  VG_CUDA(ctx, cudaSetDevice(ctx->device));
  DevTexture T;
  T.w = w; T.h = h;
  T.first_level = (int32_t)ctx->tex_levels.size();
  T.n_levels = maxlevel;
  std::vector<DevTexLevel> lv((size_t)maxlevel);
  size_t off = ctx->n_texels;
  int cw = w, ch = h;
  for (int l = 0; l < maxlevel; l++) {
    if (l > 0) {  // nwidth := int(Max(1, Ceil(width / 2))) (mipmap.go:138-139)
      // a level filtered FROM a 1-texel-high level reads row 1 (y2 := mini(2y+1, maxi(1, height-1)), mipmap.go:205,245; the x
      // axis has a wrap rule, :179-182, the y axis has none): the reference panics with an index out of range, i.e. it cannot
      // load images that are 4x or more wider than high (roughly: ceil(log2 w) - ceil(log2 h) >= 2)
      if (ch == 1)
        return ctx->fail(VG_ERR_INVALID, "vg_texture_upload: image much wider than high (the reference's stdfilter indexes row 1 of a 1-row level, mipmap.go:205,245)");
      cw = std::max(1, (cw + 1) / 2);
      ch = std::max(1, (ch + 1) / 2);
    }
    lv[(size_t)l] = DevTexLevel{(uint32_t)off, cw, ch, 0};
    off += (size_t)cw * ch;
  }
  if (off > 0xffffffffull) return ctx->fail(VG_ERR_UNSUPPORTED, "texture store exceeds 2^32 texels");
  // grow the shared texel array (device-to-device copy of what is there)
  if (off > ctx->d_texels.cap) {
    uchar4* np = nullptr;
    const size_t ncap = std::max(off, ctx->d_texels.cap * 2);
    VG_CUDA(ctx, cudaMalloc((void**)&np, ncap * sizeof(uchar4)));
    if (ctx->n_texels) VG_CUDA(ctx, cudaMemcpyAsync(np, ctx->d_texels.p, ctx->n_texels * sizeof(uchar4), cudaMemcpyDeviceToDevice, ctx->stream));
    VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_texels.p) cudaFree(ctx->d_texels.p);
    ctx->d_texels.p = np;
    ctx->d_texels.cap = ncap;
  }
  std::vector<uchar4> l0((size_t)w * h);
  for (size_t i = 0; i < l0.size(); i++) l0[i] = make_uchar4(rgb8[i * 3], rgb8[i * 3 + 1], rgb8[i * 3 + 2], 0);
  VG_CUDA(ctx, cudaMemcpyAsync(ctx->d_texels.p + lv[0].off, l0.data(), l0.size() * sizeof(uchar4), cudaMemcpyHostToDevice, ctx->stream));
  for (int l = 1; l < maxlevel; l++) VG_CUDA(ctx, launch_mip_level(ctx->d_texels.p, lv[(size_t)l - 1], lv[(size_t)l], ctx->stream));
  VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stats.kernel_launches += (uint64_t)(maxlevel - 1);
  ctx->n_texels = off;
  ctx->tex_levels.insert(ctx->tex_levels.end(), lv.begin(), lv.end());
  ctx->textures.push_back(T);
  VG_CUDA(ctx, upload(ctx->d_tex_levels, ctx->tex_levels, ctx->stream));
  VG_CUDA(ctx, upload(ctx->d_textures, ctx->textures, ctx->stream));
  VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *tex_id = (int)ctx->textures.size() - 1;
  render_invalidate(ctx);
  return VG_OK;
}

int vg_textures_clear(vg_ctx* ctx) {
  VG_LOCK(ctx);
  ctx->textures.clear();
  ctx->tex_levels.clear();
  ctx->n_texels = 0;
  for (MatTex& m : ctx->mat_tex) m = MatTex();
  render_invalidate(ctx);
  return VG_OK;
}

int vg_texture_levels(vg_ctx* ctx, int tex_id, int* n_levels) {
  VG_LOCK(ctx);
  if (tex_id < 0 || tex_id >= (int)ctx->textures.size() || !n_levels) return ctx->fail(VG_ERR_INVALID, "vg_texture_levels: bad texture id");
  *n_levels = ctx->textures[(size_t)tex_id].n_levels;
  return VG_OK;
}

int vg_texture_read_level(vg_ctx* ctx, int tex_id, int level, int* w, int* h, uint8_t* rgb8_out) {
  VG_LOCK(ctx);
  if (tex_id < 0 || tex_id >= (int)ctx->textures.size()) return ctx->fail(VG_ERR_INVALID, "vg_texture_read_level: bad texture id");
  const DevTexture& T = ctx->textures[(size_t)tex_id];
  if (level < 0 || level >= T.n_levels) return ctx->fail(VG_ERR_INVALID, "vg_texture_read_level: bad level");
  const DevTexLevel& L = ctx->tex_levels[(size_t)T.first_level + level];
  if (w) *w = L.w;
  if (h) *h = L.h;
  if (!rgb8_out) return VG_OK;
  VG_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<uchar4> tmp((size_t)L.w * L.h);
  VG_CUDA(ctx, cudaMemcpyAsync(tmp.data(), ctx->d_texels.p + L.off, tmp.size() * sizeof(uchar4), cudaMemcpyDeviceToHost, ctx->stream));
  VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < tmp.size(); i++) {
    rgb8_out[i * 3] = tmp[i].x;
    rgb8_out[i * 3 + 1] = tmp[i].y;
    rgb8_out[i * 3 + 2] = tmp[i].z;
  }
  return VG_OK;
}

int vg_material_set_texture(vg_ctx* ctx, int material, int slot, int tex_id, int chan, int filter) {
  VG_LOCK(ctx);
  if (material < 0 || material >= (int)ctx->materials.size()) return ctx->fail(VG_ERR_INVALID, "vg_material_set_texture: bad material index");
  if (slot < 0 || slot > 11 || slot == 9) return ctx->fail(VG_ERR_INVALID, "vg_material_set_texture: bad parameter slot");
  if (tex_id < 0 || tex_id >= (int)ctx->textures.size()) return ctx->fail(VG_ERR_INVALID, "vg_material_set_texture: bad texture id");
  if (chan < 0 || chan > 2) return ctx->fail(VG_ERR_INVALID, "vg_material_set_texture: channel outside [0,2] (the reference would index out of range)");
  if (filter != VG_TEXFILTER_FELINE && filter != VG_TEXFILTER_TRILINEAR) return ctx->fail(VG_ERR_INVALID, "vg_material_set_texture: bad filter");
  if (ctx->materials[(size_t)material].mask & VG_MAT_DEBUG) return ctx->fail(VG_ERR_UNSUPPORTED, "vg_material_set_texture: DebugShader.Colour as a texture map");
  ctx->mat_tex.resize(ctx->materials.size());
  TexBind& b = ctx->mat_tex[(size_t)material].slot[slot];
  b.tex = tex_id; b.chan = chan; b.filter = filter;
  ctx->materials[(size_t)material].mask |= (1u << slot);
  render_invalidate(ctx);
  return VG_OK;
}

int vg_mesh_set_uv(vg_ctx* ctx, int geom_id, const float* uv, int n_uv, const uint32_t* uvtriidx) {
  VG_LOCK(ctx);
  if (geom_id < 0 || geom_id >= (int)ctx->meshes.size()) return ctx->fail(VG_ERR_INVALID, "geom_id out of range");
  MeshStage& m = ctx->meshes[(size_t)geom_id];
  if (!m.present || m.sphere || m.instance) return ctx->fail(VG_ERR_INVALID, "vg_mesh_set_uv: upload the mesh first");
  if (!uv || n_uv <= 0 || !uvtriidx) return ctx->fail(VG_ERR_INVALID, "vg_mesh_set_uv: null/empty input");
  m.uv.assign(uv, uv + (size_t)n_uv * 2);
  m.uvtriidx.assign(uvtriidx, uvtriidx + (size_t)m.n_tris * 3);
  for (uint32_t i : m.uvtriidx)
    if (i >= (uint32_t)n_uv) return ctx->fail(VG_ERR_INVALID, "vg_mesh_set_uv: UV index out of range");
  ctx->committed = false;
  return VG_OK;
}

int vg_texture_sample_batch(vg_ctx* ctx, int tex_id, int filter, const float* coords, int64_t n, float* out) {
  VG_LOCK(ctx);
  if (tex_id < 0 || tex_id >= (int)ctx->textures.size()) return ctx->fail(VG_ERR_INVALID, "vg_texture_sample_batch: bad texture id");
  if (n < 0 || (n > 0 && (!coords || !out))) return ctx->fail(VG_ERR_INVALID, "vg_texture_sample_batch: bad input");
  if (n == 0) return VG_OK;
  VG_CUDA(ctx, cudaSetDevice(ctx->device));
  float *d_in = nullptr, *d_out = nullptr;
  VG_CUDA(ctx, cudaMalloc((void**)&d_in, (size_t)n * 8 * sizeof(float)));
  cudaError_t e = cudaMalloc((void**)&d_out, (size_t)n * 3 * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, coords, (size_t)n * 8 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = launch_texture_sample(ctx->tex_store(), tex_id, filter, d_in, n, d_out, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_in);
  if (d_out) cudaFree(d_out);
  if (e != cudaSuccess) return ctx->cuda_fail(e, "vg_texture_sample_batch");
  ctx->stats.kernel_launches += 1;
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
  else if (!std::strcmp(name, "primary_per_lane_motion")) ctx->opt_primary_per_lane_motion = value != 0;
  else if (!std::strcmp(name, "shadow_unordered")) ctx->opt_shadow_unordered = value != 0;
  else if (!std::strcmp(name, "shadow_per_lane")) ctx->opt_shadow_per_lane = value != 0;
  else if (!std::strcmp(name, "shadow_level0_per_lane")) {
    if (value < 0 || value > 2) return ctx->fail(VG_ERR_INVALID, "shadow_level0_per_lane must be 0 (cooperative kernel), 1 (per-lane loop) or 2 (measured)");
    ctx->opt_shadow_level0_per_lane = value;
  }
  else if (!std::strcmp(name, "batch_taper")) ctx->opt_batch_taper = value != 0;
  else if (!std::strcmp(name, "accumulate_wide")) ctx->opt_accumulate_wide = value != 0;
  else if (!std::strcmp(name, "accumulate_tiled")) ctx->opt_accumulate_tiled = value != 0;
  else if (!std::strcmp(name, "frame_slices_multi")) ctx->opt_frame_slices_multi = value != 0;
  else if (!std::strcmp(name, "frame_slices_min_paths_off")) ctx->opt_frame_slices_force = value != 0;
  else if (!std::strcmp(name, "frame_slices")) {
    if (value < 1 || value > 64) return ctx->fail(VG_ERR_INVALID, "frame_slices must be in [1,64]");
    ctx->opt_frame_slices = value;
  }
  else if (!std::strcmp(name, "capture_levels")) {
    ctx->opt_capture_levels = value & 31;
    if (value & 31) ctx->captured.clear();  // switching the capture on starts a new store; switching it off keeps what was captured
  }
  else if (!std::strcmp(name, "texture_coop")) ctx->opt_texture_coop = value != 0;
  else if (!std::strcmp(name, "zero_copy_batch")) ctx->opt_zero_copy_batch = value != 0;
  else if (!std::strcmp(name, "l2_persist_nodes")) ctx->opt_l2_persist_nodes = value != 0;  // takes effect at the next vg_scene_commit
  else if (!std::strcmp(name, "batch_chunk_log2")) {
    if (value < 14 || value > 24) return ctx->fail(VG_ERR_INVALID, "batch_chunk_log2 outside [14,24]");
    ctx->opt_batch_chunk_log2 = value;
  }
  else if (!std::strcmp(name, "node_order")) {  // takes effect at the next vg_scene_commit
    ctx->opt_node_order = value;
  }
  else if (!std::strcmp(name, "generic_shade")) {
    ctx->opt_generic_shade = value != 0;
    render_invalidate(ctx);
  }
  else if (!std::strcmp(name, "pixel_block")) {
    ctx->opt_pixel_block = value != 0;
    render_invalidate(ctx);
  }
  else if (!std::strcmp(name, "iter_group")) {
    if (value < 1 || value > 32 || (value & (value - 1))) return ctx->fail(VG_ERR_INVALID, "iter_group must be a power of two in [1,32]");
    ctx->opt_iter_group = value;
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
  if ((flags & VG_TRACE_COMPACT_HITS) && (ctx->dev.n_mtris > 0 || ctx->dev.n_spheres > 0 || ctx->dev.n_xforms > 0))
    return ctx->fail(VG_ERR_UNSUPPORTED, "VG_TRACE_COMPACT_HITS: the 16-byte record identifies a hit by its static triangle slot; scenes with motion meshes, sphere geoms or instances need the full VgHit");
  static int blocks_per_sm = 0;
  if (!blocks_per_sm) blocks_per_sm = trace_batch_blocks_per_sm();
  long long want = (n + kTraceBlock - 1) / kTraceBlock;
  long long grid = (long long)ctx->sm_count * blocks_per_sm;
  if (grid > want) grid = want;
  VG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  VG_CUDA(ctx, launch_trace_batch(ctx->dev, d_rays, d_hits, n, (flags & VG_TRACE_ANY_HIT) != 0, ctx->opt_traversal, ctx->d_counters.p, ctx->d_counters.p + 1, (int)grid, ctx->stream, ((flags & VG_TRACE_COMPACT_HITS) ? 1 : 0) | ((flags & VG_TRACE_RAYS_PD) ? 2 : 0)));
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
  const bool compact = (flags & VG_TRACE_COMPACT_HITS) != 0;
  if (compact && (ctx->dev.n_mtris > 0 || ctx->dev.n_spheres > 0 || ctx->dev.n_xforms > 0))
    return ctx->fail(VG_ERR_UNSUPPORTED, "VG_TRACE_COMPACT_HITS: the 16-byte record identifies a hit by its static triangle slot; scenes with motion meshes, sphere geoms or instances need the full VgHit");
  const size_t hit_bytes = compact ? sizeof(VgHitCompact) : sizeof(VgHit);
  const bool pd = (flags & VG_TRACE_RAYS_PD) != 0;  // `rays` holds 24-byte VgRayPD records
  const size_t ray_bytes = pd ? sizeof(VgRayPD) : sizeof(VgRay);
  const int mode = (compact ? 1 : 0) | (pd ? 2 : 0);
  const char* const rays_b = reinterpret_cast<const char*>(rays);
  char* const d_rays_b = reinterpret_cast<char*>(ctx->d_rays.p);
  char* const hits_b = reinterpret_cast<char*>(hits);
  char* const d_hits_b = reinterpret_cast<char*>(ctx->d_hits.p);
  const int64_t chunk = (int64_t)1 << ctx->opt_batch_chunk_log2;  // rays per pipeline stage (default 2^19 = 16 MB of rays)
  if (ctx->opt_zero_copy_batch && n >= 2 * chunk && pinned_host(rays) && pinned_host(hits)) {
    // Page-locked caller buffers, option zero_copy_batch: ONE launch whose warps read the rays from host memory and write the hits
    // back over PCIe themselves (coalesced 1-KB reads and writes per warp); no staging copies, no chunking.
    void *dr = nullptr, *dh = nullptr;
    if (cudaHostGetDevicePointer(&dr, const_cast<VgRay*>(rays), 0) == cudaSuccess && cudaHostGetDevicePointer(&dh, hits, 0) == cudaSuccess && dr && dh) {
      const int rc = trace_device_locked(ctx, reinterpret_cast<const VgRay*>(dr), n, reinterpret_cast<VgHit*>(dh), flags);
      if (rc != VG_OK) return rc;
      VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      return VG_OK;
    }
    cudaGetLastError();
  }
  const bool both_pinned = n >= 2 * chunk && pinned_host(rays) && pinned_host(hits);
  if (both_pinned) {
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
    // Stage sizes: full chunks in the middle; the first two and the last two stages are a quarter and a half chunk, so that the
    // pipeline's fill (the first upload, which nothing overlaps) and drain (the last traversal and download) are short.
    std::vector<int64_t> stages;
    {
      int64_t rem = n;
      const bool taper = ctx->opt_batch_taper && n >= 4 * chunk;
      if (taper) {
        stages.push_back(chunk / 4);
        stages.push_back(chunk / 2);
        rem -= chunk / 4 + chunk / 2;
        rem -= chunk / 2 + chunk / 4;  // kept for the tail
      }
      while (rem > 0) {
        stages.push_back(std::min(chunk, rem));
        rem -= stages.back();
      }
      if (taper) {
        stages.push_back(chunk / 2);
        stages.push_back(chunk / 4);
      }
    }
    int c = 0;
    int64_t off = 0;
    for (; c < (int)stages.size(); off += stages[c], c++) {
      const int64_t m = stages[c];
      cudaStream_t st = ctx->pipe_stream[c % 3];
      VG_CUDA(ctx, cudaMemcpyAsync(d_rays_b + (size_t)off * ray_bytes, rays_b + (size_t)off * ray_bytes, (size_t)m * ray_bytes, cudaMemcpyHostToDevice, st));
      long long grid = std::min<long long>((long long)ctx->sm_count * blocks_per_sm, (m + kTraceBlock - 1) / kTraceBlock);
      VG_CUDA(ctx, launch_trace_batch(ctx->dev, reinterpret_cast<const VgRay*>(d_rays_b + (size_t)off * ray_bytes), reinterpret_cast<VgHit*>(d_hits_b + (size_t)off * hit_bytes), m, (flags & VG_TRACE_ANY_HIT) != 0, ctx->opt_traversal,
                                      ctx->d_counters.p + 4 + (c % 3), ctx->d_counters.p + 1, (int)grid, st, mode));
      VG_CUDA(ctx, cudaMemcpyAsync(hits_b + (size_t)off * hit_bytes, d_hits_b + (size_t)off * hit_bytes, (size_t)m * hit_bytes, cudaMemcpyDeviceToHost, st));
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
  VG_CUDA(ctx, cudaMemcpyAsync(ctx->d_rays.p, rays, (size_t)n * ray_bytes, cudaMemcpyHostToDevice, ctx->stream));
  int rc = trace_device_locked(ctx, ctx->d_rays.p, n, ctx->d_hits.p, flags);
  if (rc != VG_OK) return rc;
  VG_CUDA(ctx, cudaMemcpyAsync(hits, ctx->d_hits.p, (size_t)n * hit_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  VG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VG_OK;
}

// slot -> (ElemID, geom) of every static triangle slot: what a VG_TRACE_COMPACT_HITS record's `slot` stands for
int vg_slot_table(vg_ctx* ctx, int32_t* prim_of_slot, int32_t* geom_of_slot, int64_t cap) {
  VG_LOCK(ctx);
  if (!ctx->committed) return ctx->fail(VG_ERR_INVALID, "scene not committed");
  const int64_t n = ctx->n_tri_slots;
  if (!prim_of_slot && !geom_of_slot) return (int)n;
  if (cap < n) return ctx->fail(VG_ERR_INVALID, "vg_slot_table: buffers too small");
  VG_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<float4> t((size_t)n * kTriStride);
  VG_CUDA(ctx, cudaMemcpy(t.data(), ctx->d_tris.p, t.size() * sizeof(float4), cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; i++) {
    int32_t g, pr;
    std::memcpy(&g, &t[(size_t)i * kTriStride].w, 4);
    std::memcpy(&pr, &t[(size_t)i * kTriStride + 1].w, 4);
    if (geom_of_slot) geom_of_slot[i] = g;
    if (prim_of_slot) prim_of_slot[i] = pr;
  }
  return (int)n;
}

int64_t vg_captured_rays(vg_ctx* ctx, VgRay* out, int64_t cap) {
  VG_LOCK(ctx);
  const int64_t n = (int64_t)ctx->captured.size();
  if (!out) return n;
  if (cap < n) return ctx->fail(VG_ERR_INVALID, "vg_captured_rays: buffer too small");
  if (n > 0) std::memcpy(out, ctx->captured.data(), (size_t)n * sizeof(VgRay));
  ctx->captured.clear();
  ctx->captured.shrink_to_fit();
  return n;
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
