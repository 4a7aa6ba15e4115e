// vg_ctx: owner of all device memory and of the staged (reference-format) scene data.
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/vermeer_gpu.h"
#include "device_scene.h"
#include "texture.cuh"

namespace vg {

struct MeshStage {
  bool present = false;
  bool motion = false;
  bool sphere = false;  // sphere.Sphere geom: centre/radius below, one triangle-slot-sized record, one prim
  float centre[3] = {0, 0, 0};
  float radius = 0;
  bool instance = false;  // instance.Instance: a transformed placement of geom `target`
  int target = -1;
  std::vector<VgTransformSRT> srt;
  std::vector<VgNode> nodes;
  std::vector<VgMotionNode> topo;
  std::vector<float> boxes;  // [keys][n_nodes][24]
  int keys = 1;
  std::vector<uint32_t> idxp;
  std::vector<int32_t> accel_idx;
  int n_tris = 0;
  std::vector<float> verts;  // [keys][n_verts][3]
  int n_verts = 0;
  std::vector<uint8_t> shaderidx;
  std::vector<int32_t> material_ids;
  std::vector<float> normals;
  std::vector<uint32_t> normalidx;
  float raybias = 0;
  int ref_compat = 0;
  std::vector<float> uv;           // PolyMesh.UV pairs
  std::vector<uint32_t> uvtriidx;  // 3 per triangle, leaf order
};

struct TexBind {  // one ShaderStd parameter read through maps.Texture / maps.TextureTrilinear
  int32_t tex = -1, chan = 0, filter = 0;
};
struct MatTex {
  TexBind slot[12];
  uint32_t mask = 0;  // bit k: slot k is a texture map (filled when the table goes to the device)
  bool any() const {
    for (const TexBind& b : slot)
      if (b.tex >= 0) return true;
    return false;
  }
};

struct SceneStage {
  bool present = false;
  bool motion = false;
  std::vector<VgNode> nodes;
  std::vector<VgMotionNode> topo;
  std::vector<float> boxes;
  int keys = 1;
  std::vector<int32_t> geom_of_slot;
};

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct RenderState;  // render.cu
struct CommState;    // comm.cu
struct BuildScratch;  // build_bvh.cu

}  // namespace vg

struct vg_ctx {
  std::mutex mu;
  std::string err;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t pipe_stream[3] = {nullptr, nullptr, nullptr};  // vg_trace_batch with page-locked buffers: copy/compute overlap
  cudaEvent_t pipe_done[3] = {nullptr, nullptr, nullptr};

  // staged scene (reference formats)
  std::vector<vg::MeshStage> meshes;
  vg::SceneStage scene;
  bool committed = false;

  // device scene
  vg::DevScene dev{};
  vg::DevBuf<vg::DevNode> d_nodes;
  vg::DevBuf<vg::DevMotionNode> d_mtopo;
  vg::DevBuf<float4> d_mboxes, d_tris, d_mtris, d_normals;
  vg::DevBuf<vg::DevGeom> d_geoms;
  vg::DevBuf<uint8_t> d_prim_material;
  vg::DevBuf<vg::DevXform> d_xforms;
  vg::DevBuf<vg::XfSRT> d_xf_keys;
  vg::DevBuf<vg::Mat4> d_xf_static;
  size_t scene_bytes = 0;
  int64_t n_tri_slots = 0;  // static triangle slots of the committed scene (d_tris holds kTriStride float4 each)

  // batch trace scratch
  vg::DevBuf<VgRay> d_rays;
  vg::DevBuf<VgHit> d_hits;
  vg::DevBuf<unsigned long long> d_counters;  // [0] queue head, [1..] stats

  // texture store: every level of every texture in one texel array
  std::vector<vg::DevTexture> textures;
  std::vector<vg::DevTexLevel> tex_levels;
  size_t n_texels = 0;
  vg::DevBuf<uchar4> d_texels;
  vg::DevBuf<vg::DevTexLevel> d_tex_levels;
  vg::DevBuf<vg::DevTexture> d_textures;
  vg::DevBuf<float2> d_tri_uv;
  std::vector<vg::MatTex> mat_tex;  // parallel to `materials`
  vg::DevTexStore tex_store() const { return vg::DevTexStore{d_texels.p, d_tex_levels.p, d_textures.p, (int32_t)textures.size()}; }

  vg::BuildScratch* build_scratch = nullptr;
  std::vector<VgNode> built_nodes;  // result of the last vg_build_qbvh, until vg_build_qbvh_nodes fetches it

  // shading inputs
  std::vector<VgMaterial> materials;
  std::vector<VgLight> lights;
  VgCamera camera{};
  bool have_camera = false;
  vg::DevBuf<vg::XfSRT> d_cam_keys;  // Camera.decomp when the camera has motion keys (vg_set_camera_motion)
  int cam_nkeys = 0;
  int xres = 0, yres = 0;
  int rank = 0, world = 1;
  std::vector<uint64_t> scramble;  // full frame, npix*6
  bool scramble_stale = false;     // a later vg_set_scramble went straight to the device (render.cu: render_set_scramble)
  int filter_n = 0;
  double filter_w = 0;
  std::vector<double> filter_cdf;  // cdfV[n] then cdfVU[n*n]
  int opt_trace_last_level = 1;
  int opt_iters_per_batch = 4;
  int opt_precise_trig = 0;
  int opt_primary_per_lane = 1;  // with traversal=2: camera rays (level 0) still use the per-lane loop
  int opt_primary_per_lane_motion = 0;  // camera rays of scenes with motion meshes through the per-lane loop too
  int opt_shadow_unordered = 1;  // integrator shadow queue: skip the sign-ordered push (occlusion is order independent)
  int opt_generic_shade = 0;     // 1 = always shade with the general kernel (tests: it must agree with the specialised one)
  int opt_accumulate_wide = 1;     // k_resolve_accumulate<true>: a vertex's four slots as two 256-bit loads (one lobe, S = 4, two lights)
  int opt_accumulate_tiled = 0;    // 1 = k_resolve_accumulate_t (shared-memory tile); measured slower: C2 raygen+accumulate 5.76 vs 4.74 ms per frame
  int opt_frame_slices_force = 0;  // tests: ignore the 16 M-path minimum slice size
  int opt_frame_slices_multi = 0; // 1 = slice the frame also when a multi-GPU communicator exists (measured: no gain)
  int opt_frame_slices = 4;      // vg_render_frame: slices of tile rows whose copies / exchange overlap the next slice's rendering
  int opt_capture_levels = 0;    // bit L: vg_render keeps a host copy of the level-L closest-hit ray queue (vg_captured_rays)
  std::vector<VgRay> captured;
  // level-0 shadow queue (coherent rays) through the per-lane loop without the ordered push (k_trace_queue<1,4>): 1 = always, 0 = never
  // (the cooperative kernel), 2 = whichever the first three vg_render calls measured faster on this scene. Measured: C2 shadow 31.08 ->
  // 30.12 ms and C1 0.33 -> 0.30 ms with it, but C3 248 -> 292 ms and C4 (MQBVH) 55.3 -> 70.2 ms: scene dependent, hence the measurement.
  int opt_shadow_level0_per_lane = 2;
  int opt_shadow_per_lane = 0;   // integrator shadow queue through the per-lane while-while kernel instead of the cooperative one
  int opt_iter_group = 32;       // a warp's 32 paths = (32/iter_group) pixels x iter_group iterations of the batch (render.cu: path_index)
  int opt_pixel_block = 1;       // paths of one warp cover an 8x4 pixel block of a tile (1) or a 32x1 row (0)
  int opt_batch_chunk_log2 = 19; // vg_trace_batch copy pipeline: rays per stage
  int opt_batch_taper = 1;       // ... with quarter / half stages at both ends (shorter pipeline fill and drain)
  int opt_l2_persist_nodes = 0;  // persisting-L2 access-policy window over the static node array
  int opt_zero_copy_batch = 0;   // vg_trace_batch with page-locked buffers: kernel reads rays / writes hits over PCIe itself (1) or 3-stream copy pipeline (0)
  int opt_node_order = 0;        // device order of a mesh's nodes: 0 = the reference's preorder, 1 = breadth-first (siblings adjacent)
  int opt_texture_coop = 1;      // k_surface: probes of a warp's texture lookups shared between its lanes (1) or one lane per lookup (0)
  int opt_traversal = 2;  // 0: per-lane while-while, 1: the same over a TMA-staged ray queue, 2: warp-cooperative leaves (traverse.cuh)

  vg::RenderState* rs = nullptr;
  vg::CommState* comm = nullptr;  // vg_comm_init: NCCL communicator + gather staging (comm.cu)
  VgStats stats{};

  int fail(int code, const std::string& msg) {
    err = msg;
    return code;
  }
  int cuda_fail(cudaError_t e, const char* what) {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return VG_ERR_CUDA;
  }
};

namespace vg {
// render.cu
int render_run(vg_ctx* ctx, int iter_begin, int iter_end, float* fb_out);
int render_frame(vg_ctx* ctx, const uint64_t* table, int64_t npix, int iter_begin, int iter_end, int clear_first, float* fb_out);
int render_clear(vg_ctx* ctx);
int render_fb_device(vg_ctx* ctx, float** d_fb);
void render_invalidate(vg_ctx* ctx);  // scene / frame / partition changed
int render_set_scramble(vg_ctx* ctx, const uint64_t* table, int64_t npix);
void render_destroy(vg_ctx* ctx);
void build_scratch_destroy(vg_ctx* ctx);  // build_bvh.cu
// comm.cu
void comm_destroy(vg_ctx* ctx);
int comm_gather_rows(vg_ctx* ctx, cudaStream_t st, int ty0, int ty1, float* fb_out, bool sync);
int partition_stride(int tilesX, int world);
void owned_pixels(int W, int H, int rank, int world, bool pixel_block, std::vector<int>& pix);  // tile-major pixel list of `rank`
}  // namespace vg
