/* vermeer_gpu.h — C ABI of the B200-native ray-traversal and path-integration engine.
 *
 * Two groups of entry points, both exported by libvermeer_b200.so:
 *
 *   vg_*  DEVICE LAYER.  What the reference-side cgo shim binds.  It consumes exactly the data the
 *         reference's PreRender produces (qbvh.Node[], MotionQBVH, idxp, Verts, shaderidx, the
 *         scene-level tree and geom order) and replaces the per-ray / per-frame hot path:
 *             core.TraceProbe   (core/trace.go:26)      -> vg_trace_batch
 *             core.Render/render (core/render.go:140,66) -> vg_render
 *         See INTEGRATION.md for the Go stub.
 *
 *   vh_*  HOST LAYER.  A C++ mirror of the reference's node registry and PreRender pipeline
 *         (nodes.Register nodes/register.go:26; PolyMesh.PreRender polymesh.go:73; qbvh.BuildAccel
 *         build.go:293; scene.initAccel scene.go:135; Camera.PreRender camera.go:80) for hosts
 *         without a Go toolchain.  It produces the arrays above and feeds them to vg_*.
 *
 * Conventions: every call returns 0 on success and <0 on error (vg_last_error gives the text);
 * nothing throws or aborts across the ABI; all pointers are borrowed for the duration of the call
 * (cgo rule) and copied before return; calls on one context are serialised by an internal mutex.
 * There is NO CPU fallback: compute calls fail with VG_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef VERMEER_GPU_H
#define VERMEER_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VG_OK 0
#define VG_ERR_INVALID (-1)    /* bad argument / bad state */
#define VG_ERR_NO_DEVICE (-2)  /* no usable CUDA device (there is no CPU fallback) */
#define VG_ERR_CUDA (-3)       /* a CUDA call failed */
#define VG_ERR_BUILD (-4)      /* host-side tree build failed (e.g. reference quirk f: coincident centroids) */
#define VG_ERR_UNSUPPORTED (-5)
#define VG_ERR_COMM (-6)        /* NCCL missing or an NCCL call failed (vg_comm_*, vg_gather_frame) */

/* ---- records ------------------------------------------------------------------------------ */

/* One ray of a batch: the public part of core.Ray (core/ray.go:27-34).  Ray.Setup (ray.go:103-148:
 * Kx/Ky/Kz, S, Dinv with float64 divides) runs on the device. 32 bytes. */
typedef struct VgRay {
  float o[3];
  float d[3];
  float tmax; /* Ray.Tclosest at Init: +Inf for camera/reflected rays, 1 for shadow rays */
  float time; /* Ray.Time in [0,1) */
} VgRay;

/* Result of TraceProbe for one ray. 32 bytes.
 * prim = ShaderContext.ElemID (leaf-order triangle index, polymesh/trace.go:513; original face index for
 * motion meshes, trace.go:624) or -1 on a miss; geom = index of the Geom in creation order, -1 on a miss.
 * t = Ray.Tclosest after the call (tmax on a miss). u,v,w = scaled barycentrics U,V,W of trace.go:186-188.
 * nodesT = interior-node visits (Ray.NodesT, intersect.go:114), trisT = sum of LeafCount over visited
 * triangle leaves; both feed the bytes model of the roofline. The device keeps the two counters of a ray in one
 * register, 16 bits each: exact up to 65 535 node visits / leaf triangles per ray (the BASELINE scenes need < 200). */
typedef struct VgHit {
  float t, u, v, w;
  int32_t prim;
  int32_t geom;
  int32_t nodesT;
  int32_t trisT;
} VgHit;

/* qbvh.Node, verbatim (qbvh/qbvh.go:31-36). 128 bytes. Boxes[child + 12*(0=min,1=max) + 4*axis]. */
typedef struct VgNode {
  float boxes[24];
  uint32_t axis0, axis1, axis2;
  int32_t children[4]; /* >=0 node; <0 leaf: bit31 | first<<4 | count-1; -1 empty (qbvh.go:61-91) */
  int32_t parent;
} VgNode;

/* qbvh.MotionNode, verbatim (qbvh/mqbvh.go:24-29). 40 bytes. Boxes come separately as [key][node][24]. */
typedef struct VgMotionNode {
  int32_t axis0, axis1, axis2;
  int32_t children[4];
  int32_t parent;
  uint32_t pad[2];
} VgMotionNode;

/* ShaderStd with every parameter a Constant map (builtin/shader/std.go:25-47, builtin/maps/constant.go).
 * mask bit i set <=> the corresponding parameter is non-nil in the reference. */
#define VG_MAT_EMISSION_COLOUR 1u
#define VG_MAT_EMISSION_STRENGTH 2u
#define VG_MAT_DIFFUSE_COLOUR 4u
#define VG_MAT_DIFFUSE_STRENGTH 8u
#define VG_MAT_DIFFUSE_ROUGHNESS 16u
#define VG_MAT_SPEC1_COLOUR 32u
#define VG_MAT_SPEC1_STRENGTH 64u
#define VG_MAT_SPEC1_ROUGHNESS 128u
#define VG_MAT_IOR 256u
#define VG_MAT_SPEC1_FRESNEL_MODEL 512u /* Spec1FresnelModel given ("Dielectric" | "Metal", std.go:65-73) */
#define VG_MAT_SPEC1_FRESNEL_REFL 1024u
#define VG_MAT_SPEC1_FRESNEL_EDGE 2048u
/* shader.Debug (builtin/shader/debug.go:15-49, node "DebugShader"): Eval sets OutRGB = Colour and nothing else — no lights, no
 * rays, no Level check — and EvalEmission returns black. Colour travels in diffuse_colour; every other field is ignored. */
#define VG_MAT_DEBUG 4096u
#define VG_FRESNEL_DIELECTRIC 0 /* fresnel.DielectricModel (fresnel/models.go:13-17), also the value when the string is unset */
#define VG_FRESNEL_CONDUCTOR 1  /* fresnel.ConductorModel ("Metal") */
typedef struct VgMaterial {
  uint32_t mask;
  float emission_colour[3];
  float emission_strength;
  float diffuse_colour[3];
  float diffuse_strength;
  float diffuse_roughness;
  float spec1_colour[3];
  float spec1_strength;
  float spec1_roughness; /* 0 = mirror lobe (bsdf.Specular), > 0 = GGX glossy lobe (bsdf.MicrofacetGGX, direct light only) */
  float ior;
  int32_t spec1_fresnel_model; /* VG_FRESNEL_* */
  float spec1_fresnel_refl[3];
  float spec1_fresnel_edge[3];
} VgMaterial;

/* light.Tri (builtin/light/triangle.go:18-28) after PreRender. */
typedef struct VgTriLight {
  float p0[3], p1[3], p2[3];
  int32_t samples;  /* Samples; NumSamples = 1<<samples (triangle.go:71-73) */
  int32_t material; /* index into the material table (the light's Shader) */
  int32_t geom;     /* geom id of the mesh the light created for itself (triangle.go:48-50) */
} VgTriLight;

/* Any in-scope core.Light after PreRender, in scene order (scene.AddLight order, builtin/scene/scene.go:126-129):
 *   VG_LIGHT_TRI    light.Tri    (builtin/light/triangle.go:18-28)  p0,p1,p2 = vertices
 *   VG_LIGHT_DISK   light.Disk   (builtin/light/disk.go:21-34)      p0 = P, p1 = T, p2 = B, n = N (disk.go:90-97), radius
 *   VG_LIGHT_SPHERE light.Sphere (builtin/light/sphere.go:17-28)    p0 = P, radius; geom = the sphere geom it created
 * light.Quad is not offered: both of its sampling entry points panic in the reference (quad.go:88,94). */
#define VG_LIGHT_TRI 0
#define VG_LIGHT_DISK 1
#define VG_LIGHT_SPHERE 2
typedef struct VgLight {
  int32_t type;
  int32_t samples;  /* Samples; NumSamples = 1<<samples */
  int32_t material; /* the light's Shader (EvalEmission) */
  int32_t geom;     /* geom id of the mesh / sphere the light created for itself */
  float p0[3], p1[3], p2[3];
  float n[3];
  float radius;
} VgLight;

/* camera.Camera after PreRender (builtin/camera/camera.go:80-98): single-key LocalToWorld after the
 * decompose/recompose of camera.go:188-192,225-236 (column major, math/matrix4.go:9-18). */
typedef struct VgCamera {
  float local_to_world[16];
  float tan_theta_focal;
  float aspect;
  float focal;
  float radius;
} VgCamera;

typedef struct VgStats {
  uint64_t rays;        /* every TraceProbe equivalent (core/stats.go:26-28) */
  uint64_t shadow_rays; /* of which RayTypeShadow (stats.go:31-33) */
  uint64_t nodes_t;     /* interior-node visits, closest-hit rays (vg_trace_batch* and the render's closest-hit kernel) */
  uint64_t tris_t;      /* leaf triangle counts, closest-hit rays */
  uint64_t shadow_nodes_t; /* same for the render's any-hit kernel */
  uint64_t shadow_tris_t;
  uint64_t kernel_launches;
  uint64_t closest_launches; /* traversal launches inside vg_render, per kind */
  uint64_t shadow_launches;
  double render_ms;  /* device time of the last vg_render (CUDA events on the launching stream) */
  double trace_ms;   /* device time of the last vg_trace_batch* kernel */
  double closest_ms; /* summed device time of the closest-hit traversal launches of the last vg_render */
  double shadow_ms;  /* summed device time of the any-hit traversal launches of the last vg_render */
  double shade_ms;   /* summed device time of the shading launches (k_surface + k_shade*) of the last vg_render */
  double gather_ms;  /* device time of the last vg_gather_frame exchange (pack + NCCL + scatter), without the D2H copy */
  uint64_t max_stack_depth; /* deepest traversal stack (entries) any ray of vg_render needed since vg_reset_stats; 0 unless the library is
                               a measurement build (-DVG_STACK_STATS). The reference reserves 90 entries (core/ray.go:158). */
  int64_t shadow_level0_kernel; /* any-hit kernel the last vg_render used for the level-0 shadow queue: 0 cooperative, 1 per-lane loop
                                   (option "shadow_level0_per_lane": 2 = measured on the first calls), -1 not applicable */
} VgStats;

/* ---- device layer ------------------------------------------------------------------------- */

typedef struct vg_ctx vg_ctx;

int vg_create(vg_ctx** out, int device_ordinal);
void vg_destroy(vg_ctx* ctx);
const char* vg_last_error(vg_ctx* ctx); /* ctx may be NULL: error of the last failed vg_create */
int vg_device_count(void);

/* Scene assembly. geom ids are 0..n_geoms-1 in creation order (core.AddNode order, core/core.go:77-93). */
int vg_scene_begin(vg_ctx* ctx, int n_geoms);
/* Static PolyMesh after initAccel (buildqbvh.go:85-129): nodes, leaf-ordered idxp (3*n_tris), Verts (n_verts*3),
 * optional leaf-ordered shaderidx (n_tris) mapping to material_ids[], optional Normals + leaf-ordered normalidx. */
int vg_mesh_upload(vg_ctx* ctx, int geom_id, const VgNode* nodes, int n_nodes, const uint32_t* idxp, int n_tris,
                   const float* verts, int n_verts, const uint8_t* shaderidx, const int32_t* material_ids, int n_materials,
                   const float* normals, int n_normals, const uint32_t* normalidx, float raybias);
/* Motion PolyMesh after initAccel (buildqbvh.go:19-56,148-212): topology, boxes[key][node][24], UNreordered idxp,
 * accel_idx (leaf order -> face), Verts keys*n_verts*3 (key-major, core/param/array.go:26).
 * ref_compat != 0 reproduces reference quirk (b): leaf slot i tests face i (trace.go:532-537). */
int vg_mesh_upload_motion(vg_ctx* ctx, int geom_id, const VgMotionNode* topo, int n_nodes, const float* boxes, int keys,
                          const uint32_t* idxp, const int32_t* accel_idx, int n_tris, const float* verts, int n_verts,
                          const uint8_t* shaderidx, const int32_t* material_ids, int n_materials, float raybias, int ref_compat);
/* sphere.Sphere geom (builtin/geom/sphere/sphere.go:15-28; "only used for spherical light sources"): an analytic leaf of the
 * scene-level tree, intersected like sphere/trace.go:13-109. */
int vg_sphere_upload(vg_ctx* ctx, int geom_id, const float* centre, float radius, int32_t material_id);
/* instance.Instance after PreRender (builtin/geom/instance/instance.go:36-51,117-146): a second placement of an already
 * uploaded PolyMesh geom. srt = transformSRT, one m.TransformDecomp per Transform key (math/animdecomp.go:12-17: T, R{X,Y,Z,W},
 * S column major — 23 contiguous floats, the Go struct's own layout); per ray the keys are interpolated at Ray.Time
 * (TimeKey, instance.go:16-33), recomposed, inverted, and the ray is re-Setup in object space (instance.go:86-93). The
 * instance's bounds are whatever the scene-level tree was built with (the scene file gives them). */
typedef struct VgTransformSRT {
  float T[3];
  float R[4];
  float S[16];
} VgTransformSRT;
int vg_instance_upload(vg_ctx* ctx, int geom_id, int target_geom_id, const VgTransformSRT* srt, int keys);
/* Scene-level tree (builtin/scene/scene.go:135-203): leafMax=1 nodes over geoms; geom_of_slot[i] = geom id at leaf slot i. */
int vg_scene_upload(vg_ctx* ctx, const VgNode* nodes, int n_nodes, const int32_t* geom_of_slot, int n_slots);
int vg_scene_upload_motion(vg_ctx* ctx, const VgMotionNode* topo, int n_nodes, const float* boxes, int keys,
                           const int32_t* geom_of_slot, int n_slots);
/* Flatten everything into the device layout (DESIGN.md "data layout in HBM") and copy it to HBM. */
int vg_scene_commit(vg_ctx* ctx);

/* qbvh.BuildAccel (qbvh/build.go:293-307) on the device: the reference's top-down binned-SAH build (8 bins on the longest axis
 * of the centroid bounds, cost = area * count, three binary splits per 4-wide node, leafMax clamped to [1,16]) run
 * level-synchronously over all open ranges. boxes = n x {min xyz, max xyz}, centroids = n x 3 (build.go:293: the caller's
 * arrays); idx_out receives the leaf-order permutation (what the reference leaves in `idxs`), bounds6 the root box. The nodes
 * stay in the context until vg_build_qbvh_nodes copies them out (*n_nodes says how many), in the reference's preorder
 * numbering and link encoding. Every split decision is a function of the SET of primitives in a range (min/max, integer
 * counts, a fixed-order float32 cost loop), so nodes, boxes and leaf sets are bit-identical to the reference's; the order of
 * the <= 16 primitives inside one leaf differs (stable partition here, two-pointer swap there), which can only re-order ties
 * between equal-t hits within a leaf. VG_ERR_BUILD for the inputs the reference cannot build (quirk f, non-finite centroids). */
int vg_build_qbvh(vg_ctx* ctx, const float* boxes, const float* centroids, int n, int leaf_max, int32_t* idx_out, float* bounds6, int* n_nodes);
int vg_build_qbvh_nodes(vg_ctx* ctx, VgNode* nodes_out, int nodes_cap);

int vg_set_materials(vg_ctx* ctx, const VgMaterial* mats, int n);

/* ---- texture maps ------------------------------------------------------------------------------------
 * The reference's texture store (texture/texture.go:36-47) and the maps that read it (builtin/maps/texture.go:17-83).
 * vg_texture_upload: one texture.Texture after loadTexture (texture.go:118-160): RGB8, w*h*3 bytes, row 0 = BOTTOM row of
 * the image (loadTexture flips while copying, :139). The mip pyramid of stdfilter (texture/mipmap.go:122-315) is built on the
 * device, byte-identical to the reference's. Image decoding (png/jpeg/tiff/tga) stays with the caller. *tex_id receives the
 * store index. A 1x1 image is refused (the reference panics on it, mipmap.go:127). Textures outlive vg_scene_begin. */
int vg_texture_upload(vg_ctx* ctx, const uint8_t* rgb8, int w, int h, int* tex_id);
int vg_textures_clear(vg_ctx* ctx);
/* Pyramid inspection: number of levels = ceil(log2(max(w,h))); one level as RGB8 rows (bottom-up). rgb8_out may be NULL. */
int vg_texture_levels(vg_ctx* ctx, int tex_id, int* n_levels);
int vg_texture_read_level(vg_ctx* ctx, int tex_id, int level, int* w, int* h, uint8_t* rgb8_out);
/* filter: the map type CreateRGBTextureMap / CreateFloat32TextureMap picks from the "?filter=" query (maps/texture.go:48-83) */
#define VG_TEXFILTER_FELINE 0    /* maps.Texture: texture.SampleFeline (texture/feline.go:25-151), the default */
#define VG_TEXFILTER_TRILINEAR 1 /* maps.TextureTrilinear: texture.SampleRGB (texture/texture.go:219-311) */
/* Parameter `slot` of material `material` becomes a texture map: slot = bit index of the VG_MAT_* flag of that parameter
 * (0 EmissionColour .. 8 IOR, 10 Spec1FresnelRefl, 11 Spec1FresnelEdge); chan = the channel a float parameter reads
 * ("?ch=N"). Sets the parameter's VG_MAT_* bit. vg_set_materials drops all bindings. Textured scenes hold static PolyMeshes
 * only (the reference's motion path leaves the footprint 0 and Feline then divides 0/0, trace.go:677-684, feline.go:57-61);
 * a texture on the emission of a LIGHT's shader is refused (lights evaluate it with their own lsg). */
int vg_material_set_texture(vg_ctx* ctx, int material, int slot, int tex_id, int chan, int filter);
/* PolyMesh.UV (polymesh.go:36) and the triangulated, leaf-ordered UV indices (uvtriidx after init.go:38-106 and
 * buildqbvh.go:106-116): uv = n_uv pairs, uvtriidx = 3 per triangle in the order of idxp. After vg_mesh_upload, before commit.
 * Meshes without UVs use the barycentrics (trace.go:355-358,495-501). */
int vg_mesh_set_uv(vg_ctx* ctx, int geom_id, const float* uv, int n_uv, const uint32_t* uvtriidx);
/* The two filters over a batch of lookups (host buffers): coords = n x 8 floats {U, V, Dduvdx[2], Dduvdy[2], PixelDelta[2]}
 * (core/shader.go:78-86, core/render.go:32-35), out = n x 3 floats. */
int vg_texture_sample_batch(vg_ctx* ctx, int tex_id, int filter, const float* coords, int64_t n, float* out);
int vg_set_lights(vg_ctx* ctx, const VgTriLight* lights, int n);   /* TriLights only (kept for callers that have nothing else) */
int vg_set_area_lights(vg_ctx* ctx, const VgLight* lights, int n);  /* any mix of light types, scene order */
int vg_set_camera(vg_ctx* ctx, const VgCamera* cam);
/* A camera with motion keys (builtin/camera/camera.go:225-236): decomp = Camera.decomp, one m.TransformDecomp per LocalToWorld
 * key; every camera ray lerps/slerps the keys at its Time and recomposes the matrix on the device. keys <= 1 is vg_set_camera. */
int vg_set_camera_motion(vg_ctx* ctx, const VgCamera* cam, const VgTransformSRT* decomp, int keys);
int vg_set_frame(vg_ctx* ctx, int xres, int yres);
/* Image partition across processes/GPUs: this context renders the 32x32 tiles (tx,ty) with
 * (tx + ty*stride_k) % world == rank (core/render.go:196-199 tiles). Default rank 0 of 1. */
int vg_set_partition(vg_ctx* ctx, int rank, int world);
/* framescramble (core/render.go:18-23,166-176): npix*6 uint64 {lensU,lensV,time,lambda,scramble0,scramble1}, row-major pixels. */
int vg_set_scramble(vg_ctx* ctx, const uint64_t* table, int64_t npix);
/* core.PixelFilter (core/pixelfilter.go:8) as the tables filter.CreateSampler builds (builtin/filter/filter.go:88-173):
 * n, width w, marginal CDF cdfV[n] and conditional CDFs cdfVU[n*n] (row uI). The FIS warp of filter.go:40-86 runs on the
 * device in ray generation (core/render.go:99-107). n = 0 removes the filter. */
int vg_set_filter(vg_ctx* ctx, int n, double w, const double* cdfV, const double* cdfVU);
/* Options: "trace_last_level" (1 = trace the level-4 mirror ray like the reference, std.go:243; default 1),
 * "frame_slices" (vg_render_frame's pipeline depth), "iters_per_batch" (wavefront batch depth, default 4), "iter_group" (a warp's 32 paths = 32/iter_group neighbouring pixels x
 * iter_group consecutive iterations of the batch; power of two, default 4; results do not depend on it), "precise_trig" (1 = shading trig through float64 exactly like
 * math/sincos.go; 0 = single-precision libm, default; both are within the image tolerance), "traversal" (persistent-kernel variant:
 * 2 = warp-cooperative leaf phase, default; 0 = per-lane while-while loop with coalesced LDG refill; 1 = the per-lane loop with
 * its ray queue staged into shared memory by cp.async.bulk + mbarrier; all three are bit-identical), "tma_stage" (1 = alias of
 * traversal 1, 0 = default), "shadow_level0_per_lane" (any-hit kernel of the level-0 shadow queue: 0 = cooperative leaf phase,
 * 1 = per-lane loop for coherent rays, 2 = whichever the first batches measure faster on this scene, default; the frames are the
 * same bits either way — VgStats.shadow_level0_kernel reports the choice), "batch_chunk_log2" / "batch_taper" (vg_trace_batch's
 * copy pipeline with page-locked buffers: rays per stage, shorter stages at both ends), "accumulate_wide" (256-bit slot loads in
 * the resolve + accumulate kernel; same bits). Measurement switches, none of them changes a result. */
int vg_set_option(vg_ctx* ctx, const char* name, int value);

/* TraceProbe over a batch (core/trace.go:26). flags: */
#define VG_TRACE_ANY_HIT 1u /* RayTypeShadow: return at the first leaf that reports a hit (intersect.go:228-236) */
/* `hits` is an array of 16-byte VgHitCompact records instead of VgHit: half the device->host bytes of a batch whose cost is the
 * PCIe transfer. slot = the hit triangle's global leaf-order slot (vg_slot_table maps it to ElemID and geom), -1 on a miss;
 * w is not returned (the caller recomputes what it needs from u, v). Scenes of static PolyMeshes only (VG_ERR_UNSUPPORTED
 * otherwise). t, u, v and the hit triangle are the same bits as in the full record. */
#define VG_TRACE_COMPACT_HITS 2u
typedef struct VgHitCompact {
  float t, u, v;
  int32_t slot;
} VgHitCompact;
/* `rays` is an array of 24-byte VgRayPD records {P, D} instead of VgRay: what Ray.Init(ty, P, D, maxdist = +Inf, ...) takes for a
 * camera / reflected ray (core/ray.go:56-61) at Ray.Time 0 — three quarters of the host->device bytes of a batch whose cost is the
 * PCIe transfer. Hits are the same bits as for the VgRay {P, D, +Inf, 0}. Cast the pointer to const VgRay*. */
#define VG_TRACE_RAYS_PD 4u
typedef struct VgRayPD {
  float o[3];
  float d[3];
} VgRayPD;
int vg_trace_batch(vg_ctx* ctx, const VgRay* rays, int64_t n, VgHit* hits, uint32_t flags);               /* host buffers */
int vg_trace_batch_device(vg_ctx* ctx, const VgRay* d_rays, int64_t n, VgHit* d_hits, uint32_t flags);    /* device buffers */
/* ElemID (VgHit.prim) and geom of every static triangle slot; returns the slot count (buffers may be NULL to ask for it). */
int vg_slot_table(vg_ctx* ctx, int32_t* prim_of_slot, int32_t* geom_of_slot, int64_t cap);

/* The Render loop (core/render.go:184-205) for 0-based iterations [iter_begin, iter_end) over this context's tiles,
 * continuing the running mean held on the device. fb_out (xres*yres*3 floats, row-major, may be NULL) receives the
 * full-frame buffer with non-owned pixels left 0. */
int vg_render(vg_ctx* ctx, int iter_begin, int iter_end, float* fb_out);
int vg_clear_framebuffer(vg_ctx* ctx);
/* Benchmarking aid: with option "capture_levels" = bitmask of path levels (bit 1 = the first mirror bounce ...), vg_render keeps a
 * host copy of the closest-hit ray queue of those levels (core.Trace's reflected rays, builtin/shader/std.go:219-261), appended
 * batch after batch. vg_captured_rays returns their number (out == NULL) or copies them out and clears the store. This is how
 * bench.py obtains a genuinely incoherent closest-hit batch "from the wavefront itself". */
int64_t vg_captured_rays(vg_ctx* ctx, VgRay* out, int64_t cap);
/* Device pointer to the row-major xres*yres*3 float framebuffer (for NCCL gathers done by the caller). */
int vg_framebuffer_device(vg_ctx* ctx, float** d_fb);
int vg_get_stats(vg_ctx* ctx, VgStats* out);
int vg_reset_stats(vg_ctx* ctx);

/* ---- multi-GPU: one process (or thread) per GPU, one context each -----------------------------------------
 * The reference renders its 32x32 tiles on <= 10 goroutines of one process (core/render.go:186-203) and has no exchange step;
 * here the tiles are dealt over `world` contexts and the owned pixels are brought to rank 0 once per frame.
 *   vg_comm_unique_id   rank 0 creates the NCCL id (VG_COMM_ID_BYTES bytes); the HOST distributes it to the other ranks by
 *                       whatever means it has (the Go shim: a channel between its per-GPU goroutines; bench.py: a broadcast).
 *   vg_comm_init        every rank: ncclCommInitRank on the context's device, then the image partition of
 *                       vg_set_partition(rank, world). Collective. world == 1 needs no NCCL and no id.
 *   vg_gather_frame     collective, after vg_render: every rank packs its owned pixels (own kernel), ranks > 0 ncclSend them
 *                       to rank 0, which ncclRecvs into a staging buffer and scatters them into its row-major frame (own
 *                       kernel; disjoint ownership, nothing is summed: the result is bit-identical to a 1-GPU render), then
 *                       copies the complete frame into fb_out (xres*yres*3 floats; rank 0 only, may be NULL: the frame stays
 *                       in the buffer vg_framebuffer_device returns). fb_out is ignored on the other ranks.
 * NCCL is bound at run time (dlopen libnccl.so.2, or the path in $VG_NCCL_LIB); without it these calls fail with VG_ERR_COMM
 * and everything else keeps working. */
#define VG_COMM_ID_BYTES 128
int vg_comm_unique_id(vg_ctx* ctx, void* id_out);
int vg_comm_init(vg_ctx* ctx, int rank, int world, const void* id);
int vg_comm_destroy(vg_ctx* ctx);
int vg_gather_frame(vg_ctx* ctx, float* fb_out);
int vg_nccl_version(void); /* ncclGetVersion of the library that was bound, 0 if none */
/* The whole frame step as ONE call: vg_set_scramble(table) [+ vg_clear_framebuffer] + vg_render(iter_begin, iter_end) + vg_gather_frame
 * (when a communicator exists; collective then) + the copy of the frame into fb_out (rank 0 / single GPU; may be NULL) — pipelined.
 * The image is cut into "frame_slices" (option, default 4) runs of tile rows; while slice s renders, the scramble rows of slice s+1
 * are on their way up and the finished pixels of slice s-1 on their way back (NCCL exchange + D2H of those image rows), each on its
 * own stream. This is the reference's frame loop (core/render.go:184-205) with its inputs and outputs in HOST memory, which is what
 * a Go caller has; results are bit-identical to the three separate calls. The overlap needs page-locked `table` and `fb_out`
 * (cudaHostAlloc / cudaHostRegister); with pageable buffers, or on the first call after a scene change, the plain sequence runs. */
int vg_render_frame(vg_ctx* ctx, const uint64_t* table, int64_t npix, int iter_begin, int iter_end, int clear_first, float* fb_out);
/* The tile-major list of the full-frame pixel indices `rank` of `world` owns (host only, no device needed): returns the count,
 * fills pix_out (may be NULL) if it holds at least that many. pixel_block = the "pixel_block" option (default 1). */
int vg_owned_pixels(int xres, int yres, int rank, int world, int pixel_block, int32_t* pix_out, int64_t cap);

/* ---- measured ceilings ---------------------------------------------------------------------------------------------------
 * Read bandwidth of HBM, L2 and L1 on this context's GPU, measured with streaming kernels (csrc/peaks.cu): the denominators
 * of the roofline fractions bench.py reports for the L2-/L1-resident configs. No reference counterpart. */
typedef struct VgPeaks {
  double hbm_read_gbs;     /* read-only stream over hbm_buffer_bytes (>> L2), ld.global.cg */
  double l2_read_gbs;      /* the same over l2_buffer_bytes (fits L2), re-read many times */
  double l1_read_gbs;      /* every SM re-reads its own l1_window_bytes, ld.global.ca */
  double hbm_buffer_bytes, l2_buffer_bytes, l1_window_bytes;
  int32_t sm_count;
  int32_t pad_;
} VgPeaks;
int vg_measure_peaks(vg_ctx* ctx, VgPeaks* out);

/* ---- host layer ---------------------------------------------------------------------------- */

typedef struct vh_scene vh_scene;

int vh_scene_create(vh_scene** out);
void vh_scene_destroy(vh_scene* s);
const char* vh_last_error(vh_scene* s);

/* Registered node types, as in nodes.Register (nodes/register.go:26): returns the number of names and, if
 * names != NULL, fills up to cap of them. */
int vh_registered_nodes(const char** names, int cap);

int vh_set_globals(vh_scene* s, int xres, int yres, int max_iter);
/* Host-layer options, before vh_prerender. "leaf_max" (default 16 = the reference's value, buildqbvh.go:85): leafMax of the per-mesh
 * QBVH / MQBVH builds. Any other value is the OPT-IN NON-PARITY MODE: the reference's own builder run with smaller
 * leaves gives another tree over the same triangles; the intersection routine and therefore every hit's t, u, v and face are the
 * same (up to which of two equal-t triangles wins), but NodesT / TrisT, the leaf-order ElemID numbering and the work per ray are
 * not the reference's. Parity tests and the headline numbers use 16; bench.py reports the mode beside them. */
int vh_set_option(vh_scene* s, const char* name, int value);
int vh_add_shader_std(vh_scene* s, const char* name, const VgMaterial* params);
/* Texture maps (builtin/maps/texture.go, texture/texture.go). vh_add_texture: one decoded image file under the name the
 * shaders use (rgb8 rows bottom-up, see vg_texture_upload); the first registration of a name wins, like the reference's cache.
 * vh_shader_set_texture: parameter `slot` (see vg_material_set_texture) of ShaderStd `shader` reads the file named by `value`,
 * a reference texture URL: "path[?filter=trilinear][&ch=N]" (CreateRGBTextureMap / CreateFloat32TextureMap, maps/texture.go:48-83).
 * The .vnf form is `Param rgbtex "value"` (nodes/parser.go:247-268). vh_polymesh_set_uv: PolyMesh.UV / UVIdx (polymesh.go:36-37),
 * before vh_prerender. vh_upload fails for a map whose file was never registered (the reference substitutes a checker image). */
int vh_add_texture(vh_scene* s, const char* name, int w, int h, const uint8_t* rgb8_bottom_up);
int vh_shader_set_texture(vh_scene* s, const char* shader, int slot, const char* value);
int vh_polymesh_set_uv(vh_scene* s, const char* mesh, const float* uv, int n_uv, const int32_t* uvidx, int n_uvidx);
/* DebugShader node (builtin/shader/debug.go:51-57): Colour is a constant rgb map */
int vh_add_shader_debug(vh_scene* s, const char* name, const float* colour);
int vh_add_polymesh(vh_scene* s, const char* name, const float* verts, int n_verts, int keys, const int32_t* polycount, int n_poly,
                    const int32_t* faceidx, int n_faceidx, const char* shaders_nl, const int32_t* shaderidx, int n_shaderidx,
                    const float* normals, int n_normals, const int32_t* normalidx, int n_normalidx, float raybias);
/* AiryFilter / GaussianFilter nodes (builtin/filter/airy.go:13, gauss.go:13). width/res/peak <= 0 keep the registered
 * defaults (Airy: Res 49, Width 6, Peak 4; Gaussian: Res 17, Width 2 — filter.go:14-26). */
int vh_add_filter(vh_scene* s, const char* type, const char* name, float width, int res, float peak);
int vh_add_trilight(vh_scene* s, const char* name, const float* p0, const float* p1, const float* p2, const char* shader, int samples);
/* DiskLight (builtin/light/disk.go; registered defaults Segments 20, Samples 1) and SphereLight (builtin/light/sphere.go;
 * defaults Radius 1, Samples 1). Like the reference, PreRender makes each light add its own geom (a fan mesh / a Sphere). */
int vh_add_disklight(vh_scene* s, const char* name, const float* P, const float* lookat, const float* up, float radius, const char* shader,
                     int segments, int samples);
int vh_add_spherelight(vh_scene* s, const char* name, const float* P, float radius, const char* shader, int samples);
/* GeomInstance (builtin/geom/instance/instance.go): geom = name of an earlier PolyMesh; transforms = keys x 16 floats in
 * math.Matrix4 storage (column major); bmin/bmax = n_bounds points each (only the first is used, instance.go:66-70). */
int vh_add_instance(vh_scene* s, const char* name, const char* geom, const float* bmin, const float* bmax, int n_bounds,
                    const float* transforms, int keys);
int vh_set_camera_lookat(vh_scene* s, const float* from, const float* to, const float* up, float roll, float fov, float focal,
                         float aspect, float radius);
/* The Camera node in full (camera.go:48-73): type "LookAt" with n_from / n_to / n_roll motion keys of From / To / Roll, or type
 * "Matrix" with n_mat WorldToLocal matrices (column major). */
int vh_set_camera_keys(vh_scene* s, const char* type, const float* from, int n_from, const float* to, int n_to, const float* roll, int n_roll,
                       const float* up, const float* world_to_local, int n_mat, float fov, float focal, float aspect, float radius);
/* Camera.decomp after PreRender: returns the number of keys; out (may be NULL) receives keys records. */
int vh_camera_decomp(vh_scene* s, VgTransformSRT* out);
/* nodes.Parse (nodes/parser.go:110-131): read a .vnf scene description (text in memory, or a file) and add its nodes in file
 * order, exactly like the vh_add_* calls would. In scope: Globals, Camera (LookAt or Matrix, with motion keys), ShaderStd (constant maps),
 * DebugShader, PolyMesh, GeomInstance, TriLight, DiskLight, SphereLight, Sphere, AiryFilter, GaussianFilter, OutputFloat, OutputHDR,
 * Include (its file is read by vh_prerender, like misc.Include.PreRender; parse messages of such files are left in
 * vh_last_error after a successful vh_prerender, a file that cannot be opened fails it). Returns the number
 * of parse errors the reference would have printed (0 = clean; nodes that parsed are kept, like the reference keeps them),
 * or < 0 if the file cannot be read; vh_last_error holds the "<file>:<line>:<col>: message" lines. */
int vh_parse_vnf(vh_scene* s, const char* text, size_t len, const char* filename);
int vh_load_vnf(vh_scene* s, const char* path);
int vh_globals(vh_scene* s, int32_t* out3 /* XRes, YRes, MaxIter */);
/* core.PostRender (core/core.go:63-73): run the output nodes on the finished frame (xres*yres*3 floats, row-major, top row
 * first). OutputFloat = little-endian float32 RGB rows top to bottom (builtin/driver/outputfloat.go:30-42); OutputHDR =
 * Radiance header + flat RGBE scanlines written bottom row first (image/hdr/writer.go:58-91, hdr.go:26-50). */
int vh_postrender(vh_scene* s, const float* framebuffer, int xres, int yres);
/* convertRGBToRGBE (image/hdr/hdr.go:26-50) for one pixel: known-answer tests bind this. */
int vh_rgbe(float r, float g, float b, uint8_t* out4);
/* core.PreRender (core/core.go:36-61): triangulate, build per-mesh QBVH/MQBVH, light meshes, scene tree, camera matrix. */
int vh_prerender(vh_scene* s);
/* The same, with the QBVH of every static PolyMesh of 32768+ triangles built on the device of `ctx` (vg_build_qbvh): identical
 * nodes, boxes and leaf sets; the order of the triangles inside a leaf differs (see vg_build_qbvh). */
int vh_prerender_device(vh_scene* s, vg_ctx* ctx);
/* Upload the pre-rendered scene into a device context (calls vg_scene_begin .. vg_scene_commit, vg_set_*). */
int vh_upload(vh_scene* s, vg_ctx* ctx, int motion_ref_compat);

/* Inspection of the host-built structures (parity tests compare them with the oracle's). */
int vh_num_geoms(vh_scene* s);
int vh_scene_info(vh_scene* s, int32_t* out4 /* n_nodes, is_motion, keys, n_slots */);
int vh_scene_nodes(vh_scene* s, VgNode* out);
int vh_scene_motion_nodes(vh_scene* s, VgMotionNode* topo, float* boxes);
int vh_scene_geom_order(vh_scene* s, int32_t* out);
int vh_mesh_info(vh_scene* s, int geom_id, int32_t* out6 /* n_nodes, n_tris, keys, n_verts, is_motion, has_normals */);
int vh_mesh_nodes(vh_scene* s, int geom_id, VgNode* out);
int vh_mesh_motion_nodes(vh_scene* s, int geom_id, VgMotionNode* topo, float* boxes);
int vh_mesh_idxp(vh_scene* s, int geom_id, uint32_t* idxp, int32_t* accel_idx);
int vh_camera(vh_scene* s, VgCamera* out);

#ifdef __cplusplus
}
#endif
#endif /* VERMEER_GPU_H */
