// Wavefront path integrator: the device replacement of core.Render / render / core.Trace / ShaderStd.Eval /
// ShaderContext.EvaluateLightSamples (core/render.go:66-218, core/trace.go:40-83, builtin/shader/std.go:77-296,
// core/shader.go:203-402).
//
// One batch = (owned pixels) x (iters_per_batch iterations) paths. Per batch, all on one stream with no host
// round trip (queue sizes stay in device memory):
//   k_raygen      QMC sample (RasterXY/VdC/Sobol) -> camera ray                     render.go:89-124, camera.go:221-323
//   for level 0..3:
//     k_trace_queue<closest>   persistent warps over the ray queue                  core/trace.go:26 -> traverse.cuh
//     k_shade     hit record, ShaderStd, light/BSDF samples -> shadow-ray queue (+ mirror extension-ray queue),
//                 both compacted with warp ballot + prefix popcount + one atomicAdd per warp
//     k_trace_queue<shadow>    any-hit; an occluded sample zeroes its contribution slot
//     k_resolve   per light: sum the slots in the reference's order, /total, *DiffuseColour, ... -> L[level]
//   k_accumulate  fold levels back to front, then fb = (fb*iter + C)/(iter+1) per iteration in order (render.go:127-129)
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "context.h"
#include "kernels.h"
#include "shade.cuh"
#include "traverse.cuh"

namespace vg {

struct DevLight {
  f3 p0, p1, p2;   // tri: vertices; disk: P, T, B; sphere: P
  f3 N;            // tri: normalize((P1-P0)x(P2-P0)), triangle.go:302; disk: N (disk.go:93)
  float inv_area;  // tri: 1/triangleArea, triangle.go:234; disk: 1/(Pi R^2), disk.go:125,179
  float radius;    // disk, sphere
  int type;        // VG_LIGHT_*
  f3 E;            // EvalEmission of the light's shader (std.go:299-316)
  int nsamples;    // 1 << Samples
  int geom;
  int slot_base;
  float spec[12];  // Spectrum.FromRGB(E) per Smits bin, [bin + 1] (k_spectrum_tables); a light sample's Liu is four loads
};

struct __align__(16) DevMat {  // 96 B: six 128-bit loads / stores (the per-vertex copies of textured scenes go through memory)
  f3 emission;     // EmissionColour * EmissionStrength (0 if EmissionStrength unset)
  f3 diff_colour;
  float diff_weight, spec_weight;  // normalised by their sum (std.go:137-139)
  float rough2;                    // DiffuseRoughness^2, default .5^2 (std.go:108-115, orennayar.go:25)
  f3 spec_colour;
  float spec_rough;
  float ior;
  int fresnel_model;     // VG_FRESNEL_* (std.go:172-192)
  f3 fres_refl, fres_edge;  // conductor r, g AFTER the reference's assignment quirk (std.go:187-189 writes the edge tint into refl)
  int bad;  // 1: total weight 0 (the reference panics, std.go:141-143)
  int debug;  // shader.Debug (debug.go:42-49): `emission` holds its Colour, both weights are 0 and no Level check applies
  int pad_;
};
static_assert(sizeof(DevMat) == 96, "DevMat is stored per vertex as six float4");

// ShaderStd's parameters -> the constants the kernels use (std.go:108-143,165-192,299-316). Shared by the host (constant maps,
// once per material) and the device (materials with texture maps, once per shaded vertex: k_surface).
__host__ __device__ inline DevMat derive_mat(const VgMaterial& s) {
  DevMat d;
  d.emission.x = d.emission.y = d.emission.z = 0;
  d.diff_colour = d.emission; d.spec_colour = d.emission;
  d.diff_weight = d.spec_weight = 0; d.rough2 = 0; d.spec_rough = 0; d.ior = 0; d.fresnel_model = 0;
  d.fres_refl = d.emission; d.fres_edge = d.emission;
  d.bad = 0; d.debug = 0; d.pad_ = 0;
  if (s.mask & VG_MAT_DEBUG) {
    // Eval: OutRGB = Colour (resolve_vertex's emission + 0 + 0); no lights, no lobes
    d.debug = 1;
    if (s.mask & VG_MAT_DIFFUSE_COLOUR) { d.emission.x = s.diffuse_colour[0]; d.emission.y = s.diffuse_colour[1]; d.emission.z = s.diffuse_colour[2]; }
    return d;
  }
  if (s.mask & VG_MAT_EMISSION_STRENGTH) {
    const bool hc = (s.mask & VG_MAT_EMISSION_COLOUR) != 0;
    d.emission.x = (hc ? s.emission_colour[0] : 0.0f) * s.emission_strength;
    d.emission.y = (hc ? s.emission_colour[1] : 0.0f) * s.emission_strength;
    d.emission.z = (hc ? s.emission_colour[2] : 0.0f) * s.emission_strength;
  }
  if (s.mask & VG_MAT_DIFFUSE_COLOUR) { d.diff_colour.x = s.diffuse_colour[0]; d.diff_colour.y = s.diffuse_colour[1]; d.diff_colour.z = s.diffuse_colour[2]; }
  const float dw = (s.mask & VG_MAT_DIFFUSE_STRENGTH) ? s.diffuse_strength : 0.0f;
  const float sw = (s.mask & VG_MAT_SPEC1_STRENGTH) ? s.spec1_strength : 0.0f;
  const float tot = dw + sw;
  d.diff_weight = dw / tot;
  d.spec_weight = sw / tot;
  if (tot == 0.0f) { d.bad = 1; d.diff_weight = d.spec_weight = 0; }
  const float rough = (s.mask & VG_MAT_DIFFUSE_ROUGHNESS) ? s.diffuse_roughness : 0.5f;
  d.rough2 = rough * rough;
  if (s.mask & VG_MAT_SPEC1_COLOUR) { d.spec_colour.x = s.spec1_colour[0]; d.spec_colour.y = s.spec1_colour[1]; d.spec_colour.z = s.spec1_colour[2]; }
  d.spec_rough = (s.mask & VG_MAT_SPEC1_ROUGHNESS) ? s.spec1_roughness : 0.5f;
  d.ior = (s.mask & VG_MAT_IOR) ? s.ior : 1.7f;
  // std.go:172-192: Fresnel model of the spec lobe. NewConductor(0, refl, edge) with both defaulting to .5; the reference
  // assigns Spec1FresnelEdge to `refl` (std.go:187-189), so the edge tint stays .5 whatever the scene says.
  d.fresnel_model = (s.mask & VG_MAT_SPEC1_FRESNEL_MODEL) ? s.spec1_fresnel_model : VG_FRESNEL_DIELECTRIC;
  d.fres_refl.x = d.fres_refl.y = d.fres_refl.z = 0.5f;
  d.fres_edge = d.fres_refl;
  if (s.mask & VG_MAT_SPEC1_FRESNEL_REFL) { d.fres_refl.x = s.spec1_fresnel_refl[0]; d.fres_refl.y = s.spec1_fresnel_refl[1]; d.fres_refl.z = s.spec1_fresnel_refl[2]; }
  if (s.mask & VG_MAT_SPEC1_FRESNEL_EDGE) { d.fres_refl.x = s.spec1_fresnel_edge[0]; d.fres_refl.y = s.spec1_fresnel_edge[1]; d.fres_refl.z = s.spec1_fresnel_edge[2]; }
  return d;
}

struct __align__(16) DevHit {
  float t, u, v, w;
  int32_t prim, geom, slot;
  int32_t xf;  // 1 + index of the last instance that reported a hit along this ray (0 = none): ShaderContext.Transform
};

struct RenderParams {
  DevScene sc;
  int xres, yres, nown, P;  // P = paths per batch
  // path index <-> (owned pixel, iteration of the batch): see path_index()
  int pm_B, pm_G, pm_nownB, pm_nitG, pm_A, pm_rem, niters;
  int pm_lB, pm_lG;  // log2 of pm_B, pm_G (both powers of two)
  const int* pix;           // [nown] full-frame pixel index, tile-major
  const uint64_t* scr;      // [nown*6] rows in path order, or (scr_by_pixel) the caller's whole table [xres*yres*6] in raster order
  int scr_by_pixel;
  VgCamera cam;
  const vg::XfSRT* cam_keys;  // Camera.decomp, one per LocalToWorld motion key (camera.go:188-192); cam_nkeys <= 1: cam.local_to_world
  int cam_nkeys;
  const DevMat* mats;
  // texture maps (null / 0 in scenes without them)
  DevTexStore tex;
  const VgMaterial* rawmats;  // the materials as the caller gave them: a textured one is re-derived per vertex
  const MatTex* mat_tex;      // [materials] parameter -> texture bindings
  DevMat* vmats;              // [P] per-vertex materials of TEXTURED materials (mat_tex[matid].mask != 0), indexed like hits[] (queue slot)
  float* diff;                // [12][P] ray differentials per path: DdPdx, DdPdy, DdDdx, DdDdy (core/ray.go:40-44)
  float pd0, pd1;             // core.Image.PixelDelta (camera.go:316-317)
  const DevLight* lights;
  int nlights, S, levels, trace_last_level;
  int nlobes;  // 1: diffuse light slots only (k_shade); 2: + GGX glossy slots (k_shade_generic)
  const QmcTables* qmc;      // byte-sliced RasterXY tables
  const double* filter_cdf;  // cdfV[n] then cdfVU[n*n], or null
  int filter_n;
  double filter_w;
  VgRay* rayq[2];
  int* pathq[2];
  DevHit* hits;
  float* lambda;
  float* time;
  uint8_t* v_mat;
  float* v_invtot;  // [P*nlights]
  float4* contrib;  // [P*S]
  VgRay* sray;      // [P*S]
  int* sslot;       // [P*S]
  float4* L;        // [levels*P] rgb = emission + diffuse
  float4* T;        // [levels*P] rgb = mirror throughput, w = spec weight
  int* counts;      // [0],[1] ray queue sizes, [2] shadow queue size, [3] closest fetch head, [4] shadow fetch head, [5] error flags
  unsigned long long* stats;  // [0] rays [1] shadow rays [2] nodesT [3] trisT (closest) [4] nodesT [5] trisT (shadow)
  float* fb;
};

// Path order inside a batch. A warp's 32 consecutive paths are B pixels x G iterations (B*G = 32): the samples of one pixel
// at different iterations lie within one pixel of each other, so a 4x2-pixel x 4-iteration warp spans a quarter of the image
// area an 8x4-pixel x 1-iteration warp does, stays in fewer BVH leaves, and so do the shadow rays spawned from its hits
// (results are per (pixel, iteration) and do not depend on the order). `pix` enumerates a tile's pixels in Morton order
// inside 8x4 blocks, so any B = 32, 16, 8, 4, 2 consecutive owned pixels form a compact block.
// Region A: full pixel groups x full iteration groups, warp-sized cells; the remainders (nown % B pixels, niters % G
// iterations) follow densely in iteration-major order, so queue slot == path stays a bijection on [0, nown*niters).
__device__ __forceinline__ int path_index(const RenderParams& p, int own, int it) {
  if (own < p.pm_nownB && it < p.pm_nitG)
    return ((((it >> p.pm_lG) * (p.pm_nownB >> p.pm_lB)) + (own >> p.pm_lB)) << 5) + ((it & (p.pm_G - 1)) << p.pm_lB) + (own & (p.pm_B - 1));
  if (it < p.pm_nitG) return p.pm_A + it * p.pm_rem + (own - p.pm_nownB);
  return p.pm_A + p.pm_nitG * p.pm_rem + (it - p.pm_nitG) * p.nown + own;
}
__device__ __forceinline__ void path_decode(const RenderParams& p, int path, int& own, int& it) {
  if (path < p.pm_A) {
    const int cell = path >> 5, r = path & 31, nblk = p.pm_nownB >> p.pm_lB;
    const int cq = cell / nblk;
    own = ((cell - cq * nblk) << p.pm_lB) + (r & (p.pm_B - 1));
    it = (cq << p.pm_lG) + (r >> p.pm_lB);
    return;
  }
  int q = path - p.pm_A;
  if (q < p.pm_nitG * p.pm_rem) {
    it = q / p.pm_rem;
    own = p.pm_nownB + q % p.pm_rem;
    return;
  }
  q -= p.pm_nitG * p.pm_rem;
  it = p.pm_nitG + q / p.nown;
  own = q % p.nown;
}

// One thread per (light or white, bin): the 11 values RGBToSpectrumSmits99 can take for a constant colour (shade.cuh).
__global__ void k_spectrum_tables(DevLight* lights, int nlights, QmcTables* qmc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int which = i / 12, k = i % 12;
  if (which > nlights) return;
  const int bin = k - 1;  // -1 .. 10; slot 11 is never indexed
  if (which == nlights) qmc->white_spec[k] = k <= 10 ? rgb_to_spectrum(1.0f, 1.0f, 1.0f, bin) : 0.0f;
  else lights[which].spec[k] = k <= 10 ? rgb_to_spectrum(lights[which].E.x, lights[which].E.y, lights[which].E.z, bin) : 0.0f;
}

// ---------------------------------------------------------------------------------------------------------
__global__ void k_reset(int* counts, int q0) {
  counts[0] = q0;
  counts[1] = 0;
  counts[2] = 0;
  counts[3] = 0;
  counts[4] = 0;
}
__global__ void k_next_level(int* counts, int qout) {
  counts[1 - qout] = 0;  // the queue that was just consumed becomes the next output queue
  counts[2] = 0;
  counts[3] = 0;
  counts[4] = 0;
}

// filter.Sampler.WarpSample (builtin/filter/filter.go:40-86): linear search of the marginal CDF, then of the conditional
// one; the first bin's `du / w` is the reference's formula, kept as is.
__device__ inline void warp_sample(const double* cdfV, const double* cdfVU, int n, double w, double r0, double r1, double* uo, double* vo) {
  double u = 0, v = 0;
  int uI = -1;
  for (int i = 0; i < n; i++) {
    uI = i;
    const double c = cdfV[i];
    if (r0 < c) {
      if (i == 0) u = (-w / 2) + ((r0 / c) / w);
      else {
        const double c1 = cdfV[i - 1];
        u = (-w / 2) + w * ((double)i + (r0 - c1) / (c - c1)) / (double)(n - 1);
      }
      break;
    }
  }
  const double* row = cdfVU + (size_t)uI * n;
  for (int i = 0; i < n; i++) {
    const double c = row[i];
    if (r1 < c) {
      if (i == 0) v = (-w / 2) + ((r1 / c) / w);
      else {
        const double c1 = row[i - 1];
        v = (-w / 2) + w * ((double)i + (r1 - c1) / (c - c1)) / (double)(n - 1);
      }
      *uo = u;
      *vo = v;
      return;
    }
  }
  *uo = 0;
  *vo = 0;
}

// core/render.go:89-124 + builtin/camera/camera.go:221-323 (ray differentials only in scenes with texture maps)
// MOTION: the camera has several motion keys and every ray recomposes its LocalToWorld at its own Time (camera.go:225-236).
template <bool MOTION>
__global__ void __launch_bounds__(256) k_raygen(const RenderParams p, int iter_base, int niters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.nown * niters) return;
  int own, it;
  path_decode(p, i, own, it);
  const int iter = iter_base + it + 1;  // render() receives iter+1 (render.go:192)
  const int pixel = p.pix[own];
  const int x = pixel % p.xres, y = pixel / p.xres;
  const uint64_t* scr = p.scr + (size_t)(p.scr_by_pixel ? pixel : own) * 6;
  double rasterX, rasterY;
  raster_xy12_tab(p.qmc, (uint32_t)iter, (uint32_t)x, (uint32_t)y, &rasterX, &rasterY);
  const double time = vdc((uint64_t)iter, scr[2]);
  const double lambda = (720 - 450) * vdc((uint64_t)iter, scr[3]) + 450;
  // the lens sample (render.go:96-97) is only consumed by a thin-lens camera (camera.go:254-262): a pinhole skips the two radical inverses
  double lensU = 0.0, lensV = 0.0;
  if (p.cam.radius > 0.0f) {
    lensU = vdc((uint64_t)iter, scr[0]);
    lensV = sobol((uint64_t)iter, scr[1]);
  }
  if (p.filter_cdf) {  // core/render.go:99-107
    const double fx = floor(rasterX), fy = floor(rasterY);
    double u, v;
    warp_sample(p.filter_cdf, p.filter_cdf + p.filter_n, p.filter_n, p.filter_w, rasterX - fx, rasterY - fy, &u, &v);
    rasterX = fx + 0.5 + u;
    rasterY = fy + 0.5 + v;
  }
  const float Sx = (float)(-1.0 + 2.0 * (rasterX / (double)p.xres));
  const float Sy = -(float)(-1.0 + 2.0 * (rasterY / (double)p.yres));

  const float* M = p.cam.local_to_world;
  vg::Mat4 Mt;
  if (MOTION) {
    const float k = (float)time * (float)(p.cam_nkeys - 1);
    const float fk = floorf(k);
    Mt = vg::srt_to_m4(vg::srt_lerp(p.cam_keys[(int)fk], p.cam_keys[(int)ceilf(k)], k - fk));
    M = Mt.m;
  }
  const float camu = Sx * p.cam.tan_theta_focal;
  const float camv = Sy * (p.cam.tan_theta_focal / p.cam.aspect);
  // s = camu*U + camv*V - Focal*W with the canonical basis (camera.go:248-252); the products with 0/1 are exact
  f3 s = mk3(camu, camv, 0.0f - p.cam.focal);
  f3 e = mk3(0, 0, 0);
  if (p.cam.radius > 0.0f) {
    float lx, ly;
    uniform_disk2d(p.cam.radius, (float)lensU, (float)lensV, &lx, &ly);
    e = mk3(lx, ly, 0.0f);
    s = sub3(s, e);
  }
  const f3 d = mk3(M[0] * s.x + M[4] * s.y + M[8] * s.z, M[1] * s.x + M[5] * s.y + M[9] * s.z, M[2] * s.x + M[6] * s.y + M[10] * s.z);
  const f3 D = normalize3(d);
  const f3 O = mk3(M[0] * e.x + M[4] * e.y + M[8] * e.z + M[12], M[1] * e.x + M[5] * e.y + M[9] * e.z + M[13],
                   M[2] * e.x + M[6] * e.y + M[10] * e.z + M[14]);
  float4* rp = reinterpret_cast<float4*>(p.rayq[0] + i);
  rp[0] = make_float4(O.x, O.y, O.z, D.x);
  rp[1] = make_float4(D.y, D.z, __int_as_float(0x7f800000), (float)time);
  p.pathq[0][i] = i;
  p.lambda[i] = (float)lambda;
  p.time[i] = (float)time;
  if (p.diff) {
    // camera.go:300-306: DdPdx = DdPdy = 0; DdDdx/y = d(normalize d)/d(right|up) with the UNnormalised direction d
    const f3 right = mk3(M[0], M[1], M[2]), up = mk3(M[4], M[5], M[6]);
    const float dd = dot3(d, d);
    const float k = 1 / (dd * sqrtf(dd));
    const f3 ddx = scale3(k, sub3(scale3(dd, right), scale3(dot3(d, right), d)));
    const f3 ddy = scale3(k, sub3(scale3(dd, up), scale3(dot3(d, up), d)));
    float* q = p.diff + i;
    const size_t P = (size_t)p.P;
    q[0] = 0.f; q[P] = 0.f; q[2 * P] = 0.f; q[3 * P] = 0.f; q[4 * P] = 0.f; q[5 * P] = 0.f;
    q[6 * P] = ddx.x; q[7 * P] = ddx.y; q[8 * P] = ddx.z; q[9 * P] = ddy.x; q[10 * P] = ddy.y; q[11 * P] = ddy.z;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Persistent traversal over a device-resident queue. MODE 0: closest hit -> hits[]. MODE 1: shadow rays (any hit);
// an occluded sample zeroes its contribution slot.
template <int MODE>
struct QueueIO {
  static constexpr bool kHitRecord = MODE == 0;
  static constexpr int kRefillIdleClosest = VG_REFILL_IDLE_CLOSEST_QUEUE;  // the shadow queue only wants "occluded or not": no U, V, W / ids of the hit
  const RenderParams& p;
  const VgRay* rays;
  int n;
  int* head;
  __device__ __forceinline__ long long fetch(int c) { return (long long)atomicAdd(head, c); }
  __device__ __forceinline__ long long size() const { return n; }
  __device__ __forceinline__ const VgRay* ray_ptr() const { return rays; }
  __device__ __forceinline__ void load(long long i, RayState& r) const {
    const float4* rp = reinterpret_cast<const float4*>(rays + i);
    const float4 a = __ldg(rp), b = __ldg(rp + 1);
    r.ox = a.x; r.oy = a.y; r.oz = a.z;
    r.dx = a.w; r.dy = b.x; r.dz = b.y;
    r.tclosest = b.z;
    r.time = b.w;
  }
  __device__ __forceinline__ void store(long long i, const RayState& r, const HitState& h, bool overflow) const {
    if (overflow) atomicOr(p.counts + 5, 4);
    if (MODE == 0) {
      *reinterpret_cast<float4*>(&p.hits[i]) = make_float4(r.tclosest, h.u, h.v, h.w);
      *(reinterpret_cast<int4*>(&p.hits[i]) + 1) = make_int4(h.prim, h.geom, h.slot, h.xf_last + 1);
    } else {
      if (h.prim != -1) p.contrib[p.sslot[i]] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
};

template <int MODE, int VARIANT>
__global__ void __launch_bounds__(kTraceBlock, MODE == 1 ? ((VARIANT & 64) ? VG_TRACE_MIN_BLOCKS_SHADOW_MOTION : VG_TRACE_MIN_BLOCKS_SHADOW) : ((VARIANT & 2) ? ((VARIANT & 64) ? VG_TRACE_MIN_BLOCKS_COOP_MOTION : VG_TRACE_MIN_BLOCKS_COOP) : VG_TRACE_MIN_BLOCKS)) k_trace_queue(const RenderParams p, int q) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [warps x warp_smem_bytes(VARIANT) scratch] [threads x VG_SMEM_STACK stack entries]
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5;
  Stack st;
  st.bind(smem_raw + nwarps * warp_smem_bytes(VARIANT));
  const int lane = threadIdx.x & 31;
  const int n = MODE == 0 ? p.counts[q] : p.counts[2];
  // per-thread sums of the packed per-ray counters (a persistent thread sees n / (SMs x CTAs x 128) rays: 32 bits are ample)
  unsigned nodes_acc32 = 0, tris_acc32 = 0;
  QueueIO<MODE> io{p, MODE == 0 ? p.rayq[q] : p.sray, n, p.counts + (MODE == 0 ? 3 : 4)};
  trace_persistent<MODE == 1, VARIANT>(p.sc, io, st, smem_raw + warp * warp_smem_bytes(VARIANT), nodes_acc32, tris_acc32);
  unsigned long long nodes_acc = nodes_acc32, tris_acc = tris_acc32;
  for (int o = 16; o > 0; o >>= 1) {
    nodes_acc += __shfl_down_sync(0xffffffffu, nodes_acc, o);
    tris_acc += __shfl_down_sync(0xffffffffu, tris_acc, o);
  }
  if (lane == 0) {
    atomicAdd(p.stats + (MODE == 0 ? 2 : 4), nodes_acc);
    atomicAdd(p.stats + (MODE == 0 ? 3 : 5), tris_acc);
  }
#ifdef VG_STACK_STATS
  {
    int m = st.maxsp;
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    if (lane == 0) atomicMax(p.stats + 6, (unsigned long long)m);
  }
#endif
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(p.stats + 0, (unsigned long long)n);  // every TraceProbe counts (core/stats.go:26-33)
    if (MODE == 1) atomicAdd(p.stats + 1, (unsigned long long)n);
  }
}

// ---------------------------------------------------------------------------------------------------------
struct ShadeCtx {
  f3 P, Poffset, N, Ng, DdPdu, DdPdv;
};
struct Plane4 {
  f3 n;
  float d;
};
__device__ __forceinline__ Plane4 mk4(f3 n, float d) { Plane4 r; r.n = n; r.d = d; return r; }

__device__ __forceinline__ f3 ld3(const float4* p) {
  const float4 v = __ldg(p);
  return mk3(v.x, v.y, v.z);
}

// polymesh/trace.go:276-360,504-515 (static) / :625-667 (motion), then ShaderContext.ApplyTransform (core/shader.go:129-135,
// identity transform: only the re-normalisations remain).
// APPLY_IDENTITY = false leaves N, Ng, DdPdu, DdPdv as the geom wrote them, for a caller that applies a real transform.
// FAST: the normalisations through MUFU.RSQ (shade.cuh: normalize3t) instead of a correctly rounded sqrt + divide; the
// reference's own Vec3Normalize is RSQRTSS + one Newton step, so neither form is bit-comparable with it.
template <bool APPLY_IDENTITY = true, bool FAST = false>
__device__ inline void build_context(const RenderParams& p, const DevHit& h, float time, ShadeCtx& c) {
  const DevGeom g = p.sc.geoms[h.geom];
  const float U = h.u, V = h.v, W = h.w;
  f3 E0, E1, E2;
  const bool motion = g.keys > 1;
  if (!motion) {
    const float4* tp = p.sc.tris + (size_t)h.slot * kTriStride;
    E0 = ld3(tp); E1 = ld3(tp + 1); E2 = ld3(tp + 2);
  } else {
    const float k = time * (float)(g.keys - 1);
    const float fk = floorf(k);
    const float tm = k - fk, om = 1.0f - tm;
    const int key = (int)fk, key2 = (int)ceilf(k);
    const float4* ta = p.sc.mtris + ((size_t)h.slot + (size_t)key * g.tri_key_stride) * 3;
    const float4* tb = p.sc.mtris + ((size_t)h.slot + (size_t)key2 * g.tri_key_stride) * 3;
    const f3 a0 = ld3(ta), a1 = ld3(ta + 1), a2 = ld3(ta + 2), b0 = ld3(tb), b1 = ld3(tb + 1), b2 = ld3(tb + 2);
    E0 = mk3(om * a0.x + tm * b0.x, om * a0.y + tm * b0.y, om * a0.z + tm * b0.z);
    E1 = mk3(om * a1.x + tm * b1.x, om * a1.y + tm * b1.y, om * a1.z + tm * b1.z);
    E2 = mk3(om * a2.x + tm * b2.x, om * a2.y + tm * b2.y, om * a2.z + tm * b2.z);
  }
  const float xAbs = fabsf(U * E0.x) + fabsf(V * E1.x) + fabsf(W * E2.x);
  const float yAbs = fabsf(U * E0.y) + fabsf(V * E1.y) + fabsf(W * E2.y);
  const float zAbs = fabsf(U * E0.z) + fabsf(V * E1.z) + fabsf(W * E2.z);
  const f3 e0 = sub3(E1, E0), e1 = sub3(E2, E0);
  f3 Ng = normalize3t<FAST>(cross3(e0, e1));
  f3 N = Ng;
  if (!motion && g.normal_base >= 0 && p.sc.tri_normals) {
    const float4* np = p.sc.tri_normals + (size_t)h.slot * 3;
    const f3 n0 = ld3(np), n1 = ld3(np + 1), n2 = ld3(np + 2);
    N = normalize3t<FAST>(mk3(U * n0.x + V * n1.x + W * n2.x, U * n0.y + V * n1.y + W * n2.y, U * n0.z + V * n1.z + W * n2.z));
  }
  const float g7 = (7.0f * 5.9604644775390625e-08f) / (1 - 7.0f * 5.9604644775390625e-08f);  // math/ferror.go:22-24, Gamma(7)
  const float d = g7 * xAbs * fabsf(Ng.x) + g7 * yAbs * fabsf(Ng.y) + g7 * zAbs * fabsf(Ng.z);
  c.Poffset = scale3(d, Ng);
  c.P = mk3(U * E0.x + V * E1.x + W * E2.x, U * E0.y + V * E1.y + W * E2.y, U * E0.z + V * E1.z + W * E2.z);
  if (!motion) {
    f3 axisu = sub3(mk3(1, 0, 0), scale3(Ng.x, Ng));
    if (len2_3(axisu) < 0.1f || fabsf(dot3(axisu, Ng)) > 0.3f) axisu = sub3(mk3(0, 0, 1), scale3(Ng.z, Ng));
    c.DdPdu = normalize3t<FAST>(axisu);
    c.DdPdv = cross3(Ng, c.DdPdu);
  } else {
    c.DdPdu = e0;
    c.DdPdv = e1;
  }
  if (!APPLY_IDENTITY) {
    c.N = N;
    c.Ng = Ng;
    return;
  }
  // ApplyTransform
  c.N = normalize3t<FAST>(N);
  c.Ng = normalize3t<FAST>(Ng);
  c.DdPdu = normalize3t<FAST>(c.DdPdu);
  c.DdPdv = normalize3t<FAST>(c.DdPdv);
}

// Scenes with texture maps: one pass over the hit queue between traversal and shading.
//   * texture coordinates and their screen-space derivatives at the hit: Ray.DifferentialTransfer (core/ray.go:95-104), the
//     barycentric-plane construction of polymesh/trace.go:350-502 (U, V, DdNdx/y, Dduvdx/y);
//   * every parameter of the hit material that is a maps.Texture / maps.TextureTrilinear is looked up (texture.cuh) and the
//     material constants are re-derived for this vertex (derive_mat) into vmats[queue slot];
//   * the path's ray differentials are replaced, in place, by those of the mirror ray Ray.Init(RayTypeReflected) would
//     build from this context (core/ray.go:72-87), whether or not the shader goes on to spawn it.
// Static meshes only (prepare() refuses textured scenes with other geoms).
// parameter slot k of a VgMaterial (bit index of its VG_MAT_* flag) <- the colour a texture map returned
__device__ __forceinline__ void set_param(VgMaterial& s, int k, const float c[3], int chan) {
  const float f = chan == 0 ? c[0] : (chan == 1 ? c[1] : c[2]);
  switch (k) {
    case 0: s.emission_colour[0] = c[0]; s.emission_colour[1] = c[1]; s.emission_colour[2] = c[2]; break;
    case 1: s.emission_strength = f; break;
    case 2: s.diffuse_colour[0] = c[0]; s.diffuse_colour[1] = c[1]; s.diffuse_colour[2] = c[2]; break;
    case 3: s.diffuse_strength = f; break;
    case 4: s.diffuse_roughness = f; break;
    case 5: s.spec1_colour[0] = c[0]; s.spec1_colour[1] = c[1]; s.spec1_colour[2] = c[2]; break;
    case 6: s.spec1_strength = f; break;
    case 7: s.spec1_roughness = f; break;
    case 8: s.ior = f; break;
    case 10: s.spec1_fresnel_refl[0] = c[0]; s.spec1_fresnel_refl[1] = c[1]; s.spec1_fresnel_refl[2] = c[2]; break;
    case 11: s.spec1_fresnel_edge[0] = c[0]; s.spec1_fresnel_edge[1] = c[1]; s.spec1_fresnel_edge[2] = c[2]; break;
    default: break;
  }
}

// COOP: the probes of a warp's texture lookups are shared evenly between its lanes (texture.cuh: tex_sample_warp); false = one
// lane per lookup (bit-identical; kept for the A/B measurement and as the cross-check in the tests).
#ifndef VG_SURFACE_MIN_BLOCKS
#define VG_SURFACE_MIN_BLOCKS 5
#endif
template <bool COOP>
__global__ void __launch_bounds__(128, VG_SURFACE_MIN_BLOCKS) k_surface(const RenderParams p, int level, int qin, int fast) {
  __shared__ TexWarpScratch scratch[COOP ? 4 : 1];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nq = p.counts[qin];
  if ((i & ~31) >= nq) return;  // whole warps leave together: the texture phase below is warp-collective
  int prim = -1, geom = 0, slot = 0, matid = 255;
  float T = 0, U = 0, V = 0, W = 0;
  if (i < nq) {
    const int4 h1 = *(reinterpret_cast<const int4*>(&p.hits[i]) + 1);
    prim = h1.x; geom = h1.y; slot = h1.z;
  }
  DevGeom g;
  g.normal_base = -1; g.uv_base = -1; g.prim_base = 0; g.xform = -1;
  int hxf = 0;  // 1 + the instance whose transform the context ends up with (DevHit::xf, quirk o)
  if (i < nq) hxf = (*(reinterpret_cast<const int4*>(&p.hits[i]) + 1)).w;
  if (prim >= 0) {
    const float4 h0 = *reinterpret_cast<const float4*>(&p.hits[i]);
    T = h0.x; U = h0.y; V = h0.z; W = h0.w;
    g = p.sc.geoms[geom];
    matid = p.sc.prim_material[g.prim_base + prim];
  }
  // a vertex needs this pass if its material has a texture map, or if it may spawn the mirror ray whose differentials are
  // produced here; an untextured, purely diffuse/glossy vertex ends its path's use of the differentials
  bool active = matid != 255;
  if (active) active = p.mat_tex[matid].mask != 0u || p.mats[matid].spec_weight > 0.0f;
  int path = 0;
  f3 D = mk3(0, 0, 1);
  TexCoord tc;
  tc.U = tc.V = tc.dudx = tc.dvdx = tc.dudy = tc.dvdy = 0.f;
  tc.pd0 = p.pd0;
  tc.pd1 = p.pd1;
  f3 dPdx = mk3(0, 0, 0), dPdy = dPdx, oDx = dPdx, oDy = dPdx;
  if (active) {
    path = p.pathq[qin][i];
    const float4* rp = reinterpret_cast<const float4*>(p.rayq[qin] + i);
    const float4 ra = rp[0], rb = rp[1];
    D = mk3(ra.w, rb.x, rb.y);
    // A hit inside a GeomInstance: PolyMesh.TraceElems runs while the ray is in OBJECT space (Instance.Trace, instance.go:86-95)
    // and calls Ray.DifferentialTransfer there with the object-space direction but the untouched world-space differentials
    // (trace.go:360); core.Trace then repeats the transfer with the restored world ray (core/trace.go:67). Both are mirrored:
    // Dobj feeds the texture footprint, D what the path keeps.
    f3 Dobj = D;
    if (g.xform >= 0) {
      float o6[6];
      xf_object_ray(p.sc, g.xform, rb.w, ra.x, ra.y, ra.z, D.x, D.y, D.z, o6);
      Dobj = mk3(o6[3], o6[4], o6[5]);
    }

    const float4* tp = p.sc.tris + (size_t)slot * kTriStride;
    const f3 E0 = ld3(tp), E1 = ld3(tp + 1), E2 = ld3(tp + 2);
    const f3 Ng = normalize3(cross3(sub3(E1, E0), sub3(E2, E0)));
    f3 n0 = mk3(0, 0, 0), n1 = n0, n2 = n0;
    const bool has_n = g.normal_base >= 0 && p.sc.tri_normals;
    f3 N = Ng;  // the interpolated normal BEFORE normalisation (trace.go:326-336)
    if (has_n) {
      const float4* np = p.sc.tri_normals + (size_t)slot * 3;
      n0 = ld3(np); n1 = ld3(np + 1); n2 = ld3(np + 2);
      N = mk3(U * n0.x + V * n1.x + W * n2.x, U * n0.y + V * n1.y + W * n2.y, U * n0.z + V * n1.z + W * n2.z);
    }
    float2 t0 = make_float2(0, 0), t1 = t0, t2 = t0;
    const bool has_uv = g.uv_base >= 0 && p.sc.tri_uv;
    if (has_uv) {
      const float2* up = p.sc.tri_uv + (size_t)slot * 3;
      t0 = __ldg(up); t1 = __ldg(up + 1); t2 = __ldg(up + 2);
      tc.U = U * t0.x + V * t1.x + W * t2.x;
      tc.V = U * t0.y + V * t1.y + W * t2.y;
    } else {
      tc.U = U;
      tc.V = V;
    }

    // incoming ray differentials
    const float* q = p.diff + path;
    const size_t P = (size_t)p.P;
    const f3 rPx = mk3(q[0], q[P], q[2 * P]), rPy = mk3(q[3 * P], q[4 * P], q[5 * P]);
    const f3 rDx = mk3(q[6 * P], q[7 * P], q[8 * P]), rDy = mk3(q[9 * P], q[10 * P], q[11 * P]);
    // DifferentialTransfer (core/ray.go:95-104)
    const float DNg = dot3(Dobj, Ng);
    const f3 ax = mad3(rPx, rDx, T), ay = mad3(rPy, rDy, T);
    const bool fdv = fast != 0;  // approximate reciprocals on the FAST path (tolerance-only quantities, see texture.cuh)
    const float dtdx = tex_div(-dot3(ax, Ng), DNg, fdv);
    const float dtdy = tex_div(-dot3(ay, Ng), DNg, fdv);
    dPdx = add3(ax, scale3(dtdx, Dobj));
    dPdy = add3(ay, scale3(dtdy, Dobj));

    // barycentric planes (trace.go:362-436): n = Ng x edge, normalised, then scaled so that the opposite vertex evaluates to 1
    auto plane = [&](f3 a, f3 b, f3 on) {
      f3 n = mk3(Ng.y * (a.z - b.z) - Ng.z * (a.y - b.y), Ng.z * (a.x - b.x) - Ng.x * (a.z - b.z), Ng.x * (a.y - b.y) - Ng.y * (a.x - b.x));
      const float qn = sqrtf(n.x * n.x + n.y * n.y + n.z * n.z);
      if (fast) {
        const float rq = __fdividef(1.0f, qn);
        n.x *= rq; n.y *= rq; n.z *= rq;
      } else {
        n.x /= qn; n.y /= qn; n.z /= qn;
      }
      const float d = -on.x * n.x - on.y * n.y - on.z * n.z;
      return mk4(n, d);
    };
    // alpha: edge E2-E1 through E1, opposite E0 (evaluated as E0.n + d); beta: E2-E0 through E0, opposite E1 (n.E1 + d);
    // gamma: E1-E0 through E0, opposite E2 (n.E2 + d) -- the operand order of each `l` is the reference's
    f3 na, nb, ng;
    {
      const Plane4 A = plane(E2, E1, E1);
      const float l = E0.x * A.n.x + E0.y * A.n.y + E0.z * A.n.z + A.d;
      na = fast ? scale3(__fdividef(1.0f, l), A.n) : mk3(A.n.x / l, A.n.y / l, A.n.z / l);
      const Plane4 B = plane(E2, E0, E0);
      const float lb = B.n.x * E1.x + B.n.y * E1.y + B.n.z * E1.z + B.d;
      nb = fast ? scale3(__fdividef(1.0f, lb), B.n) : mk3(B.n.x / lb, B.n.y / lb, B.n.z / lb);
      const Plane4 G = plane(E1, E0, E0);
      const float lg = G.n.x * E2.x + G.n.y * E2.y + G.n.z * E2.z + G.d;
      ng = fast ? scale3(__fdividef(1.0f, lg), G.n) : mk3(G.n.x / lg, G.n.y / lg, G.n.z / lg);
    }
    const float alphax = na.x * dPdx.x + na.y * dPdx.y + na.z * dPdx.z;
    const float betax = nb.x * dPdx.x + nb.y * dPdx.y + nb.z * dPdx.z;
    const float gammax = ng.x * dPdx.x + ng.y * dPdx.y + ng.z * dPdx.z;
    const float alphay = na.x * dPdy.x + na.y * dPdy.y + na.z * dPdy.z;
    const float betay = nb.x * dPdy.x + nb.y * dPdy.y + nb.z * dPdy.z;
    const float gammay = ng.x * dPdy.x + ng.y * dPdy.y + ng.z * dPdy.z;

    f3 dndx = mk3(0, 0, 0), dndy = mk3(0, 0, 0);
    if (has_n) {
      dndx = mk3(alphax * n0.x + betax * n1.x + gammax * n2.x, alphax * n0.y + betax * n1.y + gammax * n2.y, alphax * n0.z + betax * n1.z + gammax * n2.z);
      dndy = mk3(alphay * n0.x + betay * n1.x + gammay * n2.x, alphay * n0.y + betay * n1.y + gammay * n2.y, alphay * n0.z + betay * n1.z + gammay * n2.z);
    }
    const float NN = dot3(N, N);
    const float kN = tex_div(1.0f, NN * sqrtf(NN), fdv);
    const f3 DdNdx = scale3(kN, sub3(scale3(NN, dndx), scale3(dot3(N, dndx), N)));
    const f3 DdNdy = scale3(kN, sub3(scale3(NN, dndy), scale3(dot3(N, dndy), N)));
    if (has_uv) {
      tc.dudx = alphax * t0.x + betax * t1.x + gammax * t2.x;
      tc.dvdx = alphax * t0.y + betax * t1.y + gammax * t2.y;
      tc.dudy = alphay * t0.x + betay * t1.x + gammay * t2.x;
      tc.dvdy = alphay * t0.y + betay * t1.y + gammay * t2.y;
    } else {  // trace.go:495-501 as written
      tc.dudx = alphax * 0 + betax * 1 + gammax * 0;
      tc.dvdx = alphax * 0 + betax * 0 + gammax * 1;
      tc.dudy = alphay * 0 + betay * 1 + gammay * 0;
      tc.dvdy = alphay * 0 + betay * 0 + gammay * 1;
    }

    // differentials of the mirror ray (core/ray.go:72-87); sc.N is the shading normal after ApplyTransform (normalised once more,
    // like build_context), sc.DdDdx/y are the incoming ray's
    f3 Ns = has_n ? normalize3(normalize3(N)) : normalize3(Ng);
    if (g.xform >= 0 || hxf > 0) {
      // the second DifferentialTransfer (core/trace.go:67): world ray, the geom's own (object-space) Ng
      const float DNgW = dot3(D, Ng);
      dPdx = add3(ax, scale3(tex_div(-dot3(ax, Ng), DNgW, fdv), D));
      dPdy = add3(ay, scale3(tex_div(-dot3(ay, Ng), DNgW, fdv), D));
      // sc.N after ApplyTransform (core/shader.go:129-135) with the transform the context ends up with
      if (hxf > 0) {
        const DevXform x = p.sc.xforms[hxf - 1];
        Mat4 M, Minv;
        if (x.nkeys == 1) Minv = p.sc.xf_static[2 * (hxf - 1) + 1];
        else xf_matrices(p.sc.xf_keys + x.key_base, x.nkeys, rb.w, &M, &Minv);
        const Mat4 MinvT = m4_transpose(Minv);
        const f3 n0w = has_n ? normalize3(N) : Ng;
        float o[3];
        m4_mul_vec(MinvT, n0w.x, n0w.y, n0w.z, o);
        Ns = normalize3(mk3(o[0], o[1], o[2]));
      }
    }
    const float RdN = dot3(D, Ns);
    const float DdotNdx = dot3(rDx, Ns) + dot3(D, DdNdx);
    const float DdotNdy = dot3(rDy, Ns) + dot3(D, DdNdy);
    oDx = mad3(rDx, add3(scale3(RdN, DdNdx), scale3(DdotNdx, Ns)), -2.0f);
    oDy = mad3(rDy, add3(scale3(RdN, DdNdy), scale3(DdotNdy, Ns)), -2.0f);
  }

  // the material at this vertex: every parameter that is a texture map is looked up, then the constants are re-derived.
  // One copy of the lookup code, looped over the slots some lane of the warp needs (the unrolled form was instruction-cache bound)
  const MatTex* mt = active ? p.mat_tex + matid : nullptr;
  const unsigned mask = active ? mt->mask : 0u;
  const bool any = mask != 0u;
  unsigned wm = __reduce_or_sync(0xffffffffu, mask);
  if (wm) {
    VgMaterial s;
    if (any) s = p.rawmats[matid];
#pragma unroll 1
    while (wm) {
      const int k = __ffs(wm) - 1;
      wm &= wm - 1;
      const int tex = (mask >> k) & 1u ? mt->slot[k].tex : -1;
      float c[3] = {0.f, 0.f, 0.f};
      if (COOP) {
        TexProbeSetup su;
        su.tex = -1; su.nprobes = 0; su.feline = 0;
        su.U = su.V = su.dU = su.dV = su.lod = su.mr2 = su.n0 = 0.f;
        if (tex >= 0) su = tex_setup(p.tex, tex, mt->slot[k].filter, tc, fast != 0);
        tex_sample_warp(p.tex, su, &scratch[(threadIdx.x >> 5) & 3], c);
      } else if (tex >= 0) {
        const TexProbeSetup su = tex_setup(p.tex, tex, mt->slot[k].filter, tc, fast != 0);
        TexAccum acc;
        acc.init();
        for (int j = 0; j < su.nprobes; j++) acc.add(su, tex_probe(p.tex, su, j));
        acc.finish(su, c);
      }
      if (tex >= 0) set_param(s, k, c, mt->slot[k].chan);
    }
    if (any) p.vmats[i] = derive_mat(s);
  }

  if (active) {
    float* q = p.diff + path;
    const size_t P = (size_t)p.P;
    q[0] = dPdx.x; q[P] = dPdx.y; q[2 * P] = dPdx.z; q[3 * P] = dPdy.x; q[4 * P] = dPdy.y; q[5 * P] = dPdy.z;
    q[6 * P] = oDx.x; q[7 * P] = oDx.y; q[8 * P] = oDx.z; q[9 * P] = oDy.x; q[10 * P] = oDy.y; q[11 * P] = oDy.z;
  }
}

// ShaderContext.ApplyTransform (core/shader.go:129-135) with the transform the last hit instance left in the context
// (instance.go:107-111) at this ray's time. Poffset stays as the geom computed it, in object space, like the reference.
__device__ inline void apply_instance_transform(const RenderParams& p, int xi, float time, ShadeCtx& c) {
  const DevXform x = p.sc.xforms[xi];
  Mat4 M, Minv;
  if (x.nkeys == 1) {
    M = p.sc.xf_static[2 * xi];
    Minv = p.sc.xf_static[2 * xi + 1];
  } else {
    xf_matrices(p.sc.xf_keys + x.key_base, x.nkeys, time, &M, &Minv);
  }
  const Mat4 MinvT = m4_transpose(Minv);
  float o[3];
  m4_mul_point(M, c.P.x, c.P.y, c.P.z, o);
  c.P = mk3(o[0], o[1], o[2]);
  m4_mul_vec(MinvT, c.N.x, c.N.y, c.N.z, o);
  c.N = normalize3(mk3(o[0], o[1], o[2]));
  m4_mul_vec(MinvT, c.Ng.x, c.Ng.y, c.Ng.z, o);
  c.Ng = normalize3(mk3(o[0], o[1], o[2]));
  m4_mul_vec(M, c.DdPdu.x, c.DdPdu.y, c.DdPdu.z, o);
  c.DdPdu = normalize3(mk3(o[0], o[1], o[2]));
  m4_mul_vec(M, c.DdPdv.x, c.DdPdv.y, c.DdPdv.z, o);
  c.DdPdv = normalize3(mk3(o[0], o[1], o[2]));
}

// sphere.Sphere.Trace hit record (builtin/geom/sphere/trace.go:17-47), then ApplyTransform's re-normalisations
__device__ inline void build_context_sphere(const RenderParams& p, const DevHit& h, f3 Ro, f3 Rd, ShadeCtx& c) {
  const f3 centre = ld3(p.sc.tris + (size_t)h.slot * kTriStride);
  c.P = mad3(Ro, Rd, h.t);
  const f3 N = normalize3(sub3(c.P, centre));
  c.Poffset = scale3(0.001f, N);
  c.N = normalize3(N);
  c.Ng = normalize3(N);
  c.DdPdu = normalize3(mk3(1, 0, 0));
  c.DdPdv = normalize3(mk3(0, 0, 1));
}

// Warp-aggregated append: every lane of the warp must call this. Returns the slot index (or -1).
__device__ __forceinline__ int warp_append(int* counter, bool want) {
  const unsigned mask = __ballot_sync(0xffffffffu, want);
  if (mask == 0) return -1;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(mask) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  return want ? base + __popc(mask & ((1u << lane) - 1u)) : -1;
}

struct LightRec {
  bool valid;
  f3 Ld;
  float Ldist;
  float pdf;
};
struct BsdfRec {
  bool valid;
  f3 Ld;
  float Ldist;
  float pdf;  // float32(bsdf.PDF(wo)): the reference keeps the float64 but only ever uses it rounded (core/shader.go:323)
  float pdfLight;
};

// light.Tri.SampleArea, sample `i` of `n` (builtin/light/triangle.go:232-343)
template <bool FAST>
__device__ inline LightRec light_sample_r(const DevLight& L, const ShadeCtx& c, bool by_area, const SphTri& sph, double r0, double r1) {
  LightRec r;
  f3 Pl;
  if (by_area) {
    const double sq = FAST ? (double)sqrtf((float)(1 - r0)) : sqrt(1 - r0);
    const f3 a = scale3((float)(r1 * sq), sub3(L.p1, L.p0)), b = scale3((float)(1 - sq), sub3(L.p2, L.p0));
    Pl = mk3(L.p0.x + a.x + b.x, L.p0.y + a.y + b.y, L.p0.z + a.z + b.z);
  } else {
    const f3 x = sample_spherical_triangle<FAST>(sph, r0, r1);
    const float t = ray_plane<FAST>(c.P, x, L.p0, L.N);
    Pl = mad3(c.P, x, t);
    r.pdf = FAST ? __fdividef(1.0f, sph.area) : (float)(1 / (double)sph.area);
  }
  const f3 D = sub3(Pl, c.P);
  r.Ldist = length3t<FAST>(D);
  r.Ld = normalize3t<FAST>(D);
  r.valid = !(dot3(r.Ld, L.N) > 0 || dot3(r.Ld, c.Ng) < 0);
  if (by_area) r.pdf = divt<FAST>(L.inv_area * (r.Ldist * r.Ldist), fabsf(dot3(r.Ld, L.N)));
  return r;
}

template <bool FAST>
__device__ inline LightRec light_sample(const DevLight& L, const ShadeCtx& c, bool by_area, const SphTri& sph, long long I, int n, int i,
                                        uint64_t scr0, uint64_t scr1) {
  const uint64_t idx = (uint64_t)(I * n + i);
  return light_sample_r<FAST>(L, c, by_area, sph, vdc(idx, scr0), sobol(idx, scr1));
}

// BSDF sample `i` of `h` through Light.ValidSample (core/shader.go:212-229, triangle.go:136-230)
// The direction and pdf of an Oren-Nayar BSDF sample (core/shader.go:212-218): independent of the light.
template <bool FAST>
__device__ inline bool bsdf_dir(const Frame& fr, double r0, double r1, f3* wo, float* pdf) {
  *wo = normalize3t<FAST>(basis_expand(fr.U, fr.V, fr.N, cosine_hemisphere<FAST>(r0, r1)));
  if (FAST) {
    *pdf = oren_pdf32<true>(fr, *wo);
    return *pdf > 0;
  }
  const double pd = oren_pdf(fr, *wo);  // the reference tests the float64 value (core/shader.go:218)
  *pdf = (float)pd;
  return pd > 0;
}
// Light.ValidSample for that direction (triangle.go:136-230)
template <bool FAST>
__device__ inline BsdfRec bsdf_hit(const DevLight& L, const ShadeCtx& c, bool have_sph, const SphTri& sph, bool dir_ok, f3 wo, float pdf) {
  BsdfRec r;
  r.valid = false;
  r.pdfLight = 0.0f;
  r.Ldist = 0.0f;
  r.Ld = mk3(0, 0, 1);
  r.pdf = pdf;
  if (!dir_ok) return r;
  f3 Pl;
  if (!ray_triangle<FAST>(c.P, wo, L.p0, L.p1, L.p2, &Pl)) return r;
  // NOTE: the horizon test of ValidSample is taken at the point ON THE LIGHT (triangle.go:145-147), kept as is
  const bool by_area = dot3(c.Ng, sub3(L.p0, Pl)) < 0 || dot3(c.Ng, sub3(L.p1, Pl)) < 0 || dot3(c.Ng, sub3(L.p2, Pl)) < 0;
  float pdfl;
  if (by_area) {
    pdfl = L.inv_area;
  } else {
    const float area = have_sph ? sph.area : spherical_setup<FAST>(L.p0, L.p1, L.p2, c.P).area;
    pdfl = FAST ? __fdividef(1.0f, area) : (float)(double)(1 / area);
  }
  const f3 D = sub3(Pl, c.P);
  r.Ldist = length3t<FAST>(D);
  r.Ld = normalize3t<FAST>(D);
  if (dot3(r.Ld, L.N) > 0 || dot3(r.Ld, c.Ng) < 0) return r;
  r.pdfLight = by_area ? divt<FAST>(pdfl * (r.Ldist * r.Ldist), fabsf(dot3(r.Ld, L.N))) : pdfl;
  r.valid = true;
  return r;
}
// BSDF sample `i` of `h` through Light.ValidSample (core/shader.go:212-229)
template <bool FAST>
__device__ inline BsdfRec bsdf_sample(const DevLight& L, const ShadeCtx& c, const Frame& fr, bool have_sph, const SphTri& sph, long long I, int h,
                                      int i, uint64_t scr0, uint64_t scr1) {
  const uint64_t idx = (uint64_t)(I * h + i);
  f3 wo;
  float pdf;
  const bool ok = bsdf_dir<FAST>(fr, vdc(idx, scr0), sobol(idx, scr1), &wo, &pdf);
  return bsdf_hit<FAST>(L, c, have_sph, sph, ok, wo, pdf);
}

#ifndef VG_SHADE_MIN_BLOCKS
#define VG_SHADE_MIN_BLOCKS 4
#endif
// H1: every light takes at most one sample per strategy (NumSamples <= 2, the `Samples 1` default of the light nodes, and
// always at level > 0): pass 2 then reuses the pass-1 records and the second inlined copy of the sampling code disappears
// (the kernel is instruction-cache bound: ncu shows 20 % of its stall samples in no_instructions).
template <bool FAST, bool H1>
__global__ void __launch_bounds__(128, VG_SHADE_MIN_BLOCKS) k_shade(const RenderParams p, int level, int qin, int qout, int iter_base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = p.counts[qin];
  // whole warps past the end leave together (appends below are warp-collective)
  if ((i & ~31) >= n) return;
  bool active = i < n;
  int path = 0;
  DevHit h;
  h.prim = -1;
  f3 Ro = mk3(0, 0, 0), Rd = mk3(0, 0, 1);
  if (active) {
    path = p.pathq[qin][i];
    const float4 h0 = *reinterpret_cast<const float4*>(&p.hits[i]);
    const int4 h1 = *(reinterpret_cast<const int4*>(&p.hits[i]) + 1);
    h.t = h0.x; h.u = h0.y; h.v = h0.z; h.w = h0.w;
    h.prim = h1.x; h.geom = h1.y; h.slot = h1.z;
    const float4* rp = reinterpret_cast<const float4*>(p.rayq[qin] + i);
    const float4 a = rp[0], b = rp[1];
    Ro = mk3(a.x, a.y, a.z);
    Rd = mk3(a.w, b.x, b.y);
  }
  int matid = 255;
  if (active && h.prim >= 0) matid = p.sc.prim_material[p.sc.geoms[h.geom].prim_base + h.prim];
  if (active) {
    p.v_mat[i] = (uint8_t)matid;
    // contribution slots and 1/total are written exactly once below, for every (light, sample) of a shaded vertex; k_resolve
    // reads none of them for an unshaded one (matid 255 or level > 3)
    p.T[(size_t)level * p.P + path] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // miss, or hit without a shader (core/trace.go:63-65): nothing to shade. At Level > 3 Eval returns at once (std.go:95).
  active = active && matid != 255 && level <= 3;

  DevMat m;
  ShadeCtx c;
  Frame fr;
  f3 omegaI = mk3(0, 0, 1);
  OrenVertex ov;
  Hero hero;
  float lambda = 0, time = 0;
  long long I = 0;
  uint64_t scr0 = 0, scr1 = 0;
  if (active) {
    m = *((p.vmats && p.mat_tex[matid].mask) ? p.vmats + i : p.mats + matid);  // one load site: only the fields used are fetched
    if (m.bad) {
      atomicOr(p.counts + 5, m.bad);
      active = false;
    }
  }
  if (active) {
    lambda = p.lambda[path];
    time = p.time[path];
    int own, it;
    path_decode(p, path, own, it);
    I = (long long)(iter_base + it + 1);  // sample index I = ray.I = the 1-based iteration (render.go:123)
    build_context<true, FAST>(p, h, time, c);
    // tangent frame, std.go:98-106
    f3 V = cross3(c.N, c.DdPdu);
    if (len2_3(V) < 0.1f) V = cross3(c.N, c.DdPdv);
    V = normalize3t<FAST>(V);
    fr.U = normalize3t<FAST>(cross3(c.N, V));
    fr.V = V;
    fr.N = c.N;
    omegaI = basis_project(fr.U, fr.V, fr.N, neg3(Rd));
    hero = hero_setup(lambda);
    ov = oren_vertex<FAST>(omegaI, m.rough2, hero, p.qmc->white_spec);
    const size_t srow = p.scr_by_pixel ? (size_t)p.pix[own] : (size_t)own;
    scr0 = p.scr[srow * 6 + 4];
    scr1 = p.scr[srow * 6 + 5];
  }

  // ---- diffuse lobe: direct light with MIS (std.go:145-163, core/shader.go:203-402) ----
  const bool diffuse = active && m.diff_weight > 0.0f;
  // H1: sample index I*1+0 for both strategies and every light (shader.go:212-214, triangle.go:295-297): one pair of QMC
  // numbers and one BSDF direction per vertex
  double h_r0 = 0, h_r1 = 0;
  f3 h_wo = mk3(0, 0, 1);
  float h_pdf = 0;
  bool h_ok = false;
  if (H1 && diffuse) {
    h_r0 = vdc((uint64_t)I, scr0);
    h_r1 = sobol((uint64_t)I, scr1);
    if (level == 0) h_ok = bsdf_dir<FAST>(fr, h_r0, h_r1, &h_wo, &h_pdf);  // BSDF samples are taken at level 0 only (NS > 1)
  }
  for (int l = 0; l < p.nlights; l++) {
    const DevLight L = p.lights[l];
    const int NS = level > 0 ? 1 : L.nsamples;  // shader.go:186-191
    const bool lit = diffuse && L.geom != h.geom;  // scene.go:106-113
    Spec4 Liu;
    SphTri sph;
    bool by_area = false;
    if (lit) {
      Liu = spec_from_table(p.lights[l].spec, hero);
      by_area = dot3(c.Ng, sub3(L.p0, c.P)) < 0 || dot3(c.Ng, sub3(L.p1, c.P)) < 0 || dot3(c.Ng, sub3(L.p2, c.P)) < 0;
      if (!by_area) sph = spherical_setup<FAST>(L.p0, L.p1, L.p2, c.P);
    }
    const int hN = H1 ? 1 : (NS > 1 ? NS / 2 : NS);  // samples per strategy
    // pass 1: which strategies produced at least one sample (shader.go:241-249)
    int nB = 0, nLs = 0;
    LightRec lr0;
    BsdfRec br0;
    lr0.valid = false;
    br0.valid = false;
    if (lit) {
      if (NS > 1) {
        for (int s = 0; s < hN; s++) {
          const BsdfRec br = H1 ? bsdf_hit<FAST>(L, c, !by_area, sph, h_ok, h_wo, h_pdf) : bsdf_sample<FAST>(L, c, fr, !by_area, sph, I, hN, s, scr0, scr1);
          if (s == 0) br0 = br;
          if (br.valid) nB = hN;
        }
      }
      for (int s = 0; s < hN; s++) {
        const LightRec lr = H1 ? light_sample_r<FAST>(L, c, by_area, sph, h_r0, h_r1) : light_sample<FAST>(L, c, by_area, sph, I, hN, s, scr0, scr1);
        if (s == 0) lr0 = lr;
        if (lr.valid) nLs = hN;
      }
    }
    const int total = nB + nLs;
    if (active) p.v_invtot[(size_t)i * p.nlights + l] = !lit ? 0.0f : (NS > 1 ? (total > 0 ? 1.0f / (float)total : 0.0f) : 1.0f);

    // pass 2: contributions + shadow rays, light samples first then BSDF samples (shader.go:260-344)
    for (int s = 0; s < (NS > 1 ? 2 * hN : hN); s++) {
      const bool is_bsdf = s >= hN;
      bool want = false;
      f3 Ld = mk3(0, 0, 1);
      float Ldist = 0;
      float4 rgb4 = make_float4(0, 0, 0, 0);
      if (lit && (NS == 1 || total > 0)) {
        float p_hat;
        bool valid;
        if (!is_bsdf) {
          const LightRec lr = (H1 || s == 0) ? lr0 : light_sample<FAST>(L, c, by_area, sph, I, hN, s, scr0, scr1);
          valid = lr.valid;
          Ld = lr.Ld;
          Ldist = lr.Ldist;
          if (NS > 1) {
            p_hat = divt<FAST>((float)nB * oren_pdf32<FAST>(fr, Ld), (float)total);
            p_hat += divt<FAST>((float)nLs * lr.pdf, (float)total);
          } else {
            p_hat = lr.pdf;
          }
        } else {
          const BsdfRec br = (H1 || s == hN) ? br0 : bsdf_sample<FAST>(L, c, fr, !by_area, sph, I, hN, s - hN, scr0, scr1);
          valid = br.valid;
          Ld = br.Ld;
          Ldist = br.Ldist;
          p_hat = divt<FAST>((float)nB * br.pdf, (float)total);
          p_hat += divt<FAST>((float)nLs * br.pdfLight, (float)total);
        }
        if (valid && !(dot3(Ld, c.N) <= 0)) {
          Spec4 rho = oren_eval<FAST>(fr, ov, Ld);
          const float inv = divt<FAST>(1.0f, p_hat);
#pragma unroll
          for (int k = 0; k < 4; k++) rho.c[k] = (rho.c[k] * Liu.c[k]) * inv;
          f3 rgb = spec_to_rgb(rho, hero);
          if (NS > 1) {  // shader.go:292-296 clamps only in the MIS branch
            if (rgb.x < 0) rgb.x = 0;
            if (rgb.y < 0) rgb.y = 0;
            if (rgb.z < 0) rgb.z = 0;
          }
          rgb4 = make_float4(rgb.x, rgb.y, rgb.z, 0.f);
          want = true;
        }
      }
      const int slot = i * p.S + L.slot_base + s;
      if (active) p.contrib[slot] = rgb4;  // zero unless `want`
      const int qi = warp_append(p.counts + 2, want);
      if (want) {
        const f3 o = offset_p(c.P, c.Poffset, dot3(Ld, c.Ng) < 0 ? -1 : 1);
        const f3 d = scale3(Ldist * (1.0f - 0.0001f), Ld);  // core/ray.go:20, shader.go:269-273
        float4* rp = reinterpret_cast<float4*>(p.sray + qi);
        rp[0] = make_float4(o.x, o.y, o.z, d.x);
        rp[1] = make_float4(d.y, d.z, 1.0f, time);
        p.sslot[qi] = slot;
      }
    }
  }

  // ---- mirror lobe (std.go:194-261, bsdf/specular.go) ----
  {
    bool want = false;
    f3 wo = mk3(0, 0, 1);
    float4 T4 = make_float4(0, 0, 0, 0);
    if (active && m.spec_weight > 0.0f) {
      const f3 refl = reflect_z(omegaI);
      wo = basis_expand(fr.U, fr.V, fr.N, normalize3(refl));
      const f3 o = basis_project(fr.U, fr.V, fr.N, wo);
      const double pdf = dot3(o, refl) < 0.9999f ? 0.0 : 1.0;
      if (!(dot3(wo, c.Ng) <= 0.0f)) {
        Spec4 rho;
        rho.c[0] = rho.c[1] = rho.c[2] = rho.c[3] = 0.f;
        if (!(dot3(o, refl) < 0.9999f)) {
          const float kr = dielectric_kr(m.ior, omegaI.z);
          rho = spec_from_rgb(mk3(kr, kr, kr), hero);
          const float az = fabsf(o.z);
#pragma unroll
          for (int k = 0; k < 4; k++) rho.c[k] *= az;
        }
        const float inv = 1.0f / (float)pdf;
#pragma unroll
        for (int k = 0; k < 4; k++) rho.c[k] *= inv;
        const f3 rgb = spec_to_rgb(rho, hero);
        T4 = make_float4(rgb.x * m.spec_colour.x, rgb.y * m.spec_colour.y, rgb.z * m.spec_colour.z, m.spec_weight);
        // the level-4 ray is traced by the reference although its shader returns black (std.go:95,243)
        want = (level + 1 <= 3) || p.trace_last_level;
      }
    }
    if (active) p.T[(size_t)level * p.P + path] = T4;
    const int qi = warp_append(p.counts + qout, want);
    if (want) {
      const f3 o = offset_p(c.P, c.Poffset, 1);
      float4* rp = reinterpret_cast<float4*>(p.rayq[qout] + qi);
      rp[0] = make_float4(o.x, o.y, o.z, wo.x);
      rp[1] = make_float4(wo.y, wo.z, __int_as_float(0x7f800000), time);
      p.pathq[qout][qi] = path;
    }
  }
}

}  // namespace vg
#include "shade_generic.cuh"
namespace vg {

// Sum the slots in the reference's order and finish the diffuse / glossy terms (core/shader.go:349, std.go:157-163,269-295).
// WIDE4 (one lobe, four slots, two lights per vertex — the default light nodes of every BASELINE scene): the vertex's 64 bytes of
// slots come in as two 256-bit loads and its two 1/total factors as one 64-bit load, and the loops pick from registers. The same
// additions in the same order; what changes is the number of L1TEX wavefronts a thread-per-pixel warp spends per vertex (7 -> 4), the
// unit that bounds k_resolve_accumulate (ncu: 92 % busy).
__device__ __forceinline__ float4 pick4(const float4& a, const float4& b, const float4& c, const float4& d, int k) {
  return k == 0 ? a : (k == 1 ? b : (k == 2 ? c : d));
}
template <bool WIDE4 = false>
__device__ __forceinline__ float4 resolve_vertex(const RenderParams& p, int level, int i) {
  const int matid = p.v_mat[i];
  const int SL = p.S * p.nlobes;
  float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
  if (matid != 255 && level <= 3) {
    const DevMat m = *((p.vmats && p.mat_tex[matid].mask) ? p.vmats + i : p.mats + matid);  // one load site: only the fields used are fetched
    f3 sum[2];
    sum[0] = sum[1] = mk3(0, 0, 0);
    float4 w0, w1, w2, w3;
    float2 inv2 = make_float2(0.f, 0.f);
    if (WIDE4) {
      const float4* cp = p.contrib + (size_t)i * 4;
      asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(w0.x), "=f"(w0.y), "=f"(w0.z), "=f"(w0.w), "=f"(w1.x), "=f"(w1.y), "=f"(w1.z), "=f"(w1.w) : "l"(cp));
      asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(w2.x), "=f"(w2.y), "=f"(w2.z), "=f"(w2.w), "=f"(w3.x), "=f"(w3.y), "=f"(w3.z), "=f"(w3.w) : "l"(cp + 2));
      inv2 = *reinterpret_cast<const float2*>(p.v_invtot + (size_t)i * 2);
    }
    for (int lobe = 0; lobe < (WIDE4 ? 1 : p.nlobes); lobe++) {
      const bool on = lobe == 0 ? m.diff_weight > 0.0f : (m.spec_weight > 0.0f && m.spec_rough > 0.0f);
      if (!on) continue;
      const f3 colour = lobe == 0 ? m.diff_colour : m.spec_colour;
      f3 acc = mk3(0, 0, 0);
      for (int l = 0; l < (WIDE4 ? 2 : p.nlights); l++) {
        const DevLight& L = p.lights[l];
        const int NS = level > 0 ? 1 : L.nsamples;
        const float inv = WIDE4 ? (l == 0 ? inv2.x : inv2.y) : p.v_invtot[((size_t)i * p.nlobes + lobe) * p.nlights + l];
        if (inv == 0.0f) continue;  // light excluded (own geom) or no samples: EvaluateLightSamples returned RGB{}
        f3 col = mk3(0, 0, 0);
        for (int s = 0; s < NS; s++) {
          const float4 c = WIDE4 ? pick4(w0, w1, w2, w3, L.slot_base + s) : p.contrib[(size_t)i * SL + lobe * p.S + L.slot_base + s];
          col.x += c.x; col.y += c.y; col.z += c.z;
        }
        if (NS > 1) { col.x *= inv; col.y *= inv; col.z *= inv; }
        col.x *= colour.x; col.y *= colour.y; col.z *= colour.z;
        acc.x += col.x; acc.y += col.y; acc.z += col.z;
      }
      // the diffuse sum is scaled by diffWeight (std.go:162); the glossy direct light is NOT scaled by spec1Weight
      // (spec1Samples == 0 skips std.go:265-267)
      if (lobe == 0) { acc.x *= m.diff_weight; acc.y *= m.diff_weight; acc.z *= m.diff_weight; }
      sum[lobe] = acc;
    }
    // contrib = emission + diffuse + spec1 (std.go:287-293)
    out = make_float4((m.emission.x + sum[0].x) + sum[1].x, (m.emission.y + sum[0].y) + sum[1].y, (m.emission.z + sum[0].z) + sum[1].z, 0.f);
  }
  return out;
}
__global__ void __launch_bounds__(256) k_resolve(const RenderParams p, int level, int qin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.counts[qin]) return;
  p.L[(size_t)level * p.P + p.pathq[qin][i]] = resolve_vertex(p, level, i);
}

// Level 4 of a mirror chain (scenes that hold a DebugShader next to a mirror lobe): ShaderStd.Eval returns at once there
// (std.go:95) but Debug.Eval still sets OutRGB = Colour (debug.go:42-44), which the Level-3 mirror lobe then picks up.
__global__ void __launch_bounds__(256) k_debug_last(const RenderParams p, int level, int qin) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.counts[qin]) return;
  const DevHit h = p.hits[i];
  if (h.prim < 0) return;
  const int matid = p.sc.prim_material[p.sc.geoms[h.geom].prim_base + h.prim];
  if (matid == 255) return;
  const DevMat& m = p.mats[matid];
  if (m.debug) p.L[(size_t)level * p.P + p.pathq[qin][i]] = make_float4(m.emission.x, m.emission.y, m.emission.z, 0.f);
}

// Scenes without a mirror lobe have one level: the level-0 queue is the identity (queue slot == path), so the per-vertex sum and
// the running mean (render.go:127-129) are one kernel and L never goes to memory.
template <bool WIDE4>
__global__ void __launch_bounds__(256) k_resolve_accumulate(const RenderParams p, int iter_base, int niters) {
  const int own = blockIdx.x * blockDim.x + threadIdx.x;
  if (own >= p.nown) return;
  float* px = p.fb + (size_t)p.pix[own] * 3;
  float r = px[0], g = px[1], b = px[2];
  for (int it = 0; it < niters; it++) {
    const float4 C = resolve_vertex<WIDE4>(p, 0, path_index(p, own, it));
    const float fi = (float)(iter_base + it + 1);
    r = (r * fi + C.x) / (fi + 1.0f);
    g = (g * fi + C.y) / (fi + 1.0f);
    b = (b * fi + C.z) / (fi + 1.0f);
  }
  px[0] = r; px[1] = g; px[2] = b;
}

// The same for the default path order (a warp's 32 paths = ONE pixel x 32 iterations): a thread per pixel would stream through 32 x
// 64 B of contribution slots on its own, every 16-B load of the warp touching 32 different lines (ncu: L1TEX 89 % busy, the kernel's
// bound). Here a warp takes 32 pixels: for each of them all 32 lanes resolve the pixel's 32 iterations with coalesced loads
// (lane = iteration) into a padded shared-memory tile, then every lane runs the sequential running mean of ITS pixel out of the
// tile (lane = pixel). Same arithmetic in the same order per pixel: bit-identical (tested).
// MEASURED (round 2) and left OFF (option "accumulate_tiled"): raygen + accumulate 5.76 ms per C2 frame against 4.74 ms for the plain
// kernel — the 32 serial rounds per warp and 16 resident warps/SM (25 KB of tile per 64 threads) cost more than the coalescing gains;
// the plain kernel's per-thread streams are served well enough by L1.
__global__ void __launch_bounds__(64) k_resolve_accumulate_t(const RenderParams p, int iter_base, int niters) {
  __shared__ float tile[2][3][32 * 33];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int w = blockIdx.x * 2 + wib;  // this warp's group of 32 owned pixels
  const int own0 = w * 32;
  if (own0 >= p.nown) return;
  const int own = own0 + lane;
  float r = 0, g = 0, b = 0;
  float* px = nullptr;
  if (own < p.nown) {
    px = p.fb + (size_t)p.pix[own] * 3;
    r = px[0]; g = px[1]; b = px[2];
  }
  float(*t)[32 * 33] = tile[wib];
  for (int itg = 0; itg < (niters >> 5); itg++) {
    const int npx = min(32, p.nown - own0);
    for (int k = 0; k < npx; k++) {
      const float4 C = resolve_vertex(p, 0, ((itg * p.nown + own0 + k) << 5) + lane);  // path_index() for B = 1, G = 32
      t[0][k * 33 + lane] = C.x; t[1][k * 33 + lane] = C.y; t[2][k * 33 + lane] = C.z;
    }
    __syncwarp();
    if (own < p.nown) {
      for (int it = 0; it < 32; it++) {
        const float fi = (float)(iter_base + (itg << 5) + it + 1);
        r = (r * fi + t[0][lane * 33 + it]) / (fi + 1.0f);
        g = (g * fi + t[1][lane * 33 + it]) / (fi + 1.0f);
        b = (b * fi + t[2][lane * 33 + it]) / (fi + 1.0f);
      }
    }
    __syncwarp();
  }
  if (own < p.nown) { px[0] = r; px[1] = g; px[2] = b; }
}

// Fold the levels back to front (std.go:246-266,287-295) and continue the running mean (render.go:127-129).
__global__ void __launch_bounds__(256) k_accumulate(const RenderParams p, int iter_base, int niters) {
  const int own = blockIdx.x * blockDim.x + threadIdx.x;
  if (own >= p.nown) return;
  float* px = p.fb + (size_t)p.pix[own] * 3;
  float r = px[0], g = px[1], b = px[2];
  for (int it = 0; it < niters; it++) {
    const int path = path_index(p, own, it);
    // The path's last level: the first one whose mirror lobe did not continue it (k_shade left T = 0 there). Levels beyond it
    // hold values of earlier batches and are neither read nor needed: the fold below would multiply them by that T = 0.
    float4 Tk[5];
    int last = p.levels - 1;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      if (k < p.levels && k <= last) {
        Tk[k] = p.T[(size_t)k * p.P + path];
        if (Tk[k].x == 0.f && Tk[k].y == 0.f && Tk[k].z == 0.f && Tk[k].w == 0.f) last = k;
      }
    }
    f3 C = mk3(0, 0, 0);
#pragma unroll
    for (int k = 4; k >= 0; k--) {
      if (k > last) continue;
      const float4 L = p.L[(size_t)k * p.P + path];
      const float4 T = Tk[k];
      f3 sp = mk3(T.x * C.x, T.y * C.y, T.z * C.z);
      if (sp.x < 0 || isnan(sp.x)) sp.x = 0;
      if (sp.y < 0 || isnan(sp.y)) sp.y = 0;
      if (sp.z < 0 || isnan(sp.z)) sp.z = 0;
      sp = scale3(T.w, sp);
      C = mk3(L.x + sp.x, L.y + sp.y, L.z + sp.z);
    }
    const float fi = (float)(iter_base + it + 1);
    r = (r * fi + C.x) / (fi + 1.0f);
    g = (g * fi + C.y) / (fi + 1.0f);
    b = (b * fi + C.z) / (fi + 1.0f);
  }
  px[0] = r; px[1] = g; px[2] = b;
}

// ---------------------------------------------------------------------------------------------------------
struct RenderState {
  bool ready = false;
  int nown = 0, P = 0, S = 0, levels = 1, nlights = 0, iters = 0;
  DevBuf<int> pix;
  DevBuf<uint64_t> scr;
  DevBuf<DevMat> mats;
  DevBuf<DevMat> vmats;          // textured scenes only
  DevBuf<VgMaterial> rawmats;
  DevBuf<MatTex> mat_tex;
  DevBuf<float> diff;
  bool textured = false;
  DevBuf<DevLight> lights;
  DevBuf<VgRay> rayq0, rayq1, sray;
  DevBuf<int> pathq0, pathq1, sslot, counts;
  DevBuf<DevHit> hits;
  DevBuf<float> lambda, time, invtot, fb;
  DevBuf<double> filter;
  DevBuf<QmcTables> qmc;
  DevBuf<uint8_t> vmat;
  DevBuf<float4> contrib, L, T;
  DevBuf<unsigned long long> stats;
  int fb_w = 0, fb_h = 0;
  int trace_grid = 0;
  int coop_grid = 0, coop_grid_mot = 0;    // closest-hit kernels with cooperative leaves (VG_TRACE_MIN_BLOCKS_COOP)
  int shadow_grid = 0, shadow_grid_mot = 0;  // the any-hit kernels may be compiled for another residency (VG_TRACE_MIN_BLOCKS_SHADOW)
  int max_light_samples = 0;
  bool scr_by_pixel = false;  // the device holds the caller's whole scramble table in raster order (pinned fast path)
  // which rows rs.scr holds: valid only for this (frame, partition, pixel order); scr_valid is cleared when rs.scr is dropped
  bool scr_valid = false;
  int scr_w = 0, scr_h = 0, scr_rank = 0, scr_world = 0, scr_pixel_block = 0;
  // level-0 shadow queue: per-lane loop (1) or cooperative kernel (0), whichever the first calls measured faster on this scene
  // (bit-identical results: occlusion does not depend on the order). While undecided the batches alternate; two samples of each
  // kernel decide (minimum per iteration). Reset when the render state is rebuilt.
  int l0_choice = -1, l0_trials[2] = {0, 0};
  unsigned l0_seq = 0;
  float l0_ms_per_iter[2] = {0.f, 0.f};
  bool generic = false;  // shade with k_shade_generic (glossy lobe / conductor Fresnel / Disk or Sphere lights / sphere geoms)
  int nlobes = 1;
  std::vector<int> pix_host;
  std::vector<int> row_start;  // [tilesY + 1] index into pix of the first owned pixel of every tile row (slices of vg_render_frame)
  std::vector<cudaEvent_t> pipe_ev;  // untimed events of the frame pipeline
  cudaEvent_t pev(size_t i) {
    while (pipe_ev.size() <= i) {
      cudaEvent_t e;
      cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      pipe_ev.push_back(e);
    }
    return pipe_ev[i];
  }
  uint64_t* scr_pinned = nullptr;
  size_t scr_pinned_bytes = 0;
  float* fb_pinned = nullptr;
  size_t fb_pinned_bytes = 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  // Events around the traversal and shading launches (per-stage device time for the roofline). A fixed pool: a call that
  // launches more than kMaxTimed stages leaves the later ones untimed (VgStats says how many were timed).
  static const size_t kMaxTimedEvents = 4096;
  std::vector<cudaEvent_t> evpool;
  cudaEvent_t ev(size_t i) {
    while (evpool.size() <= i) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      evpool.push_back(e);
    }
    return evpool[i];
  }
  std::vector<cudaEvent_t> fetchpool;
  cudaEvent_t fetch_ev(int i) {
    while ((int)fetchpool.size() <= i) {
      cudaEvent_t e;
      cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      fetchpool.push_back(e);
    }
    return fetchpool[(size_t)i];
  }
  void release() {
    pix.release(); scr.release(); mats.release(); vmats.release(); rawmats.release(); mat_tex.release(); diff.release(); lights.release(); rayq0.release(); rayq1.release(); sray.release();
    pathq0.release(); pathq1.release(); sslot.release(); counts.release(); hits.release(); lambda.release(); time.release();
    invtot.release(); vmat.release(); filter.release(); qmc.release(); contrib.release(); L.release(); T.release(); stats.release();
  }
};

void render_invalidate(vg_ctx* ctx) {
  if (!ctx->rs) return;
  ctx->rs->ready = false;
  ctx->rs->l0_choice = -1;
  ctx->rs->l0_trials[0] = ctx->rs->l0_trials[1] = 0;
  ctx->rs->l0_seq = 0;
}
void render_destroy(vg_ctx* ctx) {
  if (!ctx->rs) return;
  ctx->rs->release();
  ctx->rs->fb.release();
  if (ctx->rs->e0) cudaEventDestroy(ctx->rs->e0);
  if (ctx->rs->e1) cudaEventDestroy(ctx->rs->e1);
  for (cudaEvent_t e : ctx->rs->evpool) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->rs->fetchpool) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->rs->pipe_ev) cudaEventDestroy(e);
  if (ctx->rs->scr_pinned) cudaFreeHost(ctx->rs->scr_pinned);
  if (ctx->rs->fb_pinned) cudaFreeHost(ctx->rs->fb_pinned);
  delete ctx->rs;
  ctx->rs = nullptr;
}

#define RCUDA(call)                                          \
  do {                                                       \
    cudaError_t e_ = (call);                                 \
    if (e_ != cudaSuccess) return ctx->cuda_fail(e_, #call); \
  } while (0)

static inline f3 h3(const float* p) { f3 r; r.x = p[0]; r.y = p[1]; r.z = p[2]; return r; }

// host-side normalize with the reference's operation order (math/vec3_amd64.s:11-43); the reciprocal square root is
// correctly rounded here.
static f3 host_normalize(f3 a) {
  float x0 = a.x * a.x, x1 = a.y * a.y, x2 = a.z * a.z;
  x1 = x1 + x0;
  x1 = x1 + x2;
  const float r = (float)(1.0 / std::sqrt((double)x1));
  f3 o; o.x = a.x * r; o.y = a.y * r; o.z = a.z * r;
  return o;
}

static int ensure_fb(vg_ctx* ctx) {
  if (!ctx->rs) {
    ctx->rs = new RenderState();
    cudaEventCreate(&ctx->rs->e0);
    cudaEventCreate(&ctx->rs->e1);
  }
  RenderState& rs = *ctx->rs;
  if (ctx->xres <= 0 || ctx->yres <= 0) return ctx->fail(VG_ERR_INVALID, "frame size not set (vg_set_frame)");
  if (rs.fb_w != ctx->xres || rs.fb_h != ctx->yres) {
    RCUDA(rs.fb.reserve((size_t)ctx->xres * ctx->yres * 3));
    RCUDA(cudaMemsetAsync(rs.fb.p, 0, (size_t)ctx->xres * ctx->yres * 3 * sizeof(float), ctx->stream));
    rs.fb_w = ctx->xres;
    rs.fb_h = ctx->yres;
  }
  return VG_OK;
}

// Partition of framescramble (core/render.go:166-176) at upload: only the rows of owned pixels go to the device.
// rows of the caller's (page-locked, device-visible) scramble table -> path order; thread = one 16-byte third of a row
__global__ void k_gather_scramble(const uint4* __restrict__ table, const int* __restrict__ pix, int nown, uint4* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)nown * 3) return;
  const int own = (int)(i / 3), part = (int)(i % 3);
  out[i] = table[(size_t)pix[own] * 3 + part];
}

// true if `p` is page-locked host memory the GPU can DMA from/to directly (cudaHostAlloc / cudaHostRegister / torch pin_memory)
static bool is_pinned_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

static int upload_scramble_rows(vg_ctx* ctx, const uint64_t* table);
static int upload_scramble(vg_ctx* ctx, const uint64_t* table) {
  RenderState& rs = *ctx->rs;
  rs.scr_valid = false;
  const int rc = upload_scramble_rows(ctx, table);
  if (rc == VG_OK) {
    rs.scr_valid = true;
    rs.scr_w = ctx->xres; rs.scr_h = ctx->yres; rs.scr_rank = ctx->rank; rs.scr_world = ctx->world; rs.scr_pixel_block = ctx->opt_pixel_block;
  }
  return rc;
}
static int upload_scramble_rows(vg_ctx* ctx, const uint64_t* table) {
  RenderState& rs = *ctx->rs;
  if (rs.nown == 0) return VG_OK;
  rs.scr_by_pixel = false;
  if (ctx->world == 1 && is_pinned_host(table)) {
    // The caller's table is page-locked and this context owns every pixel: one DMA of the table as it is, no host pass;
    // the kernels index it by raster pixel instead of by path order.
    const size_t all = (size_t)ctx->xres * ctx->yres * 6;
    RCUDA(rs.scr.reserve(all));
    RCUDA(cudaMemcpyAsync(rs.scr.p, table, all * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    RCUDA(cudaStreamSynchronize(ctx->stream));
    rs.scr_by_pixel = true;
    return VG_OK;
  }
  if (is_pinned_host(table)) {
    // Page-locked table, this context owns a subset of the pixels: a kernel gathers the owned rows straight from host memory
    // over PCIe (zero-copy read of 48-B rows, runs of 8 or 32 rows contiguous), no host pass and no staging copy.
    void* dview = nullptr;
    if (cudaHostGetDevicePointer(&dview, const_cast<uint64_t*>(table), 0) == cudaSuccess && dview) {
      RCUDA(rs.scr.reserve((size_t)rs.nown * 6));
      const long long n2 = (long long)rs.nown * 3;  // one thread per 16 bytes
      k_gather_scramble<<<(unsigned)((n2 + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(dview), rs.pix.p, rs.nown,
                                                                                reinterpret_cast<uint4*>(rs.scr.p));
      RCUDA(cudaGetLastError());
      RCUDA(cudaStreamSynchronize(ctx->stream));
      return VG_OK;
    }
    cudaGetLastError();
  }
  // a context that owns every pixel keeps the table as it is (raster order, kernels index it by pixel): no gather, and the
  // device copy can restore the host copy when the pixel order changes later (prepare())
  const bool whole = ctx->world == 1;
  const size_t n = whole ? (size_t)ctx->xres * ctx->yres : (size_t)rs.nown;
  const size_t bytes = n * 48;
  if (whole) RCUDA(rs.scr.reserve(n * 6));
  if (rs.scr_pinned_bytes < bytes) {
    if (rs.scr_pinned) cudaFreeHost(rs.scr_pinned);
    rs.scr_pinned = nullptr;
    RCUDA(cudaMallocHost((void**)&rs.scr_pinned, bytes));
    rs.scr_pinned_bytes = bytes;
  }
  const int* pix = rs.pix_host.data();
  // Owned pixels come in runs of 8 (or 32) consecutive pixels of one raster row. A few host threads gather their slice of the
  // rows run by run into pinned memory, chunk by chunk, and each issues the H2D copy of a chunk as soon as it is gathered, so
  // the PCIe transfer overlaps the gathering of the next chunks (one core's memcpy, ~10 GB/s, is slower than the link).
  const int nthreads = n >= (1u << 16) ? 8 : 1;
  const size_t chunk = 1u << 15;  // pixels per copy: 1.5 MB
  uint64_t* dst_dev = rs.scr.p;
  cudaStream_t stream = ctx->stream;
  const int device = ctx->device;
  std::vector<cudaError_t> errs((size_t)nthreads, cudaSuccess);
  auto work = [&](int t, size_t b, size_t e) {
    if (t > 0) cudaSetDevice(device);
    for (size_t c0 = b; c0 < e; c0 += chunk) {
      const size_t c1 = std::min(e, c0 + chunk);
      size_t i = c0;
      if (whole) {
        std::memcpy(rs.scr_pinned + c0 * 6, table + c0 * 6, (c1 - c0) * 48);
        i = c1;
      }
      while (i < c1) {
        size_t j = i + 1;
        while (j < c1 && pix[j] == pix[j - 1] + 1) j++;
        std::memcpy(rs.scr_pinned + i * 6, table + (size_t)pix[i] * 6, (j - i) * 48);
        i = j;
      }
      const cudaError_t e2 = cudaMemcpyAsync(dst_dev + c0 * 6, rs.scr_pinned + c0 * 6, (c1 - c0) * 48, cudaMemcpyHostToDevice, stream);
      if (e2 != cudaSuccess) errs[(size_t)t] = e2;
    }
  };
  {
    std::vector<std::thread> th;
    const size_t per = (n + nthreads - 1) / nthreads;
    for (int t = 1; t < nthreads; t++)
      if ((size_t)t * per < n) th.emplace_back(work, t, (size_t)t * per, std::min(n, (size_t)(t + 1) * per));
    work(0, 0, std::min(n, per));
    for (auto& t : th) t.join();
  }
  for (cudaError_t e2 : errs) RCUDA(e2);
  RCUDA(cudaStreamSynchronize(ctx->stream));
  rs.scr_by_pixel = whole;
  return VG_OK;
}

static int prepare(vg_ctx* ctx) {
  int rc = ensure_fb(ctx);
  if (rc != VG_OK) return rc;
  RenderState& rs = *ctx->rs;
  if (rs.ready) return VG_OK;
  if (!ctx->committed) return ctx->fail(VG_ERR_INVALID, "scene not committed");
  if (!ctx->have_camera) return ctx->fail(VG_ERR_INVALID, "no camera (core.Render: ErrNoCamera)");
  const int W = ctx->xres, H = ctx->yres;
  if ((int64_t)ctx->scramble.size() != (int64_t)W * H * 6) return ctx->fail(VG_ERR_INVALID, "scramble table missing or of the wrong size (vg_set_scramble)");

  // owned pixels: the reference's 32x32 tiles (render.go:196-199) dealt round-robin over the ranks (comm.cu: owned_pixels)
  std::vector<int> pix;
  owned_pixels(W, H, ctx->rank, ctx->world, ctx->opt_pixel_block != 0, pix);
  rs.nown = (int)pix.size();
  rs.pix_host = pix;
  {
    const int tilesY = (H + 31) / 32;
    rs.row_start.assign((size_t)tilesY + 1, (int)pix.size());
    for (int i = (int)pix.size() - 1; i >= 0; i--) rs.row_start[(size_t)((pix[(size_t)i] / W) / 32)] = i;
    for (int ty = tilesY - 1; ty >= 0; ty--) rs.row_start[(size_t)ty] = std::min(rs.row_start[(size_t)ty], rs.row_start[(size_t)ty + 1]);
  }

  // materials
  std::vector<DevMat> mats(ctx->materials.size());
  bool any_mirror = false, any_glossy = false, any_conductor = false, any_other_light = false, any_debug = false;
  bool textured = false;
  for (size_t i = 0; i < mats.size(); i++) {
    const VgMaterial& s = ctx->materials[i];
    DevMat& d = mats[i];
    d = derive_mat(s);
    if (d.debug) {
      any_debug = true;
      continue;
    }
    const MatTex* mt = i < ctx->mat_tex.size() && ctx->mat_tex[i].any() ? &ctx->mat_tex[i] : nullptr;
    textured |= mt != nullptr;
    // a textured strength / roughness can take any value at a vertex: keep every kernel path it may reach
    const bool spec_maybe = d.spec_weight > 0.0f || (mt && (mt->slot[6].tex >= 0 || mt->slot[3].tex >= 0));
    if (mt && mt->slot[6].tex >= 0 && d.bad) d.bad = 0;
    if (spec_maybe) {
      const bool rough_tex = mt && mt->slot[7].tex >= 0;
      if (d.spec_rough == 0.0f || rough_tex) any_mirror = true;
      if (d.spec_rough != 0.0f || rough_tex) any_glossy = true;
      if (d.fresnel_model != VG_FRESNEL_DIELECTRIC) any_conductor = true;
    }
  }
  if (textured) {
    for (const MeshStage& m : ctx->meshes)
      if (m.present && (m.motion || m.sphere))
        return ctx->fail(VG_ERR_UNSUPPORTED, "texture maps need a scene of static PolyMeshes and GeomInstances of them (the reference's motion path leaves "
                                             "the texture footprint 0: trace.go:677-684, feline.go:57-61; a Sphere geom keeps whatever footprint the last mesh leaf left)");
    for (const MeshStage& m : ctx->meshes) {
      if (!m.present || !m.instance) continue;
      int t = m.target;
      while (ctx->meshes[(size_t)t].instance) t = ctx->meshes[(size_t)t].target;
      if (ctx->meshes[(size_t)t].motion) return ctx->fail(VG_ERR_UNSUPPORTED, "texture maps: a GeomInstance of a motion mesh (see above)");
    }
    for (const VgLight& l : ctx->lights)
      if (l.material >= 0 && l.material < (int)ctx->mat_tex.size() && (ctx->mat_tex[(size_t)l.material].slot[0].tex >= 0 || ctx->mat_tex[(size_t)l.material].slot[1].tex >= 0))
        return ctx->fail(VG_ERR_UNSUPPORTED, "a texture map on the emission of a light's shader (the light evaluates it with its own lsg)");
  }
  rs.textured = textured;
  // lights
  std::vector<DevLight> lights(ctx->lights.size());
  int S = 0;
  rs.max_light_samples = 0;
  for (size_t i = 0; i < lights.size(); i++) {
    const VgLight& s = ctx->lights[i];
    DevLight& d = lights[i];
    d.type = s.type;
    d.radius = s.radius;
    d.p0 = h3(s.p0); d.p1 = h3(s.p1); d.p2 = h3(s.p2);
    if (s.type == VG_LIGHT_TRI) {
      f3 e1; e1.x = d.p1.x - d.p0.x; e1.y = d.p1.y - d.p0.y; e1.z = d.p1.z - d.p0.z;
      f3 e2; e2.x = d.p2.x - d.p0.x; e2.y = d.p2.y - d.p0.y; e2.z = d.p2.z - d.p0.z;
      f3 cr; cr.x = e1.y * e2.z - e1.z * e2.y; cr.y = e1.z * e2.x - e1.x * e2.z; cr.z = e1.x * e2.y - e1.y * e2.x;
      d.N = host_normalize(cr);
      float x0 = cr.x * cr.x, x1 = cr.y * cr.y, x2 = cr.z * cr.z;
      x1 = x1 + x0; x1 = x1 + x2;
      const float area = 0.5f * std::sqrt(x1);
      d.inv_area = 1 / area;
    } else {
      any_other_light = true;
      d.N = h3(s.n);
      d.inv_area = 1.0f / (3.14159265358f * s.radius * s.radius);  // disk.go:125,179 (float32 expression)
    }
    d.E.x = d.E.y = d.E.z = 0;
    // Debug.EvalEmission returns black (debug.go:49)
    if (s.material >= 0 && s.material < (int)mats.size() && !mats[s.material].debug) d.E = mats[s.material].emission;
    if (s.samples < 0 || s.samples > 8) return ctx->fail(VG_ERR_INVALID, "TriLight.Samples outside [0,8]");
    d.nsamples = 1 << s.samples;
    rs.max_light_samples = std::max(rs.max_light_samples, d.nsamples);
    d.geom = s.geom;
    d.slot_base = S;
    S += d.nsamples;
  }
  if (S == 0) S = 1;
  rs.S = S;
  // sphere geoms need the analytic hit record; glossy lobes, conductor Fresnel and non-Tri lights the general kernel
  bool any_sphere_geom = false;
  for (const MeshStage& m : ctx->meshes) any_sphere_geom |= m.sphere;
  rs.generic = any_glossy || any_conductor || any_other_light || any_sphere_geom || ctx->dev.n_xforms > 0 || ctx->opt_generic_shade;
  rs.nlobes = any_glossy ? 2 : 1;
  rs.nlights = (int)lights.size();
  // a mirror chain's Level-4 ray is shaded by nothing but a DebugShader (ShaderStd.Eval returns at Level > 3, std.go:95; Debug.Eval
  // has no such test): scenes that hold both keep a fifth level for that colour
  rs.levels = any_mirror ? (any_debug ? 5 : 4) : 1;
  rs.iters = ctx->opt_iters_per_batch;
  {
    // The batch depth is a request: queue slots are 32-bit indices (paths x contribution slots must stay below 2^31) and the
    // wavefront state (~230 B + 52 B per contribution slot + 32 B per level and path) has to fit the free device memory.
    const size_t per_path = 230 + (textured ? sizeof(DevMat) + 48 : 0) + (size_t)S * rs.nlobes * 52 + (size_t)rs.levels * 32 + (size_t)std::max(1, (int)lights.size()) * rs.nlobes * 4;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = (size_t)16 << 30;
    // the wavefront buffers of an earlier prepare() are re-used (DevBuf::reserve keeps what is large enough): they count as free
    free_b += rs.rayq0.cap * sizeof(VgRay) * 2 + rs.sray.cap * sizeof(VgRay) + (rs.pathq0.cap * 2 + rs.sslot.cap) * sizeof(int) +
              rs.hits.cap * sizeof(DevHit) + (rs.invtot.cap + rs.diff.cap) * sizeof(float) + rs.vmat.cap +
              (rs.contrib.cap + rs.L.cap + rs.T.cap) * sizeof(float4) + rs.vmats.cap * sizeof(DevMat);
    const size_t by_mem = (free_b / 10 * 8) / per_path / (size_t)std::max(1, rs.nown);
    const size_t by_idx = (((size_t)1 << 31) - 1) / ((size_t)S * rs.nlobes) / (size_t)std::max(1, rs.nown);
    const size_t cap = std::max<size_t>(1, std::min(by_mem, by_idx));
    if ((size_t)rs.iters > cap) rs.iters = (int)cap;
  }
  rs.P = rs.nown * rs.iters;
  const size_t P = (size_t)rs.P;

  {
    // byte-sliced tables of the two XOR enumerations in RasterXY (shade.cuh: raster_xy12_tab)
    static const uint32_t inv12[24] = {0xf0f000, 0x505000, 0x303000, 0x101000, 0xff0000, 0x550000, 0x330000, 0x110000,
                                       0xf0000,  0x50000,  0x30000,  0x10000,  0x888800, 0x444400, 0x222200, 0x111100,
                                       0x800080, 0x400040, 0x200020, 0x100010, 0x80008,  0x40004,  0x20002,  0x10001};
    std::vector<QmcTables> tv(1);
    QmcTables& T = tv[0];
    uint64_t v[56];
    v[0] = 1ull << 51;
    for (int k = 1; k < 56; k++) v[k] = v[k - 1] ^ (v[k - 1] >> 1);
    for (int byte = 0; byte < 7; byte++)
      for (int x = 0; x < 256; x++) {
        uint64_t r = 0;
        for (int j = 0; j < 8; j++)
          if (x & (1 << j)) r ^= v[byte * 8 + j];
        T.sob[byte][x] = r;
      }
    for (int byte = 0; byte < 3; byte++)
      for (int x = 0; x < 256; x++) {
        uint32_t r = 0;
        for (int j = 0; j < 8; j++)
          if (x & (1 << j)) r ^= inv12[byte * 8 + j];
        T.rinv[byte][x] = r;
      }
    RCUDA(rs.qmc.reserve(1));
    RCUDA(cudaMemcpyAsync(rs.qmc.p, tv.data(), sizeof(QmcTables), cudaMemcpyHostToDevice, ctx->stream));
    RCUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (ctx->filter_n > 0) {
    RCUDA(rs.filter.reserve(ctx->filter_cdf.size()));
    RCUDA(cudaMemcpyAsync(rs.filter.p, ctx->filter_cdf.data(), ctx->filter_cdf.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  RCUDA(rs.pix.reserve(pix.size()));
  {
    const size_t need = ctx->world == 1 ? (size_t)W * H * 6 : (size_t)rs.nown * 6;
    if (need > rs.scr.cap) rs.scr_valid = false;  // reserve() reallocates: the rows on the device are gone
    RCUDA(rs.scr.reserve(need));
  }
  RCUDA(rs.mats.reserve(mats.size()));
  RCUDA(rs.lights.reserve(lights.size()));
  if (!pix.empty()) {
    RCUDA(cudaMemcpyAsync(rs.pix.p, pix.data(), pix.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  if (!mats.empty()) RCUDA(cudaMemcpyAsync(rs.mats.p, mats.data(), mats.size() * sizeof(DevMat), cudaMemcpyHostToDevice, ctx->stream));
  if (!lights.empty()) RCUDA(cudaMemcpyAsync(rs.lights.p, lights.data(), lights.size() * sizeof(DevLight), cudaMemcpyHostToDevice, ctx->stream));
  {
    const int nt = ((int)lights.size() + 1) * 12;
    k_spectrum_tables<<<(nt + 127) / 128, 128, 0, ctx->stream>>>(rs.lights.p, (int)lights.size(), rs.qmc.p);
    RCUDA(cudaGetLastError());
  }
  RCUDA(rs.rayq0.reserve(P)); RCUDA(rs.rayq1.reserve(P)); RCUDA(rs.pathq0.reserve(P)); RCUDA(rs.pathq1.reserve(P));
  RCUDA(rs.hits.reserve(P)); RCUDA(rs.lambda.reserve(P)); RCUDA(rs.time.reserve(P)); RCUDA(rs.vmat.reserve(P));
  RCUDA(rs.invtot.reserve(P * std::max(1, rs.nlights) * rs.nlobes));
  const size_t SL = (size_t)S * rs.nlobes;
  RCUDA(rs.contrib.reserve(P * SL)); RCUDA(rs.sray.reserve(P * SL)); RCUDA(rs.sslot.reserve(P * SL));
  RCUDA(rs.L.reserve(P * rs.levels)); RCUDA(rs.T.reserve(P * rs.levels));
  if (rs.textured) {
    RCUDA(rs.vmats.reserve(P)); RCUDA(rs.diff.reserve(P * 12));
    RCUDA(rs.rawmats.reserve(ctx->materials.size())); RCUDA(rs.mat_tex.reserve(ctx->materials.size()));
    std::vector<MatTex> mt(ctx->materials.size());
    for (size_t i = 0; i < mt.size() && i < ctx->mat_tex.size(); i++) {
      mt[i] = ctx->mat_tex[i];
      mt[i].mask = 0;
      for (int k = 0; k < 12; k++)
        if (mt[i].slot[k].tex >= 0) mt[i].mask |= 1u << k;
    }
    RCUDA(cudaMemcpyAsync(rs.rawmats.p, ctx->materials.data(), ctx->materials.size() * sizeof(VgMaterial), cudaMemcpyHostToDevice, ctx->stream));
    RCUDA(cudaMemcpyAsync(rs.mat_tex.p, mt.data(), mt.size() * sizeof(MatTex), cudaMemcpyHostToDevice, ctx->stream));
    RCUDA(cudaStreamSynchronize(ctx->stream));
  }
  RCUDA(rs.counts.reserve(16)); RCUDA(rs.stats.reserve(8));
  RCUDA(cudaMemsetAsync(rs.counts.p, 0, 16 * sizeof(int), ctx->stream));
  RCUDA(cudaMemsetAsync(rs.stats.p, 0, 8 * sizeof(unsigned long long), ctx->stream));
  RCUDA(cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope

  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_queue<0, 0>, kTraceBlock, trace_smem_bytes(0));
  rs.trace_grid = ctx->sm_count * std::max(1, nb);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_queue<0, 2>, kTraceBlock, trace_smem_bytes(2));
  rs.coop_grid = ctx->sm_count * std::max(1, nb);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_queue<0, 66>, kTraceBlock, trace_smem_bytes(66));
  rs.coop_grid_mot = ctx->sm_count * std::max(1, nb);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_queue<1, 3>, kTraceBlock, trace_smem_bytes(3));
  rs.shadow_grid = ctx->sm_count * std::max(1, nb);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_queue<1, 67>, kTraceBlock, trace_smem_bytes(67));
  rs.shadow_grid_mot = ctx->sm_count * std::max(1, nb);
  // The scramble rows. The steady-state vg_set_scramble uploads straight from the caller's table and does not refresh the host
  // copy (a 100 MB memcpy per frame); if the device still holds the rows of exactly this frame / partition / pixel order they
  // are the latest ones and stay. Otherwise they are re-gathered from the host copy, which must then be current.
  const bool resident = rs.scr_valid && rs.scr_w == W && rs.scr_h == H && rs.scr_rank == ctx->rank && rs.scr_world == ctx->world &&
                        (rs.scr_by_pixel || rs.scr_pixel_block == ctx->opt_pixel_block);
  if (!resident) {
    if (ctx->scramble_stale) {
      if (rs.scr_valid && rs.scr_by_pixel && rs.scr_w == W && rs.scr_h == H) {
        // the device holds the caller's latest whole table: bring the host copy up to date from it
        RCUDA(cudaMemcpyAsync(ctx->scramble.data(), rs.scr.p, (size_t)W * H * 48, cudaMemcpyDeviceToHost, ctx->stream));
        RCUDA(cudaStreamSynchronize(ctx->stream));
        ctx->scramble_stale = false;
      } else {
        return ctx->fail(VG_ERR_INVALID, "the image partition or pixel order changed after a vg_set_scramble that went straight to the device: "
                                         "call vg_set_scramble again");
      }
    }
    rc = upload_scramble(ctx, ctx->scramble.data());
    if (rc != VG_OK) return rc;
  }
  rs.ready = true;
  return VG_OK;
}

int render_set_scramble(vg_ctx* ctx, const uint64_t* table, int64_t npix) {
  if ((int64_t)ctx->xres * ctx->yres != npix && ctx->xres > 0) return ctx->fail(VG_ERR_INVALID, "vg_set_scramble: table size != XRes*YRes");
  if (ctx->rs && ctx->rs->ready) {
    // steady state: this context's rows go straight from the caller's table to the device. The host copy is NOT refreshed
    // (100 MB per frame at 1080p); it is marked stale and prepare() either keeps the device rows or refuses to fall back on it.
    if ((int64_t)ctx->scramble.size() != npix * 6) {
      ctx->scramble.assign(table, table + (size_t)npix * 6);
      ctx->scramble_stale = false;
    } else {
      ctx->scramble_stale = true;
    }
    return upload_scramble(ctx, table);
  }
  ctx->scramble.assign(table, table + (size_t)npix * 6);
  ctx->scramble_stale = false;
  if (ctx->rs) ctx->rs->scr_valid = false;
  return VG_OK;
}

int render_clear(vg_ctx* ctx) {
  int rc = ensure_fb(ctx);
  if (rc != VG_OK) return rc;
  RCUDA(cudaMemsetAsync(ctx->rs->fb.p, 0, (size_t)ctx->xres * ctx->yres * 3 * sizeof(float), ctx->stream));
  RCUDA(cudaStreamSynchronize(ctx->stream));
  return VG_OK;
}

int render_fb_device(vg_ctx* ctx, float** d_fb) {
  int rc = ensure_fb(ctx);
  if (rc != VG_OK) return rc;
  *d_fb = ctx->rs->fb.p;
  return VG_OK;
}

// vg_render_frame: the frame step as ONE pipelined call. The image is cut into slices of tile rows; the scramble rows of slice
// s+1 go up (copy stream) and the finished pixels of slice s-1 come back — multi-GPU exchange (comm.cu) and D2H of the slice's
// image rows (out stream) — while slice s renders on the main stream. Every (pixel, iteration) sample is what vg_render
// computes, so the frame is bit-identical to vg_set_scramble + vg_render + vg_gather_frame.
struct FramePipe {
  const uint64_t* table;  // the caller's framescramble table (page-locked for the overlapped upload)
  bool clear;
  int slices;
};
static int render_run_impl(vg_ctx* ctx, int iter_begin, int iter_end, float* fb_out, const FramePipe* fp);
int render_run(vg_ctx* ctx, int iter_begin, int iter_end, float* fb_out) { return render_run_impl(ctx, iter_begin, iter_end, fb_out, nullptr); }

int render_frame(vg_ctx* ctx, const uint64_t* table, int64_t npix, int iter_begin, int iter_end, int clear_first, float* fb_out) {
  if (!table || (int64_t)ctx->xres * ctx->yres != npix) return ctx->fail(VG_ERR_INVALID, "vg_render_frame: table size != XRes*YRes");
  int rc = ensure_fb(ctx);
  if (rc != VG_OK) return rc;
  RenderState& rs = *ctx->rs;
  const bool pipelined = rs.ready && is_pinned_host(table) && (!fb_out || is_pinned_host(fb_out)) && rs.nown > 0;
  if (!pipelined) {
    // first call (nothing prepared yet) or pageable buffers: the plain sequence, same result
    rc = render_set_scramble(ctx, table, npix);
    if (rc != VG_OK) return rc;
    if (clear_first && (rc = render_clear(ctx)) != VG_OK) return rc;
    if (ctx->comm) {
      rc = render_run_impl(ctx, iter_begin, iter_end, nullptr, nullptr);
      return rc != VG_OK ? rc : comm_gather_rows(ctx, ctx->stream, 0, (ctx->yres + 31) / 32, fb_out, true);
    }
    return render_run_impl(ctx, iter_begin, iter_end, fb_out, nullptr);
  }
  // Slices pay when a slice is still a large batch: each one adds a set of kernel launches with their ramp-up and tail.
  // Measured on B200 (C2, 64 spp): one GPU 77.81 -> 77.19 ms per frame step with 4 slices (device time 75.4 ms; the exposed copies
  // shrink from 2.4 to 1.6 ms); two GPUs 39.52 vs 39.65 ms, i.e. no gain once the per-rank slice is a quarter of half a frame and
  // every slice carries its own NCCL exchange; a 512x512 16-spp frame (0.75 ms of rendering) loses 0.1 ms. Hence: at most
  // "frame_slices" slices, none smaller than 16 M paths, and the multi-GPU case keeps the whole frame as one slice.
  const long long paths = (long long)rs.nown * std::max(0, iter_end - iter_begin);
  int slices = (int)std::min<long long>(std::max(1, ctx->opt_frame_slices), std::max<long long>(1, paths / (16ll << 20)));
  if (ctx->opt_frame_slices_force) slices = std::max(1, ctx->opt_frame_slices);
  if (ctx->world > 1 && !ctx->opt_frame_slices_multi) slices = 1;
  FramePipe fp{table, clear_first != 0, slices};
  return render_run_impl(ctx, iter_begin, iter_end, fb_out, &fp);
}

static int render_run_impl(vg_ctx* ctx, int iter_begin, int iter_end, float* fb_out, const FramePipe* fp) {
  if (iter_begin < 0 || iter_end < iter_begin) return ctx->fail(VG_ERR_INVALID, "vg_render: bad iteration range");
  if (iter_end >= (1 << 28)) return ctx->fail(VG_ERR_INVALID, "vg_render: iteration index beyond the 28 frame bits of RasterXY(12,...)");
  int rc = prepare(ctx);
  if (rc != VG_OK) return rc;
  RenderState& rs = *ctx->rs;
  cudaStream_t st = ctx->stream;
  const int tilesY = (ctx->yres + 31) / 32;
  const int S = fp ? std::min(fp->slices, tilesY) : 1;
  cudaStream_t st_in = nullptr, st_out = nullptr;
  if (fp) {
    for (int k = 0; k < 2; k++) {
      if (!ctx->pipe_stream[k]) RCUDA(cudaStreamCreateWithFlags(&ctx->pipe_stream[k], cudaStreamNonBlocking));
    }
    st_in = ctx->pipe_stream[0];
    st_out = ctx->pipe_stream[1];
  }

  RenderParams p;
  std::memset(&p, 0, sizeof(p));
  p.sc = ctx->dev;
  p.xres = ctx->xres; p.yres = ctx->yres; p.nown = rs.nown; p.P = rs.P;
  p.pix = rs.pix.p; p.qmc = rs.qmc.p; p.scr = rs.scr.p; p.scr_by_pixel = rs.scr_by_pixel ? 1 : 0; p.cam = ctx->camera; p.cam_keys = ctx->d_cam_keys.p; p.cam_nkeys = ctx->cam_nkeys; p.mats = rs.mats.p; p.lights = rs.lights.p;
  p.filter_cdf = ctx->filter_n > 0 ? rs.filter.p : nullptr; p.filter_n = ctx->filter_n; p.filter_w = ctx->filter_w;
  p.nlights = rs.nlights; p.S = rs.S; p.levels = rs.levels; p.trace_last_level = (ctx->opt_trace_last_level || rs.levels == 5) ? 1 : 0;
  p.nlobes = rs.nlobes;
  p.rayq[0] = rs.rayq0.p; p.rayq[1] = rs.rayq1.p; p.pathq[0] = rs.pathq0.p; p.pathq[1] = rs.pathq1.p;
  p.hits = rs.hits.p; p.lambda = rs.lambda.p; p.time = rs.time.p; p.v_mat = rs.vmat.p; p.v_invtot = rs.invtot.p;
  p.contrib = rs.contrib.p; p.sray = rs.sray.p; p.sslot = rs.sslot.p; p.L = rs.L.p; p.T = rs.T.p;
  p.counts = rs.counts.p; p.stats = rs.stats.p; p.fb = rs.fb.p;
  if (rs.textured) {
    p.tex = ctx->tex_store();
    p.rawmats = rs.rawmats.p; p.mat_tex = rs.mat_tex.p; p.vmats = rs.vmats.p; p.diff = rs.diff.p;
    // sc.Image.PixelDelta (camera.go:316-317)
    p.pd0 = 2 * ctx->camera.tan_theta_focal / (float)ctx->xres;
    p.pd1 = 2 * ctx->camera.tan_theta_focal / (ctx->camera.aspect * (float)ctx->yres);
  }

  const int variant = ctx->opt_traversal;
  const bool xf = ctx->dev.n_xforms > 0;    // kernels that carry the instance enter/leave code (VARIANT & 16)
  const bool sph = ctx->dev.n_spheres > 0;
  const bool mot = ctx->dev.n_mtris > 0;    // kernels whose cooperative leaf phase takes motion triangles (VARIANT & 64)  // kernels that carry the analytic sphere leaf (traverse.cuh: VARIANT & 8)
  uint64_t launches = 0;
  size_t nev = 0;
  struct Timed { int kind; size_t e0, e1; int iters; };  // kind 0 closest-hit traversal, 1 any-hit traversal, 2 shading (k_surface + k_shade*), 3 / 4 any-hit traversal of level 0 with the cooperative / per-lane kernel
  // level-0 shadow queue kernel, per batch: option 0 / 1 = fixed, 2 = measured — while undecided the batches alternate between the
  // two kernels and their times per iteration are compared after the call (RenderState::l0_choice)
  const bool l0_candidate = variant == 2 && !xf && !sph && ctx->opt_shadow_unordered;
  const bool l0_tuning = l0_candidate && ctx->opt_shadow_level0_per_lane == 2 && rs.l0_choice < 0;
  int l0_kernel = !l0_candidate ? 0 : (ctx->opt_shadow_level0_per_lane != 2 ? ctx->opt_shadow_level0_per_lane : std::max(rs.l0_choice, 0));
  std::vector<Timed> timed;
  bool untimed = false;
  // one event between consecutive stages: the end of one stage is the start of the next
  auto mark = [&]() -> size_t {
    if (nev >= RenderState::kMaxTimedEvents) { untimed = true; return (size_t)-1; }
    cudaEventRecord(rs.ev(nev), st);
    return nev++;
  };
  auto stage = [&](int kind, size_t a, size_t b, int iters = 0) {
    if (a != (size_t)-1 && b != (size_t)-1) timed.push_back(Timed{kind, a, b, iters});
  };
  RCUDA(cudaMemsetAsync(rs.stats.p, 0, 8 * sizeof(unsigned long long), st));
  RCUDA(cudaEventRecord(rs.e0, st));
  auto slice_rows = [&](int sl, int* ty0, int* ty1) {
    *ty0 = (int)((long long)tilesY * sl / S);
    *ty1 = (int)((long long)tilesY * (sl + 1) / S);
  };
  if (fp) {
    // the copy stream starts where the main stream is (earlier work on the buffers is done), then takes the slices' rows in order
    RCUDA(cudaEventRecord(rs.pev(0), st));
    RCUDA(cudaStreamWaitEvent(st_in, rs.pev(0), 0));
    RCUDA(cudaStreamWaitEvent(st_out, rs.pev(0), 0));
    if (fp->clear) RCUDA(cudaMemsetAsync(rs.fb.p, 0, (size_t)ctx->xres * ctx->yres * 3 * sizeof(float), st));
    void* dview = nullptr;
    if (ctx->world > 1 && (cudaHostGetDevicePointer(&dview, const_cast<uint64_t*>(fp->table), 0) != cudaSuccess || !dview)) {
      cudaGetLastError();
      return ctx->fail(VG_ERR_CUDA, "vg_render_frame: the page-locked table has no device mapping");
    }
    for (int sl = 0; sl < S; sl++) {
      int ty0, ty1;
      slice_rows(sl, &ty0, &ty1);
      if (ctx->world == 1) {  // whole table, raster order: the slice is a run of image rows
        const size_t r0 = (size_t)std::min(ctx->yres, ty0 * 32) * ctx->xres, r1 = (size_t)std::min(ctx->yres, ty1 * 32) * ctx->xres;
        if (r1 > r0) RCUDA(cudaMemcpyAsync(rs.scr.p + r0 * 6, fp->table + r0 * 6, (r1 - r0) * 48, cudaMemcpyHostToDevice, st_in));
      } else {                // owned rows only, gathered straight from host memory by a kernel
        const int o0 = rs.row_start[(size_t)ty0], o1 = rs.row_start[(size_t)ty1];
        const long long n2 = (long long)(o1 - o0) * 3;
        if (n2 > 0) k_gather_scramble<<<(unsigned)((n2 + 255) / 256), 256, 0, st_in>>>(reinterpret_cast<const uint4*>(dview), rs.pix.p + o0, o1 - o0,
                                                                                         reinterpret_cast<uint4*>(rs.scr.p) + (size_t)o0 * 3);
      }
      RCUDA(cudaEventRecord(rs.pev(1 + (size_t)sl), st_in));
    }
    RCUDA(cudaGetLastError());
    rs.scr_by_pixel = ctx->world == 1;
    rs.scr_valid = true;
    rs.scr_w = ctx->xres; rs.scr_h = ctx->yres; rs.scr_rank = ctx->rank; rs.scr_world = ctx->world; rs.scr_pixel_block = ctx->opt_pixel_block;
    ctx->scramble_stale = true;
    p.scr_by_pixel = rs.scr_by_pixel ? 1 : 0;
  }
  for (int sl = 0; sl < S; sl++) {
    int ty0 = 0, ty1 = tilesY;
    slice_rows(sl, &ty0, &ty1);
    const int o0 = rs.row_start[(size_t)ty0], o1 = rs.row_start[(size_t)ty1];
    const int sn = o1 - o0;  // owned pixels of this slice
    p.pix = rs.pix.p + o0;
    p.scr = rs.scr_by_pixel ? rs.scr.p : rs.scr.p + (size_t)o0 * 6;
    p.nown = sn;
    if (fp) RCUDA(cudaStreamWaitEvent(st, rs.pev(1 + (size_t)sl), 0));
    for (int ib = iter_begin; ib < iter_end && sn > 0; ib += rs.iters) {
      const int niters = std::min(rs.iters, iter_end - ib);
      if (l0_tuning) l0_kernel = (rs.l0_seq++) & 1;
      const int np = sn * niters;
      {
        int G = 1;  // the largest power of two within the option, the warp size and this batch's iteration count
        while (G * 2 <= ctx->opt_iter_group && G * 2 <= 32 && G * 2 <= niters) G *= 2;
        p.pm_G = G;
        p.pm_B = 32 / G;
        p.pm_lG = 0;
        while ((1 << p.pm_lG) < G) p.pm_lG++;
        p.pm_lB = 5 - p.pm_lG;
        p.pm_nownB = sn - sn % p.pm_B;
        p.pm_nitG = niters - niters % G;
        p.pm_A = p.pm_nownB * p.pm_nitG;
        p.pm_rem = sn - p.pm_nownB;
        p.niters = niters;
      }
      k_reset<<<1, 1, 0, st>>>(rs.counts.p, np);
      // L and T need no clearing between batches: every path that reaches level k has L[k] (k_resolve) and T[k] (k_shade: zero
      // unless the mirror lobe continues the path) written at that level, and k_accumulate's fold multiplies everything beyond a
      // path's last level by that level's T = 0 with the reference's own clamp (NaN -> 0, std.go:255-259) in between, so stale
      // values of deeper levels cannot reach the pixel. (C3: two 4 GB memsets per batch, 23 ms of a 880 ms frame.) The one
      // exception is the DebugShader level, written only where a level-4 ray lands on such a surface.
      if (rs.levels == 5) {
        RCUDA(cudaMemsetAsync(rs.L.p + (size_t)4 * rs.P, 0, (size_t)rs.P * sizeof(float4), st));
        RCUDA(cudaMemsetAsync(rs.T.p + (size_t)4 * rs.P, 0, (size_t)rs.P * sizeof(float4), st));
      }
      if (p.cam_nkeys > 1) k_raygen<true><<<(np + 255) / 256, 256, 0, st>>>(p, ib, niters);
      else k_raygen<false><<<(np + 255) / 256, 256, 0, st>>>(p, ib, niters);
      launches += 2;
      int qin = 0;
      // levels 0..3 are shaded; with trace_last_level the level-4 rays are traced (and counted) but not shaded
      const int nlev = rs.levels == 1 ? 1 : (p.trace_last_level ? 5 : 4);
      for (int level = 0; level < nlev; level++) {
        const int qout = 1 - qin;
        if (ctx->opt_capture_levels & (1 << level)) {
          // debug / benchmarking aid: keep a host copy of this level's closest-hit input queue (vg_captured_rays)
          int cnt = 0;
          RCUDA(cudaMemcpyAsync(&cnt, rs.counts.p + qin, sizeof(int), cudaMemcpyDeviceToHost, st));
          RCUDA(cudaStreamSynchronize(st));
          const size_t old = ctx->captured.size();
          ctx->captured.resize(old + (size_t)cnt);
          if (cnt > 0) RCUDA(cudaMemcpy(ctx->captured.data() + old, p.rayq[qin], (size_t)cnt * sizeof(VgRay), cudaMemcpyDeviceToHost));
        }
        const size_t ev_a = mark();
        // camera rays are coherent: the per-lane loop with its compile-time axis specialisation is faster there (measured
        // 3.81 vs 3.48 Grays/s); every later level is incoherent and takes the cooperative leaf phase. Motion meshes are the
        // exception: their leaves cost 2-3x a static one (two keys to load and lerp), so the cooperative phase wins at level 0
        // too (C4: 49.6 vs 56.7 ms per frame)
        if (xf) k_trace_queue<0, 26><<<rs.coop_grid, kTraceBlock, trace_smem_bytes(26), st>>>(p, qin);
        else if (sph) {
          if (variant == 2 && (level > 0 || !ctx->opt_primary_per_lane)) k_trace_queue<0, 10><<<rs.coop_grid, kTraceBlock, trace_smem_bytes(10), st>>>(p, qin);
          else k_trace_queue<0, 8><<<rs.trace_grid, kTraceBlock, trace_smem_bytes(8), st>>>(p, qin);
        } else if (mot && variant == 2 && (level > 0 || !ctx->opt_primary_per_lane_motion)) k_trace_queue<0, 66><<<rs.coop_grid_mot, kTraceBlock, trace_smem_bytes(66), st>>>(p, qin);
        else if (variant == 1) k_trace_queue<0, 1><<<rs.trace_grid, kTraceBlock, trace_smem_bytes(1), st>>>(p, qin);
        else if (variant == 2 && (level > 0 || !ctx->opt_primary_per_lane)) k_trace_queue<0, 2><<<rs.coop_grid, kTraceBlock, trace_smem_bytes(2), st>>>(p, qin);
        else k_trace_queue<0, 0><<<rs.trace_grid, kTraceBlock, trace_smem_bytes(0), st>>>(p, qin);
        const size_t ev_b = mark();
        stage(0, ev_a, ev_b);
        launches++;
        if (level <= 3) {
          if (rs.textured) {
            const int fast = ctx->opt_precise_trig ? 0 : 1;
            if (ctx->opt_texture_coop) k_surface<true><<<(np + 127) / 128, 128, 0, st>>>(p, level, qin, fast);
            else k_surface<false><<<(np + 127) / 128, 128, 0, st>>>(p, level, qin, fast);
            launches++;
          }
          const bool h1 = rs.max_light_samples <= 2 || level > 0;
          if (rs.generic) {
            if (ctx->opt_precise_trig) k_shade_generic<false><<<(np + 127) / 128, 128, 0, st>>>(p, level, qin, qout, ib);
            else k_shade_generic<true><<<(np + 127) / 128, 128, 0, st>>>(p, level, qin, qout, ib);
          } else if (ctx->opt_precise_trig) {
            if (h1) k_shade<false, true><<<(np + 127) / 128, 128, 0, st>>>(p, level, qin, qout, ib);
            else k_shade<false, false><<<(np + 127) / 128, 128, 0, st>>>(p, level, qin, qout, ib);
          } else {
            if (h1) k_shade<true, true><<<(np + 127) / 128, 128, 0, st>>>(p, level, qin, qout, ib);
            else k_shade<true, false><<<(np + 127) / 128, 128, 0, st>>>(p, level, qin, qout, ib);
          }
          const size_t ev_c = mark();
          stage(2, ev_b, ev_c);
          if (xf) {
            if (ctx->opt_shadow_unordered) k_trace_queue<1, 27><<<rs.shadow_grid, kTraceBlock, trace_smem_bytes(27), st>>>(p, 0);
            else k_trace_queue<1, 26><<<rs.shadow_grid, kTraceBlock, trace_smem_bytes(26), st>>>(p, 0);
          } else if (sph) {
            if (variant == 2 && ctx->opt_shadow_unordered) k_trace_queue<1, 11><<<rs.shadow_grid, kTraceBlock, trace_smem_bytes(11), st>>>(p, 0);
            else if (variant == 2) k_trace_queue<1, 10><<<rs.shadow_grid, kTraceBlock, trace_smem_bytes(10), st>>>(p, 0);
            else k_trace_queue<1, 8><<<rs.shadow_grid, kTraceBlock, trace_smem_bytes(8), st>>>(p, 0);
          } else if (level == 0 && l0_kernel == 1) {
            // the shadow rays of camera hits are as coherent as the camera rays (a warp = one pixel x 32 iterations towards one light):
            // the per-lane loop, here without the ordered push
            if (mot) k_trace_queue<1, 68><<<rs.shadow_grid_mot, kTraceBlock, trace_smem_bytes(68), st>>>(p, 0);
            else k_trace_queue<1, 4><<<rs.shadow_grid, kTraceBlock, trace_smem_bytes(4), st>>>(p, 0);
          } else if (mot && variant == 2) {
            if (ctx->opt_shadow_unordered) k_trace_queue<1, 67><<<rs.shadow_grid_mot, kTraceBlock, trace_smem_bytes(67), st>>>(p, 0);
            else k_trace_queue<1, 66><<<rs.shadow_grid_mot, kTraceBlock, trace_smem_bytes(66), st>>>(p, 0);
          } else if (variant == 1) k_trace_queue<1, 1><<<rs.shadow_grid, kTraceBlock, trace_smem_bytes(1), st>>>(p, 0);
          else if (variant == 2 && ctx->opt_shadow_per_lane) k_trace_queue<1, 0><<<rs.shadow_grid, kTraceBlock, trace_smem_bytes(0), st>>>(p, 0);
          else if (variant == 2 && ctx->opt_shadow_unordered) k_trace_queue<1, 3><<<rs.shadow_grid, kTraceBlock, trace_smem_bytes(3), st>>>(p, 0);
          else if (variant == 2) k_trace_queue<1, 2><<<rs.shadow_grid, kTraceBlock, trace_smem_bytes(2), st>>>(p, 0);
          else k_trace_queue<1, 0><<<rs.shadow_grid, kTraceBlock, trace_smem_bytes(0), st>>>(p, 0);
          stage(level == 0 ? 3 + l0_kernel : 1, ev_c, mark(), np);  // (np paths in this batch: slices of a frame differ in size)
          if (rs.levels > 1) {
            k_resolve<<<(np + 255) / 256, 256, 0, st>>>(p, level, qin);
            launches++;
          }
          launches += 2;
        }
        if (level == 4 && rs.levels == 5) {
          k_debug_last<<<(np + 255) / 256, 256, 0, st>>>(p, level, qin);
          launches++;
        }
        if (level + 1 < nlev) {
          k_next_level<<<1, 1, 0, st>>>(rs.counts.p, qout);
          launches++;
        }
        qin = qout;
      }
      if (rs.levels > 1) k_accumulate<<<(sn + 255) / 256, 256, 0, st>>>(p, ib, niters);
      else if (p.pm_G == 32 && (niters & 31) == 0 && ctx->opt_accumulate_tiled) k_resolve_accumulate_t<<<(sn + 63) / 64, 64, 0, st>>>(p, ib, niters);
      else if (p.nlobes == 1 && p.S == 4 && p.nlights == 2 && ctx->opt_accumulate_wide) k_resolve_accumulate<true><<<(sn + 255) / 256, 256, 0, st>>>(p, ib, niters);
      else k_resolve_accumulate<false><<<(sn + 255) / 256, 256, 0, st>>>(p, ib, niters);
      launches++;
    }
    if (fp) {
      // this slice's pixels are final: send them on their way while the next slice renders
      RCUDA(cudaEventRecord(rs.pev(1 + (size_t)S + (size_t)sl), st));
      RCUDA(cudaStreamWaitEvent(st_out, rs.pev(1 + (size_t)S + (size_t)sl), 0));
      if (ctx->comm) {
        rc = comm_gather_rows(ctx, st_out, ty0, ty1, fb_out, false);
        if (rc != VG_OK) return rc;
      } else if (fb_out) {
        const size_t r0 = (size_t)std::min(ctx->yres, ty0 * 32) * ctx->xres * 3, r1 = (size_t)std::min(ctx->yres, ty1 * 32) * ctx->xres * 3;
        if (r1 > r0) RCUDA(cudaMemcpyAsync(fb_out + r0, rs.fb.p + r0, (r1 - r0) * sizeof(float), cudaMemcpyDeviceToHost, st_out));
      }
    }
  }
  RCUDA(cudaGetLastError());
  RCUDA(cudaEventRecord(rs.e1, st));
  unsigned long long hstats[7] = {0, 0, 0, 0, 0, 0, 0};
  int flags = 0;
  RCUDA(cudaMemcpyAsync(hstats, rs.stats.p, sizeof(hstats), cudaMemcpyDeviceToHost, st));
  RCUDA(cudaMemcpyAsync(&flags, rs.counts.p + 5, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (fp) {
    RCUDA(cudaStreamSynchronize(st_out));
    RCUDA(cudaStreamSynchronize(st_in));
  } else if (fb_out && is_pinned_host(fb_out)) {
    // the caller's buffer is page-locked: DMA straight into it
    RCUDA(cudaMemcpyAsync(fb_out, rs.fb.p, (size_t)ctx->xres * ctx->yres * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  } else if (fb_out) {
    // device -> pinned staging -> caller's (pageable) buffer
    const size_t bytes = (size_t)ctx->xres * ctx->yres * 3 * sizeof(float);
    if (rs.fb_pinned_bytes < bytes) {
      if (rs.fb_pinned) cudaFreeHost(rs.fb_pinned);
      rs.fb_pinned = nullptr;
      RCUDA(cudaMallocHost((void**)&rs.fb_pinned, bytes));
      rs.fb_pinned_bytes = bytes;
    }
    // D2H in chunks, each followed by an event; host threads copy a chunk from the pinned staging buffer into the caller's
    // (pageable) buffer as soon as its event has fired, so the host memcpy overlaps the rest of the transfer
    const int nchunks = 8, nthreads = 4;
    const size_t per = ((bytes + nchunks - 1) / nchunks + 255) & ~(size_t)255;
    for (int c = 0; c < nchunks; c++) {
      const size_t off = (size_t)c * per;
      if (off < bytes) RCUDA(cudaMemcpyAsync((char*)rs.fb_pinned + off, (const char*)rs.fb.p + off, std::min(per, bytes - off), cudaMemcpyDeviceToHost, st));
      RCUDA(cudaEventRecord(rs.fetch_ev(c), st));
    }
    {
      char* dst = (char*)fb_out;
      const char* src = (const char*)rs.fb_pinned;
      const int device = ctx->device;
      auto work = [&, dst, src](int t) {
        if (t > 0) cudaSetDevice(device);
        for (int c = t; c < nchunks; c += nthreads) {
          cudaEventSynchronize(rs.fetch_ev(c));
          const size_t off = (size_t)c * per;
          if (off < bytes) std::memcpy(dst + off, src + off, std::min(per, bytes - off));
        }
      };
      for (int c = 0; c < nchunks; c++) rs.fetch_ev(c);  // create before the threads start
      std::vector<std::thread> th;
      for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
      work(0);
      for (auto& t : th) t.join();
    }
  }
  RCUDA(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, rs.e0, rs.e1);
  ctx->stats.render_ms = ms;
  // per-stage times and launch counts of THIS call (stages beyond the event pool are not timed; `untimed` says so)
  ctx->stats.closest_ms = ctx->stats.shadow_ms = ctx->stats.shade_ms = 0;
  ctx->stats.closest_launches = ctx->stats.shadow_launches = 0;
  for (const Timed& k : timed) {
    float t = 0;
    cudaEventElapsedTime(&t, rs.ev(k.e0), rs.ev(k.e1));
    if (k.kind == 0) { ctx->stats.closest_ms += t; ctx->stats.closest_launches++; }
    else if (k.kind == 1 || k.kind == 3 || k.kind == 4) {
      ctx->stats.shadow_ms += t;
      ctx->stats.shadow_launches++;
      if (l0_tuning && k.kind >= 3 && k.iters > 0) {  // the first sample of a kernel carries its module load: keep the minimum
        const int kk = k.kind - 3;
        const float per = t / (float)k.iters;
        rs.l0_ms_per_iter[kk] = rs.l0_trials[kk] == 0 ? per : std::min(rs.l0_ms_per_iter[kk], per);
        rs.l0_trials[kk]++;
      }
    }
    else ctx->stats.shade_ms += t;
  }
  if (l0_tuning && rs.l0_trials[0] >= 2 && rs.l0_trials[1] >= 2) rs.l0_choice = rs.l0_ms_per_iter[1] < rs.l0_ms_per_iter[0] ? 1 : 0;
  ctx->stats.shadow_level0_kernel = l0_candidate ? l0_kernel : -1;
  (void)untimed;
  ctx->stats.rays += hstats[0];
  ctx->stats.shadow_rays += hstats[1];
  ctx->stats.nodes_t += hstats[2];
  ctx->stats.tris_t += hstats[3];
  ctx->stats.shadow_nodes_t += hstats[4];
  ctx->stats.shadow_tris_t += hstats[5];
  ctx->stats.max_stack_depth = std::max<uint64_t>(ctx->stats.max_stack_depth, hstats[6]);
  ctx->stats.kernel_launches += launches;
  if (flags) {
    cudaMemsetAsync(rs.counts.p + 5, 0, sizeof(int), st);
    if (flags & 1) return ctx->fail(VG_ERR_INVALID, "a shaded ShaderStd has no weight (DiffuseStrength + Spec1Strength == 0; the reference panics: std.go:141-143)");
    if (flags & 4) return ctx->fail(VG_ERR_INVALID, "traversal stack overflow (the reference's 90-entry stack would panic)");
  }
  return VG_OK;
}

}  // namespace vg
