// Per-ray two-level QBVH / MQBVH traversal for sm_100a. One thread owns one ray.
//
// Semantics are those of the reference's stack traversal, so that hit/miss, primitive id and t are
// bit-identical on identical rays:
//   core.Ray.Setup                core/ray.go:103-148        -> ray_setup()
//   intersectBoxes (SSE 4-box)    qbvh/intersect_amd64.s:13-100 -> box4()
//   qbvh.Trace                    qbvh/intersect.go:91-246   -> trace_ray() static-node branch
//   qbvh.TraceMotion              qbvh/motionintersect.go:22-127 -> trace_ray() motion-node branch
//   PolyMesh.TraceElems           builtin/geom/polymesh/trace.go:108-194 -> tri_test<false>()
//   PolyMesh.TraceMotionElems     builtin/geom/polymesh/trace.go:520-615 -> tri_test<true>()
//   Scene.Trace / TraceElems      builtin/scene/scene.go:30-78 -> geom-leaf branch (unified stack)
//
// This translation unit must be compiled with -fmad=false: Go/amd64 never contracts a*b+c, and every
// float operation below has to round exactly like the SSE scalar code. Min/max are written as
// compare-selects in the operand order of the x86 MINPS/MAXPS they replace (second operand wins on NaN).
//
// B200 mapping: the 128-B node is fetched with 8 LDG.128 through the read-only path; a leaf's
// triangles are 3 LDG.128 each from one contiguous run; the traversal stack lives in shared memory
// (kSmemStack entries per thread, interleaved by lane so a warp's accesses are conflict free) and
// spills to local memory only beyond that.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_scene.h"

namespace vg {

struct RayState {
  float ox, oy, oz;
  float dx, dy, dz;
  float idx, idy, idz;  // Dinv
  float s0, s1, s2;     // S
  float pkx, pky, pkz;  // origin permuted by (Kx,Ky,Kz)
  int kx, ky, kz;
  uint32_t signbits;  // bit a set <=> D[a] < 0
  uint32_t xsign;     // 0x80000000 if the reference would have swapped Kx/Ky (D[Kz] < 0)
  bool special;       // non-finite origin/direction/Dinv: NaNs can reach the slab test, use the exact x86 min/max emulation
  float tclosest;
  float time;
};

struct HitState {
  float u, v, w;
  int32_t prim;  // -1 = none
  int32_t geom;
  int32_t slot;  // global triangle slot of the hit (key-0 slot for motion meshes)
  uint32_t cnt;  // per-ray counters packed into one register: NodesT (interior-node visits) in bits 0-15, TrisT (leaf triangle counts) in bits 16-31
  int32_t xf_hit;   // instance (index into DevScene::xforms) inside which the current closest hit was found, or -1
  int32_t xf_last;  // last instance that reported a hit: the transform ShaderContext ends up with (instance.go:107-111), or -1
};

__device__ __forceinline__ float sel3(float x, float y, float z, int k) { return k == 0 ? x : (k == 1 ? y : z); }

// core/ray.go:103-148
__device__ __forceinline__ void ray_setup(RayState& r) {
  int kz = 0;
  if (fabsf(r.dy) > fabsf(r.dz)) {
    if (fabsf(r.dy) > fabsf(r.dx)) kz = 1;
  } else {
    if (fabsf(r.dz) > fabsf(r.dx)) kz = 2;
  }
  int kx = kz + 1;
  if (kx == 3) kx = 0;
  int ky = kx + 1;
  if (ky == 3) ky = 0;
  float dkz = sel3(r.dx, r.dy, r.dz, kz);
  // ray.go:132-136 swaps Kx and Ky when D[Kz] < 0. With the swapped axes the reference's edge functions are
  // fU = Cy*Bx - Cx*By etc. in terms of the UNswapped sheared coordinates, which is what (-Cx)*By - Cy*(-Bx) evaluates
  // to bit for bit (negation is exact and IEEE addition commutes, signed zeros included). So the device keeps the cyclic
  // order (kx,ky,kz) = (kz+1, kz+2, kz) mod 3 — which is what makes the compile-time axis specialisation possible — and
  // flips the sign of the three sheared x coordinates instead (one XOR each).
  r.xsign = dkz < 0.0f ? 0x80000000u : 0u;
  double z = (double)dkz;
  r.s2 = (float)(1.0 / z);
  r.s0 = (float)((double)sel3(r.dx, r.dy, r.dz, kx) / z);
  r.s1 = (float)((double)sel3(r.dx, r.dy, r.dz, ky) / z);
  r.idx = (float)(1.0 / (double)r.dx);
  r.idy = (float)(1.0 / (double)r.dy);
  r.idz = (float)(1.0 / (double)r.dz);
  r.kx = kx;
  r.ky = ky;
  r.kz = kz;
  r.pkx = sel3(r.ox, r.oy, r.oz, kx);
  r.pky = sel3(r.ox, r.oy, r.oz, ky);
  r.pkz = sel3(r.ox, r.oy, r.oz, kz);
  r.signbits = (r.dx < 0.0f ? 1u : 0u) | (r.dy < 0.0f ? 2u : 0u) | (r.dz < 0.0f ? 4u : 0u);
  const float chk = (r.ox + r.oy + r.oz) * 0.0f + (r.dx + r.dy + r.dz) * 0.0f + (r.idx + r.idy + r.idz) * 0.0f;  // NaN iff anything is Inf/NaN
  r.special = !(chk == 0.0f) || !(fabsf(r.idx) <= 3.4028234e38f) || !(fabsf(r.idy) <= 3.4028234e38f) || !(fabsf(r.idz) <= 3.4028234e38f);
}

// x86 MINPS dst,src: dst = (dst < src) ? dst : src ; MAXPS: dst = (dst > src) ? dst : src
__device__ __forceinline__ float x86min(float dst, float src) { return dst < src ? dst : src; }
__device__ __forceinline__ float x86max(float dst, float src) { return dst > src ? dst : src; }

// One child of intersect_amd64.s:13-100. Returns tNear; *hit = tNear <= tmax.
// EXACT: compare-selects in the operand order of MINPS/MAXPS (the second operand wins on NaN). !EXACT: FMNMX — identical
// bits whenever no operand is NaN, which ray_setup() guarantees for rays without Inf/NaN in origin or Dinv (node boxes are
// finite or +-Inf, never NaN, and Inf * finite-nonzero Dinv is Inf); only the sign of a zero can differ, and no decision or
// stored value depends on it.
template <bool EXACT>
__device__ __forceinline__ float box1(const RayState& r, float lx, float ly, float lz, float hx, float hy, float hz, bool* hit) {
  float t1 = (lx - r.ox) * r.idx;
  float t2 = (hx - r.ox) * r.idx;
  float x6 = EXACT ? x86min(t2, t1) : fminf(t2, t1);
  float x7 = EXACT ? x86max(t2, t1) : fmaxf(t2, t1);
  t1 = (ly - r.oy) * r.idy;
  t2 = (hy - r.oy) * r.idy;
  float x1 = EXACT ? x86min(t2, t1) : fminf(t2, t1);
  float x0 = EXACT ? x86max(t2, t1) : fmaxf(t2, t1);
  x6 = EXACT ? x86max(x6, x1) : fmaxf(x6, x1);
  x7 = EXACT ? x86min(x7, x0) : fminf(x7, x0);
  t1 = (lz - r.oz) * r.idz;
  t2 = (hz - r.oz) * r.idz;
  x1 = EXACT ? x86min(t2, t1) : fminf(t2, t1);
  x0 = EXACT ? x86max(t2, t1) : fmaxf(t2, t1);
  x6 = EXACT ? x86max(x6, x1) : fmaxf(x6, x1);
  x7 = EXACT ? x86min(x7, x0) : fminf(x7, x0);
  float tn = EXACT ? x86max(0.0f, x6) : fmaxf(0.0f, x6);
  *hit = tn <= x7;
  return tn;
}
struct Box4Out {
  float t0, t1, t2, t3;
  bool h0, h1, h2, h3;
};

// The same slab test with the subtractions and multiplications issued as packed pairs (sm_100 FADD2 / FMUL2: __fadd2_rn,
// __fmul2_rn — two IEEE round-to-nearest float32 operations per instruction, bit-identical to the scalar ones; a - b is issued as
// a + (-b), the same operation). The traversal kernels are issue-bound (ncu: 75 % of issue slots active), and the 24 FADD +
// 24 FMUL of a node visit are its largest single block of instructions: 24 packed ones replace them. A 128-bit node load
// delivers children (0,1) and (2,3) of one plane in adjacent registers, which is exactly the pair layout the packed form wants.
// MEASURED on B200 (round 2) and left OFF: C2 closest / shadow 19.23 / 32.93 ms with it, 19.16 / 33.04 ms without, C3 308.3 / 281.1
// vs 311.0 / 279.2 ms — no difference beyond noise. The packed instructions evidently occupy the FP32 pipe for two issue cycles, so
// the issue slots they free cannot be used by the dependent min/max chain, and the duplicated (-O, Dinv) pairs cost 6 registers
// under the 64/72-register caps (+8..32 B of spills). Kept as a build switch (bit-identical, parity-tested once).
#ifndef VG_BOX_F32X2
#define VG_BOX_F32X2 0
#endif
template <bool EXACT>
__device__ __forceinline__ void box_axis2(float2 lo, float2 hi, float2 no, float2 id, float2& tn, float2& tf, bool first) {
  const float2 t1 = __fmul2_rn(__fadd2_rn(lo, no), id);
  const float2 t2 = __fmul2_rn(__fadd2_rn(hi, no), id);
  const float nx = EXACT ? x86min(t2.x, t1.x) : fminf(t2.x, t1.x), ny = EXACT ? x86min(t2.y, t1.y) : fminf(t2.y, t1.y);
  const float fx = EXACT ? x86max(t2.x, t1.x) : fmaxf(t2.x, t1.x), fy = EXACT ? x86max(t2.y, t1.y) : fmaxf(t2.y, t1.y);
  if (first) {
    tn = make_float2(nx, ny);
    tf = make_float2(fx, fy);
  } else {
    tn.x = EXACT ? x86max(tn.x, nx) : fmaxf(tn.x, nx);
    tn.y = EXACT ? x86max(tn.y, ny) : fmaxf(tn.y, ny);
    tf.x = EXACT ? x86min(tf.x, fx) : fminf(tf.x, fx);
    tf.y = EXACT ? x86min(tf.y, fy) : fminf(tf.y, fy);
  }
}
template <bool EXACT>
__device__ __forceinline__ void box4_packed(const RayState& r, const float4& lx, const float4& ly, const float4& lz, const float4& hx, const float4& hy,
                                            const float4& hz, Box4Out& o) {
  const float2 nox = make_float2(-r.ox, -r.ox), noy = make_float2(-r.oy, -r.oy), noz = make_float2(-r.oz, -r.oz);
  const float2 ix = make_float2(r.idx, r.idx), iy = make_float2(r.idy, r.idy), iz = make_float2(r.idz, r.idz);
  float2 n01, f01, n23, f23;
  box_axis2<EXACT>(make_float2(lx.x, lx.y), make_float2(hx.x, hx.y), nox, ix, n01, f01, true);
  box_axis2<EXACT>(make_float2(lx.z, lx.w), make_float2(hx.z, hx.w), nox, ix, n23, f23, true);
  box_axis2<EXACT>(make_float2(ly.x, ly.y), make_float2(hy.x, hy.y), noy, iy, n01, f01, false);
  box_axis2<EXACT>(make_float2(ly.z, ly.w), make_float2(hy.z, hy.w), noy, iy, n23, f23, false);
  box_axis2<EXACT>(make_float2(lz.x, lz.y), make_float2(hz.x, hz.y), noz, iz, n01, f01, false);
  box_axis2<EXACT>(make_float2(lz.z, lz.w), make_float2(hz.z, hz.w), noz, iz, n23, f23, false);
  o.t0 = EXACT ? x86max(0.0f, n01.x) : fmaxf(0.0f, n01.x);
  o.t1 = EXACT ? x86max(0.0f, n01.y) : fmaxf(0.0f, n01.y);
  o.t2 = EXACT ? x86max(0.0f, n23.x) : fmaxf(0.0f, n23.x);
  o.t3 = EXACT ? x86max(0.0f, n23.y) : fmaxf(0.0f, n23.y);
  o.h0 = o.t0 <= f01.x;
  o.h1 = o.t1 <= f01.y;
  o.h2 = o.t2 <= f23.x;
  o.h3 = o.t3 <= f23.y;
}
template <bool EXACT>
__device__ __forceinline__ void box4(const RayState& r, const float4& lx, const float4& ly, const float4& lz, const float4& hx, const float4& hy,
                                     const float4& hz, Box4Out& o) {
#if VG_BOX_F32X2
  if (!EXACT) {  // (the NaN-exact path keeps the scalar form: rare rays, and its selects are written against MINPS/MAXPS operand order)
    box4_packed<false>(r, lx, ly, lz, hx, hy, hz, o);
    return;
  }
#endif
  o.t0 = box1<EXACT>(r, lx.x, ly.x, lz.x, hx.x, hy.x, hz.x, &o.h0);
  o.t1 = box1<EXACT>(r, lx.y, ly.y, lz.y, hx.y, hy.y, hz.y, &o.h1);
  o.t2 = box1<EXACT>(r, lx.z, ly.z, lz.z, hx.z, hy.z, hz.z, &o.h2);
  o.t3 = box1<EXACT>(r, lx.w, ly.w, lz.w, hx.w, hy.w, hz.w, &o.h3);
}

// Watertight ray/triangle test, trace.go:127-192 (static: bias `<` with eps+RayBias folded into `bias`)
// and trace.go:556-622 (motion: `<=` with RayBias). Returns true and updates tclosest,u,v,w on accept.
// KZ = 0,1,2: the ray's dominant axis is known at compile time (warp-uniform fast path, no selects); KZ = -1: per-lane selects.
template <int KZ>
__device__ __forceinline__ float pick(float x, float y, float z, int k) {
  if (KZ < 0) return sel3(x, y, z, k);
  const int kk = k;  // k is (KZ + const) % 3 at the call sites below; resolved at compile time
  return kk == 0 ? x : (kk == 1 ? y : z);
}
// The leaf-phase view of a ray: what the triangle tests read. The persistent kernels publish it to shared memory once per ray
// (coop_publish) and load it back at a leaf, so the eight values do not occupy registers (or local memory) during the node loop.
// The occlusion-only per-lane kernel leaves a leaf at its first accepted triangle (the reference finishes the leaf and then returns,
// intersect.go:231-236: the same "occluded"). Measured: C2 shadow 29.12 -> 28.49 ms. Also measured on that kernel and left as they
// were: refill below 8 / 16 / 24 / 28 active lanes 29.09 / 29.12 / 29.32 / 30.51 ms, 256-bit node loads 29.31 ms, 9 CTAs per SM
// (56 registers) 30.44 ms.
#ifndef VG_OCCL_LEAF_EXIT
#define VG_OCCL_LEAF_EXIT 1
#endif
struct LeafRay {
  float pkx, pky, pkz, s0, s1, s2;
  uint32_t xsign;
  int kx, ky, kz;
};
template <bool MOTION, int KZ>
__device__ __forceinline__ bool tri_test(const LeafRay& r, float& tclosest, float3 p0, float3 p1, float3 p2, float bias, float* U, float* V, float* W) {
  const int kz = KZ < 0 ? r.kz : KZ;
  const int kx = KZ < 0 ? r.kx : (KZ + 1) % 3;
  const int ky = KZ < 0 ? r.ky : (KZ + 2) % 3;
  const float AKz = pick<KZ>(p0.x, p0.y, p0.z, kz) - r.pkz;
  const float BKz = pick<KZ>(p1.x, p1.y, p1.z, kz) - r.pkz;
  const float CKz = pick<KZ>(p2.x, p2.y, p2.z, kz) - r.pkz;
  const float Cx = __uint_as_float(__float_as_uint((pick<KZ>(p2.x, p2.y, p2.z, kx) - r.pkx) - r.s0 * CKz) ^ r.xsign);
  const float By = (pick<KZ>(p1.x, p1.y, p1.z, ky) - r.pky) - r.s1 * BKz;
  const float Cy = (pick<KZ>(p2.x, p2.y, p2.z, ky) - r.pky) - r.s1 * CKz;
  const float Bx = __uint_as_float(__float_as_uint((pick<KZ>(p1.x, p1.y, p1.z, kx) - r.pkx) - r.s0 * BKz) ^ r.xsign);
  const float Ax = __uint_as_float(__float_as_uint((pick<KZ>(p0.x, p0.y, p0.z, kx) - r.pkx) - r.s0 * AKz) ^ r.xsign);
  const float Ay = (pick<KZ>(p0.x, p0.y, p0.z, ky) - r.pky) - r.s1 * AKz;
  float fU = Cx * By - Cy * Bx;
  float fV = Ax * Cy - Ay * Cx;
  float fW = Bx * Ay - By * Ax;
  if (fU == 0.0f || fV == 0.0f || fW == 0.0f) {
    double CxBy = (double)Cx * (double)By;
    double CyBx = (double)Cy * (double)Bx;
    fU = (float)(CxBy - CyBx);
    double AxCy = (double)Ax * (double)Cy;
    double AyCx = (double)Ay * (double)Cx;
    fV = (float)(AxCy - AyCx);
    double BxAy = (double)Bx * (double)Ay;
    double ByAx = (double)By * (double)Ax;
    fW = (float)(BxAy - ByAx);
  }
  if ((fU < 0.0f || fV < 0.0f || fW < 0.0f) && (fU > 0.0f || fV > 0.0f || fW > 0.0f)) return false;
  const float det = fU + fV + fW;
  if (det == 0.0f) return false;
  const float T = r.s2 * (fU * AKz + fV * BKz + fW * CKz);
  const uint32_t sgn = __float_as_uint(det) & 0x80000000u;
  const float Ts = __uint_as_float(__float_as_uint(T) ^ sgn);
  const float ds = __uint_as_float(__float_as_uint(det) ^ sgn);
  if (MOTION) {
    if (Ts <= bias * ds || Ts > tclosest * ds) return false;
  } else {
    if (Ts < bias * ds || Ts > tclosest * ds) return false;
  }
  const float rcp = 1.0f / det;
  *U = fU * rcp;
  *V = fV * rcp;
  *W = fW * rcp;
  tclosest = T * rcp;
  return true;
}

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// 256-bit read-only global load (sm_100: LDG.E.256.CONSTANT; PTX ld.global.nc.v8.f32, 32-byte aligned address). A divergent warp
// pays one L1TEX wavefront per lane and load whatever its width, so a 128-B node costs 4 wavefronts per lane instead of 8.
#ifndef VG_LDG256
#define VG_LDG256 1
#endif
__device__ __forceinline__ void ldg8(const void* p, float4& a, float4& b) {
#if VG_LDG256
#ifndef VG_NODE_HINT
#define VG_NODE_HINT ""
#endif
  asm volatile("ld.global.nc" VG_NODE_HINT ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
#else
  a = __ldg(reinterpret_cast<const float4*>(p));
  b = __ldg(reinterpret_cast<const float4*>(p) + 1);
#endif
}
// one static triangle record
__device__ __forceinline__ void ld_tri(const float4* tp, float4& v0, float4& v1, float4& v2) {
  if (kTriStride == 4) {
    float4 pad;
    ldg8(tp, v0, v1);
    ldg8(tp + 2, v2, pad);
  } else {
    v0 = __ldg(tp); v1 = __ldg(tp + 1); v2 = __ldg(tp + 2);
  }
}

// Traversal stack: kSmemStack entries per thread in shared memory, the rest in local memory.
#ifndef VG_SMEM_STACK
#define VG_SMEM_STACK 8
#endif
// Measured with a -DVG_STACK_STATS build (scripts/stack_depth.py, B200, 8 iterations of each config): the deepest stack any ray
// needed is 3 entries on C1, 11 on C2, 10 on C4 and 15 on the two-level 10 M-triangle C3 / C5 scene. 8 shared + 40 local entries
// leave a 3x margin over that (the reference reserves 90 and panics beyond, core/ray.go:158; here an overflow raises the error flag).
#ifndef VG_LOCAL_STACK
#define VG_LOCAL_STACK 40
#endif

// The shared part is addressed as a 32-bit shared-state-space byte address with a compile-time stride (every traversal kernel runs
// VG_TRACE_BLOCK threads per CTA): round 2's SASS showed the generic 64-bit base pointer and the runtime stride spilled to local
// memory and re-read (LDL + LDL.64) for every push of every node step of these L1TEX-wavefront-bound kernels.
#ifndef VG_TRACE_BLOCK
#define VG_TRACE_BLOCK 128
#endif
static const uint32_t kStackStrideBytes = VG_TRACE_BLOCK * 8;
__device__ __forceinline__ void sts_u2(uint32_t addr, uint2 v) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
// ONE register holds the stack: `top` = the shared-space address the next push of this thread would write, i.e.
// region + threadIdx.x * 8 + depth * kStackStrideBytes. The stride is 1024 B and a thread's offset inside an entry row is < 1024,
// so depth = (top - region) >> 10 with `region` CTA-uniform (the compiler keeps it in a uniform register / rematerialises it).
// The depth counter and the base address it used to be paired with were both spilled in the 64-register kernels.
struct Stack {
  uint32_t top;
  uint32_t region;  // CTA-uniform: shared-space address of the stack area
  uint2 local[VG_LOCAL_STACK];
  bool overflow;
#ifdef VG_STACK_STATS
  int maxsp = 0;  // measurement build only: deepest stack this thread has seen (VgStats.max_stack_depth)
#endif
  static_assert(kStackStrideBytes == 1024, "depth() assumes a 1024-byte entry row (128 threads x 8 B)");
  // `stack_region`: the CTA's stack area in dynamic shared memory (VG_TRACE_BLOCK x VG_SMEM_STACK entries)
  __device__ __forceinline__ void bind(const void* stack_region) {
    region = (uint32_t)__cvta_generic_to_shared(stack_region);
    top = region + threadIdx.x * 8u;
  }
  __device__ __forceinline__ int depth() const { return (int)((top - region) >> 10); }
  __device__ __forceinline__ bool empty() const { return top - region < kStackStrideBytes; }
  __device__ __forceinline__ void reset() { top = region + ((top - region) & (kStackStrideBytes - 1u)); }
  __device__ __forceinline__ void push(float t, int32_t node) {
    uint2 e = make_uint2(__float_as_uint(t), (uint32_t)node);
    const uint32_t off = top - region;
    if (off < VG_SMEM_STACK * kStackStrideBytes) sts_u2(top, e);
    else if (off < (VG_SMEM_STACK + VG_LOCAL_STACK) * kStackStrideBytes) local[(int)(off >> 10) - VG_SMEM_STACK] = e;
    else { overflow = true; return; }
    top += kStackStrideBytes;
#ifdef VG_STACK_STATS
    if (depth() > maxsp) maxsp = depth();
#endif
  }
  __device__ __forceinline__ uint2 pop() {
    top -= kStackStrideBytes;
    const uint32_t off = top - region;
    return off < VG_SMEM_STACK * kStackStrideBytes ? lds_u2(top) : local[(int)(off >> 10) - VG_SMEM_STACK];
  }
};

// Push order of intersect.go:137-216 / motionintersect.go:55-97, expressed as the pop order (reverse of push).
// With s0 = D[axis0]<0, s1 = D[axis1]<0, s2 = D[axis2]<0 the push sequence is
//   s0 ? {pair01, pair23} : {pair23, pair01},  pair01 = s1 ? (0,1) : (1,0),  pair23 = s2 ? (2,3) : (3,2).
struct Child4 {
  int32_t c[4];
  float t[4];
};

__device__ __forceinline__ void cswap(bool doit, int32_t& ca, float& ta, int32_t& cb, float& tb) {
  int32_t c0 = doit ? cb : ca, c1 = doit ? ca : cb;
  float t0 = doit ? tb : ta, t1 = doit ? ta : tb;
  ca = c0; cb = c1; ta = t0; tb = t1;
}

// Per-lane traversal state that survives across refills of the warp.
struct TravState {
  RayState r;
  HitState h;
  int32_t cur;  // node to process next; -1 = none (stack empty and nothing pending => the ray is finished)
};

__device__ __forceinline__ void trav_begin(const DevScene& sc, TravState& t, Stack& st) {
  t.h.prim = -1;
  t.h.geom = -1;
  t.h.slot = -1;
  t.h.u = t.h.v = t.h.w = 0.0f;
  t.h.cnt = 0;
  t.h.xf_hit = -1;
  t.h.xf_last = -1;
  st.reset();
  t.cur = sc.root;  // qbvh.Trace pushes the root with T = Tclosest and pops it at once (intersect.go:93-104)
}

// Pop until an entry survives the cull of intersect.go:106 (Tclosest < T); -1 when the stack is empty.
// !ORDERED (occlusion-only rays): Tclosest never changes before the ray terminates, so nothing is ever culled.
template <bool ORDERED = true>
__device__ __forceinline__ int32_t pop_next(const RayState& r, Stack& st) {
  while (!st.empty()) {
    const uint2 e = st.pop();
    if (!ORDERED || !(r.tclosest < __uint_as_float(e.x))) return (int32_t)e.y;
  }
  return -1;
}

// One interior node (static or motion): 4 box tests, ordered push, next node. intersect.go:113-216, motionintersect.go:44-97
// ORDERED = false is for rays whose only result is "occluded or not" (the integrator's shadow queue): every box-hit
// child is visited whatever the order until the first accepted triangle, so the sign-ordered push sequence is skipped.
// The set of nodes/leaves visited by an unoccluded ray, hence its NodesT/TrisT, is unchanged.
// WIDE: the node comes in as four 256-bit loads instead of eight 128-bit ones. Incoherent rays (the cooperative kernels: every
// lane in another node, L1TEX wavefronts the busiest unit at 89 %): +8 % (2145 -> 2323 Mrays/s on the incoherent batch);
// coherent camera rays (the per-lane kernel) lose 1 %, so that kernel keeps the 128-bit loads.
template <bool ORDERED = true, bool WIDE = false>
__device__ __forceinline__ void node_step(const DevScene& sc, TravState& t, Stack& st) {
  RayState& r = t.r;
  const int32_t node = t.cur;
  t.h.cnt++;
  int32_t c0, c1, c2, c3;
  float t0, t1, t2, t3;
  bool h0, h1, h2, h3;
  uint32_t a0, a1, a2;
  if (node < sc.n_static) {
    const DevNode* nd = sc.nodes + node;
    float4 lx, ly, lz, hx, hy, hz;
    uint4 m0, m1;
    if (WIDE) {
      float4 mm0, mm1;
      ldg8(&nd->lo_x, lx, ly);
      ldg8(&nd->lo_z, lz, hx);
      ldg8(&nd->hi_y, hy, hz);
      ldg8(&nd->m0, mm0, mm1);
      m0 = make_uint4(__float_as_uint(mm0.x), __float_as_uint(mm0.y), __float_as_uint(mm0.z), __float_as_uint(mm0.w));
      m1 = make_uint4(__float_as_uint(mm1.x), __float_as_uint(mm1.y), __float_as_uint(mm1.z), __float_as_uint(mm1.w));
    } else {
      lx = ldg4(&nd->lo_x); ly = ldg4(&nd->lo_y); lz = ldg4(&nd->lo_z);
      hx = ldg4(&nd->hi_x); hy = ldg4(&nd->hi_y); hz = ldg4(&nd->hi_z);
      m0 = __ldg(&nd->m0); m1 = __ldg(&nd->m1);
    }
    Box4Out bo;
    if (r.special) box4<true>(r, lx, ly, lz, hx, hy, hz, bo);
    else box4<false>(r, lx, ly, lz, hx, hy, hz, bo);
    t0 = bo.t0; t1 = bo.t1; t2 = bo.t2; t3 = bo.t3;
    h0 = bo.h0; h1 = bo.h1; h2 = bo.h2; h3 = bo.h3;
    a0 = m0.x; a1 = m0.y; a2 = m0.z;
    c0 = (int32_t)m0.w; c1 = (int32_t)m1.x; c2 = (int32_t)m1.y; c3 = (int32_t)m1.z;
  } else {
    // motionintersect.go:44-51: lerp the 24 box floats between the two keys, then the same box test
    const DevMotionNode mn = *(sc.mtopo + (node - sc.n_static));
    const int keys = (int)(mn.axes_keys >> 8);
    const float k = r.time * (float)(keys - 1);  // polymesh/trace.go:79-84, scene.go:49-54
    const float fk = floorf(k);
    const float tm = k - fk;
    const int key = (int)fk, key2 = (int)ceilf(k);
    const float4* b0 = sc.mboxes + (size_t)(mn.box_base + key * mn.box_key_stride) * 6;
    const float4* b1 = sc.mboxes + (size_t)(mn.box_base + key2 * mn.box_key_stride) * 6;
    const float om = 1.0f - tm;
    float4 bx[6];
    float4 pw[2], qw[2];
#pragma unroll
    for (int i = 0; i < 6; i++) {
      float4 p, q;
      if (WIDE) {  // 96-B box sets are 32-B aligned: three 256-bit loads per key
        if ((i & 1) == 0) {
          ldg8(b0 + i, pw[0], pw[1]);
          ldg8(b1 + i, qw[0], qw[1]);
        }
        p = pw[i & 1];
        q = qw[i & 1];
      } else {
        p = ldg4(b0 + i);
        q = ldg4(b1 + i);
      }
      bx[i].x = om * p.x + tm * q.x;
      bx[i].y = om * p.y + tm * q.y;
      bx[i].z = om * p.z + tm * q.z;
      bx[i].w = om * p.w + tm * q.w;
    }
    Box4Out bo;
    if (r.special) box4<true>(r, bx[0], bx[1], bx[2], bx[3], bx[4], bx[5], bo);
    else box4<false>(r, bx[0], bx[1], bx[2], bx[3], bx[4], bx[5], bo);
    t0 = bo.t0; t1 = bo.t1; t2 = bo.t2; t3 = bo.t3;
    h0 = bo.h0; h1 = bo.h1; h2 = bo.h2; h3 = bo.h3;
    a0 = mn.axes_keys & 3u; a1 = (mn.axes_keys >> 2) & 3u; a2 = (mn.axes_keys >> 4) & 3u;
    c0 = mn.child[0]; c1 = mn.child[1]; c2 = mn.child[2]; c3 = mn.child[3];
  }
  // Children that miss, are empty, or already lie beyond Tclosest would be culled at pop time
  // (intersect.go:106, Tclosest only shrinks): drop them now. NodesT is unaffected.
  if (!h0 || t0 > r.tclosest) c0 = -1;
  if (!h1 || t1 > r.tclosest) c1 = -1;
  if (!h2 || t2 > r.tclosest) c2 = -1;
  if (!h3 || t3 > r.tclosest) c3 = -1;
  if (ORDERED) {
    const bool s0 = (r.signbits >> a0) & 1u, s1 = (r.signbits >> a1) & 1u, s2 = (r.signbits >> a2) & 1u;
    // arrange (e0,e1,e2,e3) = push sequence
    cswap(!s1, c0, t0, c1, t1);  // pair01 = s1 ? (0,1) : (1,0)
    cswap(!s2, c2, t2, c3, t3);  // pair23 = s2 ? (2,3) : (3,2)
    cswap(!s0, c0, t0, c2, t2);  // s0 ? {pair01,pair23} : {pair23,pair01}
    cswap(!s0, c1, t1, c3, t3);
  }
  // The entry pushed last is the one the reference pops next: keep it in a register, push the others. Branch-free:
  // entry i is pushed iff it is valid and a valid entry follows it.
  const bool v0 = c0 != -1, v1 = c1 != -1, v2 = c2 != -1, v3 = c3 != -1;
  const int32_t next = v3 ? c3 : (v2 ? c2 : (v1 ? c1 : c0));
  const bool p0 = v0 && (v1 || v2 || v3), p1 = v1 && (v2 || v3), p2 = v2 && v3;
  if (st.top - st.region < (VG_SMEM_STACK - 2) * kStackStrideBytes) {  // depth + 3 <= VG_SMEM_STACK: all three in shared memory
    const uint32_t a0 = st.top, a1 = a0 + (p0 ? kStackStrideBytes : 0u), a2 = a1 + (p1 ? kStackStrideBytes : 0u);
    if (p0) sts_u2(a0, make_uint2(__float_as_uint(t0), (uint32_t)c0));
    if (p1) sts_u2(a1, make_uint2(__float_as_uint(t1), (uint32_t)c1));
    if (p2) sts_u2(a2, make_uint2(__float_as_uint(t2), (uint32_t)c2));
    st.top = a2 + (p2 ? kStackStrideBytes : 0u);
#ifdef VG_STACK_STATS
    if (st.depth() > st.maxsp) st.maxsp = st.depth();
#endif
  } else {
    if (p0) st.push(t0, c0);
    if (p1) st.push(t1, c1);
    if (p2) st.push(t2, c2);
  }
  t.cur = next != -1 ? next : pop_next<ORDERED>(r, st);
}

// Per warp (every persistent kernel): 32 per-LANE blocks of two float4 — the leaf-phase parameters of the lane's current ray
// {pkx, pky, pkz, s0} {s1, s2, xsign, Tclosest}, written ONCE when the ray is set up (coop_publish). The cooperative kernels add 32
// per-RANK records {kz | motion << 8 | owner lane << 16, leaf base, item start, Ray.Time} written by the lanes that hold a leaf in the
// current phase. Round 2 (first half) kept these values in RayState for the whole traversal: the 64/72-register kernels spilled
// them and re-read them (5 LDL + 3 STS.128 per lane and leaf phase in the SASS of the cooperative kernels).
struct CoopSmem {
  float4* rp;
};
__device__ __forceinline__ void coop_publish(const CoopSmem& cs, const RayState& r) {
  float4* b = cs.rp + (threadIdx.x & 31) * 2;
  b[0] = make_float4(r.pkx, r.pky, r.pkz, r.s0);
  b[1] = make_float4(r.s1, r.s2, __uint_as_float(r.xsign), r.tclosest);
}
__device__ __forceinline__ LeafRay leaf_ray(const CoopSmem& cs, int kz) {
  const float4* b = cs.rp + (threadIdx.x & 31) * 2;
  const float4 q0 = b[0], q1 = b[1];
  LeafRay lr;
  lr.pkx = q0.x; lr.pky = q0.y; lr.pkz = q0.z; lr.s0 = q0.w;
  lr.s1 = q1.x; lr.s2 = q1.y; lr.xsign = __float_as_uint(q1.z);
  lr.kz = kz;
  lr.kx = kz == 2 ? 0 : kz + 1;
  lr.ky = lr.kx == 2 ? 0 : lr.kx + 1;
  return lr;
}

// Triangle loops of one leaf. The next triangle's loads are issued before the current one is tested. An accepted triangle only
// records its slot (h.prim = -2 static / -3 motion): U, V, W and the ids are recomputed from the slot when the ray is stored
// (finalize_hit), so they are not carried through the traversal loops.
// How the next triangle of a leaf is fetched ahead (VG_LEAF_PIPE): 1 = into a second register set that is copied into the current one
// every iteration (12 MOV per triangle: `if (i + 1 < count)` is the largest line of the ncu source page, 6.5 %), 2 = a
// prefetch.global.L1 of the next record and plain loads (no second set, no copies), 0 = nothing. An unroll by two over two
// alternating register sets was MEASURED out at compile time: under the 64/72-register caps it spills both sets (stack frame
// 368 -> 576 B, LDL/STL inside the triangle loop).
// MEASURED (C2, closest / shadow ms): 1 -> 17.85 / 30.48, 0 -> 18.22 / 28.33, 2 -> 18.12 / 29.04: the closest-hit per-lane kernels
// keep the second register set, the occlusion-only kernel (64 registers) fetches nothing ahead.
#ifndef VG_LEAF_PIPE
#define VG_LEAF_PIPE 1
#endif
#ifndef VG_LEAF_PIPE_OCCL
#define VG_LEAF_PIPE_OCCL 0
#endif
template <int KZ, bool DEFER>
__device__ __forceinline__ bool leaf_static(const DevScene& sc, const LeafRay& lr, float& tclosest, HitState& h, int base, int count) {
  bool leafhit = false;
  const float4* tp = sc.tris + (size_t)base * kTriStride;
  auto test = [&](const float4& v0, const float4& v1, const float4& v2, int i) -> bool {
    float U, V, W;
    if (tri_test<false, KZ>(lr, tclosest, make_float3(v0.x, v0.y, v0.z), make_float3(v1.x, v1.y, v1.z), make_float3(v2.x, v2.y, v2.z), v2.w, &U, &V, &W)) {
      h.slot = base + i;
      if (DEFER) {
        h.prim = -2;
      } else {
        h.u = U; h.v = V; h.w = W;
        h.geom = __float_as_int(v0.w);
        h.prim = __float_as_int(v1.w);
      }
      leafhit = true;
      return true;
    }
    return false;
  };
  constexpr int PIPE = DEFER ? VG_LEAF_PIPE_OCCL : VG_LEAF_PIPE;
  if (PIPE != 1) {
    for (int i = 0; i < count; i++, tp += kTriStride) {
      float4 v0, v1, v2;
      ld_tri(tp, v0, v1, v2);
      if (PIPE == 2 && i + 1 < count) asm volatile("prefetch.global.L1 [%0];" ::"l"(tp + kTriStride));
      const bool hit = test(v0, v1, v2, i);
      if (VG_OCCL_LEAF_EXIT && DEFER && hit) break;
    }
    return leafhit;
  }
  float4 n0, n1, n2;
  ld_tri(tp, n0, n1, n2);
  for (int i = 0; i < count; i++) {
    const float4 v0 = n0, v1 = n1, v2 = n2;
    if (i + 1 < count) {
      tp += kTriStride;
      ld_tri(tp, n0, n1, n2);
    }
    const bool hit = test(v0, v1, v2, i);
    if (VG_OCCL_LEAF_EXIT && DEFER && hit) break;
  }
  return leafhit;
}

// trace.go:547-554: the three vertices of motion slot `slot` lerped between the keys of its mesh at `time`; *bias = the record's RayBias
__device__ __forceinline__ void motion_tri(const DevScene& sc, size_t slot, float time, float3& p0, float3& p1, float3& p2, float* bias) {
  const float4* k0 = sc.mtris + slot * 3;  // key-0 record of the slot: geom id in [0].w, RayBias in [2].w
  const DevGeom gm = sc.geoms[__float_as_int(ldg4(k0).w)];
  const float k = time * (float)(gm.keys - 1);
  const float fk = floorf(k);
  const float tm = k - fk, om = 1.0f - tm;
  const int key = (int)fk, key2 = (int)ceilf(k);
  const float4* ta = sc.mtris + (slot + (size_t)key * gm.tri_key_stride) * 3;
  const float4* tb = sc.mtris + (slot + (size_t)key2 * gm.tri_key_stride) * 3;
  const float4 a0 = ldg4(ta), a1 = ldg4(ta + 1), a2 = ldg4(ta + 2);
  const float4 b0 = ldg4(tb), b1 = ldg4(tb + 1), b2 = ldg4(tb + 2);
  p0 = make_float3(om * a0.x + tm * b0.x, om * a0.y + tm * b0.y, om * a0.z + tm * b0.z);
  p1 = make_float3(om * a1.x + tm * b1.x, om * a1.y + tm * b1.y, om * a1.z + tm * b1.z);
  p2 = make_float3(om * a2.x + tm * b2.x, om * a2.y + tm * b2.y, om * a2.z + tm * b2.z);
  *bias = a2.w;  // the w fields (geom id, face, RayBias) are the same in every key's record
}

template <int KZ, bool DEFER>
__device__ __forceinline__ bool leaf_motion(const DevScene& sc, const LeafRay& lr, float time, float& tclosest, HitState& h, int base, int count) {
  bool leafhit = false;
  // trace.go:547-554: lerp the three vertices between the keys of this mesh
  const float4 g0 = ldg4(sc.mtris + (size_t)base * 3);  // key-0 record of the first slot: geom id in w
  const DevGeom gm = sc.geoms[__float_as_int(g0.w)];
  const float k = time * (float)(gm.keys - 1);
  const float fk = floorf(k);
  const float tm = k - fk, om = 1.0f - tm;
  const int key = (int)fk, key2 = (int)ceilf(k);
  const float4* ta = sc.mtris + ((size_t)base + (size_t)key * gm.tri_key_stride) * 3;
  const float4* tb = sc.mtris + ((size_t)base + (size_t)key2 * gm.tri_key_stride) * 3;
  for (int i = 0; i < count; i++, ta += 3, tb += 3) {
    const float4 a0 = ldg4(ta), a1 = ldg4(ta + 1), a2 = ldg4(ta + 2);
    const float4 b0 = ldg4(tb), b1 = ldg4(tb + 1), b2 = ldg4(tb + 2);
    const float3 p0 = make_float3(om * a0.x + tm * b0.x, om * a0.y + tm * b0.y, om * a0.z + tm * b0.z);
    const float3 p1 = make_float3(om * a1.x + tm * b1.x, om * a1.y + tm * b1.y, om * a1.z + tm * b1.z);
    const float3 p2 = make_float3(om * a2.x + tm * b2.x, om * a2.y + tm * b2.y, om * a2.z + tm * b2.z);
    float U, V, W;
    if (tri_test<true, KZ>(lr, tclosest, p0, p1, p2, a2.w, &U, &V, &W)) {  // (geom id, face and RayBias in .w are the same in every key's record)
      h.slot = base + i;
      if (DEFER) {
        h.prim = -3;
      } else {
        h.u = U; h.v = V; h.w = W;
        h.geom = __float_as_int(a0.w);
        h.prim = __float_as_int(a1.w);
      }
      leafhit = true;
    }
  }
  return leafhit;
}

// One triangle leaf. Returns true if any triangle of the leaf was accepted. trace.go:116-194 / :528-667
// If every lane that arrived here together has the same dominant axis Kz (coherent camera / shadow rays), the
// component selection is resolved at compile time; otherwise per-lane selects.
// SMEM: the leaf parameters come from the lane's published block and an accepted triangle only records its slot (the occlusion
// kernel V = 4, compiled for 64 registers); otherwise they stay in RayState and the hit record is written at accept time (the
// closest-hit per-lane kernels at 72 registers: MEASURED, the published-block form cost them 17.85 -> 19.57 ms on the C2 camera rays —
// 4 KB more shared memory per CTA moves the carve-out from 64 to 100 KB, and a coherent warp re-reads its blocks at every leaf).
template <bool SMEM, bool MOTK = false>
__device__ __forceinline__ bool leaf_step(const DevScene& sc, TravState& t, uint32_t un, const CoopSmem& cs) {
  RayState& r = t.r;
  HitState& h = t.h;
  const int base = (int)((un >> 4) & kLeafBaseMask);
  const int count = (int)(un & 15u) + 1;
  h.cnt += (uint32_t)count << 16;
  LeafRay lr;
  if (SMEM) {
    lr = leaf_ray(cs, r.kz);
  } else {
    lr.pkx = r.pkx; lr.pky = r.pky; lr.pkz = r.pkz; lr.s0 = r.s0; lr.s1 = r.s1; lr.s2 = r.s2; lr.xsign = r.xsign; lr.kx = r.kx; lr.ky = r.ky; lr.kz = r.kz;
  }
  if (!MOTK && (un & kMotionTriBit)) return leaf_motion<-1, SMEM>(sc, lr, r.time, r.tclosest, h, base, count);
  const unsigned am = __activemask();
  const int k0 = __shfl_sync(am, r.kz, __ffs(am) - 1);
  if (MOTK) {  // kernels launched for scenes with motion meshes (VARIANT & 64): the motion leaf loop gets the compile-time axis too
    const bool mo = (un & kMotionTriBit) != 0;
    if (__all_sync(am, mo && r.kz == k0)) {
      if (k0 == 0) return leaf_motion<0, SMEM>(sc, lr, r.time, r.tclosest, h, base, count);
      if (k0 == 1) return leaf_motion<1, SMEM>(sc, lr, r.time, r.tclosest, h, base, count);
      return leaf_motion<2, SMEM>(sc, lr, r.time, r.tclosest, h, base, count);
    }
    if (mo) return leaf_motion<-1, SMEM>(sc, lr, r.time, r.tclosest, h, base, count);
    return leaf_static<-1, SMEM>(sc, lr, r.tclosest, h, base, count);
  }
  if (__all_sync(am, r.kz == k0)) {
    if (k0 == 0) return leaf_static<0, SMEM>(sc, lr, r.tclosest, h, base, count);
    if (k0 == 1) return leaf_static<1, SMEM>(sc, lr, r.tclosest, h, base, count);
    return leaf_static<2, SMEM>(sc, lr, r.tclosest, h, base, count);
  }
  return leaf_static<-1, SMEM>(sc, lr, r.tclosest, h, base, count);
}

// U, V, W of the accepted triangle recomputed from its slot (the same arithmetic on the same inputs as at accept time: the same bits),
// and the ids of its record. Called once per ray that hit, when the ray is stored.
__device__ __forceinline__ void hit_uvw_static(const DevScene& sc, const LeafRay& lr, HitState& h) {
  float4 v0, v1, v2;
  ld_tri(sc.tris + (size_t)h.slot * kTriStride, v0, v1, v2);
  float tc = __int_as_float(0x7f800000);
  tri_test<false, -1>(lr, tc, make_float3(v0.x, v0.y, v0.z), make_float3(v1.x, v1.y, v1.z), make_float3(v2.x, v2.y, v2.z), v2.w, &h.u, &h.v, &h.w);
  h.geom = __float_as_int(v0.w);
  h.prim = __float_as_int(v1.w);
}
__device__ __forceinline__ void hit_uvw_motion(const DevScene& sc, const LeafRay& lr, float time, HitState& h, bool ids) {
  float3 p0, p1, p2;
  float bias;
  motion_tri(sc, (size_t)h.slot, time, p0, p1, p2, &bias);
  float tc = __int_as_float(0x7f800000);
  tri_test<true, -1>(lr, tc, p0, p1, p2, bias, &h.u, &h.v, &h.w);
  if (ids) {
    const float4* tp = sc.mtris + (size_t)h.slot * 3;
    h.geom = __float_as_int(ldg4(tp).w);
    h.prim = __float_as_int(ldg4(tp + 1).w);
  }
}
__device__ __forceinline__ void finalize_hit(const DevScene& sc, const CoopSmem& cs, TravState& t) {
  if (t.h.prim == -2) hit_uvw_static(sc, leaf_ray(cs, t.r.kz), t.h);
  else if (t.h.prim == -3) hit_uvw_motion(sc, leaf_ray(cs, t.r.kz), t.r.time, t.h, true);
}

// Analytic sphere geom at the scene level (builtin/geom/sphere/trace.go:13-109): solveQuadratic / raySphereIntersect in the
// reference's float32 operation order (-fmad=false), accepted iff t < Tclosest. Rare (only sphere lights create these), so
// the arithmetic is kept out of line and takes/returns scalars only: the traversal state stays in registers and the hot
// loops pay one predicate for it. Returns the accepted t, or -1.
static __device__ __noinline__ float sphere_hit_t(float4 c0, float radius, float ox, float oy, float oz, float dx, float dy, float dz, float tclosest) {
  const float Lx = ox - c0.x, Ly = oy - c0.y, Lz = oz - c0.z;
  const float a = dx * dx + dy * dy + dz * dz;
  const float b = 2 * (dx * Lx + dy * Ly + dz * Lz);
  const float c = (Lx * Lx + Ly * Ly + Lz * Lz) - radius * radius;
  const float discr = b * b - 4 * a * c;
  if (discr < 0) return -1.0f;
  float x0, x1;
  if (discr == 0) {
    x1 = -0.5f * b / a;
    x0 = x1;
  } else {
    const float q = b > 0 ? -0.5f * (b + sqrtf(discr)) : -0.5f * (b - sqrtf(discr));
    x0 = q / a;
    x1 = c / q;
  }
  if (x0 > x1) { const float tmp = x0; x0 = x1; x1 = tmp; }
  if (x0 < 0) {
    x0 = x1;
    if (x0 < 0) return -1.0f;
  }
  if (!(x0 < tclosest)) return -1.0f;
  return x0;
}
__device__ __forceinline__ bool sphere_leaf(const DevScene& sc, TravState& t, uint32_t un) {
  const int slot = (int)(un & kLeafBaseMask);
  const float4 c0 = ldg4(sc.tris + (size_t)slot * kTriStride), c1 = ldg4(sc.tris + (size_t)slot * kTriStride + 1);
  const float th = sphere_hit_t(c0, c1.x, t.r.ox, t.r.oy, t.r.oz, t.r.dx, t.r.dy, t.r.dz, t.r.tclosest);
  if (th < 0.0f) return false;
  t.r.tclosest = th;
  t.h.u = t.h.v = t.h.w = 0.0f;
  t.h.geom = __float_as_int(c0.w);
  t.h.prim = 0;
  t.h.slot = slot;
  return true;
}

// while-while traversal of the lane's ray until it finishes or, if `min_active` > 0, until fewer than `min_active`
// lanes of the warp are still traversing (the caller then refills the idle lanes and comes back).
// Returns true when this lane's ray is finished.
// SPH: the scene holds analytic sphere geoms (kernels for scenes without them do not carry the call).
template <bool ANY_HIT, bool SPH = false, bool ORDERED = true, bool SMEM = false, bool MOTK = false>
__device__ __forceinline__ bool trav_run(const DevScene& sc, TravState& t, Stack& st, const CoopSmem& cs, int min_active) {
  while (t.cur != -1) {
    // 256-bit node loads in the per-lane kernel too: re-measured with pixel-coherent warps, C2 closest 19.09 vs 19.15 ms, C3 299.6 vs
    // 300.5 ms — within noise, so it keeps the 128-bit loads (round 1: 28.59 vs 28.32 ms).
#ifndef VG_PERLANE_WIDE
#define VG_PERLANE_WIDE 0
#endif
#ifndef VG_OCCL_WIDE
#define VG_OCCL_WIDE VG_PERLANE_WIDE
#endif
    while (t.cur >= 0) node_step<ORDERED, (SMEM ? VG_OCCL_WIDE : VG_PERLANE_WIDE) != 0>(sc, t, st);
    while (t.cur < -1) {
      const uint32_t un = (uint32_t)t.cur;
      if (un & kGeomBit) {
        if (SPH && (un & kSphereBit)) {  // scene.go:61-78 -> sphere.Trace
          if (sphere_leaf(sc, t, un) && ANY_HIT) {
            st.reset();
            t.cur = -1;
            return true;
          }
          t.cur = pop_next<ORDERED>(t.r, st);
          continue;
        }
        // scene.go:61-78 -> Geom.Trace -> qbvh.Trace pushes the mesh root with T = Tclosest and pops it at once
        t.cur = (int32_t)(un & kGeomRootMask);
        break;
      }
      const bool leafhit = leaf_step<SMEM, MOTK>(sc, t, un, cs);
      if (ANY_HIT && leafhit) {  // intersect.go:231-236: shadow rays return at the first leaf reporting a hit
        st.reset();
        t.cur = -1;
        return true;
      }
      t.cur = pop_next<ORDERED>(t.r, st);
    }
    if (min_active > 0 && __popc(__activemask()) < min_active) break;
  }
  return t.cur == -1;
}

// ---- TMA-staged ray queue ------------------------------------------------------------------------------------
// Each warp owns two 1-KB shared-memory slots of 32 VgRay records and one mbarrier per slot. A slot is filled by ONE
// bulk asynchronous copy (cp.async.bulk.shared.global -> SASS UBLKCP) issued by lane 0 and completed on the mbarrier, so
// the ~1 us HBM latency of the ray queue (the queues are streamed, never L2-resident) overlaps traversal of the rays
// the warp already holds; lanes that run dry take their next ray from shared memory.
struct WarpStage {
  float4* buf;         // 2 slots x 32 rays x 2 float4
  unsigned long long* bar;  // 2 mbarriers
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Persistent warp loop with refill (Aila & Laine style "dynamic fetch") over a TMA-staged queue: lanes whose ray finished
// take new rays as soon as fewer than VG_REFILL_BELOW lanes are busy, so a warp is never held by its slowest rays.
// IO: long long fetch(int count)  — claim `count` consecutive queue slots, returns the first (lane 0 calls it)
//     long long size();  const VgRay* ray_ptr()
//     void store(long long i, const RayState& r, const HitState& h, bool overflow)
// round 2, pixel-coherent warps (path_index): 16 / 24 / 30 -> C2 closest 19.08 / 19.16 / 19.43 ms, C3 closest 305.7 / 311.1 / 329.4 ms
#ifndef VG_REFILL_BELOW
#define VG_REFILL_BELOW 16
#endif
template <bool ANY_HIT, class IO>
__device__ __forceinline__ void trace_persistent_tma(const DevScene& sc, IO& io, Stack& st, WarpStage ws, const CoopSmem& cs, unsigned& nodes_acc,
                                                     unsigned& tris_acc) {
  const int lane = threadIdx.x & 31;
  const long long n = io.size();
  const char* rays = reinterpret_cast<const char*>(io.ray_ptr());
  TravState t;
  t.cur = -1;
  long long my = -1;  // queue index of the ray this lane is tracing
  st.reset();
  st.overflow = false;

  // warp-uniform staging state, kept in scalars (indexing small arrays by `cur` would push them to local memory)
  long long base0 = 0, base1 = 0;
  int cnt0 = 0, cnt1 = 0;
  uint32_t phase0 = 0, phase1 = 0;
  int cur = 0, off = 0;

  auto claim = [&](int b, long long& base_out, int& cnt_out) {  // claim the next 32 queue slots, start their copy into slot b
    long long bs = 0;
    if (lane == 0) bs = io.fetch(32);
    bs = __shfl_sync(0xffffffffu, bs, 0);
    long long c = n - bs;
    c = c < 0 ? 0 : (c > 32 ? 32 : c);
    base_out = bs;
    cnt_out = (int)c;
    if (c > 0 && lane == 0) {
      mbar_expect_tx(ws.bar + b, (uint32_t)c * 32u);
      bulk_g2s(ws.buf + b * 64, rays + bs * 32, (uint32_t)c * 32u, ws.bar + b);
    }
  };

  if (lane == 0) {
    mbar_init(ws.bar + 0, 1);
    mbar_init(ws.bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  claim(0, base0, cnt0);
  claim(1, base1, cnt1);

  while (true) {
#pragma unroll 1
    for (int rep = 0; rep < 2; rep++) {
      const unsigned idle = __ballot_sync(0xffffffffu, my < 0);
      const int ccnt = cur ? cnt1 : cnt0;
      if (idle == 0 || ccnt == 0) break;
      while (!mbar_try_wait(ws.bar + cur, cur ? phase1 : phase0)) {}
      const int avail = ccnt - off;
      const int want = __popc(idle);
      const int take = want < avail ? want : avail;
      const int rank = __popc(idle & ((1u << lane) - 1u));
      if (my < 0 && rank < take) {
        const float4* rp = ws.buf + cur * 64 + (off + rank) * 2;
        const float4 a = rp[0], b = rp[1];
        my = (cur ? base1 : base0) + off + rank;
        t.r.ox = a.x; t.r.oy = a.y; t.r.oz = a.z;
        t.r.dx = a.w; t.r.dy = b.x; t.r.dz = b.y;
        t.r.tclosest = b.z;
        t.r.time = b.w;
        ray_setup(t.r);
        trav_begin(sc, t, st);
      }
      off += take;
      if (off == ccnt) {  // slot consumed: recycle it for the chunk after next
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads before the next async write
        if (cur) { phase1 ^= 1u; claim(1, base1, cnt1); }
        else { phase0 ^= 1u; claim(0, base0, cnt0); }
        cur ^= 1;
        off = 0;
      }
    }
    if (__ballot_sync(0xffffffffu, my >= 0) == 0) break;  // nothing in flight and the queue is drained
    if (my >= 0) {
      if (trav_run<ANY_HIT>(sc, t, st, cs, (cur ? cnt1 : cnt0) == 0 ? 0 : VG_REFILL_BELOW)) {
        io.store(my, t.r, t.h, st.overflow);
        nodes_acc += t.h.cnt & 0xffffu;
        tris_acc += t.h.cnt >> 16;
        st.overflow = false;
        my = -1;
      }
    }
  }
}

// Default variant: the idle lanes claim exactly as many queue slots as they need and read their records with two
// coalesced LDG.128 each. Measured faster than the TMA-staged variant on B200 (3.83 vs 3.16 Grays/s on C2 primary rays,
// profiles/README.md): the fetch is <4 % of a ray's loads and 28 resident warps already hide its latency, while the
// staging state costs registers under the 72-register cap.
template <bool ANY_HIT, bool SPH, bool ORDERED, bool SMEM, bool MOTK, class IO>
__device__ __forceinline__ void trace_persistent_ldg(const DevScene& sc, IO& io, Stack& st, const CoopSmem& cs, unsigned& nodes_acc, unsigned& tris_acc) {
  const int lane = threadIdx.x & 31;
  const long long n = io.size();
  TravState t;
  t.cur = -1;
  long long my = -1;   // queue index of the ray this lane is tracing
  bool exhausted = false;
  st.reset();
  st.overflow = false;
  while (true) {
    const unsigned idle = __ballot_sync(0xffffffffu, my < 0);
    if (idle != 0 && !exhausted) {
      const int want = __popc(idle);
      long long base = 0;
      if (lane == 0) base = io.fetch(want);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + want >= n) exhausted = true;
      if (my < 0) {
        const long long i = base + __popc(idle & ((1u << lane) - 1u));
        if (i < n) {
          my = i;
          io.load(i, t.r);
          ray_setup(t.r);
          if (SMEM) coop_publish(cs, t.r);
          trav_begin(sc, t, st);
        }
      }
    }
    if (__ballot_sync(0xffffffffu, my >= 0) == 0) break;
    if (my >= 0) {
#ifndef VG_REFILL_BELOW_OCCL
#define VG_REFILL_BELOW_OCCL VG_REFILL_BELOW
#endif
      if (trav_run<ANY_HIT, SPH, ORDERED, SMEM, MOTK>(sc, t, st, cs, exhausted ? 0 : (SMEM ? VG_REFILL_BELOW_OCCL : VG_REFILL_BELOW))) {
        if (SMEM && IO::kHitRecord) finalize_hit(sc, cs, t);
        io.store(my, t.r, t.h, st.overflow);
        nodes_acc += t.h.cnt & 0xffffu;
        tris_acc += t.h.cnt >> 16;
        st.overflow = false;
        my = -1;
      }
    }
  }
}

// ---- Instances (builtin/geom/instance/instance.go:73-114) ------------------------------------------------------------
// Entering an instance leaf replaces the lane's ray by its object-space image (Matrix4MulPoint / MulVec with the inverse
// transform at Ray.Time, then Ray.Setup) and pushes a sentinel under the target mesh's root; popping the sentinel restores
// the world-space ray from the queue record. t is the same parameter in both spaces (D is not re-normalised), so Tclosest
// and the T keys of the stack entries pushed before entering stay valid. Rare and heavy (a 4x4 inverse, and a slerp when
// the transform has motion keys): kept out of line, scalars in, six floats out.
static __device__ __noinline__ void xf_object_ray(const DevScene& sc, int xi, float time, float ox, float oy, float oz, float dx, float dy, float dz,
                                                 float* out6) {
  out6[0] = ox; out6[1] = oy; out6[2] = oz;
  out6[3] = dx; out6[4] = dy; out6[5] = dz;
  // an instance of an instance applies the outer inverse, then the inner one, each on the result of the last — exactly the nested
  // Instance.Trace calls (instance.go:86-95); the Ray.Setup between them has no effect on P and D
  for (int hops = 0; xi >= 0 && hops < 8; hops++) {
    const DevXform x = sc.xforms[xi];
    Mat4 M, Minv;
    if (x.nkeys == 1) Minv = sc.xf_static[2 * xi + 1];
    else xf_matrices(sc.xf_keys + x.key_base, x.nkeys, time, &M, &Minv);
    float p[3], d[3];
    m4_mul_point(Minv, out6[0], out6[1], out6[2], p);
    m4_mul_vec(Minv, out6[3], out6[4], out6[5], d);
    out6[0] = p[0]; out6[1] = p[1]; out6[2] = p[2];
    out6[3] = d[0]; out6[4] = d[1]; out6[5] = d[2];
    xi = x.inner;
  }
}

// ---- Warp-cooperative leaves ------------------------------------------------------------------------------------
// The per-lane while-while loop above leaves most of a warp idle on incoherent rays: lanes sit at leaves of 1..16
// triangles while others still walk nodes, and inside the leaf phase every lane runs its own trip count (measured
// 8-15 of 32 lanes active per instruction). Here the node phase stays per lane, but the leaf phase is flattened:
// all (ray, triangle) pairs pending in the warp are enumerated with a warp prefix sum and tested 32 at a time by ALL
// lanes (also those that hold no leaf), reading consecutive 48-B triangle records. A lane that needs the node phase
// is therefore never idle during the leaf phase, and the node phase can stop as soon as fewer than VG_NODE_MIN lanes
// still want it.
//
// Exactness (polymesh/trace.go:116-194 is a sequential loop over the leaf): everything up to the sign/det/bias tests is
// independent of Tclosest; the only dependent test is `T^sign > Tclosest*|det|`. Since Tclosest only shrinks inside the
// loop and rounding is monotone, a triangle rejected against the leaf-entry Tclosest is rejected in the sequential
// loop too. So the parallel pass tests against the entry value and yields "candidates"; the owner lane then replays
// its candidates in triangle order against its live Tclosest (one shuffle round per candidate, almost always <= 1 per
// leaf), which reproduces the sequential accept/reject decisions and the final (T,U,V,W,idx) bit for bit.
// Node phase: keep stepping nodes until fewer than VG_NODE_MIN lanes want one (and a leaf is pending). Refill: fetch new rays
// once that many lanes are idle — occlusion rays (coherent in queue order) do best refilled in big groups, closest-hit bounce
// rays in smaller ones. Measured on C2 (shadow queue ms per frame / incoherent Mrays/s): NODE_MIN 8,4,2 -> 42.3, 41.8, 40.7(*);
// REFILL_IDLE 8,16,24,32 -> 42.3, 41.0, 39.9(*), 40.3 / 2110, 2144, 2107, - (* with NODE_MIN 4).
// Re-measured in round 2 with pixel-coherent warps: NODE_MIN 2 / 4 / 8 / 16 -> incoherent batch 2290 / 2360 / 2375 / 2289 Mrays/s, C3 closest
// 318.2 / 311.1 / 306.1 / 307.0 ms, C2 shadow 33.14 / 33.06 / 33.25 / 33.79 ms; REFILL_IDLE_ANYHIT 8 / 16 / 24 / 32 -> C2 shadow 35.28 / 33.80 /
// 33.06 / 33.94 ms, C3 shadow 311.9 / 288.6 / 279.7 / 302.8 ms.
#ifndef VG_NODE_MIN
#define VG_NODE_MIN 8
#endif
#ifndef VG_REFILL_IDLE_ANYHIT
#define VG_REFILL_IDLE_ANYHIT 24
#endif
#ifndef VG_REFILL_IDLE_CLOSEST
#define VG_REFILL_IDLE_CLOSEST 16
#endif
// ... and 24 for the integrator's closest-hit queues of levels >= 1 (IO::kRefillIdleClosest): measured C3 closest 277.5 -> 273.9 ms
// with 24, while the shuffled TraceProbe batch loses 2.3 % with it (2593 vs 2655 Mrays/s) and keeps 16.
#ifndef VG_REFILL_IDLE_CLOSEST_QUEUE
#define VG_REFILL_IDLE_CLOSEST_QUEUE 24
#endif

struct TriCand {
  float fU, fV, fW, det, T;
};
// The per-rank leaf record of a cooperative leaf phase: x = kz (bits 0-1) | motion leaf (bit 2) | owner lane (bits 3-7) | key count of
// the leaf's mesh (bits 8-15, motion leaves) | first item of the leaf in the phase's enumeration (bits 16-25); y = first triangle slot;
// z = the mesh's key stride in slots (motion leaves); w = Ray.Time. The leaf's lane looks the mesh up ONCE; round 2 (first half) had
// every (ray, triangle) item of a motion leaf repeat the two dependent loads (slot record -> geom record) before its own six.
__device__ __forceinline__ int rec_kz(int x) { return x & 3; }
__device__ __forceinline__ bool rec_motion(int x) { return (x & 4) != 0; }
__device__ __forceinline__ int rec_owner(int x) { return (x >> 3) & 31; }
__device__ __forceinline__ int rec_keys(int x) { return (x >> 8) & 255; }
__device__ __forceinline__ int rec_start(int x) { return (x >> 16) & 1023; }

// Tclosest-independent part of the watertight test + the entry-Tclosest cull. Returns "candidate".
template <bool MOTION, int KZ>
__device__ __forceinline__ bool tri_candidate(float pkx, float pky, float pkz, float s0, float s1, float s2, uint32_t xsign, int kzr, float tcl,
                                              float3 p0, float3 p1, float3 p2, float bias, TriCand& o) {
  const int kz = KZ < 0 ? kzr : KZ;
  const int kx = KZ < 0 ? (kzr == 2 ? 0 : kzr + 1) : (KZ + 1) % 3;
  const int ky = KZ < 0 ? (kx == 2 ? 0 : kx + 1) : (KZ + 2) % 3;
  const float AKz = pick<KZ>(p0.x, p0.y, p0.z, kz) - pkz;
  const float BKz = pick<KZ>(p1.x, p1.y, p1.z, kz) - pkz;
  const float CKz = pick<KZ>(p2.x, p2.y, p2.z, kz) - pkz;
  const float Cx = __uint_as_float(__float_as_uint((pick<KZ>(p2.x, p2.y, p2.z, kx) - pkx) - s0 * CKz) ^ xsign);
  const float By = (pick<KZ>(p1.x, p1.y, p1.z, ky) - pky) - s1 * BKz;
  const float Cy = (pick<KZ>(p2.x, p2.y, p2.z, ky) - pky) - s1 * CKz;
  const float Bx = __uint_as_float(__float_as_uint((pick<KZ>(p1.x, p1.y, p1.z, kx) - pkx) - s0 * BKz) ^ xsign);
  const float Ax = __uint_as_float(__float_as_uint((pick<KZ>(p0.x, p0.y, p0.z, kx) - pkx) - s0 * AKz) ^ xsign);
  const float Ay = (pick<KZ>(p0.x, p0.y, p0.z, ky) - pky) - s1 * AKz;
  float fU = Cx * By - Cy * Bx;
  float fV = Ax * Cy - Ay * Cx;
  float fW = Bx * Ay - By * Ax;
  if (fU == 0.0f || fV == 0.0f || fW == 0.0f) {
    double CxBy = (double)Cx * (double)By;
    double CyBx = (double)Cy * (double)Bx;
    fU = (float)(CxBy - CyBx);
    double AxCy = (double)Ax * (double)Cy;
    double AyCx = (double)Ay * (double)Cx;
    fV = (float)(AxCy - AyCx);
    double BxAy = (double)Bx * (double)Ay;
    double ByAx = (double)By * (double)Ax;
    fW = (float)(BxAy - ByAx);
  }
  if ((fU < 0.0f || fV < 0.0f || fW < 0.0f) && (fU > 0.0f || fV > 0.0f || fW > 0.0f)) return false;
  const float det = fU + fV + fW;
  if (det == 0.0f) return false;
  const float T = s2 * (fU * AKz + fV * BKz + fW * CKz);
  const uint32_t sgn = __float_as_uint(det) & 0x80000000u;
  const float Ts = __uint_as_float(__float_as_uint(T) ^ sgn);
  const float ds = __uint_as_float(__float_as_uint(det) ^ sgn);
  if (MOTION) {
    if (Ts <= bias * ds || Ts > tcl * ds) return false;
  } else {
    if (Ts < bias * ds || Ts > tcl * ds) return false;
  }
  o.fU = fU; o.fV = fV; o.fW = fW; o.det = det; o.T = T;
  return true;
}

// The (ray, triangle) items of one 32-wide window. KZ >= 0: every item's ray has dominant axis KZ (static component access).
template <int KZ>
__device__ __forceinline__ bool coop_item(const DevScene& sc, const float4 q0, const float4 q1, const float4 q2, int j, TriCand& tc) {
  const int i = j - rec_start(__float_as_int(q2.x));
  const float4* tp = sc.tris + (size_t)(__float_as_int(q2.y) + i) * kTriStride;
  float4 v0, v1, v2;
  ld_tri(tp, v0, v1, v2);
  return tri_candidate<false, KZ>(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, __float_as_uint(q1.z), rec_kz(__float_as_int(q2.x)), q1.w, make_float3(v0.x, v0.y, v0.z),
                                  make_float3(v1.x, v1.y, v1.z), make_float3(v2.x, v2.y, v2.z), v2.w, tc);
}

// A motion-triangle item (polymesh/trace.go:547-554,556-622): the two keys' vertices are lerped at the ray's time; the key count
// and key stride come from the item's geom record.
template <int KZ>
__device__ __forceinline__ bool coop_item_motion(const DevScene& sc, const float4 q0, const float4 q1, const float4 q2, int j, TriCand& tc) {
  const int x = __float_as_int(q2.x);
  const int i = j - rec_start(x);
  const size_t slot = (size_t)(__float_as_int(q2.y) + i);
  const size_t stride = (size_t)__float_as_int(q2.z);
  const float k = q2.w * (float)(rec_keys(x) - 1);
  const float fk = floorf(k);
  const float tm = k - fk, om = 1.0f - tm;
  const int key = (int)fk, key2 = (int)ceilf(k);
  const float4* ta = sc.mtris + (slot + (size_t)key * stride) * 3;
  const float4* tb = sc.mtris + (slot + (size_t)key2 * stride) * 3;
  const float4 a0 = ldg4(ta), a1 = ldg4(ta + 1), a2 = ldg4(ta + 2);
  const float4 b0 = ldg4(tb), b1 = ldg4(tb + 1), b2 = ldg4(tb + 2);
  const float3 p0 = make_float3(om * a0.x + tm * b0.x, om * a0.y + tm * b0.y, om * a0.z + tm * b0.z);
  const float3 p1 = make_float3(om * a1.x + tm * b1.x, om * a1.y + tm * b1.y, om * a1.z + tm * b1.z);
  const float3 p2 = make_float3(om * a2.x + tm * b2.x, om * a2.y + tm * b2.y, om * a2.z + tm * b2.z);
  return tri_candidate<true, KZ>(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, __float_as_uint(q1.z), rec_kz(x), q1.w, p0, p1, p2, a2.w, tc);  // the w fields (geom id, face, RayBias) are the same in every key's record
}

// One cooperative leaf phase. `isleaf`: this lane's t.cur is a triangle leaf that takes part (static; also motion leaves when
// MOT). Returns leafhit for the lane.
template <bool MOT, bool LIVE_T>
__device__ __forceinline__ bool coop_leaves(const DevScene& sc, TravState& t, bool isleaf, const CoopSmem& cs) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  RayState& r = t.r;
  HitState& h = t.h;
  const uint32_t un = (uint32_t)t.cur;
  const int count = isleaf ? (int)(un & 15u) + 1 : 0;
  const int base = (int)((un >> 4) & kLeafBaseMask);
  int incl = count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const int start = incl - count;
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  const uint32_t leafmask = __ballot_sync(0xffffffffu, isleaf);
  if (isleaf) {
    h.cnt += (uint32_t)count << 16;
    if (LIVE_T) reinterpret_cast<float*>(cs.rp + lane * 2 + 1)[3] = r.tclosest;  // (occlusion rays: Tclosest never changes before the ray ends)
    const bool mot = MOT && (un & kMotionTriBit);
    int keys = 0, stride = 0;
    if (mot) {  // polymesh/trace.go:79-84: the key count and key stride of the leaf's mesh, looked up once per leaf
      const DevGeom gm = sc.geoms[__float_as_int(ldg4(sc.mtris + (size_t)base * 3).w)];
      keys = gm.keys;
      stride = gm.tri_key_stride;
    }
    cs.rp[64 + __popc(leafmask & lt)] = make_float4(__int_as_float(r.kz | (mot ? 4 : 0) | (lane << 3) | (keys << 8) | (start << 16)), __int_as_float(base), __int_as_float(stride), r.time);
  }
  const int kz0 = __shfl_sync(0xffffffffu, r.kz, __ffs(leafmask) - 1);
  // all leaves of this phase static / all motion and one dominant axis: the item test is a template instance with static component
  // access (a phase that mixes static and motion leaves, or axes, takes the per-lane selects)
  const bool any_mot = MOT && __any_sync(0xffffffffu, isleaf && (un & kMotionTriBit));
  const bool kz_uniform = !any_mot && __all_sync(0xffffffffu, !isleaf || r.kz == kz0);
  bool kz_uniform_mot = false;
  if (MOT) kz_uniform_mot = __all_sync(0xffffffffu, !isleaf || ((un & kMotionTriBit) && r.kz == kz0));
  __syncwarp();
  bool leafhit = false;
  for (int w = 0; w < total; w += 32) {
    const int j = w + lane;
    // owner rank of item j = (#segments starting at or before j) - 1: segments that start before this window (warp-uniform)
    // plus the segment heads inside the window at slots <= lane (one REDUX.OR builds the head bitmask)
    const int f = start - w;
    const uint32_t heads = __reduce_or_sync(0xffffffffu, (isleaf && f >= 0 && f < 32) ? (1u << f) : 0u);
    const int before = __popc(__ballot_sync(0xffffffffu, isleaf && f < 0));
    bool cand = false;
    TriCand tc;
    tc.fU = tc.fV = tc.fW = tc.det = tc.T = 0.0f;
    if (j < total) {
      const int R = before + __popc(heads & (lt | (1u << lane))) - 1;
      const float4 q2 = cs.rp[64 + R];
      const float4* b = cs.rp + rec_owner(__float_as_int(q2.x)) * 2;
      const float4 q0 = b[0], q1 = b[1];
      if (kz_uniform) {
        if (kz0 == 0) cand = coop_item<0>(sc, q0, q1, q2, j, tc);
        else if (kz0 == 1) cand = coop_item<1>(sc, q0, q1, q2, j, tc);
        else cand = coop_item<2>(sc, q0, q1, q2, j, tc);
      } else if (MOT && kz_uniform_mot) {
        if (kz0 == 0) cand = coop_item_motion<0>(sc, q0, q1, q2, j, tc);
        else if (kz0 == 1) cand = coop_item_motion<1>(sc, q0, q1, q2, j, tc);
        else cand = coop_item_motion<2>(sc, q0, q1, q2, j, tc);
      } else if (MOT && rec_motion(__float_as_int(q2.x))) {
        cand = coop_item_motion<-1>(sc, q0, q1, q2, j, tc);
      } else {
        cand = coop_item<-1>(sc, q0, q1, q2, j, tc);
      }
    }
    const uint32_t cm = __ballot_sync(0xffffffffu, cand);
    if (cm == 0) continue;
    // this lane's items inside the window: item slots [start-w, start+count-w) clipped to [0,32)
    uint32_t mine = 0;
    if (isleaf) {
      const int lo = f, hi = f + count;
      if (hi > 0 && lo < 32) {
        const uint32_t mlo = lo <= 0 ? 0xffffffffu : (0xffffffffu << lo);
        const uint32_t mhi = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
        mine = cm & mlo & mhi;
      }
    }
    while (__any_sync(0xffffffffu, mine != 0)) {
      const int src = mine ? (__ffs(mine) - 1) : lane;
      const float fU = __shfl_sync(0xffffffffu, tc.fU, src);
      const float fV = __shfl_sync(0xffffffffu, tc.fV, src);
      const float fW = __shfl_sync(0xffffffffu, tc.fW, src);
      const float det = __shfl_sync(0xffffffffu, tc.det, src);
      const float T = __shfl_sync(0xffffffffu, tc.T, src);
      if (mine) {
        mine &= mine - 1;
        const uint32_t sgn = __float_as_uint(det) & 0x80000000u;
        const float Ts = __uint_as_float(__float_as_uint(T) ^ sgn);
        const float ds = __uint_as_float(__float_as_uint(det) ^ sgn);
        if (!(Ts > r.tclosest * ds)) {  // trace.go:182, the Tclosest half of the test against the live value
          const float rcp = 1.0f / det;
          h.u = fU * rcp;
          h.v = fV * rcp;
          h.w = fW * rcp;
          r.tclosest = T * rcp;
          h.slot = base + (src - f);
          h.prim = (MOT && (un & kMotionTriBit)) ? -3 : -2;  // geom / prim ids are read from the (motion) triangle record once, when the ray is stored
          leafhit = true;
        }
      }
    }
  }
  __syncwarp();  // the records are rewritten by the next leaf phase, the blocks by the next refill
  return leafhit;
}

template <bool ANY_HIT, bool ORDERED, bool SPH, bool XF, bool MOT, class IO>
__device__ __forceinline__ void trace_persistent_coop(const DevScene& sc, IO& io, Stack& st, const CoopSmem& cs, unsigned& nodes_acc,
                                                      unsigned& tris_acc) {
  const int lane = threadIdx.x & 31;
  const long long n = io.size();
  TravState t;
  t.cur = -1;
  long long my = -1;
  bool exhausted = false;
  int cur_xf = -1;  // XF: instance the lane's ray is currently inside of
  st.reset();
  st.overflow = false;
  while (true) {
    const unsigned idle = __ballot_sync(0xffffffffu, my < 0);
    if (!exhausted && __popc(idle) >= (ANY_HIT ? VG_REFILL_IDLE_ANYHIT : IO::kRefillIdleClosest)) {
      const int want = __popc(idle);
      long long base = 0;
      if (lane == 0) base = io.fetch(want);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + want >= n) exhausted = true;
      if (my < 0) {
        const long long i = base + __popc(idle & ((1u << lane) - 1u));
        if (i < n) {
          my = i;
          io.load(i, t.r);
          ray_setup(t.r);
          coop_publish(cs, t.r);
          trav_begin(sc, t, st);
        }
      }
    }
    if (__ballot_sync(0xffffffffu, my >= 0) == 0) break;
    // node phase: per lane, until fewer than VG_NODE_MIN lanes want it and at least one leaf is pending
    while (true) {
      while (t.cur < -1 && ((uint32_t)t.cur & kGeomBit)) {  // scene.go:61-78
        if (XF && ((uint32_t)t.cur & kXformBit)) {          // -> instance.Trace
          if ((uint32_t)t.cur & kSphereBit) {               // the sentinel: leave the instance (instance.go:98-105)
            const float tcl = t.r.tclosest;
            io.load(my, t.r);
            t.r.tclosest = tcl;
            ray_setup(t.r);
            coop_publish(cs, t.r);
            cur_xf = -1;
            t.cur = pop_next<ORDERED>(t.r, st);
          } else {                                          // enter (instance.go:86-96)
            const int xi = (int)((uint32_t)t.cur & kXformMask);
            st.push(__int_as_float(0xff800000), (int32_t)(kLeafBit | kGeomBit | kXformBit | kSphereBit | (uint32_t)xi));  // T = -Inf: never culled
            float o6[6];
            xf_object_ray(sc, xi, t.r.time, t.r.ox, t.r.oy, t.r.oz, t.r.dx, t.r.dy, t.r.dz, o6);
            t.r.ox = o6[0]; t.r.oy = o6[1]; t.r.oz = o6[2];
            t.r.dx = o6[3]; t.r.dy = o6[4]; t.r.dz = o6[5];
            ray_setup(t.r);
            coop_publish(cs, t.r);
            cur_xf = xi;
            t.cur = sc.xforms[xi].root;
          }
        } else if (SPH && ((uint32_t)t.cur & kSphereBit)) {  // -> sphere.Trace
          const bool sh = sphere_leaf(sc, t, (uint32_t)t.cur);
          if (XF && sh) t.h.xf_hit = -1;
          if (sh && ANY_HIT) {
            st.reset();
            t.cur = -1;
          } else {
            t.cur = pop_next<ORDERED>(t.r, st);
          }
        } else {
          t.cur = (int32_t)((uint32_t)t.cur & kGeomRootMask);  // -> mesh root next
        }
      }
      const unsigned nm = __ballot_sync(0xffffffffu, t.cur >= 0);
      if (nm == 0) break;
      if (__popc(nm) < VG_NODE_MIN && __any_sync(0xffffffffu, t.cur < -1)) break;
      // (MEASURED and removed, round 2: an occlusion-only lane that holds a triangle leaf while the node phase goes on trading it for the
      // interior node on top of its stack instead of idling — C2 shadow (cooperative kernel) 31.02 -> 32.04 ms, C3 248.2 -> 262.9 ms, C4 55.3 ->
      // 56.5 ms: the deferred leaves arrive in thinner leaf phases and occluded rays find their occluder later.)
      if (t.cur >= 0) node_step<ORDERED, VG_LDG256 != 0>(sc, t, st);
    }
    // leaf phase
    const bool leaf = t.cur < -1;
    const bool mleaf = leaf && ((uint32_t)t.cur & kMotionTriBit);
    if (__any_sync(0xffffffffu, leaf)) {
      bool leafhit = false;
// (A hybrid that ran the plain per-lane loop when >= 16 / 24 / 28 lanes of the warp sat in the SAME static leaf was MEASURED in
      // round 2 and removed: C2 shadow 34.25 / 34.24 / 34.28 vs 33.33 ms, C3 shadow 290.8 vs 280.3 ms, incoherent batch 2331 vs 2361 Mrays/s —
      // the second inlined leaf loop cost more in code size and registers across the vote than the bookkeeping it saved.)
      leafhit = coop_leaves<MOT, !ANY_HIT>(sc, t, leaf && (MOT || !mleaf), cs);
      if (!MOT && mleaf) {  // (kernels without the motion items in the cooperative phase: per lane, from the lane's published block)
        const uint32_t un = (uint32_t)t.cur;
        const int count = (int)(un & 15u) + 1;
        t.h.cnt += (uint32_t)count << 16;
        const LeafRay lr = leaf_ray(cs, t.r.kz);
        leafhit = leaf_motion<-1, false>(sc, lr, t.r.time, t.r.tclosest, t.h, (int)((un >> 4) & kLeafBaseMask), count);
      }
      if (leaf) {
        if (XF && leafhit) {
          t.h.xf_hit = cur_xf;
          if (cur_xf >= 0) t.h.xf_last = cur_xf;
        }
        if (ANY_HIT && leafhit) {  // intersect.go:231-236
          st.reset();
          t.cur = -1;
        } else {
          t.cur = pop_next<ORDERED>(t.r, st);
        }
      }
    }
    if (my >= 0 && t.cur == -1) {
      if (MOT && t.h.prim == -3) {
        const float4* tp = sc.mtris + (size_t)t.h.slot * 3;
        t.h.geom = __float_as_int(ldg4(tp).w);
        t.h.prim = __float_as_int(ldg4(tp + 1).w);
      } else if (t.h.prim == -2) {
        const float4* tp = sc.tris + (size_t)t.h.slot * kTriStride;
        t.h.geom = __float_as_int(ldg4(tp).w);
        t.h.prim = __float_as_int(ldg4(tp + 1).w);
      }
      if (XF && t.h.xf_hit >= 0 && t.h.prim >= 0) t.h.geom = sc.xforms[t.h.xf_hit].geom;  // scene.go:65: sc.Geom = the Instance
      if (XF) cur_xf = -1;
      io.store(my, t.r, t.h, st.overflow);
      nodes_acc += t.h.cnt & 0xffffu;
      tris_acc += t.h.cnt >> 16;
      st.overflow = false;
      my = -1;
    }
  }
}

// VARIANT & 7: 0 = per-lane while-while with coalesced LDG refill, 1 = the same over the TMA-staged queue, 2 = warp-cooperative
// leaves, 3 = cooperative leaves without the ordered push (occlusion-only any-hit rays), 4 = the per-lane loop without the ordered push. VARIANT & 8: the scene holds analytic
// sphere geoms (variants 0, 2, 3 only; the launchers map variant 1 to 0 for such scenes). VARIANT & 16: the scene holds
// instances (cooperative variants 2 and 3 only, always together with & 8).
template <bool ANY_HIT, int VARIANT, class IO>
__device__ __forceinline__ void trace_persistent(const DevScene& sc, IO& io, Stack& st, unsigned char* warp_smem, unsigned& nodes_acc,
                                                 unsigned& tris_acc) {
  constexpr int V = VARIANT & 7;
  constexpr bool SPH = (VARIANT & 8) != 0;
  constexpr bool XF = (VARIANT & 16) != 0;
  constexpr bool MOT = (VARIANT & 64) != 0 || XF;  // motion-triangle leaves take part in the cooperative leaf phase (variants 2, 3)
  CoopSmem cs;
  cs.rp = reinterpret_cast<float4*>(warp_smem);  // variants 2, 3, 4: the per-lane leaf-parameter blocks at [0, 1024)
  if (V == 1) {
    WarpStage ws;
    ws.buf = reinterpret_cast<float4*>(warp_smem);
    ws.bar = reinterpret_cast<unsigned long long*>(warp_smem + 2048);
    trace_persistent_tma<ANY_HIT>(sc, io, st, ws, cs, nodes_acc, tris_acc);
  } else if (V == 2) {
    trace_persistent_coop<ANY_HIT, true, SPH, XF, MOT>(sc, io, st, cs, nodes_acc, tris_acc);
  } else if (V == 3) {
    trace_persistent_coop<ANY_HIT, false, SPH, XF, MOT>(sc, io, st, cs, nodes_acc, tris_acc);
  } else if (V == 4) {  // per-lane loop without the ordered push: occlusion-only rays that are coherent (the integrator's level-0 shadow queue)
    static_assert(V != 4 || ANY_HIT, "the unordered per-lane loop is for occlusion-only rays");
    trace_persistent_ldg<ANY_HIT, SPH, false, true, (VARIANT & 64) != 0>(sc, io, st, cs, nodes_acc, tris_acc);
  } else {
    trace_persistent_ldg<ANY_HIT, SPH, true, false, false>(sc, io, st, cs, nodes_acc, tris_acc);
  }
}

}  // namespace vg
