// Batch TraceProbe kernels (core/trace.go:26) — persistent warps over a ray queue.
//
// Grid = (SMs x resident CTAs); every warp claims rays from the queue with one warp-aggregated atomicAdd and
// refills the lanes whose ray has finished (traverse.cuh: trace_persistent), so neither the kernel's tail nor a
// warp's slowest ray holds the others.
// The 32-byte VgRay / VgHit records are read and written as two 128-bit accesses each; consecutive
// lanes touch consecutive records, so a warp's fetch is one contiguous 1-KB burst.
#include "kernels.h"
#include "traverse.cuh"

namespace vg {

struct BatchIO {
  static constexpr bool kHitRecord = true;
  static constexpr int kRefillIdleClosest = VG_REFILL_IDLE_CLOSEST;
  const VgRay* rays;
  VgHit* hits;
  long long n;
  unsigned long long* counter;
  int compact;  // bit 0 = VG_TRACE_COMPACT_HITS: 16-byte records {t, u, v, slot}; bit 1 = VG_TRACE_RAYS_PD: 24-byte {P, D} ray records
  __device__ __forceinline__ long long fetch(int c) { return (long long)atomicAdd(counter, (unsigned long long)c); }
  __device__ __forceinline__ long long size() const { return n; }
  __device__ __forceinline__ const VgRay* ray_ptr() const { return rays; }
  __device__ __forceinline__ void load(long long i, RayState& r) const {
    if (compact & 2) {  // VgRayPD: Ray.Init(ty, P, D, +Inf, ...) at Time 0 (core/ray.go:56-65); 24-byte records are 8-byte aligned: 3 x LDG.64
      const float2* rp = reinterpret_cast<const float2*>(reinterpret_cast<const char*>(rays) + i * 24);
      const float2 a = __ldg(rp), b = __ldg(rp + 1), c = __ldg(rp + 2);
      r.ox = a.x; r.oy = a.y; r.oz = b.x;
      r.dx = b.y; r.dy = c.x; r.dz = c.y;
      r.tclosest = __int_as_float(0x7f800000);
      r.time = 0.0f;
      return;
    }
    const float4* rp = reinterpret_cast<const float4*>(rays + i);
    const float4 a = __ldg(rp), b = __ldg(rp + 1);
    r.ox = a.x; r.oy = a.y; r.oz = a.z;
    r.dx = a.w; r.dy = b.x; r.dz = b.y;
    r.tclosest = b.z;
    r.time = b.w;
  }
  __device__ __forceinline__ void store(long long i, const RayState& r, const HitState& h, bool overflow) const {
    if (compact & 1) {
      reinterpret_cast<float4*>(hits)[i] = make_float4(r.tclosest, h.u, h.v, __int_as_float(overflow ? -2 : (h.prim == -1 ? -1 : h.slot)));
    } else {
      float4* hp = reinterpret_cast<float4*>(hits + i);
      hp[0] = make_float4(r.tclosest, h.u, h.v, h.w);
      reinterpret_cast<int4*>(hp)[1] = make_int4(overflow ? -2 : h.prim, h.geom, (int)(h.cnt & 0xffffu), (int)(h.cnt >> 16));
    }
  }
};

template <bool ANY_HIT, int VARIANT>
// residency per kernel family as for the queue kernels (kernels.h); the grid is sized for the static cooperative variant, the
// default of vg_trace_batch: a persistent kernel that fits fewer CTAs just leaves the surplus ones to find the counter exhausted
__global__ void __launch_bounds__(kTraceBlock, ((VARIANT & 2) && !(VARIANT & 64)) ? VG_TRACE_MIN_BLOCKS_BATCH_COOP : VG_TRACE_MIN_BLOCKS) k_trace_batch(const DevScene sc, const VgRay* __restrict__ rays, VgHit* __restrict__ hits,
                                                             long long n, unsigned long long* __restrict__ counter,
                                                             unsigned long long* __restrict__ stats, int compact) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [warps x warp_smem_bytes(VARIANT) (TMA ray slots + mbarriers, or cooperative-leaf blocks)] [threads x VG_SMEM_STACK stack entries]
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5;
  Stack st;
  st.bind(smem_raw + nwarps * warp_smem_bytes(VARIANT));
  const int lane = threadIdx.x & 31;
  // per-thread sums of the packed per-ray counters (a persistent thread sees n / (SMs x CTAs x 128) rays: 32 bits are ample)
  unsigned nodes_acc32 = 0, tris_acc32 = 0;
  BatchIO io{rays, hits, n, counter, compact};
  trace_persistent<ANY_HIT, VARIANT>(sc, io, st, smem_raw + warp * warp_smem_bytes(VARIANT), nodes_acc32, tris_acc32);
  unsigned long long nodes_acc = nodes_acc32, tris_acc = tris_acc32;
  // warp-aggregated statistics (core/stats.go keeps global atomics per ray; one atomic per warp here)
  for (int o = 16; o > 0; o >>= 1) {
    nodes_acc += __shfl_down_sync(0xffffffffu, nodes_acc, o);
    tris_acc += __shfl_down_sync(0xffffffffu, tris_acc, o);
  }
  if (lane == 0 && stats) {
    atomicAdd(stats + 0, nodes_acc);
    atomicAdd(stats + 1, tris_acc);
  }
}

cudaError_t launch_trace_batch(const DevScene& sc, const VgRay* d_rays, VgHit* d_hits, long long n, bool any_hit, int variant,
                               unsigned long long* d_counter, unsigned long long* d_stats, int grid, cudaStream_t stream, int mode) {
  if ((mode & 2) && variant == 1) variant = 0;  // the TMA-staged queue copies 32-byte records; {P, D} records take the LDG refill
  cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
#define VG_LAUNCH(A, V) k_trace_batch<A, V><<<grid, kTraceBlock, trace_smem_bytes(V), stream>>>(sc, d_rays, d_hits, n, d_counter, d_stats, mode)
  if (sc.n_xforms > 0) {  // instances: cooperative kernels with the transform enter/leave code (traverse.cuh: VARIANT & 16)
    if (any_hit) VG_LAUNCH(true, 26);
    else VG_LAUNCH(false, 26);
  } else if (sc.n_spheres > 0) {  // kernels that carry the analytic sphere leaf (traverse.cuh: VARIANT & 8)
    if (any_hit) {
      if (variant == 2) VG_LAUNCH(true, 10);
      else VG_LAUNCH(true, 8);
    } else {
      if (variant == 2) VG_LAUNCH(false, 10);
      else VG_LAUNCH(false, 8);
    }
  } else if (sc.n_mtris > 0 && variant == 2) {  // motion triangles in the cooperative leaf phase (VARIANT & 64)
    if (any_hit) VG_LAUNCH(true, 66);
    else VG_LAUNCH(false, 66);
  } else if (any_hit) {
    if (variant == 1) VG_LAUNCH(true, 1);
    else if (variant == 2) VG_LAUNCH(true, 2);
    else VG_LAUNCH(true, 0);
  } else {
    if (variant == 1) VG_LAUNCH(false, 1);
    else if (variant == 2) VG_LAUNCH(false, 2);
    else VG_LAUNCH(false, 0);
  }
#undef VG_LAUNCH
  return cudaGetLastError();
}

int trace_batch_blocks_per_sm() {
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace_batch<false, 2>, kTraceBlock, trace_smem_bytes(2));
  return nb > 0 ? nb : 1;
}

}  // namespace vg
