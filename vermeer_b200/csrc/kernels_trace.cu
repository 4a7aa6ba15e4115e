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

// STREAMED (vg_trace_batch with page-locked host buffers): the ray array is still arriving over PCIe while the kernel runs. The
// host's copy stream raises `ready` (chunks landed so far) after every chunk with a stream memory operation; a lane whose claimed
// ray lies beyond the frontier waits for it. Finished rays count into `done[chunk]`, which the host's download stream waits on
// (cuStreamWaitValue32) before it copies that chunk's hits back: ONE persistent launch overlaps upload, traversal and download
// with no per-chunk launch ramp or tail.
struct StreamSync {
  const unsigned* ready;  // per chunk: non-zero once the chunk's rays are resident
  unsigned* done;         // per chunk: rays finished (hits stored and fenced)
  unsigned* err;          // set if the frontier did not move for kStreamTimeoutNs (host-side failure): the kernel must not hang
  int chunk_log2;
};
static const unsigned long long kStreamTimeoutNs = 4000000000ull;
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <bool STREAMED = false>
struct BatchIO {
  static constexpr bool kHitRecord = true;
  const VgRay* rays;
  VgHit* hits;
  long long n;
  unsigned long long* counter;
  int compact;  // bit 0 = VG_TRACE_COMPACT_HITS: 16-byte records {t, u, v, slot}; bit 1 = VG_TRACE_RAYS_PD: 24-byte {P, D} ray records
  __device__ __forceinline__ long long fetch(int c) { return (long long)atomicAdd(counter, (unsigned long long)c); }
  __device__ __forceinline__ long long size() const { return n; }
  __device__ __forceinline__ const VgRay* ray_ptr() const { return rays; }
  StreamSync ss;
  mutable unsigned ready_seen = 0;  // per thread: 1 + the chunk last seen resident (rays are claimed in order, so the flag is rarely re-read)
  __device__ __forceinline__ void wait_for(long long i) const {
    const unsigned c = (unsigned)(i >> ss.chunk_log2);
    if (ready_seen == c + 1u) return;
    const unsigned* flag = ss.ready + c;
    if (ld_relaxed_u32(flag) == 0u) {
      const unsigned long long t0 = globaltimer_ns();
      while (true) {
        __nanosleep(400);
        if (ld_relaxed_u32(flag) != 0u) break;
        if (globaltimer_ns() - t0 > kStreamTimeoutNs) {  // give up waiting: the call fails on the host, the kernel drains
          atomicExch(ss.err, 1u);
          break;
        }
      }
    }
    ready_seen = c + 1u;
  }
  __device__ __forceinline__ void load(long long i, RayState& r) const {
    if (STREAMED) {
      wait_for(i);
      // L2 loads (ld.global.cg): the copy engine wrote these lines after this kernel started
      if (compact & 2) {
        const float2* rp = reinterpret_cast<const float2*>(reinterpret_cast<const char*>(rays) + i * 24);
        const float2 a = __ldcg(rp), b = __ldcg(rp + 1), c = __ldcg(rp + 2);
        r.ox = a.x; r.oy = a.y; r.oz = b.x;
        r.dx = b.y; r.dy = c.x; r.dz = c.y;
        r.tclosest = __int_as_float(0x7f800000);
        r.time = 0.0f;
      } else {
        const float4* rp = reinterpret_cast<const float4*>(rays + i);
        const float4 a = __ldcg(rp), b = __ldcg(rp + 1);
        r.ox = a.x; r.oy = a.y; r.oz = a.z;
        r.dx = a.w; r.dy = b.x; r.dz = b.y;
        r.tclosest = b.z;
        r.time = b.w;
      }
      return;
    }
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
    if (STREAMED) {
      // Count the finished rays into their chunk: one release-reduction per group of lanes that finish together in the same chunk.
      // red.release.gpu = MEMBAR.ALL.GPU + REDG (the hits are visible before the count that releases the chunk's download) and, unlike
      // __threadfence() / fence.acq_rel (MEMBAR + CCTL.IVALL), it does not invalidate the L1 the node and triangle fetches live in.
      // The warp barrier orders the other lanes' hit stores before the leader's release (cumulativity).
      const unsigned am = __activemask();
      const int leader = __ffs(am) - 1;
      const unsigned c = (unsigned)(i >> ss.chunk_log2);
      const unsigned c0 = __shfl_sync(am, c, leader);
      const unsigned same = __ballot_sync(am, c == c0);
      __syncwarp(am);
      if (c != c0) red_release_add(ss.done + c, 1u);  // (a claim that straddles a chunk boundary)
      else if ((int)(threadIdx.x & 31) == leader) red_release_add(ss.done + c0, (unsigned)__popc(same));
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
  BatchIO<false> io{rays, hits, n, counter, compact};
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

// The same kernel over a ray array that is still being uploaded (BatchIO<true>). Static PolyMesh scenes, per-lane (0) and
// cooperative (2) variants.
template <bool ANY_HIT, int VARIANT>
__global__ void __launch_bounds__(kTraceBlock, (VARIANT & 2) ? VG_TRACE_MIN_BLOCKS_BATCH_COOP : VG_TRACE_MIN_BLOCKS) k_trace_batch_streamed(const DevScene sc, const VgRay* __restrict__ rays, VgHit* __restrict__ hits,
                                                             long long n, unsigned long long* __restrict__ counter,
                                                             unsigned long long* __restrict__ stats, int compact, const StreamSync ss) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5;
  Stack st;
  st.bind(smem_raw + nwarps * warp_smem_bytes(VARIANT));
  const int lane = threadIdx.x & 31;
  unsigned nodes_acc32 = 0, tris_acc32 = 0;
  BatchIO<true> io{rays, hits, n, counter, compact, ss};
  trace_persistent<ANY_HIT, VARIANT>(sc, io, st, smem_raw + warp * warp_smem_bytes(VARIANT), nodes_acc32, tris_acc32);
  unsigned long long nodes_acc = nodes_acc32, tris_acc = tris_acc32;
  for (int o = 16; o > 0; o >>= 1) {
    nodes_acc += __shfl_down_sync(0xffffffffu, nodes_acc, o);
    tris_acc += __shfl_down_sync(0xffffffffu, tris_acc, o);
  }
  if (lane == 0 && stats) {
    atomicAdd(stats + 0, nodes_acc);
    atomicAdd(stats + 1, tris_acc);
  }
}

cudaError_t launch_trace_batch_streamed(const DevScene& sc, const VgRay* d_rays, VgHit* d_hits, long long n, bool any_hit, int variant,
                                        unsigned long long* d_counter, unsigned long long* d_stats, int grid, cudaStream_t stream, int mode,
                                        const unsigned* d_ready, unsigned* d_done, unsigned* d_err, int chunk_log2) {
  cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return e;
  StreamSync ss{d_ready, d_done, d_err, chunk_log2};
#define VG_LAUNCH(A, V) k_trace_batch_streamed<A, V><<<grid, kTraceBlock, trace_smem_bytes(V), stream>>>(sc, d_rays, d_hits, n, d_counter, d_stats, mode, ss)
  if (any_hit) {
    if (variant == 2) VG_LAUNCH(true, 2);
    else VG_LAUNCH(true, 0);
  } else {
    if (variant == 2) VG_LAUNCH(false, 2);
    else VG_LAUNCH(false, 0);
  }
#undef VG_LAUNCH
  return cudaGetLastError();
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
