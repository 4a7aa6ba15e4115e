// Host-visible launchers of the CUDA kernels (implemented in kernels_*.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vermeer_gpu.h"
#include "device_scene.h"

namespace vg {

#ifndef VG_TRACE_BLOCK
#define VG_TRACE_BLOCK 128
#endif
#ifndef VG_TRACE_MIN_BLOCKS
#define VG_TRACE_MIN_BLOCKS 7
#endif
// resident CTAs per SM the shadow (any-hit) queue kernels are compiled for; its own switch so that the two can be tuned apart
#ifndef VG_TRACE_MIN_BLOCKS_SHADOW
#define VG_TRACE_MIN_BLOCKS_SHADOW 8
#endif
// closest-hit queue kernels with the cooperative leaf phase (VARIANT & 2: every level after the camera rays)
// measured: static 8 vs 7 -> C3 closest 453.2 -> 441.8 ms; motion (VARIANT & 64) 8 vs 7 -> C4 closest 48.1 -> 49.0 ms, so it stays at 7
#ifndef VG_TRACE_MIN_BLOCKS_COOP
#define VG_TRACE_MIN_BLOCKS_COOP 8
#endif
#ifndef VG_TRACE_MIN_BLOCKS_COOP_MOTION
#define VG_TRACE_MIN_BLOCKS_COOP_MOTION VG_TRACE_MIN_BLOCKS
#endif
#ifndef VG_TRACE_MIN_BLOCKS_BATCH_COOP
#define VG_TRACE_MIN_BLOCKS_BATCH_COOP 8
#endif
#ifndef VG_TRACE_MIN_BLOCKS_SHADOW_MOTION
#define VG_TRACE_MIN_BLOCKS_SHADOW_MOTION VG_TRACE_MIN_BLOCKS_SHADOW
#endif
static const int kTraceBlock = VG_TRACE_BLOCK;
#ifndef VG_SMEM_STACK
#define VG_SMEM_STACK 8
#endif
// Dynamic shared memory of the traversal kernels: a per-warp scratch region, then the per-thread stacks. The scratch depends on
// the kernel variant (VARIANT & 7): none for the per-lane loop, 2 x 1 KB ray slots + 2 mbarriers for the TMA-staged queue,
// 32 x 48 B ray-parameter blocks for the cooperative leaf phase. Shared memory not taken here stays L1: the scene data these
// kernels re-read lives there, and the measured optimum is a SHORT shared stack (8 entries/thread: C2 frame 105.6 -> 102.9 ms
// against 16; 4 and 2 lose again to local-memory spills of the stack).
// Per warp: nothing for the per-lane loop (0); 2 x 1 KB ray slots + 2 mbarriers for the TMA-staged queue (1); 32 x 32 B per-lane
// leaf-parameter blocks (traverse.cuh: coop_publish) for the occlusion-only per-lane loop (4), + 32 x 16 B per-rank leaf records for
// the cooperative leaf phase (2, 3).
__host__ __device__ constexpr int warp_smem_bytes(int variant) {
  return (variant & 7) == 1 ? 2048 + 16 : (((variant & 7) == 2 || (variant & 7) == 3) ? 1024 + 32 * 16 : ((variant & 7) == 4 ? 1024 : 0));
}
inline size_t trace_smem_bytes(int variant) { return (size_t)(kTraceBlock / 32) * warp_smem_bytes(variant) + (size_t)kTraceBlock * VG_SMEM_STACK * 8; }

// kernels_trace.cu
cudaError_t launch_trace_batch(const DevScene& sc, const VgRay* d_rays, VgHit* d_hits, long long n, bool any_hit, int variant,
                               unsigned long long* d_counter, unsigned long long* d_stats, int grid, cudaStream_t stream, int mode = 0);  // mode: bit 0 = 16-byte compact hits, bit 1 = 24-byte {P, D} ray records
int trace_batch_blocks_per_sm();

}  // namespace vg
