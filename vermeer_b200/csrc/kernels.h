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
static const int kTraceBlock = VG_TRACE_BLOCK;
#ifndef VG_SMEM_STACK
#define VG_SMEM_STACK 16
#endif
// dynamic shared memory of the traversal kernels: a per-warp scratch region (TMA ray slots + mbarriers, or the ray-parameter
// blocks + owner table of the cooperative leaf phase), then the per-thread stacks
static const int kWarpSmemBytes = 2048 + 16;
inline size_t trace_smem_bytes() { return (size_t)(kTraceBlock / 32) * kWarpSmemBytes + (size_t)kTraceBlock * VG_SMEM_STACK * 8; }

// kernels_trace.cu
cudaError_t launch_trace_batch(const DevScene& sc, const VgRay* d_rays, VgHit* d_hits, long long n, bool any_hit, int variant,
                               unsigned long long* d_counter, unsigned long long* d_stats, int grid, cudaStream_t stream);
int trace_batch_blocks_per_sm();

}  // namespace vg
