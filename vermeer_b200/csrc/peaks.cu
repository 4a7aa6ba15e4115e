// Measured read-bandwidth ceilings of the memory levels the traversal kernels are served from (vg_measure_peaks).
//
// The 1 M-triangle configs (C2, C4) keep their nodes and triangles in L2 and, to 70 %, in L1 (ncu: l1tex hit 69-71 %,
// lts hit 73-78 %), so the HBM peak in MEASURED_PEAKS.json is not the ceiling that can bind them. bench.py divides the
// kernels' measured L2 / L1 traffic by THESE numbers, measured on the same GPU in the same run, instead of by a guide
// figure. There is no reference counterpart: this file is measurement infrastructure only.
//
//   hbm   read-only stream over a buffer far larger than L2 (4 x 126 MB), ld.global.cg.v4 (no L1 allocation)
//   l2    the same kernel over a 48 MB buffer, re-read `passes` times: after the first pass every sector is an L2 hit
//   l1    every CTA re-reads its own 64 KB window (one CTA per SM, 1024 threads, ld.global.ca.v4): L1 hits only
//
// Each kernel keeps 8 independent 128-bit loads in flight per thread and folds them into one XOR so that nothing is
// eliminated; the result is written only if it has an impossible value.
#include <algorithm>

#include "context.h"

namespace vg {

__device__ __forceinline__ uint4 ld_cg(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_ca(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// grid-stride read of n uint4, `passes` times
__global__ void __launch_bounds__(512) k_peak_stream(const uint4* __restrict__ buf, size_t n, int passes, unsigned* sink) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int p = 0; p < passes; p++) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n; i += 8 * stride) {
      uint4 v[8];
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = ld_cg(buf + i + k * stride);
#pragma unroll
      for (int k = 0; k < 8; k++) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
    for (; i < n; i += stride) {
      const uint4 v = ld_cg(buf + i);
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (acc == 0x9e3779b9u) *sink = acc;
}

// every CTA re-reads its own window of `win` uint4 (win a multiple of 8 * blockDim.x)
__global__ void __launch_bounds__(1024) k_peak_l1(const uint4* __restrict__ buf, int win, int passes, unsigned* sink) {
  const uint4* w = buf + (size_t)blockIdx.x * win;
  unsigned acc = 0;
  for (int p = 0; p < passes; p++) {
    for (int i = threadIdx.x; i < win; i += 8 * blockDim.x) {
      uint4 v[8];
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = ld_ca(w + i + k * blockDim.x);
#pragma unroll
      for (int k = 0; k < 8; k++) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
  }
  if (acc == 0x9e3779b9u) *sink = acc;
}

}  // namespace vg

#define PCUDA(call)                                          \
  do {                                                       \
    cudaError_t e_ = (call);                                 \
    if (e_ != cudaSuccess) {                                 \
      if (buf) cudaFree(buf);                                \
      return ctx->cuda_fail(e_, #call);                      \
    }                                                        \
  } while (0)

extern "C" int vg_measure_peaks(vg_ctx* ctx, VgPeaks* out) {
  if (!ctx || !out) return VG_ERR_INVALID;
  std::lock_guard<std::mutex> lock(ctx->mu);
  uint4* buf = nullptr;
  PCUDA(cudaSetDevice(ctx->device));
  const size_t hbm_bytes = (size_t)512 << 20, l2_bytes = (size_t)48 << 20, l1_win_bytes = (size_t)64 << 10;
  PCUDA(cudaMalloc((void**)&buf, hbm_bytes + 16));
  PCUDA(cudaMemsetAsync(buf, 0x5a, hbm_bytes, ctx->stream));
  unsigned* sink = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(buf) + hbm_bytes);
  cudaEvent_t e0 = ctx->ev0, e1 = ctx->ev1;
  const int sms = ctx->sm_count;
  auto timed = [&](auto launch, double bytes, double* gbs) -> cudaError_t {
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0, ctx->stream);
      launch();
      cudaEventRecord(e1, ctx->stream);
      cudaError_t e = cudaEventSynchronize(e1);
      if (e != cudaSuccess) return e;
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms > 0) best = std::max(best, bytes / (ms * 1e-3) / 1e9);  // the first repetition warms the caches
    }
    *gbs = best;
    return cudaGetLastError();
  };
  const int hbm_passes = 4, l2_passes = 40, l1_passes = 400;
  PCUDA(timed([&] { vg::k_peak_stream<<<sms * 4, 512, 0, ctx->stream>>>(buf, hbm_bytes / 16, hbm_passes, sink); }, (double)hbm_bytes * hbm_passes, &out->hbm_read_gbs));
  PCUDA(timed([&] { vg::k_peak_stream<<<sms * 4, 512, 0, ctx->stream>>>(buf, l2_bytes / 16, l2_passes, sink); }, (double)l2_bytes * l2_passes, &out->l2_read_gbs));
  PCUDA(timed([&] { vg::k_peak_l1<<<sms, 1024, 0, ctx->stream>>>(buf, (int)(l1_win_bytes / 16), l1_passes, sink); }, (double)l1_win_bytes * sms * l1_passes, &out->l1_read_gbs));
  out->hbm_buffer_bytes = (double)hbm_bytes;
  out->l2_buffer_bytes = (double)l2_bytes;
  out->l1_window_bytes = (double)l1_win_bytes;
  out->sm_count = sms;
  cudaFree(buf);
  return VG_OK;
}
