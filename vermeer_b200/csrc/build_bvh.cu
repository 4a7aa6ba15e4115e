// qbvh.BuildAccel on the device: the reference's top-down binned-SAH build (qbvh/build.go:22-307), run
// level-synchronously over ALL open ranges of a level at once instead of one recursive goroutine.
//
// The reference's tree is a function of SETS, not of order: a binary split looks at the centroid bounds of its range (min/max),
// bins every primitive (`int(k1*(c-k0))`), unions boxes and counts per bin (min/max and integer adds), evaluates 7 candidate
// planes in a fixed order in float32, and partitions. None of that depends on the order of the primitives inside the range, so a
// build that keeps the same ranges reproduces the same splits, the same node boxes and the same leaf sets bit for bit. What
// does depend on order is the permutation INSIDE a leaf: the reference partitions in place with a two-pointer swap, this build
// with a stable scan-and-scatter, so the (at most 16) triangles of a leaf come out in a different order. Traversal visits the
// same nodes and tests the same triangles; only a tie between two equal-t hits inside one leaf can resolve differently.
// The exception is the reference's degenerate branch (all centroids of a range equal on the split axis: it returns len/2+1
// without partitioning, build.go:35-43), which is order dependent; it is kept (stable order) and bounded by the depth guard.
//
// Per round (one binary split of every open range; a 4-wide node is two rounds: the range, then its two halves):
//   k_seg_ids      primitive -> range (binary search over the sorted range table)
//   k_seg_bounds   centroid bounds per range: block reduction when a block sits inside one range, float atomics otherwise
//   k_seg_prepare  longest axis, flat test, k0/k1 (one thread per range)
//   k_seg_bin      8 bins per range: counts and boxes (shared-memory bins when a block sits inside one range)
//   k_seg_sah      the reference's prefix/suffix sweep and cost loop, one thread per range -> best plane, pivot
//   k_flags + scan + k_scatter   stable partition of every range at once (one global exclusive scan)
// The host keeps the (small) range tables, numbers the nodes in the reference's preorder and unions the boxes bottom-up.
#include <algorithm>
#include <cstring>
#include <limits>
#include <vector>

#include "context.h"

namespace vg {
namespace {

const int kBins = 8;  // build.go:22

struct SegState {  // one open range of the current round
  int lo, hi;
  int cb[6];        // centroid bounds as ordered ints: lo xyz, hi xyz
  int axis, flat;
  float k0, k1;
  int cnt[kBins];
  int bb[kBins][6];  // bin boxes as ordered ints
  int best, pivot;
  int scan_lo;       // exclusive scan value at lo
  int pad;
};

// monotonic float <-> int mapping for atomicMin/atomicMax on floats
__host__ __device__ inline int f2o(float f) {
  int i;
#ifdef __CUDA_ARCH__
  i = __float_as_int(f);
#else
  std::memcpy(&i, &f, 4);
#endif
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ inline float o2f(int i) {
  i = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
  return __int_as_float(i);
#else
  float f;
  std::memcpy(&f, &i, 4);
  return f;
#endif
}

struct Prim {  // 48 B per primitive, moved by every partition
  float4 lo;   // box min, w unused
  float4 hi;   // box max
  float4 c;    // centroid, w = original index (int bits)
};

__device__ inline int find_seg(const SegState* segs, int nseg, int i) {
  int lo = 0, hi = nseg - 1, ans = -1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (segs[mid].lo <= i) { ans = mid; lo = mid + 1; }
    else hi = mid - 1;
  }
  if (ans >= 0 && i < segs[ans].hi) return ans;
  return -1;
}

__global__ void __launch_bounds__(256) k_seg_ids(const SegState* segs, int nseg, int n, int* seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) seg[i] = find_seg(segs, nseg, i);
}

__global__ void __launch_bounds__(256) k_seg_init(SegState* segs, int nseg, const int2* ranges) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  SegState& S = segs[s];
  S.lo = ranges[s].x;
  S.hi = ranges[s].y;
  const int pinf = f2o(__int_as_float(0x7f800000)), ninf = f2o(__int_as_float(0xff800000));
  for (int a = 0; a < 3; a++) { S.cb[a] = pinf; S.cb[3 + a] = ninf; }
  for (int b = 0; b < kBins; b++) {
    S.cnt[b] = 0;
    for (int a = 0; a < 3; a++) { S.bb[b][a] = pinf; S.bb[b][3 + a] = ninf; }
  }
  S.best = -1;
  S.pivot = S.lo;
}

__device__ inline float warp_min(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ inline float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// build.go:151-165: centroid bounds of every open range
__global__ void __launch_bounds__(256) k_seg_bounds(SegState* segs, const int* seg, const Prim* prims, int n) {
  __shared__ float red[6][8];
  __shared__ int s_first, s_last;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = i < n ? seg[i] : -1;
  if (threadIdx.x == 0) s_first = s;
  if (threadIdx.x == blockDim.x - 1) s_last = s;
  __syncthreads();
  const float inf = __int_as_float(0x7f800000);
  float c[3] = {inf, inf, inf};
  if (s >= 0) { const float4 v = prims[i].c; c[0] = v.x; c[1] = v.y; c[2] = v.z; }
  if (s_first >= 0 && s_first == s_last) {
    // ranges are contiguous: first == last means the whole block is inside one range
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int a = 0; a < 3; a++) {
      const float mn = warp_min(c[a]), mx = warp_max(s >= 0 ? c[a] : -inf);
      if (lane == 0) { red[a][w] = mn; red[3 + a][w] = mx; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      float v = red[threadIdx.x][0];
      for (int k = 1; k < 8; k++) v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][k]) : fmaxf(v, red[threadIdx.x][k]);
      if (threadIdx.x < 3) atomicMin(&segs[s_first].cb[threadIdx.x], f2o(v));
      else atomicMax(&segs[s_first].cb[threadIdx.x], f2o(v));
    }
  } else if (s >= 0) {
    for (int a = 0; a < 3; a++) {
      atomicMin(&segs[s].cb[a], f2o(c[a]));
      atomicMax(&segs[s].cb[3 + a], f2o(c[a]));
    }
  }
}

// build.go:28-47: MaxDim of the centroid bounds, the flat test, and the bin scale
__global__ void __launch_bounds__(256) k_seg_prepare(SegState* segs, int nseg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  SegState& S = segs[s];
  float lo[3], hi[3];
  for (int a = 0; a < 3; a++) { lo[a] = o2f(S.cb[a]); hi[a] = o2f(S.cb[3 + a]); }
  const float d0 = hi[0] - lo[0], d1 = hi[1] - lo[1], d2 = hi[2] - lo[2];
  int axis;  // math/boundingbox.go MaxDim, same comparisons as the host builder
  if (d0 < d1) axis = d1 < d2 ? 2 : 1;
  else axis = d0 < d2 ? 2 : 0;
  S.axis = axis;
  S.flat = hi[axis] == lo[axis] ? 1 : 0;
  S.k0 = lo[axis];
  S.k1 = S.flat ? 0.0f : (float)kBins * (float)(1.0 - 0.00006) / (hi[axis] - lo[axis]);
}

__device__ inline int prim_bin(const SegState& S, const float4 c) {
  const float v = S.axis == 0 ? c.x : (S.axis == 1 ? c.y : c.z);
  return (int)(S.k1 * (v - S.k0));
}

// build.go:49-66: bin counts and bin boxes
__global__ void __launch_bounds__(256) k_seg_bin(SegState* segs, const int* seg, const Prim* prims, int n, int* err) {
  __shared__ int sh_cnt[kBins];
  __shared__ int sh_bb[kBins][6];
  __shared__ int s_first, s_last;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = i < n ? seg[i] : -1;
  if (threadIdx.x == 0) s_first = s;
  if (threadIdx.x == blockDim.x - 1) s_last = s;
  const int pinf = f2o(__int_as_float(0x7f800000)), ninf = f2o(__int_as_float(0xff800000));
  if (threadIdx.x < kBins) {
    sh_cnt[threadIdx.x] = 0;
    for (int a = 0; a < 3; a++) { sh_bb[threadIdx.x][a] = pinf; sh_bb[threadIdx.x][3 + a] = ninf; }
  }
  __syncthreads();
  const bool uniform = s_first >= 0 && s_first == s_last;
  int bin = -1;
  Prim p;
  if (s >= 0 && !segs[s].flat) {
    p = prims[i];
    bin = prim_bin(segs[s], p.c);
    if (bin < 0 || bin > kBins - 1) {  // non-finite centroid: the reference indexes out of range and panics
      atomicOr(err, 1);
      bin = -1;
    }
  }
  if (bin >= 0) {
    int* cnt = uniform ? sh_cnt : segs[s].cnt;
    int(*bb)[6] = uniform ? sh_bb : segs[s].bb;
    atomicAdd(&cnt[bin], 1);
    atomicMin(&bb[bin][0], f2o(p.lo.x)); atomicMin(&bb[bin][1], f2o(p.lo.y)); atomicMin(&bb[bin][2], f2o(p.lo.z));
    atomicMax(&bb[bin][3], f2o(p.hi.x)); atomicMax(&bb[bin][4], f2o(p.hi.y)); atomicMax(&bb[bin][5], f2o(p.hi.z));
  }
  __syncthreads();
  if (uniform && threadIdx.x < kBins && sh_cnt[threadIdx.x] > 0) {
    SegState& S = segs[s_first];
    const int b = threadIdx.x;
    atomicAdd(&S.cnt[b], sh_cnt[b]);
    for (int a = 0; a < 3; a++) {
      atomicMin(&S.bb[b][a], sh_bb[b][a]);
      atomicMax(&S.bb[b][3 + a], sh_bb[b][3 + a]);
    }
  }
}

struct FBox {
  float lo[3], hi[3];
  __device__ void reset() {
    for (int a = 0; a < 3; a++) { lo[a] = __int_as_float(0x7f800000); hi[a] = __int_as_float(0xff800000); }
  }
  // math/boundingbox.go GrowBox with the x86 MINSS/MAXSS operand order of the host builder (hmath.h: fmin_x86(lo, p.lo))
  __device__ void grow(const FBox& p) {
    for (int a = 0; a < 3; a++) {
      lo[a] = lo[a] < p.lo[a] ? lo[a] : p.lo[a];
      hi[a] = hi[a] > p.hi[a] ? hi[a] : p.hi[a];
    }
  }
  __device__ float dim(int a) const { return hi[a] - lo[a]; }
  __device__ float area() const { return dim(0) * dim(1) * 2.0f + dim(1) * dim(2) * 2 + dim(0) * dim(2) * 2; }
};

// build.go:68-107: left/right sweeps and the cost loop, in the reference's order; then the size of the left side
__global__ void __launch_bounds__(128) k_seg_sah(SegState* segs, int nseg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  SegState& S = segs[s];
  const int n = S.hi - S.lo;
  if (S.flat) {  // build.go:35-43: len/2+1, no partition
    S.best = -2;
    S.pivot = S.lo + (n / 2 + 1);
    return;
  }
  FBox bb[kBins], lbox[kBins], rbox[kBins];
  int ln[kBins], rn[kBins];
  for (int b = 0; b < kBins; b++)
    for (int a = 0; a < 3; a++) { bb[b].lo[a] = o2f(S.bb[b][a]); bb[b].hi[a] = o2f(S.bb[b][3 + a]); }
  FBox acc;
  acc.reset();
  int cnt = 0;
  for (int i = 0; i < kBins; i++) { acc.grow(bb[i]); cnt += S.cnt[i]; lbox[i] = acc; ln[i] = cnt; }
  acc.reset();
  cnt = 0;
  for (int i = kBins - 1; i >= 0; i--) { acc.grow(bb[i]); cnt += S.cnt[i]; rbox[i] = acc; rn[i] = cnt; }
  int best = -1;
  float best_cost = __int_as_float(0x7f800000);
  for (int i = 1; i < kBins; i++) {
    const float cost = lbox[i - 1].area() * (float)ln[i - 1] + rbox[i].area() * (float)rn[i];
    if (cost < best_cost) { best = i; best_cost = cost; }
  }
  S.best = best;
  S.pivot = S.lo + (best >= 1 ? ln[best - 1] : 0);
}
__global__ void __launch_bounds__(256) k_seg_results(const SegState* segs, int nseg, int2* out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg) out[s] = make_int2(segs[s].axis, segs[s].pivot);
}

// build.go:109-134: which side each primitive goes to (the reference recomputes the bin here too)
__global__ void __launch_bounds__(256) k_flags(const SegState* segs, const int* seg, const Prim* prims, int n, int* flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = seg[i];
  int f = 0;
  if (s >= 0) {
    const SegState& S = segs[s];
    if (S.flat) f = (i - S.lo) < (S.pivot - S.lo) ? 1 : 0;
    else f = prim_bin(S, prims[i].c) < S.best ? 1 : 0;
  }
  flag[i] = f;
}

// ---- exclusive scan of ints: 1024 elements per block, block sums scanned recursively --------------------------------------
__global__ void __launch_bounds__(256) k_scan_block(const int* in, int* out, int n, int* block_sums) {
  __shared__ int warp_sums[8];
  const int base = blockIdx.x * 1024 + threadIdx.x * 4;
  int v[4];
  for (int k = 0; k < 4; k++) v[k] = base + k < n ? in[base + k] : 0;
  const int t = v[0] + v[1] + v[2] + v[3];
  int incl = t;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) warp_sums[w] = incl;
  __syncthreads();
  int woff = 0;
  for (int k = 0; k < w; k++) woff += warp_sums[k];
  int run = woff + incl - t;
  for (int k = 0; k < 4; k++) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == 255 && block_sums) block_sums[blockIdx.x] = run;
}
__global__ void __launch_bounds__(256) k_scan_add(int* out, int n, const int* block_offs) {
  const int i = blockIdx.x * 1024 + threadIdx.x * 4;
  const int off = block_offs[blockIdx.x];
  for (int k = 0; k < 4; k++)
    if (i + k < n) out[i + k] += off;
}

struct Scan {
  std::vector<DevBuf<int>> sums;  // one per recursion level
  cudaError_t run(const int* in, int* out, int n, cudaStream_t st, int level = 0) {
    const int nb = (n + 1023) / 1024;
    if ((int)sums.size() <= level) sums.resize(level + 1);
    cudaError_t e = sums[level].reserve((size_t)nb);
    if (e != cudaSuccess) return e;
    k_scan_block<<<nb, 256, 0, st>>>(in, out, n, nb > 1 ? sums[level].p : nullptr);
    if (nb > 1) {
      e = run(sums[level].p, sums[level].p, nb, st, level + 1);  // in place: each block reads its inputs before it writes
      if (e != cudaSuccess) return e;
      k_scan_add<<<nb, 256, 0, st>>>(out, n, sums[level].p);
    }
    return cudaGetLastError();
  }
  void release() {
    for (auto& b : sums) b.release();
  }
};

__global__ void __launch_bounds__(128) k_seg_scan_lo(SegState* segs, int nseg, const int* scan) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg) segs[s].scan_lo = scan[segs[s].lo];
}

// stable partition of every open range at once; primitives outside open ranges stay where they are
__global__ void __launch_bounds__(256) k_scatter(const SegState* segs, const int* seg, const int* flag, const int* scan, const Prim* src, Prim* dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = seg[i];
  int d = i;
  if (s >= 0) {
    const SegState& S = segs[s];
    const int left_rank = scan[i] - S.scan_lo;
    d = flag[i] ? S.lo + left_rank : S.pivot + ((i - S.lo) - left_rank);
  }
  dst[d] = src[i];
}

__global__ void __launch_bounds__(256) k_load_prims(const float* boxes, const float* cent, int n, Prim* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Prim p;
  p.lo = make_float4(boxes[i * 6 + 0], boxes[i * 6 + 1], boxes[i * 6 + 2], 0.f);
  p.hi = make_float4(boxes[i * 6 + 3], boxes[i * 6 + 4], boxes[i * 6 + 5], 0.f);
  p.c = make_float4(cent[i * 3 + 0], cent[i * 3 + 1], cent[i * 3 + 2], __int_as_float(i));
  out[i] = p;
}

// union of the primitive boxes of every leaf range (build.go:138-149 calcBox); also the permutation
__global__ void __launch_bounds__(128) k_leaf_boxes(const int2* leaves, int nleaf, const Prim* prims, float* out6) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nleaf) return;
  FBox b;
  b.reset();
  for (int i = leaves[l].x; i < leaves[l].y; i++) {
    FBox p;
    const float4 lo = prims[i].lo, hi = prims[i].hi;
    p.lo[0] = lo.x; p.lo[1] = lo.y; p.lo[2] = lo.z;
    p.hi[0] = hi.x; p.hi[1] = hi.y; p.hi[2] = hi.z;
    b.grow(p);
  }
  for (int a = 0; a < 3; a++) { out6[l * 6 + a] = b.lo[a]; out6[l * 6 + 3 + a] = b.hi[a]; }
}
__global__ void __launch_bounds__(256) k_store_idx(const Prim* prims, int n, int32_t* idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = __float_as_int(prims[i].c.w);
}

}  // namespace

// Device and page-locked scratch of the builder, kept in the context between builds (cudaMalloc / cudaFree per build cost more
// than the build itself).
struct BuildScratch {
  DevBuf<float> d_boxes, d_cent, d_leafbox;
  DevBuf<Prim> d_prims[2];
  DevBuf<int> d_seg, d_flag, d_scan, d_err;
  DevBuf<SegState> d_segs;
  DevBuf<int2> d_leaves, d_ranges, d_results;
  DevBuf<int32_t> d_idx;
  Scan scan;
  int2* h_pinned = nullptr;  // ranges up / results down
  size_t h_pinned_cap = 0;
  cudaError_t pinned(size_t n) {
    if (n <= h_pinned_cap) return cudaSuccess;
    if (h_pinned) cudaFreeHost(h_pinned);
    h_pinned = nullptr;
    h_pinned_cap = 0;
    const size_t cap = std::max<size_t>(n, 4096);
    cudaError_t e = cudaMallocHost((void**)&h_pinned, cap * sizeof(int2));
    if (e == cudaSuccess) h_pinned_cap = cap;
    return e;
  }
  void release() {
    d_boxes.release(); d_cent.release(); d_leafbox.release(); d_prims[0].release(); d_prims[1].release();
    d_seg.release(); d_flag.release(); d_scan.release(); d_err.release(); d_segs.release(); d_leaves.release();
    d_ranges.release(); d_results.release(); d_idx.release();
    scan.release();
    if (h_pinned) cudaFreeHost(h_pinned);
    h_pinned = nullptr;
    h_pinned_cap = 0;
  }
};
void build_scratch_destroy(vg_ctx* ctx) {
  if (ctx->build_scratch) {
    ctx->build_scratch->release();
    delete ctx->build_scratch;
    ctx->build_scratch = nullptr;
  }
}

namespace {

struct QNode {  // one 4-wide node while the build runs (breadth-first numbering)
  int lo, hi;
  int a[3];
  int p[3];       // absolute pivots: top, left half, right half
  int child[4];   // breadth-first index of the child node, or -1 (leaf / empty)
  int leaf[4];    // index into the leaf-range list, or -1
  float box[4][6];
};

}  // namespace

// Returns VG_OK or an error code with ctx->err set. Launch count is added to ctx->stats.kernel_launches.
int build_qbvh_device(vg_ctx* ctx, const float* boxes, const float* cent, int n, int leaf_max, std::vector<VgNode>& out, int32_t* idx_out,
                      float* bounds6) {
#define BCUDA(call)                                                      \
  do {                                                                   \
    cudaError_t e_ = (call);                                             \
    if (e_ != cudaSuccess) { rc = ctx->cuda_fail(e_, #call); goto done; } \
  } while (0)
  int rc = VG_OK;
  if (leaf_max > 16) leaf_max = 16;
  if (leaf_max < 1) leaf_max = 1;
  cudaStream_t st = ctx->stream;
  if (!ctx->build_scratch) ctx->build_scratch = new BuildScratch();
  BuildScratch& B = *ctx->build_scratch;
  DevBuf<float>&d_boxes = B.d_boxes, &d_cent = B.d_cent, &d_leafbox = B.d_leafbox;
  DevBuf<Prim>* d_prims = B.d_prims;
  DevBuf<int>&d_seg = B.d_seg, &d_flag = B.d_flag, &d_scan = B.d_scan, &d_err = B.d_err;
  DevBuf<SegState>& d_segs = B.d_segs;
  DevBuf<int2>& d_leaves = B.d_leaves;
  DevBuf<int32_t>& d_idx = B.d_idx;
  Scan& scan = B.scan;
  std::vector<QNode> nodes;
  std::vector<int2> leaves;
  std::vector<int> act;
  uint64_t launches = 0;
  int cur = 0;
  const int gridN = (n + 255) / 256;

  // one round: binary-split every range of `ranges` (lo, hi); ranges with n <= leaf_max answer (axis 0, pivot = hi) at once
  // (build.go:152-154). Returns axis / pivot per range.
  struct Split { int axis, pivot; };
  std::vector<Split> result;
  auto split_round = [&](const std::vector<int2>& ranges) -> int {
    result.assign(ranges.size(), Split{0, 0});
    std::vector<int> where;
    size_t T0 = 0;
    for (size_t r = 0; r < ranges.size(); r++)
      if (ranges[r].y - ranges[r].x > leaf_max) T0++;
    if (T0 == 0) {
      for (size_t r = 0; r < ranges.size(); r++) result[r] = Split{0, ranges[r].y};
      return VG_OK;
    }
    cudaError_t e;
    if ((e = B.pinned(T0)) != cudaSuccess) return ctx->cuda_fail(e, "build: pinned tables");
    for (size_t r = 0; r < ranges.size(); r++) {
      const int cn = ranges[r].y - ranges[r].x;
      if (cn <= leaf_max) { result[r] = Split{0, ranges[r].y}; continue; }
      B.h_pinned[where.size()] = ranges[r];
      where.push_back((int)r);
    }
    const int T = (int)where.size();
    if ((e = d_segs.reserve((size_t)T)) != cudaSuccess || (e = B.d_ranges.reserve((size_t)T)) != cudaSuccess ||
        (e = B.d_results.reserve((size_t)T)) != cudaSuccess)
      return ctx->cuda_fail(e, "build: range tables");
    if ((e = cudaMemcpyAsync(B.d_ranges.p, B.h_pinned, (size_t)T * sizeof(int2), cudaMemcpyHostToDevice, st)) != cudaSuccess) return ctx->cuda_fail(e, "build: range table copy");
    const int gridT = (T + 127) / 128;
    k_seg_init<<<(T + 255) / 256, 256, 0, st>>>(d_segs.p, T, B.d_ranges.p);
    k_seg_ids<<<gridN, 256, 0, st>>>(d_segs.p, T, n, d_seg.p);
    k_seg_bounds<<<gridN, 256, 0, st>>>(d_segs.p, d_seg.p, d_prims[cur].p, n);
    k_seg_prepare<<<(T + 255) / 256, 256, 0, st>>>(d_segs.p, T);
    k_seg_bin<<<gridN, 256, 0, st>>>(d_segs.p, d_seg.p, d_prims[cur].p, n, d_err.p);
    k_seg_sah<<<gridT, 128, 0, st>>>(d_segs.p, T);
    k_flags<<<gridN, 256, 0, st>>>(d_segs.p, d_seg.p, d_prims[cur].p, n, d_flag.p);
    if ((e = scan.run(d_flag.p, d_scan.p, n, st)) != cudaSuccess) return ctx->cuda_fail(e, "build: scan");
    k_seg_scan_lo<<<gridT, 128, 0, st>>>(d_segs.p, T, d_scan.p);
    k_scatter<<<gridN, 256, 0, st>>>(d_segs.p, d_seg.p, d_flag.p, d_scan.p, d_prims[cur].p, d_prims[1 - cur].p, n);
    k_seg_results<<<(T + 255) / 256, 256, 0, st>>>(d_segs.p, T, B.d_results.p);
    launches += 13;
    cur = 1 - cur;
    if ((e = cudaMemcpyAsync(B.h_pinned, B.d_results.p, (size_t)T * sizeof(int2), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return ctx->cuda_fail(e, "build: range results");
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return ctx->cuda_fail(e, "build: round");
    for (int t = 0; t < T; t++) result[(size_t)where[(size_t)t]] = Split{B.h_pinned[t].x, B.h_pinned[t].y};
    return VG_OK;
  };

  BCUDA(cudaSetDevice(ctx->device));
  BCUDA(d_boxes.reserve((size_t)n * 6)); BCUDA(d_cent.reserve((size_t)n * 3));
  BCUDA(d_prims[0].reserve((size_t)n)); BCUDA(d_prims[1].reserve((size_t)n));
  BCUDA(d_seg.reserve((size_t)n)); BCUDA(d_flag.reserve((size_t)n)); BCUDA(d_scan.reserve((size_t)n)); BCUDA(d_err.reserve(1));
  BCUDA(d_idx.reserve((size_t)n));
  BCUDA(cudaMemsetAsync(d_err.p, 0, sizeof(int), st));
  BCUDA(cudaMemcpyAsync(d_boxes.p, boxes, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice, st));
  BCUDA(cudaMemcpyAsync(d_cent.p, cent, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
  k_load_prims<<<gridN, 256, 0, st>>>(d_boxes.p, d_cent.p, n, d_prims[0].p);
  launches++;

  {
    QNode root;
    std::memset(&root, 0, sizeof(root));
    root.lo = 0;
    root.hi = n;
    nodes.push_back(root);
    act.push_back(0);
  }
  for (int depth = 0; !act.empty(); depth++) {
    if (depth > 256) {
      rc = ctx->fail(VG_ERR_BUILD, "qbvh.BuildAccel: unbounded recursion — more than leafMax primitives share one centroid (the reference "
                                   "overflows its stack here: build.go:35-43)");
      goto done;
    }
    std::vector<int2> ranges;
    for (int q : act) ranges.push_back(make_int2(nodes[(size_t)q].lo, nodes[(size_t)q].hi));
    if ((rc = split_round(ranges)) != VG_OK) goto done;
    for (size_t t = 0; t < act.size(); t++) {
      QNode& Q = nodes[(size_t)act[t]];
      Q.a[0] = result[t].axis;
      Q.p[0] = result[t].pivot;
    }
    ranges.clear();
    for (int q : act) {
      const QNode& Q = nodes[(size_t)q];
      ranges.push_back(make_int2(Q.lo, Q.p[0]));
      ranges.push_back(make_int2(Q.p[0], Q.hi));
    }
    if ((rc = split_round(ranges)) != VG_OK) goto done;
    std::vector<int> next;
    for (size_t t = 0; t < act.size(); t++) {
      const int q = act[t];
      nodes[(size_t)q].a[1] = result[2 * t].axis;
      nodes[(size_t)q].p[1] = result[2 * t].pivot;
      nodes[(size_t)q].a[2] = result[2 * t + 1].axis;
      nodes[(size_t)q].p[2] = result[2 * t + 1].pivot;
      const QNode Q = nodes[(size_t)q];
      const int lo[4] = {Q.lo, Q.p[1], Q.p[0], Q.p[2]};
      const int hi[4] = {Q.p[1], Q.p[0], Q.p[2], Q.hi};
      for (int k = 0; k < 4; k++) {
        const int cn = hi[k] - lo[k];
        nodes[(size_t)q].child[k] = -1;
        nodes[(size_t)q].leaf[k] = -1;
        if (cn > leaf_max) {
          QNode C;
          std::memset(&C, 0, sizeof(C));
          C.lo = lo[k];
          C.hi = hi[k];
          nodes[(size_t)q].child[k] = (int)nodes.size();
          next.push_back((int)nodes.size());
          nodes.push_back(C);
        } else if (cn > 0) {
          nodes[(size_t)q].leaf[k] = (int)leaves.size();
          leaves.push_back(make_int2(lo[k], hi[k]));
        }
      }
    }
    act.swap(next);
  }

  // leaf boxes and the permutation
  {
    std::vector<float> leafbox(leaves.size() * 6);
    if (!leaves.empty()) {
      BCUDA(d_leaves.reserve(leaves.size())); BCUDA(d_leafbox.reserve(leaves.size() * 6));
      BCUDA(cudaMemcpyAsync(d_leaves.p, leaves.data(), leaves.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
      k_leaf_boxes<<<(unsigned)((leaves.size() + 127) / 128), 128, 0, st>>>(d_leaves.p, (int)leaves.size(), d_prims[cur].p, d_leafbox.p);
      BCUDA(cudaMemcpyAsync(leafbox.data(), d_leafbox.p, leafbox.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
      launches++;
    }
    k_store_idx<<<gridN, 256, 0, st>>>(d_prims[cur].p, n, d_idx.p);
    launches++;
    BCUDA(cudaMemcpyAsync(idx_out, d_idx.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    int herr = 0;
    BCUDA(cudaMemcpyAsync(&herr, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    BCUDA(cudaStreamSynchronize(st));
    if (herr) {
      rc = ctx->fail(VG_ERR_BUILD, "calcMinCost: bin out of range (non-finite centroid?)");
      goto done;
    }

    // boxes bottom-up (children have larger breadth-first indices than their parent), then the reference's preorder numbering
    const float inf = std::numeric_limits<float>::infinity();
    std::vector<float> nodebox(nodes.size() * 6);
    for (int q = (int)nodes.size() - 1; q >= 0; q--) {
      QNode& Q = nodes[(size_t)q];
      float nb[6] = {inf, inf, inf, -inf, -inf, -inf};
      for (int k = 0; k < 4; k++) {
        const float* src = nullptr;
        if (Q.child[k] >= 0) src = &nodebox[(size_t)Q.child[k] * 6];
        else if (Q.leaf[k] >= 0) src = &leafbox[(size_t)Q.leaf[k] * 6];
        if (!src) {  // qbvh.go:67-79: empty leaf = -1 with an InfBox
          for (int a = 0; a < 6; a++) Q.box[k][a] = inf;
          continue;
        }
        for (int a = 0; a < 6; a++) Q.box[k][a] = src[a];
        for (int a = 0; a < 3; a++) {
          nb[a] = nb[a] < src[a] ? nb[a] : src[a];
          nb[3 + a] = nb[3 + a] > src[3 + a] ? nb[3 + a] : src[3 + a];
        }
      }
      std::memcpy(&nodebox[(size_t)q * 6], nb, sizeof(nb));
    }
    if (bounds6) std::memcpy(bounds6, &nodebox[0], 6 * sizeof(float));
    std::vector<int> order(nodes.size(), -1), stack;
    int next_id = 0;
    stack.push_back(0);
    while (!stack.empty()) {
      const int q = stack.back();
      stack.pop_back();
      order[(size_t)q] = next_id++;
      for (int k = 3; k >= 0; k--)
        if (nodes[(size_t)q].child[k] >= 0) stack.push_back(nodes[(size_t)q].child[k]);
    }
    out.assign(nodes.size(), VgNode{});
    for (size_t q = 0; q < nodes.size(); q++) {
      const QNode& Q = nodes[q];
      VgNode& N = out[(size_t)order[q]];
      N.axis0 = (uint32_t)Q.a[0]; N.axis1 = (uint32_t)Q.a[1]; N.axis2 = (uint32_t)Q.a[2];
      const int lo[4] = {Q.lo, Q.p[1], Q.p[0], Q.p[2]};
      const int hi[4] = {Q.p[1], Q.p[0], Q.p[2], Q.hi};
      for (int k = 0; k < 4; k++) {
        for (int a = 0; a < 3; a++) { N.boxes[k + a * 4] = Q.box[k][a]; N.boxes[k + 12 + a * 4] = Q.box[k][3 + a]; }
        if (Q.child[k] >= 0) N.children[k] = order[(size_t)Q.child[k]];
        else if (Q.leaf[k] >= 0) N.children[k] = (int32_t)((1u << 31) | (((uint32_t)lo[k] << 4) & 0xfffffff0u) | ((uint32_t)(hi[k] - lo[k] - 1) & 0xf));  // qbvh.go:82-91
        else N.children[k] = -1;
      }
    }
  }
done:
  ctx->stats.kernel_launches += launches;
  return rc;
#undef BCUDA
}

}  // namespace vg

extern "C" int vg_build_qbvh(vg_ctx* ctx, const float* boxes, const float* centroids, int n, int leaf_max, int32_t* idx_out, float* bounds6,
                             int* n_nodes) {
  if (!ctx) return VG_ERR_INVALID;
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (!boxes || !centroids || n <= 0 || !n_nodes || !idx_out) return ctx->fail(VG_ERR_INVALID, "vg_build_qbvh: null/empty input");
  ctx->built_nodes.clear();
  const int rc = vg::build_qbvh_device(ctx, boxes, centroids, n, leaf_max, ctx->built_nodes, idx_out, bounds6);
  if (rc != VG_OK) {
    ctx->built_nodes.clear();
    return rc;
  }
  *n_nodes = (int)ctx->built_nodes.size();
  return VG_OK;
}

extern "C" int vg_build_qbvh_nodes(vg_ctx* ctx, VgNode* nodes_out, int nodes_cap) {
  if (!ctx) return VG_ERR_INVALID;
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (!nodes_out || nodes_cap < (int)ctx->built_nodes.size()) return ctx->fail(VG_ERR_INVALID, "vg_build_qbvh_nodes: buffer smaller than the last build");
  std::memcpy(nodes_out, ctx->built_nodes.data(), ctx->built_nodes.size() * sizeof(VgNode));
  return VG_OK;
}
