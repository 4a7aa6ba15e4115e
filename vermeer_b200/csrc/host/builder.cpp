// Host-side QBVH / MQBVH construction for the vh_* layer.
//
// Produces node arrays that are bit-identical to what the reference's builder emits for the same
// input (same binned-SAH decisions, same in-place partition order, same preorder node numbering), so
// traversal order, NodesT and tie-breaks on the device match the reference:
//   qbvh.BuildAccel / buildAccelRec / binarySplit / calcMinCost / calcBox   qbvh/build.go:22-307
//   qbvh.BuildAccelMotion / buildAccelMotionRec                              qbvh/motionbuild.go:12-122
// Unlike the reference (single recursive goroutine appending to one slice) the four sub-ranges of the top
// levels are built concurrently into private node vectors and stitched in preorder afterwards: the
// ranges are disjoint, the partition is in place, and preorder numbering makes every subtree a
// contiguous block, so the result does not depend on the schedule.
#include "builder.h"

#include <future>
#include <stdexcept>
#include <thread>
#include <utility>

namespace vh {
namespace {

const int kBins = 8;  // build.go:22
const int kMaxDepth = 256;

struct Range {
  Box* boxes;
  V3* cent;
  int32_t* idx;
  int n;
  Range sub(int lo, int hi) const { return Range{boxes + lo, cent + lo, idx + lo, hi - lo}; }
};

struct BuildError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// build.go:28-136: choose the SAH split among 8 bins on the longest axis of the centroid bounds and
// partition the range in place (two-pointer swap from both ends, as the reference does).
void sah_partition(const Box& cb, Range r, int* axis_out, int* pivot_out) {
  const int axis = cb.max_dim();
  if (cb.hi[axis] == cb.lo[axis]) {  // flat: the reference returns len/2+1 WITHOUT partitioning (build.go:35-43)
    *axis_out = axis;
    *pivot_out = r.n / 2 + 1;
    return;
  }
  const float k1 = (float)kBins * (float)(1.0 - 0.00006) / (cb.hi[axis] - cb.lo[axis]);
  const float k0 = cb.lo[axis];
  Box bb[kBins];
  int32_t bn[kBins] = {0};
  for (int i = 0; i < kBins; i++) bb[i].reset();
  for (int i = 0; i < r.n; i++) {
    const int bin = (int)(k1 * (r.cent[i][axis] - k0));
    if (bin < 0 || bin > kBins - 1) throw BuildError("calcMinCost: bin out of range (non-finite centroid?)");
    bn[bin]++;
    bb[bin].grow_box(r.boxes[i]);
  }
  Box lbox[kBins], rbox[kBins];
  int32_t ln[kBins], rn[kBins];
  Box acc;
  acc.reset();
  int32_t cnt = 0;
  for (int i = 0; i < kBins; i++) { acc.grow_box(bb[i]); cnt += bn[i]; lbox[i] = acc; ln[i] = cnt; }
  acc.reset();
  cnt = 0;
  for (int i = kBins - 1; i >= 0; i--) { acc.grow_box(bb[i]); cnt += bn[i]; rbox[i] = acc; rn[i] = cnt; }
  int best = -1;
  float best_cost = kInf;
  for (int i = 1; i < kBins; i++) {
    const float cost = lbox[i - 1].area() * (float)ln[i - 1] + rbox[i].area() * (float)rn[i];
    if (cost < best_cost) { best = i; best_cost = cost; }
  }
  int left = 0, right = r.n - 1;
  while (left <= right) {
    const int bin = (int)(k1 * (r.cent[left][axis] - k0));
    if (bin < best) {
      left++;
    } else {
      std::swap(r.idx[left], r.idx[right]);
      std::swap(r.cent[left], r.cent[right]);
      std::swap(r.boxes[left], r.boxes[right]);
      right--;
    }
  }
  *axis_out = axis;
  *pivot_out = left;
}

// build.go:151-178
void binary_split(Range r, int leafMax, int* axis, int* pivot) {
  if (r.n <= leafMax) { *axis = 0; *pivot = r.n; return; }
  Box cb;
  cb.reset();
  for (int i = 0; i < r.n; i++) cb.grow_point(r.cent[i].x, r.cent[i].y, r.cent[i].z);
  sah_partition(cb, r, axis, pivot);
}

inline uint32_t leaf_link(uint32_t first, uint32_t count) {  // qbvh.go:82-91
  return (1u << 31) | ((first << 4) & 0xfffffff0u) | ((count - 1) & 0xf);
}

template <class NodeT>
struct Policy;
template <>
struct Policy<VgNode> {
  static void set_axes(VgNode& n, int a0, int a1, int a2) { n.axis0 = (uint32_t)a0; n.axis1 = (uint32_t)a1; n.axis2 = (uint32_t)a2; }
  static void set_box(VgNode& n, int k, const Box& b) {
    for (int a = 0; a < 3; a++) { n.boxes[k + a * 4] = b.lo[a]; n.boxes[k + 12 + a * 4] = b.hi[a]; }
  }
  static const bool kBoxes = true;
};
template <>
struct Policy<VgMotionNode> {
  static void set_axes(VgMotionNode& n, int a0, int a1, int a2) { n.axis0 = a0; n.axis1 = a1; n.axis2 = a2; }
  static void set_box(VgMotionNode&, int, const Box&) {}
  static const bool kBoxes = false;
};

template <class NodeT>
struct Builder {
  int leafMax;
  size_t par_threshold;

  // Appends the subtree of `r` to `out` in preorder; returns its root index in `out`.
  int32_t build(std::vector<NodeT>& out, Range r, int base, Box* box_out, int depth, int par_levels) {
    if (depth > kMaxDepth)
      throw BuildError("qbvh.BuildAccel: unbounded recursion — more than leafMax primitives share one centroid "
                       "(the reference overflows its stack here: build.go:35-43)");
    int a0, p0, a1, p1, a2, p2;
    binary_split(r, leafMax, &a0, &p0);
    binary_split(r.sub(0, p0), leafMax, &a1, &p1);
    binary_split(r.sub(p0, r.n), leafMax, &a2, &p2);
    const int32_t me = (int32_t)out.size();
    out.push_back(NodeT{});
    Policy<NodeT>::set_axes(out[me], a0, a1, a2);
    const int lo[4] = {0, p1, p0, p0 + p2};
    const int hi[4] = {p1, p0, p0 + p2, r.n};

    const bool parallel = par_levels > 0 && (size_t)r.n >= par_threshold;
    std::vector<NodeT> sub[4];
    std::future<void> fut[4];
    Box cbox[4];
    bool is_leaf[4];
    for (int k = 0; k < 4; k++) {
      const int cn = hi[k] - lo[k];
      is_leaf[k] = cn <= leafMax;
      if (is_leaf[k]) continue;
      if (parallel) {
        fut[k] = std::async(std::launch::async, [this, &sub, &cbox, r, lo, hi, base, depth, par_levels, k] {
          build(sub[k], r.sub(lo[k], hi[k]), base + lo[k], &cbox[k], depth + 1, par_levels - 1);
        });
      }
    }
    Box nodebox;
    nodebox.reset();
    std::exception_ptr first_error;
    for (int k = 0; k < 4; k++) {
      const int cn = hi[k] - lo[k];
      if (is_leaf[k]) {
        if (cn == 0) {  // qbvh.go:67-79: empty leaf = -1 with an InfBox
          out[me].children[k] = -1;
          Box inf;
          for (int a = 0; a < 3; a++) inf.lo[a] = inf.hi[a] = kInf;
          Policy<NodeT>::set_box(out[me], k, inf);
        } else {
          Box cb;
          cb.reset();
          if (Policy<NodeT>::kBoxes)
            for (int i = lo[k]; i < hi[k]; i++) cb.grow_box(r.boxes[i]);
          Policy<NodeT>::set_box(out[me], k, cb);
          out[me].children[k] = (int32_t)leaf_link((uint32_t)(base + lo[k]), (uint32_t)cn);
          if (Policy<NodeT>::kBoxes) nodebox.grow_box(cb);
        }
        continue;
      }
      int32_t child;
      if (parallel) {
        try {
          fut[k].get();
        } catch (...) {
          if (!first_error) first_error = std::current_exception();
          continue;
        }
        // stitch: the private subtree becomes the contiguous block starting at out.size()
        const int32_t off = (int32_t)out.size();
        for (NodeT& n : sub[k]) {
          for (int c = 0; c < 4; c++)
            if (n.children[c] >= 0) n.children[c] += off;
          out.push_back(n);
        }
        sub[k].clear();
        sub[k].shrink_to_fit();
        child = off;
      } else {
        child = build(out, r.sub(lo[k], hi[k]), base + lo[k], &cbox[k], depth + 1, 0);
      }
      Policy<NodeT>::set_box(out[me], k, cbox[k]);
      out[me].children[k] = child;
      if (Policy<NodeT>::kBoxes) nodebox.grow_box(cbox[k]);
    }
    if (first_error) std::rethrow_exception(first_error);
    if (box_out) *box_out = nodebox;
    return me;
  }
};

template <class NodeT>
int run(Box* boxes, V3* cent, int32_t* idx, int n, int leafMax, std::vector<NodeT>& out, Box* bounds, std::string* err) {
  if (leafMax > 16) leafMax = 16;
  if (leafMax < 1) leafMax = 1;
  out.clear();
  Builder<NodeT> b;
  b.leafMax = leafMax;
  b.par_threshold = 65536;
  const unsigned hw = std::thread::hardware_concurrency();
  const int par_levels = hw >= 4 ? 2 : 0;
  try {
    Box bb;
    bb.reset();
    b.build(out, Range{boxes, cent, idx, n}, 0, &bb, 0, par_levels);
    if (bounds) *bounds = bb;
  } catch (const std::exception& e) {
    if (err) *err = e.what();
    return -1;
  }
  return 0;
}

}  // namespace

int build_qbvh(Box* boxes, V3* cent, int32_t* idx, int n, int leafMax, std::vector<VgNode>& out, Box* bounds, std::string* err) {
  return run<VgNode>(boxes, cent, idx, n, leafMax, out, bounds, err);
}
int build_mqbvh(Box* boxes, V3* cent, int32_t* idx, int n, int leafMax, std::vector<VgMotionNode>& out, std::string* err) {
  return run<VgMotionNode>(boxes, cent, idx, n, leafMax, out, nullptr, err);
}

}  // namespace vh
