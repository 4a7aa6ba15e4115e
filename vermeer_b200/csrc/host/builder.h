// QBVH / MQBVH construction (see builder.cpp).
#pragma once
#include <string>
#include <vector>

#include "../../../include/vermeer_gpu.h"
#include "hmath.h"

namespace vh {

// Builds the 4-wide tree over n primitives given by (boxes, centroids, idx); partitions the three arrays
// in place so that leaf (base,count) ranges index them directly (qbvh/build.go:293-307). Returns 0 or -1 (+err).
int build_qbvh(Box* boxes, V3* centroids, int32_t* idx, int n, int leafMax, std::vector<VgNode>& out, Box* bounds, std::string* err);
// Topology only (qbvh/motionbuild.go:108-122); per-key boxes are filled by the caller.
int build_mqbvh(Box* boxes, V3* centroids, int32_t* idx, int n, int leafMax, std::vector<VgMotionNode>& out, std::string* err);

}  // namespace vh
