// ORACLE — TEST INFRASTRUCTURE ONLY. See qbvh.h for the reference files this restates.
#include "qbvh.h"

#include <emmintrin.h>
#include <xmmintrin.h>

#include <utility>

namespace orc {

static const int nbins = 8;  // qbvh/build.go:22

// qbvh/build.go:28-136. Returns (axis, pivot); partitions indxs/centroids/boxes in place.
static void calcMinCost(const BoundingBox& bounds, Vec3* centroids, BoundingBox* boxes, int32_t* indxs, int n, int* axis_out, int* pivot_out) {
  BoundingBox binBounds[nbins];
  int32_t binN[nbins] = {0};

  int axis = bounds.MaxDim();

  if (bounds.b[1][axis] == bounds.b[0][axis]) {  // build.go:35-43 flat box: median-ish split, no partition
    *axis_out = axis;
    *pivot_out = n / 2 + 1;
    return;
  }

  float k1 = (float)nbins * (float)(1.0 - 0.00006) / (bounds.b[1][axis] - bounds.b[0][axis]);
  float k0 = bounds.b[0][axis];

  for (int i = 0; i < nbins; i++) binBounds[i].Reset();

  for (int i = 0; i < n; i++) {
    float c = centroids[i][axis];
    int bin = (int)(k1 * (c - k0));
    if (bin < 0) throw std::runtime_error("calcMinCost: bin < 0");
    if (bin > nbins - 1) throw std::runtime_error("calcMinCost: bin > Nbins-1");
    binN[bin]++;
    binBounds[bin].GrowBox(boxes[i]);
  }

  BoundingBox lbox[nbins], rbox[nbins];
  int32_t lN[nbins], rN[nbins];

  BoundingBox box;
  box.Reset();
  int32_t cnt = 0;
  for (int i = 0; i < nbins; i++) {
    box.GrowBox(binBounds[i]);
    cnt += binN[i];
    lbox[i] = box;
    lN[i] = cnt;
  }
  box.Reset();
  cnt = 0;
  for (int i = 0; i < nbins; i++) {
    box.GrowBox(binBounds[nbins - 1 - i]);
    cnt += binN[nbins - 1 - i];
    rbox[nbins - 1 - i] = box;
    rN[nbins - 1 - i] = cnt;
  }

  int binMinCost = -1;
  float minCost = kInfPos;
  for (int i = 1; i < nbins; i++) {
    float cost = lbox[i - 1].SurfaceArea() * (float)lN[i - 1] + rbox[i].SurfaceArea() * (float)rN[i];
    if (cost < minCost) {
      binMinCost = i;
      minCost = cost;
    }
  }

  int left = 0;
  int right = n - 1;
  while (left <= right) {
    float c = centroids[left][axis];
    int bin = (int)(k1 * (c - k0));
    if (bin < binMinCost) {
      left++;
    } else {
      std::swap(indxs[left], indxs[right]);
      std::swap(centroids[left], centroids[right]);
      std::swap(boxes[left], boxes[right]);
      right--;
    }
  }
  *axis_out = axis;
  *pivot_out = left;
}

// qbvh/build.go:138-149
static BoundingBox calcBox(const BoundingBox* boxes, int n) {
  BoundingBox box;
  box.Reset();
  for (int i = 0; i < n; i++) box.GrowBox(boxes[i]);
  return box;
}

// qbvh/build.go:151-178
static void binarySplit(BoundingBox* boxes, Vec3* centroids, int leafMax, int32_t* indxs, int n, int* axis, int* pivot) {
  BoundingBox bounds;
  bounds.Reset();
  for (int i = 0; i < n; i++) bounds.GrowVec3(centroids[i]);
  if (n <= leafMax) {
    *axis = 0;
    *pivot = n;
    return;
  }
  calcMinCost(bounds, centroids, boxes, indxs, n, axis, pivot);
}

// qbvh/build.go:180-288
// NOTE (reference quirk f): when >leafMax primitives share one centroid, calcMinCost's flat-axis case
// (build.go:35-43) returns pivot n/2+1 without partitioning, child 0 keeps all n primitives and the Go code
// recurses until the goroutine stack overflows. The oracle reports that as an error instead of crashing.
static int32_t buildAccelRec(std::vector<Node>* nodes, BoundingBox* boxes, Vec3* centroids, int32_t* indxs, int n, int leafMax, int baseidx, BoundingBox* outbox, int depth = 0) {
  if (depth > 256) throw std::runtime_error("qbvh.BuildAccel: unbounded recursion (coincident centroids, reference would overflow its stack)");
  int axis0, pivot0, axis1, pivot1, axis2, pivot2;
  binarySplit(boxes, centroids, leafMax, indxs, n, &axis0, &pivot0);
  binarySplit(boxes, centroids, leafMax, indxs, pivot0, &axis1, &pivot1);
  binarySplit(boxes + pivot0, centroids + pivot0, leafMax, indxs + pivot0, n - pivot0, &axis2, &pivot2);

  int32_t nodei = (int32_t)nodes->size();
  nodes->push_back(Node{});
  (*nodes)[nodei].Axis0 = (uint32_t)axis0;
  (*nodes)[nodei].Axis1 = (uint32_t)axis1;
  (*nodes)[nodei].Axis2 = (uint32_t)axis2;

  // child k covers [lo[k], hi[k])
  const int lo[4] = {0, pivot1, pivot0, pivot0 + pivot2};
  const int hi[4] = {pivot1, pivot0, pivot0 + pivot2, n};
  for (int k = 0; k < 4; k++) {
    int cn = hi[k] - lo[k];
    if (cn <= leafMax) {
      BoundingBox cb = calcBox(boxes + lo[k], cn);
      (*nodes)[nodei].SetBounds(k, cb);
      (*nodes)[nodei].SetLeaf(k, (uint32_t)(baseidx + lo[k]), (uint32_t)cn);
    } else {
      BoundingBox cb;
      int32_t child = buildAccelRec(nodes, boxes + lo[k], centroids + lo[k], indxs + lo[k], cn, leafMax, baseidx + lo[k], &cb, depth + 1);
      (*nodes)[nodei].SetBounds(k, cb);
      (*nodes)[nodei].Children[k] = child;
    }
  }

  BoundingBox nodebox;
  nodebox.Reset();
  for (int i = 0; i < 4; i++)
    if ((*nodes)[nodei].Children[i] != -1) nodebox.GrowBox((*nodes)[nodei].Bounds(i));
  *outbox = nodebox;
  return nodei;
}

// qbvh/build.go:293-307
std::vector<Node> BuildAccel(BoundingBox* boxes, Vec3* centroids, int32_t* indxs, int n, int leafMax, BoundingBox* bounds) {
  if (leafMax > 16) leafMax = 16;
  if (leafMax < 1) leafMax = 1;
  std::vector<Node> nodes;
  buildAccelRec(&nodes, boxes, centroids, indxs, n, leafMax, 0, bounds);
  return nodes;
}

// qbvh/motionbuild.go:12-100
static int32_t buildAccelMotionRec(std::vector<MotionNode>* nodes, BoundingBox* boxes, Vec3* centroids, int32_t* indxs, int n, int leafMax, int baseidx, int depth = 0) {
  if (depth > 256) throw std::runtime_error("qbvh.BuildAccelMotion: unbounded recursion (coincident centroids, reference would overflow its stack)");
  int axis0, pivot0, axis1, pivot1, axis2, pivot2;
  binarySplit(boxes, centroids, leafMax, indxs, n, &axis0, &pivot0);
  binarySplit(boxes, centroids, leafMax, indxs, pivot0, &axis1, &pivot1);
  binarySplit(boxes + pivot0, centroids + pivot0, leafMax, indxs + pivot0, n - pivot0, &axis2, &pivot2);

  int32_t nodei = (int32_t)nodes->size();
  nodes->push_back(MotionNode{});
  (*nodes)[nodei].Axis0 = axis0;
  (*nodes)[nodei].Axis1 = axis1;
  (*nodes)[nodei].Axis2 = axis2;

  const int lo[4] = {0, pivot1, pivot0, pivot0 + pivot2};
  const int hi[4] = {pivot1, pivot0, pivot0 + pivot2, n};
  for (int k = 0; k < 4; k++) {
    int cn = hi[k] - lo[k];
    if (cn <= leafMax) {
      (*nodes)[nodei].SetLeaf(k, (uint32_t)(baseidx + lo[k]), (uint32_t)cn);
    } else {
      int32_t child = buildAccelMotionRec(nodes, boxes + lo[k], centroids + lo[k], indxs + lo[k], cn, leafMax, baseidx + lo[k], depth + 1);
      (*nodes)[nodei].Children[k] = child;
    }
  }
  return nodei;
}

// qbvh/motionbuild.go:108-122
std::vector<MotionNode> BuildAccelMotion(BoundingBox* boxes, Vec3* centroids, int32_t* indxs, int n, int leafMax) {
  if (leafMax > 16) leafMax = 16;
  if (leafMax < 1) leafMax = 1;
  std::vector<MotionNode> nodes;
  buildAccelMotionRec(&nodes, boxes, centroids, indxs, n, leafMax, 0);
  return nodes;
}

// qbvh/intersect_amd64.s:13-100. Go-asm `OP src,dst`; _mm_min_ps(a,b) is MINPS with dst=a, src=b,
// i.e. (a < b) ? a : b — the SECOND operand is returned on NaN / equal zeros.
void intersectBoxes(const Ray* ray, const float* boxes, int32_t* hits, float* t) {
  __m128 X3 = _mm_set1_ps(ray->P[0]);
  __m128 X4 = _mm_set1_ps(ray->Dinv[0]);
  __m128 X2 = _mm_load_ps(boxes + 0);
  X2 = _mm_sub_ps(X2, X3);
  X2 = _mm_mul_ps(X2, X4);  // t1 = (boxmin.x - O.x) * Dinv.x
  __m128 X6 = _mm_load_ps(boxes + 12);
  X6 = _mm_sub_ps(X6, X3);
  X6 = _mm_mul_ps(X6, X4);  // t2
  __m128 X7 = X6;
  X6 = _mm_min_ps(X6, X2);  // MINPS X2,X6
  X7 = _mm_max_ps(X7, X2);  // MAXPS X2,X7

  X3 = _mm_set1_ps(ray->P[1]);
  X4 = _mm_set1_ps(ray->Dinv[1]);
  X2 = _mm_load_ps(boxes + 4);
  X2 = _mm_sub_ps(X2, X3);
  X2 = _mm_mul_ps(X2, X4);
  __m128 X0 = _mm_load_ps(boxes + 16);
  X0 = _mm_sub_ps(X0, X3);
  X0 = _mm_mul_ps(X0, X4);
  __m128 X1 = X0;
  X1 = _mm_min_ps(X1, X2);  // MINPS X2,X1
  X0 = _mm_max_ps(X0, X2);  // MAXPS X2,X0
  X6 = _mm_max_ps(X6, X1);  // MAXPS X1,X6
  X7 = _mm_min_ps(X7, X0);  // MINPS X0,X7

  X3 = _mm_set1_ps(ray->P[2]);
  X4 = _mm_set1_ps(ray->Dinv[2]);
  X2 = _mm_load_ps(boxes + 8);
  X2 = _mm_sub_ps(X2, X3);
  X2 = _mm_mul_ps(X2, X4);
  X0 = _mm_load_ps(boxes + 20);
  X0 = _mm_sub_ps(X0, X3);
  X0 = _mm_mul_ps(X0, X4);
  X1 = X0;
  X1 = _mm_min_ps(X1, X2);
  X0 = _mm_max_ps(X0, X2);
  X6 = _mm_max_ps(X6, X1);
  X7 = _mm_min_ps(X7, X0);

  X0 = _mm_setzero_ps();
  X0 = _mm_max_ps(X0, X6);  // MAXPS X6,X0 : tNear = max(0, tmin)
  _mm_store_ps(t, X0);
  X0 = _mm_cmple_ps(X0, X7);  // CMPPS X7,X0,$2 : tNear <= tmax
  _mm_store_si128((__m128i*)hits, _mm_castps_si128(X0));
}

// qbvh/intersect.go:17-50 — the reference's OTHER scalar version: planes picked by the sign of Dinv instead of min/max, and
// tFar additionally clamped by Tclosest (:39). The reference's commented-out self-check (:116-133) compares the asm's hits with
// this one; tests/test_oracle_traversal.py restates that check and pins where the two may differ.
void intersectBoxesSlow(const Ray* ray, const float* boxes, int32_t* hits, float* t) {
  uint8_t sign[3] = {0, 0, 0};
  for (int k = 0; k < 3; k++)
    if (ray->Dinv[k] < 0.0f) sign[k] = 1;
  for (int idx = 0; idx < 4; idx++) {
    float tmin = (boxes[idx + (sign[0] * 12) + 0] - ray->P[0]) * ray->Dinv[0];
    float tmax = (boxes[idx + ((1 - sign[0]) * 12) + 0] - ray->P[0]) * ray->Dinv[0];
    float tymin = (boxes[idx + (sign[1] * 12) + 4] - ray->P[1]) * ray->Dinv[1];
    float tymax = (boxes[idx + ((1 - sign[1]) * 12) + 4] - ray->P[1]) * ray->Dinv[1];
    float tzmin = (boxes[idx + (sign[2] * 12) + 8] - ray->P[2]) * ray->Dinv[2];
    float tzmax = (boxes[idx + ((1 - sign[2]) * 12) + 8] - ray->P[2]) * ray->Dinv[2];
    float tNear = Max(Max(tmin, tymin), Max(0.0f, tzmin));
    float tFar = Min(Min(tmax, tymax), Min(ray->Tclosest, tzmax));
    t[idx] = tNear;
    hits[idx] = (tNear <= tFar) ? -1 : 0;
  }
}

// qbvh/intersect.go:52-87 — the reference's scalar twin of the asm.
void intersectBoxesSlow2(const Ray* ray, const float* boxes, int32_t* hits, float* t) {
  for (int idx = 0; idx < 4; idx++) {
    float tx1 = (boxes[idx + (0 * 12) + 0] - ray->P[0]) * ray->Dinv[0];
    float tx2 = (boxes[idx + (1 * 12) + 0] - ray->P[0]) * ray->Dinv[0];
    float tmin = Min(tx1, tx2);
    float tmax = Max(tx1, tx2);
    float ty1 = (boxes[idx + (0 * 12) + 4] - ray->P[1]) * ray->Dinv[1];
    float ty2 = (boxes[idx + (1 * 12) + 4] - ray->P[1]) * ray->Dinv[1];
    tmin = Max(tmin, Min(ty1, ty2));
    tmax = Min(tmax, Max(ty1, ty2));
    float tz1 = (boxes[idx + (0 * 12) + 8] - ray->P[2]) * ray->Dinv[2];
    float tz2 = (boxes[idx + (1 * 12) + 8] - ray->P[2]) * ray->Dinv[2];
    tmin = Max(tmin, Min(tz1, tz2));
    tmax = Min(tmax, Max(tz1, tz2));
    t[idx] = Max(0, tmin);
    hits[idx] = (tmax >= Max(0, tmin)) ? -1 : 0;
  }
}

// qbvh/intersect.go:91-246
bool QTrace(const std::vector<Node>& qbvh, Primitive* prim, Ray* ray, ShaderContext* sg) {
  auto& tr = ray->Task->Traversal;
  int32_t stackTop = tr.StackTop;
  tr.Stack[stackTop].Node = 0;
  tr.Stack[stackTop].T = ray->Tclosest;
  bool hit = false;
  stackTop++;

  auto push = [&](const Node& nd, int k) {
    tr.Stack[stackTop].Node = nd.Children[k];
    tr.Stack[stackTop].T = tr.T[k];
    stackTop -= tr.Hits[k];
  };

  while (stackTop > tr.StackTop) {
    stackTop--;
    int32_t node = tr.Stack[stackTop].Node;
    if (ray->Tclosest < tr.Stack[stackTop].T || node == -1) continue;

    if (node >= 0) {
      ray->NodesT++;
      const Node& nd = qbvh[node];
      intersectBoxes(ray, nd.Boxes, tr.Hits, tr.T);
      // intersect.go:137-216. The Go array-bounds check would panic past Stack[89]; mirror it.
      if (stackTop + 4 > 90) throw std::runtime_error("qbvh.Trace: traversal stack overflow");
      if (ray->D[nd.Axis0] < 0) {
        if (ray->D[nd.Axis1] < 0) { push(nd, 0); push(nd, 1); } else { push(nd, 1); push(nd, 0); }
        if (ray->D[nd.Axis2] < 0) { push(nd, 2); push(nd, 3); } else { push(nd, 3); push(nd, 2); }
      } else {
        if (ray->D[nd.Axis2] < 0) { push(nd, 2); push(nd, 3); } else { push(nd, 3); push(nd, 2); }
        if (ray->D[nd.Axis1] < 0) { push(nd, 0); push(nd, 1); } else { push(nd, 1); push(nd, 0); }
      }
    } else {
      ray->LeafsT++;
      int32_t tmp = tr.StackTop;
      tr.StackTop = stackTop + 1;
      if (prim->TraceElems(ray, sg, LeafBase(node), LeafCount(node))) {
        hit = true;
        if (ray->Type & RayTypeShadow) {
          tr.StackTop = tmp;
          return true;
        }
      }
      tr.StackTop = tmp;
    }
  }
  return hit;
}

// qbvh/motionintersect.go:22-127
bool QTraceMotion(const MotionQBVH& qbvh, float time, int key, int key2, MotionPrimitive* prim, Ray* ray, ShaderContext* sg) {
  auto& tr = ray->Task->Traversal;
  int32_t stackTop = tr.StackTop;
  tr.Stack[stackTop].Node = 0;
  tr.Stack[stackTop].T = ray->Tclosest;
  bool hit = false;

  while (stackTop >= tr.StackTop) {
    int32_t node = tr.Stack[stackTop].Node;
    float T = tr.Stack[stackTop].T;
    stackTop--;
    if (ray->Tclosest < T) node = -1;

    if (node >= 0) {
      const MotionNode* pnode = &qbvh.Nodes[node];
      ray->NodesT++;  // oracle-only counter (the reference does not count in TraceMotion)
      for (int i = 0; i < 24; i++) tr.Boxes[i] = (1.0f - time) * qbvh.Boxes[key][node].v[i] + time * qbvh.Boxes[key2][node].v[i];
      intersectBoxes(ray, tr.Boxes, tr.Hits, tr.T);

      int order[4] = {0, 1, 2, 3};
      if (ray->D[pnode->Axis0] < 0) {
        if (ray->D[pnode->Axis2] < 0) { order[3] = 3; order[2] = 2; } else { order[3] = 2; order[2] = 3; }
        if (ray->D[pnode->Axis1] < 0) { order[1] = 1; order[0] = 0; } else { order[1] = 0; order[0] = 1; }
      } else {
        if (ray->D[pnode->Axis2] < 0) { order[1] = 3; order[0] = 2; } else { order[1] = 2; order[0] = 3; }
        if (ray->D[pnode->Axis1] < 0) { order[3] = 1; order[2] = 0; } else { order[3] = 0; order[2] = 1; }
      }
      for (int j = 0; j < 4; j++) {
        int k = order[j];
        if (tr.Hits[k] != 0) {
          stackTop++;
          if (stackTop >= 90) throw std::runtime_error("qbvh.TraceMotion: traversal stack overflow");
          tr.Stack[stackTop].Node = pnode->Children[k];
          tr.Stack[stackTop].T = tr.T[k];
        }
      }
    } else if (node < -1) {
      int leafBase = LeafBase(node);
      int leafCount = LeafCount(node);
      ray->LeafsT++;
      int32_t tmp = tr.StackTop;
      tr.StackTop = stackTop + 1;
      if (prim->TraceMotionElems(time, key, key2, ray, sg, leafBase, leafCount)) {
        hit = true;
        if (ray->Type & RayTypeShadow) {
          tr.StackTop = tmp;
          return true;
        }
      }
      tr.StackTop = tmp;
    }
  }
  return hit;
}

}  // namespace orc
