// ORACLE — TEST INFRASTRUCTURE ONLY. See scene.h / core.h for the reference files this restates.
#include "scene.h"

#include <cmath>

#include "ldseq.h"

namespace orc {

// ---------------------------------------------------------------------------------------------
// core/ray.go:103-148
void Ray::Setup() {
  Kz = 0;
  if (Abs(D[1]) > Abs(D[2])) {
    if (Abs(D[1]) > Abs(D[0])) Kz = 1;
  } else {
    if (Abs(D[2]) > Abs(D[0])) Kz = 2;
  }
  Kx = Kz + 1;
  if (Kx == 3) Kx = 0;
  Ky = Kx + 1;
  if (Ky == 3) Ky = 0;
  if (D[Kz] < 0.0f) {
    int32_t tmp = Kx;
    Kx = Ky;
    Ky = tmp;
  }
  double z = (double)D[Kz];
  S[2] = (float)(1.0 / z);
  S[0] = (float)((double)D[Kx] / z);
  S[1] = (float)((double)D[Ky] / z);
  Dinv[0] = (float)(1.0 / (double)D[0]);
  Dinv[1] = (float)(1.0 / (double)D[1]);
  Dinv[2] = (float)(1.0 / (double)D[2]);
}

// core/ray.go:56-93
void Ray::Init(uint32_t ty, Vec3 P_, Vec3 D_, float maxdist, uint8_t level, const ShaderContext* sc) {
  P = P_;
  D = D_;
  Tclosest = maxdist;
  Type = ty;
  Setup();
  Level = level;
  Lambda = sc->Lambda;
  Time = sc->Time;
  NodesT = 0;
  LeafsT = 0;
  TrisT = 0;
  Scramble[0] = sc->Scramble[0];
  Scramble[1] = sc->Scramble[1];
  I = sc->I;
  // Compute ray differentials for reflection (ray.go:72-87)
  if (ty & RayTypeReflected) {
    DdPdx = sc->DdPdx;
    DdPdy = sc->DdPdy;
    const float DdotNdx = Vec3Dot(sc->DdDdx, sc->N) + Vec3Dot(sc->Rd, sc->DdNdx);
    const float DdotNdy = Vec3Dot(sc->DdDdy, sc->N) + Vec3Dot(sc->Rd, sc->DdNdy);
    DdDdx = Vec3Mad(sc->DdDdx, Vec3Add(Vec3Scale(Vec3Dot(sc->Rd, sc->N), sc->DdNdx), Vec3Scale(DdotNdx, sc->N)), -2);
    DdDdy = Vec3Mad(sc->DdDdy, Vec3Add(Vec3Scale(Vec3Dot(sc->Rd, sc->N), sc->DdNdy), Vec3Scale(DdotNdy, sc->N)), -2);
  }
}

// core/ray.go:95-104
void Ray::DifferentialTransfer(ShaderContext* sc) const {
  const float dtdx = -Vec3Dot(Vec3Mad(DdPdx, DdDdx, Tclosest), sc->Ng) / Vec3Dot(D, sc->Ng);
  const float dtdy = -Vec3Dot(Vec3Mad(DdPdy, DdDdy, Tclosest), sc->Ng) / Vec3Dot(D, sc->Ng);
  sc->DdPdx = Vec3Add(Vec3Mad(DdPdx, DdDdx, Tclosest), Vec3Scale(dtdx, D));
  sc->DdPdy = Vec3Add(Vec3Mad(DdPdy, DdDdy, Tclosest), Vec3Scale(dtdy, D));
  sc->DdDdx = DdDdx;
  sc->DdDdy = DdDdy;
}

// ---------------------------------------------------------------------------------------------
// builtin/geom/polymesh/init.go:12-134
void PolyMesh::init() {
  const bool hasNormals = !Normals.Elems.empty();
  const bool hasUV = !UV.empty();  // init.go:38-48,77-83,100-106
  if (hasPolyCount) {
    uint32_t basei = 0;
    for (size_t k = 0; k < PolyCount.size(); k++) {
      uint32_t i = 0;
      for (int j = 0; j < (int)(PolyCount[k] - 2); j++) {
        i++;
        idxp.push_back((uint32_t)FaceIdx[basei]);
        idxp.push_back((uint32_t)FaceIdx[basei + i]);
        idxp.push_back((uint32_t)FaceIdx[basei + i + 1]);
        if (hasNormals) {
          const std::vector<int32_t>& src = hasNormalIdx ? NormalIdx : FaceIdx;
          normalidx.push_back((uint32_t)src[basei]);
          normalidx.push_back((uint32_t)src[basei + i]);
          normalidx.push_back((uint32_t)src[basei + i + 1]);
        }
        if (hasUV) {
          const std::vector<int32_t>& src = hasUVIdx ? UVIdx : FaceIdx;
          uvtriidx.push_back((uint32_t)src[basei]);
          uvtriidx.push_back((uint32_t)src[basei + i]);
          uvtriidx.push_back((uint32_t)src[basei + i + 1]);
        }
        if (!ShaderIdx.empty()) shaderidx.push_back((uint8_t)ShaderIdx[k]);
      }
      basei += (uint32_t)PolyCount[k];
    }
  } else {
    if (hasFaceIdx) {
      for (size_t j = 0; j < FaceIdx.size(); j++) {
        idxp.push_back((uint32_t)FaceIdx[j]);
        if (hasNormals) normalidx.push_back(hasNormalIdx ? (uint32_t)NormalIdx[j] : (uint32_t)FaceIdx[j]);
        if (hasUV) uvtriidx.push_back(hasUVIdx ? (uint32_t)UVIdx[j] : (uint32_t)FaceIdx[j]);
      }
    } else {
      for (int j = 0; j < Verts.ElemsPerKey; j++) {
        idxp.push_back((uint32_t)j);
        if (hasNormals) normalidx.push_back(hasNormalIdx ? (uint32_t)NormalIdx[j] : (uint32_t)j);
        if (hasUV) uvtriidx.push_back(hasUVIdx ? (uint32_t)UVIdx[j] : (uint32_t)j);
      }
    }
    for (int32_t idx : ShaderIdx) shaderidx.push_back((uint8_t)idx);
  }
  FaceIdx.clear();
  PolyCount.clear();
  NormalIdx.clear();
  ShaderIdx.clear();
  UVIdx.clear();
}

// builtin/geom/polymesh/buildqbvh.go:14-144
void PolyMesh::initAccel() {
  std::vector<BoundingBox> boxes(facecount);
  std::vector<Vec3> centroids(facecount);
  std::vector<int32_t> idxs(facecount);

  if (Verts.MotionKeys > 1) {
    float k = 0.5f * (float)(Verts.MotionKeys - 1);
    float time = k - Floor(k);
    int key = (int)Floor(k);
    int key2 = (int)Ceil(k);
    for (int i = 0; i < facecount; i++) {
      Vec3 V0 = Vec3Lerp(Verts.Elems[(int)idxp[i * 3 + 0] + Verts.ElemsPerKey * key], Verts.Elems[(int)idxp[i * 3 + 0] + Verts.ElemsPerKey * key2], time);
      Vec3 V1 = Vec3Lerp(Verts.Elems[(int)idxp[i * 3 + 1] + Verts.ElemsPerKey * key], Verts.Elems[(int)idxp[i * 3 + 1] + Verts.ElemsPerKey * key2], time);
      Vec3 V2 = Vec3Lerp(Verts.Elems[(int)idxp[i * 3 + 2] + Verts.ElemsPerKey * key], Verts.Elems[(int)idxp[i * 3 + 2] + Verts.ElemsPerKey * key2], time);
      boxes[i].Reset();
      boxes[i].GrowVec3(V0);
      boxes[i].GrowVec3(V1);
      boxes[i].GrowVec3(V2);
      for (int c = 0; c < 3; c++) centroids[i][c] = (V0[c] + V1[c] + V2[c]) / 3;
      idxs[i] = (int32_t)i;
    }
    accel.mqbvh.Nodes = BuildAccelMotion(boxes.data(), centroids.data(), idxs.data(), facecount, 16);
    accel.idx = idxs;
    initMotionBoxes();
    return;
  }

  for (int i = 0; i < facecount; i++) {
    Vec3 V0 = Verts.Elems[(int)idxp[i * 3 + 0]];
    Vec3 V1 = Verts.Elems[(int)idxp[i * 3 + 1]];
    Vec3 V2 = Verts.Elems[(int)idxp[i * 3 + 2]];
    boxes[i].Reset();
    boxes[i].GrowVec3(V0);
    boxes[i].GrowVec3(V1);
    boxes[i].GrowVec3(V2);
    for (int c = 0; c < 3; c++) centroids[i][c] = (V0[c] + V1[c] + V2[c]) / 3;
    idxs[i] = (int32_t)i;
  }

  accel.qbvh = BuildAccel(boxes.data(), centroids.data(), idxs.data(), facecount, 16, &bounds);
  accel.idx = idxs;

  std::vector<uint32_t> nidxp(idxp.size());
  for (size_t i = 0; i < idxs.size(); i++) {
    nidxp[i * 3 + 0] = idxp[idxs[i] * 3 + 0];
    nidxp[i * 3 + 1] = idxp[idxs[i] * 3 + 1];
    nidxp[i * 3 + 2] = idxp[idxs[i] * 3 + 2];
  }
  if (!shaderidx.empty()) {
    std::vector<uint8_t> ns(shaderidx.size());
    for (size_t i = 0; i < idxs.size(); i++) ns[i] = shaderidx[idxs[i]];
    shaderidx = ns;
  }
  if (!normalidx.empty()) {
    std::vector<uint32_t> nn(normalidx.size());
    for (size_t i = 0; i < idxs.size(); i++) {
      nn[i * 3 + 0] = normalidx[idxs[i] * 3 + 0];
      nn[i * 3 + 1] = normalidx[idxs[i] * 3 + 1];
      nn[i * 3 + 2] = normalidx[idxs[i] * 3 + 2];
    }
    normalidx = nn;
  }
  if (!uvtriidx.empty()) {  // buildqbvh.go:106-116
    std::vector<uint32_t> uu(uvtriidx.size());
    for (size_t i = 0; i < idxs.size(); i++) {
      uu[i * 3 + 0] = uvtriidx[idxs[i] * 3 + 0];
      uu[i * 3 + 1] = uvtriidx[idxs[i] * 3 + 1];
      uu[i * 3 + 2] = uvtriidx[idxs[i] * 3 + 2];
    }
    uvtriidx = uu;
  }
  idxp = nidxp;
}

// builtin/geom/polymesh/buildqbvh.go:148-169
void PolyMesh::initMotionBoxes() {
  if (Verts.MotionKeys < 2) throw std::runtime_error("initMotionBoxes: can't init with < 2 motion keys");
  accel.mqbvh.Boxes.assign(Verts.MotionKeys, {});
  BoundingBox fullbox;
  fullbox.Reset();
  for (int k = 0; k < Verts.MotionKeys; k++) {
    accel.mqbvh.Boxes[k].assign(accel.mqbvh.Nodes.size(), MotionNodeBoxes{});
    BoundingBox box = initMotionBoxesRec(k, 0);
    motionBounds.push_back(box);
    fullbox.GrowBox(box);
  }
  bounds = fullbox;
}

// builtin/geom/polymesh/buildqbvh.go:171-212
BoundingBox PolyMesh::initMotionBoxesRec(int key, int32_t node) {
  BoundingBox nodebox;
  nodebox.Reset();
  for (int k = 0; k < 4; k++) {
    int32_t ch = accel.mqbvh.Nodes[node].Children[k];
    if (ch < 0) {
      if (ch == -1) continue;
      BoundingBox box;
      box.Reset();
      int leafBase = LeafBase(ch);
      int leafCount = LeafCount(ch);
      for (int i = leafBase; i < leafBase + leafCount; i++) {
        int faceidx = (int)accel.idx[i];
        box.GrowVec3(Verts.Elems[(int)idxp[faceidx * 3 + 0] + (Verts.ElemsPerKey * key)]);
        box.GrowVec3(Verts.Elems[(int)idxp[faceidx * 3 + 1] + (Verts.ElemsPerKey * key)]);
        box.GrowVec3(Verts.Elems[(int)idxp[faceidx * 3 + 2] + (Verts.ElemsPerKey * key)]);
      }
      accel.mqbvh.Boxes[key][node].SetBounds(k, box);
      nodebox.GrowBox(box);
    } else {
      BoundingBox box = initMotionBoxesRec(key, ch);
      accel.mqbvh.Boxes[key][node].SetBounds(k, box);
      nodebox.GrowBox(box);
    }
  }
  return nodebox;
}

// builtin/geom/polymesh/bounds.go:25-53
BoundingBox PolyMesh::Bounds(float time) const {
  if (!accel.qbvh.empty()) return bounds;
  float k = time * (float)((int)motionBounds.size() - 1);
  float t = k - Floor(k);
  int key = (int)Floor(k);
  int key2 = (int)Ceil(k);
  return BoundingBoxLerp(motionBounds[key], motionBounds[key2], t);
}

// builtin/geom/polymesh/trace.go:15-104, the object Transform path included exactly as written: the ray goes to object space
// (:22-55), and AFTER the traversal — hit or miss — the context's Transform / InvTransform are overwritten and sg.P, sg.N are
// re-derived from whatever sg.Po / sg.N currently hold (:62-74, :88-101). Bounds() ignores the transform (bounds.go:26,
// `if false && ...`), so the scene tree culls the mesh by its UNtransformed box. DESIGN.md quirk q.
bool PolyMesh::Trace(Ray* ray, ShaderContext* sg) {
  Vec3 Rp, Rd, Rdinv;
  float S[3] = {0, 0, 0};
  int32_t Kx = 0, Ky = 0, Kz = 0;
  Matrix4 transform, invTransform;
  const bool xf = !Transform.empty();
  if (xf) {
    Rp = ray->P; Rd = ray->D; Rdinv = ray->Dinv;
    S[0] = ray->S[0]; S[1] = ray->S[1]; S[2] = ray->S[2];
    Kx = ray->Kx; Ky = ray->Ky; Kz = ray->Kz;
    if (Transform.size() > 1) {
      float k = ray->Time * (float)((int)Transform.size() - 1);
      float time = k - Floor(k);
      int key = (int)Floor(k), key2 = (int)Ceil(k);
      transform = TransformDecompToMatrix4(TransformDecompLerp(transformSRT[key], transformSRT[key2], time));
    } else {
      transform = Transform[0];
    }
    Matrix4Inverse(transform, &invTransform);
    ray->P = Matrix4MulPoint(invTransform, Rp);
    ray->D = Matrix4MulVec(invTransform, Rd);
    ray->Setup();
  }
  bool hit;
  if (!accel.qbvh.empty()) {
    hit = QTrace(accel.qbvh, this, ray, sg);
  } else {
    float k = ray->Time * (float)(Verts.MotionKeys - 1);
    float time = k - Floor(k);
    int key = (int)Floor(k);
    int key2 = (int)Ceil(k);
    hit = QTraceMotion(accel.mqbvh, time, key, key2, this, ray, sg);
  }
  if (xf) {
    ray->P = Rp; ray->D = Rd; ray->Dinv = Rdinv;
    ray->S[0] = S[0]; ray->S[1] = S[1]; ray->S[2] = S[2];
    ray->Kx = Kx; ray->Ky = Ky; ray->Kz = Kz;
    sg->Transform = transform;       // hit or not (trace.go:70-73)
    sg->InvTransform = invTransform;
    sg->transformSet = true;
    sg->P = Matrix4MulPoint(transform, sg->Po);
    sg->N = Matrix4MulVec(Matrix4Transpose(invTransform), sg->N);
  }
  return hit;
}

// The watertight edge functions of trace.go:127-155 for one triangle; returns false if rejected.
// Shared by the static (bias test `<`, eps+RayBias) and motion (`<=`, RayBias) variants.
static inline bool triTest(const Ray* ray, const Vec3& P0, const Vec3& P1, const Vec3& P2, float biasTerm, bool motionBias,
                           float* U, float* V, float* W, float* Tout) {
  const int Kx = ray->Kx, Ky = ray->Ky, Kz = ray->Kz;
  float AKz = P0[Kz] - ray->P[Kz];
  float BKz = P1[Kz] - ray->P[Kz];
  float CKz = P2[Kz] - ray->P[Kz];
  float fU, fV, fW;
  {
    float Cx = (P2[Kx] - ray->P[Kx]) - ray->S[0] * CKz;
    float By = (P1[Ky] - ray->P[Ky]) - ray->S[1] * BKz;
    float Cy = (P2[Ky] - ray->P[Ky]) - ray->S[1] * CKz;
    float Bx = (P1[Kx] - ray->P[Kx]) - ray->S[0] * BKz;
    float Ax = (P0[Kx] - ray->P[Kx]) - ray->S[0] * AKz;
    float Ay = (P0[Ky] - ray->P[Ky]) - ray->S[1] * AKz;
    fU = Cx * By - Cy * Bx;
    fV = Ax * Cy - Ay * Cx;
    fW = Bx * Ay - By * Ax;
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
  }
  if ((fU < 0.0f || fV < 0.0f || fW < 0.0f) && (fU > 0.0f || fV > 0.0f || fW > 0.0f)) return false;
  float det = fU + fV + fW;
  if (det == 0.0f) return false;
  float T = ray->S[2] * (fU * AKz + fV * BKz + fW * CKz);
  uint32_t detSign = SignMask(det);
  if (!motionBias) {
    if (Xorf(T, detSign) < biasTerm * Xorf(det, detSign) || Xorf(T, detSign) > ray->Tclosest * Xorf(det, detSign)) return false;
  } else {
    if (Xorf(T, detSign) <= biasTerm * Xorf(det, detSign) || Xorf(T, detSign) > ray->Tclosest * Xorf(det, detSign)) return false;
  }
  float rcpDet = 1.0f / det;
  *U = fU * rcpDet;
  *V = fV * rcpDet;
  *W = fW * rcpDet;
  *Tout = T * rcpDet;
  return true;
}

// builtin/geom/polymesh/trace.go:108-194 (intersection) + :276-360, :504-515 (hit record).
// mesh.idxp is always non-nil after init(), so the `else` branch (:195-270) is dead and not restated.
bool PolyMesh::TraceElems(Ray* ray, ShaderContext* sg, int base, int count) {
  int32_t idx = -1;
  float U = 0, V = 0, W = 0;
  ray->TrisT += count;
  const float biasTerm = kEpsilonFloat32 + RayBias;
  for (int i = base; i < base + count; i++) {
    int i0 = (int)idxp[i * 3 + 0], i1 = (int)idxp[i * 3 + 1], i2 = (int)idxp[i * 3 + 2];
    float u, v, w, t;
    if (!triTest(ray, Verts.Elems[i0], Verts.Elems[i1], Verts.Elems[i2], biasTerm, false, &u, &v, &w, &t)) continue;
    U = u; V = v; W = w;
    ray->Tclosest = t;
    idx = (int32_t)i;
  }
  if (idx == -1) return false;

  int32_t i0 = (int32_t)idxp[idx * 3 + 0], i1 = (int32_t)idxp[idx * 3 + 1], i2 = (int32_t)idxp[idx * 3 + 2];
  const Vec3& E0 = Verts.Elems[i0];
  const Vec3& E1 = Verts.Elems[i1];
  const Vec3& E2 = Verts.Elems[i2];

  float xAbsSum = Abs(U * E0[0]) + Abs(V * E1[0]) + Abs(W * E2[0]);
  float yAbsSum = Abs(U * E0[1]) + Abs(V * E1[1]) + Abs(W * E2[1]);
  float zAbsSum = Abs(U * E0[2]) + Abs(V * E1[2]) + Abs(W * E2[2]);

  uint8_t shaderIdx = 0;
  if (!shaderidx.empty()) shaderIdx = shaderidx[idx];
  sg->shader = shader.empty() ? nullptr : shader[shaderIdx];

  float e00 = E1[0] - E0[0], e01 = E1[1] - E0[1], e02 = E1[2] - E0[2];
  float e10 = E2[0] - E0[0], e11 = E2[1] - E0[1], e12 = E2[2] - E0[2];

  sg->Ng[0] = e01 * e12 - e02 * e11;
  sg->Ng[1] = e02 * e10 - e00 * e12;
  sg->Ng[2] = e00 * e11 - e01 * e10;
  sg->Ng = Vec3Normalize(sg->Ng);

  Vec3 N;  // the interpolated normal BEFORE normalisation (trace.go:326-336), used by the normal differentials
  if (!Normals.Elems.empty()) {
    for (int k = 0; k < 3; k++)
      sg->N[k] = U * Normals.Elems[normalidx[(idx * 3) + 0]][k] + V * Normals.Elems[normalidx[(idx * 3) + 1]][k] + W * Normals.Elems[normalidx[(idx * 3) + 2]][k];
    N = sg->N;
    sg->N = Vec3Normalize(sg->N);
  } else {
    N = sg->Ng;
    sg->N = sg->Ng;
  }

  float d = Gamma(7) * xAbsSum * Abs(sg->Ng[0]) + Gamma(7) * yAbsSum * Abs(sg->Ng[1]) + Gamma(7) * zAbsSum * Abs(sg->Ng[2]);
  sg->Poffset = Vec3Scale(d, sg->Ng);
  sg->Bu = U;
  sg->Bv = V;
  sg->Bw = W;
  for (int k = 0; k < 3; k++) sg->P[k] = U * E0[k] + V * E1[k] + W * E2[k];
  sg->Po = sg->P;

  // trace.go:350-358
  if (!UV.empty()) {
    const float* t0 = UV[uvtriidx[(idx * 3) + 0]].v;
    const float* t1 = UV[uvtriidx[(idx * 3) + 1]].v;
    const float* t2 = UV[uvtriidx[(idx * 3) + 2]].v;
    sg->U = U * t0[0] + V * t1[0] + W * t2[0];
    sg->V = U * t0[1] + V * t1[1] + W * t2[1];
  } else {
    sg->U = U;
    sg->V = V;
  }
  ray->DifferentialTransfer(sg);  // trace.go:360

  // trace.go:362-502: barycentric planes -> d(alpha,beta,gamma)/dx,dy -> normal and texture-coordinate differentials
  {
    auto sq = [](float x) { return x * x; };
    const Vec3& Ng = sg->Ng;
    float nalphax = Ng[1] * (E2[2] - E1[2]) - Ng[2] * (E2[1] - E1[1]);
    float nalphay = Ng[2] * (E2[0] - E1[0]) - Ng[0] * (E2[2] - E1[2]);
    float nalphaz = Ng[0] * (E2[1] - E1[1]) - Ng[1] * (E2[0] - E1[0]);
    float q = Sqrt(sq(nalphax) + sq(nalphay) + sq(nalphaz));
    nalphax /= q; nalphay /= q; nalphaz /= q;
    float nalphad = -E1[0] * nalphax - E1[1] * nalphay - E1[2] * nalphaz;
    float l = E0[0] * nalphax + E0[1] * nalphay + E0[2] * nalphaz + nalphad;
    nalphax /= l; nalphay /= l; nalphaz /= l; nalphad /= l;

    float nbetax = Ng[1] * (E2[2] - E0[2]) - Ng[2] * (E2[1] - E0[1]);
    float nbetay = Ng[2] * (E2[0] - E0[0]) - Ng[0] * (E2[2] - E0[2]);
    float nbetaz = Ng[0] * (E2[1] - E0[1]) - Ng[1] * (E2[0] - E0[0]);
    q = Sqrt(sq(nbetax) + sq(nbetay) + sq(nbetaz));
    nbetax /= q; nbetay /= q; nbetaz /= q;
    float nbetad = -E0[0] * nbetax - E0[1] * nbetay - E0[2] * nbetaz;
    l = nbetax * E1[0] + nbetay * E1[1] + nbetaz * E1[2] + nbetad;
    nbetax /= l; nbetay /= l; nbetaz /= l; nbetad /= l;

    float ngammax = Ng[1] * (E1[2] - E0[2]) - Ng[2] * (E1[1] - E0[1]);
    float ngammay = Ng[2] * (E1[0] - E0[0]) - Ng[0] * (E1[2] - E0[2]);
    float ngammaz = Ng[0] * (E1[1] - E0[1]) - Ng[1] * (E1[0] - E0[0]);
    q = Sqrt(sq(ngammax) + sq(ngammay) + sq(ngammaz));
    ngammax /= q; ngammay /= q; ngammaz /= q;
    float ngammad = -E0[0] * ngammax - E0[1] * ngammay - E0[2] * ngammaz;
    l = ngammax * E2[0] + ngammay * E2[1] + ngammaz * E2[2] + ngammad;
    ngammax /= l; ngammay /= l; ngammaz /= l; ngammad /= l;
    (void)nalphad; (void)nbetad; (void)ngammad;

    const float alphax = nalphax * sg->DdPdx[0] + nalphay * sg->DdPdx[1] + nalphaz * sg->DdPdx[2];
    const float betax = nbetax * sg->DdPdx[0] + nbetay * sg->DdPdx[1] + nbetaz * sg->DdPdx[2];
    const float gammax = ngammax * sg->DdPdx[0] + ngammay * sg->DdPdx[1] + ngammaz * sg->DdPdx[2];
    const float alphay = nalphax * sg->DdPdy[0] + nalphay * sg->DdPdy[1] + nalphaz * sg->DdPdy[2];
    const float betay = nbetax * sg->DdPdy[0] + nbetay * sg->DdPdy[1] + nbetaz * sg->DdPdy[2];
    const float gammay = ngammax * sg->DdPdy[0] + ngammay * sg->DdPdy[1] + ngammaz * sg->DdPdy[2];

    Vec3 dndx = V3(0, 0, 0), dndy = V3(0, 0, 0);
    if (!Normals.Elems.empty()) {
      const Vec3& n0 = Normals.Elems[normalidx[(idx * 3) + 0]];
      const Vec3& n1 = Normals.Elems[normalidx[(idx * 3) + 1]];
      const Vec3& n2 = Normals.Elems[normalidx[(idx * 3) + 2]];
      for (int k = 0; k < 3; k++) {
        dndx[k] = alphax * n0[k] + betax * n1[k] + gammax * n2[k];
        dndy[k] = alphay * n0[k] + betay * n1[k] + gammay * n2[k];
      }
    }
    sg->DdNdx = Vec3Sub(Vec3Scale(Vec3Dot(N, N), dndx), Vec3Scale(Vec3Dot(N, dndx), N));
    sg->DdNdx = Vec3Scale(1 / (Vec3Dot(N, N) * Sqrt(Vec3Dot(N, N))), sg->DdNdx);
    sg->DdNdy = Vec3Sub(Vec3Scale(Vec3Dot(N, N), dndy), Vec3Scale(Vec3Dot(N, dndy), N));
    sg->DdNdy = Vec3Scale(1 / (Vec3Dot(N, N) * Sqrt(Vec3Dot(N, N))), sg->DdNdy);

    if (!UV.empty()) {
      const float* t0 = UV[uvtriidx[(idx * 3) + 0]].v;
      const float* t1 = UV[uvtriidx[(idx * 3) + 1]].v;
      const float* t2 = UV[uvtriidx[(idx * 3) + 2]].v;
      for (int k = 0; k < 2; k++) {
        sg->Dduvdx[k] = alphax * t0[k] + betax * t1[k] + gammax * t2[k];
        sg->Dduvdy[k] = alphay * t0[k] + betay * t1[k] + gammay * t2[k];
      }
    } else {  // trace.go:495-501, as written (U above is the FIRST barycentric, these follow the second and third)
      sg->Dduvdx[0] = alphax * 0 + betax * 1 + gammax * 0;
      sg->Dduvdx[1] = alphax * 0 + betax * 0 + gammax * 1;
      sg->Dduvdy[0] = alphay * 0 + betay * 1 + gammay * 0;
      sg->Dduvdy[1] = alphay * 0 + betay * 0 + gammay * 1;
    }
  }

  // trace.go:504-511
  Vec3 axisu = Vec3Sub(V3(1, 0, 0), Vec3Scale(Vec3Dot(V3(1, 0, 0), sg->Ng), sg->Ng));
  if (Vec3Length2(axisu) < 0.1f || Abs(Vec3Dot(axisu, sg->Ng)) > 0.3f)
    axisu = Vec3Sub(V3(0, 0, 1), Vec3Scale(Vec3Dot(V3(0, 0, 1), sg->Ng), sg->Ng));
  sg->DdPdu = Vec3Normalize(axisu);
  sg->DdPdv = Vec3Cross(sg->Ng, sg->DdPdu);
  sg->ElemID = (uint32_t)idx;
  return true;
}

// Brute force: every triangle in leaf order through the same routine (one "leaf" of facecount tris).
bool PolyMesh::TraceBrute(Ray* ray, ShaderContext* sg) {
  bool hit = false;
  if (Verts.MotionKeys > 1) {
    float k = ray->Time * (float)(Verts.MotionKeys - 1);
    float time = k - Floor(k);
    int key = (int)Floor(k), key2 = (int)Ceil(k);
    // iterate in leaf order so that "fixed" mode visits faces accel.idx[i]
    return TraceMotionElems(time, key, key2, ray, sg, 0, facecount);
  }
  for (int b = 0; b < facecount; b += 16) {
    int c = facecount - b < 16 ? facecount - b : 16;
    if (TraceElems(ray, sg, b, c)) hit = true;
  }
  return hit;
}

// builtin/geom/polymesh/trace.go:520-719
bool PolyMesh::TraceMotionElems(float time, int key, int key2, Ray* ray, ShaderContext* sg, int base, int count) {
  int32_t idx = -1;
  float U = 0, V = 0, W = 0;
  ray->TrisT += count;
  for (int i = base; i < base + count; i++) {
    int32_t faceidx = (int32_t)i;
    int i0, i1, i2;
    if (ref_compat_motion) {
      // trace.go:532-537: idxp != nil, so face i is tested although the leaf box bounds face accel.idx[i] (quirk b)
      i0 = (int)idxp[i * 3 + 0]; i1 = (int)idxp[i * 3 + 1]; i2 = (int)idxp[i * 3 + 2];
    } else {
      faceidx = accel.idx[i];
      i0 = (int)idxp[faceidx * 3 + 0]; i1 = (int)idxp[faceidx * 3 + 1]; i2 = (int)idxp[faceidx * 3 + 2];
    }
    Vec3 V0 = Vec3Lerp(Verts.Elems[i0 + Verts.ElemsPerKey * key], Verts.Elems[i0 + Verts.ElemsPerKey * key2], time);
    Vec3 V1 = Vec3Lerp(Verts.Elems[i1 + Verts.ElemsPerKey * key], Verts.Elems[i1 + Verts.ElemsPerKey * key2], time);
    Vec3 V2 = Vec3Lerp(Verts.Elems[i2 + Verts.ElemsPerKey * key], Verts.Elems[i2 + Verts.ElemsPerKey * key2], time);

    float u, v, w, t;
    if (!triTest(ray, V0, V1, V2, RayBias, true, &u, &v, &w, &t)) continue;
    U = u; V = v; W = w;
    ray->Tclosest = t;
    idx = faceidx;

    for (int k = 0; k < 3; k++) sg->Po[k] = U * V0[k] + V * V1[k] + W * V2[k];
    float xAbsSum = Abs(U * V0[0]) + Abs(V * V1[0]) + Abs(W * V2[0]);
    float yAbsSum = Abs(U * V0[1]) + Abs(V * V1[1]) + Abs(W * V2[1]);
    float zAbsSum = Abs(U * V0[2]) + Abs(V * V1[2]) + Abs(W * V2[2]);
    float e00 = V1[0] - V0[0], e01 = V1[1] - V0[1], e02 = V1[2] - V0[2];
    float e10 = V2[0] - V0[0], e11 = V2[1] - V0[1], e12 = V2[2] - V0[2];
    sg->Ng[0] = e01 * e12 - e02 * e11;
    sg->Ng[1] = e02 * e10 - e00 * e12;
    sg->Ng[2] = e00 * e11 - e01 * e10;
    sg->Ng = Vec3Normalize(sg->Ng);
    sg->DdPdu = V3(e00, e01, e02);
    sg->DdPdv = V3(e10, e11, e12);
    float d = Gamma(7) * xAbsSum * Abs(sg->Ng[0]) + Gamma(7) * yAbsSum * Abs(sg->Ng[1]) + Gamma(7) * zAbsSum * Abs(sg->Ng[2]);
    sg->Poffset = Vec3Scale(d, sg->Ng);
    sg->Bu = U;
    sg->Bv = V;
    sg->Bw = W;
  }
  if (idx == -1) return false;

  uint8_t shaderIdx = 0;
  if (!shaderidx.empty()) shaderIdx = shaderidx[idx];
  sg->shader = shader.empty() ? nullptr : shader[shaderIdx];
  sg->P = sg->Po;

  // trace.go:677-684: surface parameters only; the motion path transfers no differentials (Dduvdx/y stay 0)
  if (!UV.empty()) {
    const float* t0 = UV[uvtriidx[(idx * 3) + 0]].v;
    const float* t1 = UV[uvtriidx[(idx * 3) + 1]].v;
    const float* t2 = UV[uvtriidx[(idx * 3) + 2]].v;
    sg->U = U * t0[0] + V * t1[0] + W * t2[0];
    sg->V = U * t0[1] + V * t1[1] + W * t2[1];
  } else {
    sg->U = U;
    sg->V = V;
  }

  if (!Normals.Elems.empty() && Normals.MotionKeys == Verts.MotionKeys) {
    Vec3 N0 = Vec3Lerp(Normals.Elems[(int)((idx * 3) + 0) + Normals.ElemsPerKey * key], Normals.Elems[(int)((idx * 3) + 0) + Normals.ElemsPerKey * key2], time);
    Vec3 N1 = Vec3Lerp(Normals.Elems[(int)((idx * 3) + 1) + Normals.ElemsPerKey * key], Normals.Elems[(int)((idx * 3) + 1) + Normals.ElemsPerKey * key2], time);
    Vec3 N2 = Vec3Lerp(Normals.Elems[(int)((idx * 3) + 2) + Normals.ElemsPerKey * key], Normals.Elems[(int)((idx * 3) + 2) + Normals.ElemsPerKey * key2], time);
    for (int k = 0; k < 3; k++) sg->N[k] = U * N0[k] + V * N1[k] + W * N2[k];
    sg->N = Vec3Normalize(sg->N);
  } else {
    sg->N = sg->Ng;
  }
  sg->ElemID = (uint32_t)idx;
  return true;
}

// ---------------------------------------------------------------------------------------------
// builtin/geom/instance/instance.go
void Instance::PreRender() {  // :117-146
  for (const Matrix4& m : Transform) transformSRT.push_back(TransformDecompMatrix4(m));
  for (size_t i = 0; i < BMin.size(); i++) {
    BoundingBox box;  // zero value, NOT Reset(): the box always contains the origin (instance.go:124)
    for (int k = 0; k < 3; k++) box.b[0][k] = box.b[1][k] = 0.0f;
    box.GrowVec3(BMin[i]);
    box.GrowVec3(BMax[i]);
    bounds.push_back(box);
  }
}
TransformDecomp Instance::TimeKey(float time) const {  // :16-33
  if (transformSRT.size() > 1) {
    float k = time * (float)((int)transformSRT.size() - 1);
    float timeFrac = k - Floor(k);
    int key = (int)Floor(k);
    int key2 = (int)Ceil(k);
    return TransformDecompLerp(transformSRT[key], transformSRT[key2], timeFrac);
  }
  return transformSRT[0];
}
bool Instance::Trace(Ray* ray, ShaderContext* sg) {  // :73-114
  Vec3 Rp = ray->P, Rd = ray->D, Rdinv = ray->Dinv;
  float S[3] = {ray->S[0], ray->S[1], ray->S[2]};
  int32_t Kx = ray->Kx, Ky = ray->Ky, Kz = ray->Kz;
  Matrix4 transform = TransformDecompToMatrix4(TimeKey(ray->Time));
  Matrix4 invTransform;
  Matrix4Inverse(transform, &invTransform);
  ray->P = Matrix4MulPoint(invTransform, Rp);
  ray->D = Matrix4MulVec(invTransform, Rd);
  ray->Setup();
  bool hit = geom->Trace(ray, sg);
  ray->P = Rp;
  ray->D = Rd;
  ray->Dinv = Rdinv;
  ray->S[0] = S[0]; ray->S[1] = S[1]; ray->S[2] = S[2];
  ray->Kx = Kx; ray->Ky = Ky; ray->Kz = Kz;
  if (hit) {
    // "we know that this ray has hit this geom as closest point so ok to overwrite" — and nothing ever resets it: a later,
    // closer hit on an un-instanced geom is still shaded with this transform (kept; DESIGN.md quirk o)
    sg->Transform = transform;
    sg->InvTransform = invTransform;
    sg->transformSet = true;
  }
  return hit;
}

// ---------------------------------------------------------------------------------------------
// builtin/scene/scene.go:30-59
bool Scene::Trace(Ray* ray, ShaderContext* sg) {
  if (!qbvh.empty()) return QTrace(qbvh, this, ray, sg);
  float k = ray->Time * (float)((int)mqbvh.Boxes.size() - 1);
  float time = k - Floor(k);
  int key = (int)Floor(k);
  int key2 = (int)Ceil(k);
  return QTraceMotion(mqbvh, time, key, key2, this, ray, sg);
}

// builtin/scene/scene.go:61-78
bool Scene::TraceElems(Ray* ray, ShaderContext* sc, int base, int count) {
  bool hit = false;
  for (int i = base; i < base + count; i++) {
    if (geoms[i]->Trace(ray, sc)) {
      sc->geom = geoms[i];
      hit = true;
      if (ray->Type & RayTypeShadow) return true;
    }
  }
  return hit;
}
bool Scene::TraceMotionElems(float, int, int, Ray* ray, ShaderContext* sc, int base, int count) {
  return TraceElems(ray, sc, base, count);
}

// builtin/scene/scene.go:100-116
void Scene::LightsPrepare(ShaderContext* sg) {
  sg->Lights.clear();
  for (Light* l : lights) {
    if (l->GetGeom() != sg->geom) sg->Lights.push_back(l);
  }
}

// builtin/scene/scene.go:135-203
void Scene::initAccel() {
  std::vector<BoundingBox> boxes;
  std::vector<int32_t> indices;
  std::vector<Vec3> centroids;
  int maxKeys = 0;
  for (Geom* g : geoms) if (g->MotionKeys() > maxKeys) maxKeys = g->MotionKeys();

  if (maxKeys == 1) {
    for (size_t i = 0; i < geoms.size(); i++) {
      BoundingBox box = geoms[i]->Bounds(0);
      boxes.push_back(box);
      indices.push_back((int32_t)i);
      centroids.push_back(box.Centroid());
    }
    qbvh = BuildAccel(boxes.data(), centroids.data(), indices.data(), (int)indices.size(), 1, &bounds);
    std::vector<Geom*> ngeoms(indices.size());
    for (size_t i = 0; i < indices.size(); i++) ngeoms[i] = geoms[indices[i]];
    geoms = ngeoms;
  } else {
    for (size_t i = 0; i < geoms.size(); i++) {
      BoundingBox box = geoms[i]->Bounds(0.5f);
      boxes.push_back(box);
      indices.push_back((int32_t)i);
      centroids.push_back(box.Centroid());
    }
    mqbvh.Nodes = BuildAccelMotion(boxes.data(), centroids.data(), indices.data(), (int)indices.size(), 1);
    std::vector<Geom*> ngeoms(indices.size());
    for (size_t i = 0; i < indices.size(); i++) ngeoms[i] = geoms[indices[i]];
    geoms = ngeoms;
    bounds.Reset();
    initMotionBoxes(maxKeys);
  }
}

// builtin/scene/scene.go:207-226
void Scene::initMotionBoxes(int keys) {
  if (keys < 2) throw std::runtime_error("s.initMotionBoxes: can't init with < 2 motion keys");
  mqbvh.Boxes.assign(keys, {});
  BoundingBox fullbox;
  fullbox.Reset();
  for (int k = 0; k < keys; k++) {
    mqbvh.Boxes[k].assign(mqbvh.Nodes.size(), MotionNodeBoxes{});
    BoundingBox box = initMotionBoxesRec(k, 0);
    fullbox.GrowBox(box);
  }
  bounds = fullbox;
}

// builtin/scene/scene.go:228-268
BoundingBox Scene::initMotionBoxesRec(int key, int32_t node) {
  BoundingBox nodebox;
  nodebox.Reset();
  for (int k = 0; k < 4; k++) {
    int32_t ch = mqbvh.Nodes[node].Children[k];
    if (ch < 0) {
      if (ch == -1) continue;
      BoundingBox box;
      box.Reset();
      int leafBase = LeafBase(ch);
      int leafCount = LeafCount(ch);
      for (int i = leafBase; i < leafBase + leafCount; i++) {
        float time = (float)key / (float)((int)mqbvh.Boxes.size() - 1);
        box.GrowBox(geoms[i]->Bounds(time));
      }
      mqbvh.Boxes[key][node].SetBounds(k, box);
      nodebox.GrowBox(box);
    } else {
      BoundingBox box = initMotionBoxesRec(key, ch);
      mqbvh.Boxes[key][node].SetBounds(k, box);
      nodebox.GrowBox(box);
    }
  }
  return nodebox;
}

// ---------------------------------------------------------------------------------------------
// core/trace.go:26-35
bool TraceProbe(Ray* ray, ShaderContext* sg) {
  if (ray->Task->sharedRayCount) {  // stats.incRayCount / incShadowRayCount: global atomics (core/stats.go:26-33)
    ray->Task->sharedRayCount->fetch_add(1);
    if (ray->Type & RayTypeShadow) ray->Task->sharedShadowRayCount->fetch_add(1);
  } else {
    ray->Task->rayCount++;
    if (ray->Type & RayTypeShadow) ray->Task->shadowRayCount++;
  }
  return ray->Task->scene->Trace(ray, sg);
}

// core/trace.go:40-83
bool Trace(Ray* ray, TraceSample* samp) {
  ShaderContext sg;
  sg.Ro = ray->P;
  sg.Rd = ray->D;
  sg.Level = ray->Level;
  sg.Lambda = ray->Lambda;
  sg.I = ray->I;
  sg.Time = ray->Time;
  sg.task = ray->Task;
  sg.Scramble[0] = ray->Scramble[0];
  sg.Scramble[1] = ray->Scramble[1];
  if (ray->Task) {  // Image: image (trace.go:55) — the package-level constants the camera wrote (camera.go:316-317)
    sg.PixelDelta[0] = ray->Task->PixelDelta[0];
    sg.PixelDelta[1] = ray->Task->PixelDelta[1];
  }

  if (TraceProbe(ray, &sg)) {
    if (sg.shader == nullptr) return false;
    ray->DifferentialTransfer(&sg);  // trace.go:67 (a second time for meshes: same inputs, same result)
    sg.ApplyTransform();
    sg.shader->Eval(&sg);
    if (samp != nullptr) {
      samp->Colour = sg.OutRGB;
      samp->Point = sg.P;
      samp->ElemID = sg.ElemID;
      samp->geom = sg.geom;
    }
    return true;
  }
  return false;
}

// core/shader.go:129-135 with Transform = InvTransform = identity (object transforms out of scope):
// the matrix products are exact, the re-normalisations are kept.
void ShaderContext::ApplyTransform() {
  if (transformSet) {  // core/shader.go:129-135
    P = Matrix4MulPoint(Transform, Po);
    N = Vec3Normalize(Matrix4MulVec(Matrix4Transpose(InvTransform), N));
    Ng = Vec3Normalize(Matrix4MulVec(Matrix4Transpose(InvTransform), Ng));
    DdPdu = Vec3Normalize(Matrix4MulVec(Transform, DdPdu));
    DdPdv = Vec3Normalize(Matrix4MulVec(Transform, DdPdv));
    return;
  }
  // Transform == identity (core/trace.go:57-58): the products with 0/1 are exact, only the re-normalisations remain
  P = Po;
  N = Vec3Normalize(N);
  Ng = Vec3Normalize(Ng);
  DdPdu = Vec3Normalize(DdPdu);
  DdPdv = Vec3Normalize(DdPdv);
}

// core/shader.go:139-162
Vec3 ShaderContext::OffsetP(int dir) const {
  Vec3 pofs = dir < 0 ? Vec3Neg(Poffset) : Poffset;
  Vec3 po = Vec3Add(P, pofs);
  for (int i = 0; i < 3; i++) {
    if (pofs[i] > 0) po[i] = NextFloatUp(po[i]);
    else if (pofs[i] < -0.0f) po[i] = NextFloatDown(po[i]);
  }
  return po;
}

// core/shader.go:166-176
void ShaderContext::LightsPrepare() {
  task->scene->LightsPrepare(this);
  Lidx = -1;
  Lsamples.clear();
}

// core/shader.go:179-197
bool ShaderContext::NextLight() {
  Lidx++;
  if (Lidx < (int)Lights.size()) {
    Lp = Lights[Lidx];
    NSamples = Lp->NumSamples(this);
    if (Level > 0) NSamples = 1;
    return true;
  }
  return false;
}

// core/shader.go:203-402
RGB ShaderContext::EvaluateLightSamples(BSDF* bsdf) {
  std::vector<BSDFSample> bsdfSamples;
  RGB col;

  if (NSamples > 1) {
    for (int i = 0; i < NSamples / 2; i++) {
      uint64_t idx = (uint64_t)(I * (NSamples / 2) + i);
      double r0 = VanDerCorput(idx, Scramble[0]);
      double r1 = Sobol(idx, Scramble[1]);
      Vec3 omegaO = Vec3Normalize(bsdf->Sample(r0, r1));
      BSDFSample s{};
      s.D = omegaO;
      s.Pdf = bsdf->PDF(omegaO);
      if (s.Pdf <= 0) continue;
      if (Lp->ValidSample(this, &s)) bsdfSamples.push_back(s);
    }
    Lsamples.clear();
    Lp->SampleArea(this, NSamples / 2);

    int nBSDFSamples = NSamples / 2;
    int nLightSamples = NSamples / 2;
    if (Lsamples.empty()) nLightSamples = 0;
    if (bsdfSamples.empty()) nBSDFSamples = 0;
    int totalSamples = nBSDFSamples + nLightSamples;
    if (totalSamples == 0) return RGB{};

    Ray ray;
    ray.Task = task;
    ShaderContext chsc;
    chsc.task = task;

    for (const LightSample& ls : Lsamples) {
      if (Vec3Dot(ls.Ld, N) <= 0) continue;
      if (Vec3Dot(ls.Ld, Ng) < 0)
        ray.Init(RayTypeShadow, OffsetP(-1), Vec3Scale(ls.Ldist * (1.0f - ShadowRayEpsilon), ls.Ld), 1.0f, 0, this);
      else
        ray.Init(RayTypeShadow, OffsetP(1), Vec3Scale(ls.Ldist * (1.0f - ShadowRayEpsilon), ls.Ld), 1.0f, 0, this);
      if (!TraceProbe(&ray, &chsc)) {
        Spectrum rho = bsdf->Eval(ls.Ld);
        rho.Mul(ls.Liu);
        float p_hat = (float)nBSDFSamples * (float)bsdf->PDF(ls.Ld) / (float)totalSamples;
        p_hat += (float)nLightSamples * ls.Pdf / (float)totalSamples;
        rho.Scale(1.0f / p_hat);
        RGB rgb = rho.ToRGB();
        for (int k = 0; k < 3; k++) if (rgb[k] < 0) rgb[k] = 0;
        col.Add(rgb);
      }
    }
    for (const BSDFSample& bs : bsdfSamples) {
      if (Vec3Dot(bs.Ld, N) <= 0) continue;
      if (Vec3Dot(bs.Ld, Ng) < 0)
        ray.Init(RayTypeShadow, OffsetP(-1), Vec3Scale(bs.Ldist * (1.0f - ShadowRayEpsilon), bs.Ld), 1.0f, 0, this);
      else
        ray.Init(RayTypeShadow, OffsetP(1), Vec3Scale(bs.Ldist * (1.0f - ShadowRayEpsilon), bs.Ld), 1.0f, 0, this);
      if (!TraceProbe(&ray, &chsc)) {
        Spectrum rho = bsdf->Eval(bs.Ld);
        rho.Mul(bs.Liu);
        float p_hat = (float)nBSDFSamples * (float)bs.Pdf / (float)totalSamples;
        p_hat += (float)nLightSamples * bs.PdfLight / (float)totalSamples;
        rho.Scale(1.0f / p_hat);
        RGB rgb = rho.ToRGB();
        for (int k = 0; k < 3; k++) if (rgb[k] < 0) rgb[k] = 0;
        col.Add(rgb);
      }
    }
    col.Scale(1.0f / (float)totalSamples);
  } else {
    Lsamples.clear();
    Lp->SampleArea(this, NSamples);
    for (const LightSample& ls : Lsamples) {
      Ray ray;
      ray.Task = task;
      ShaderContext chsc;
      chsc.task = task;
      if (Vec3Dot(ls.Ld, N) <= 0) continue;
      if (Vec3Dot(ls.Ld, Ng) < 0)
        ray.Init(RayTypeShadow, OffsetP(-1), Vec3Scale(ls.Ldist * (1.0f - ShadowRayEpsilon), ls.Ld), 1.0f, 0, this);
      else
        ray.Init(RayTypeShadow, OffsetP(1), Vec3Scale(ls.Ldist * (1.0f - ShadowRayEpsilon), ls.Ld), 1.0f, 0, this);
      if (!TraceProbe(&ray, &chsc)) {
        Spectrum rho = bsdf->Eval(ls.Ld);
        rho.Mul(ls.Liu);
        rho.Scale(1.0f / ls.Pdf);
        RGB rgb = rho.ToRGB();
        col.Add(rgb);
      }
    }
  }
  return col;
}

}  // namespace orc
