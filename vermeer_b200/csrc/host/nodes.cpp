// Host mirror of the reference's PreRender pipeline — see nodes.h for the interface map.
#include "nodes.h"

#include <atomic>
#include <thread>

#include <cmath>
#include <mutex>

#include "builder.h"

namespace vh {

// ---- registry (nodes/register.go) -------------------------------------------------------------
namespace {
std::map<std::string, CreateFn>& registry() {
  static std::map<std::string, CreateFn> r;
  return r;
}
std::mutex& registry_mu() {
  static std::mutex m;
  return m;
}
}  // namespace

int Register(const std::string& name, CreateFn create) {
  std::lock_guard<std::mutex> l(registry_mu());
  if (registry().count(name)) return -1;  // "node type already registered"
  registry()[name] = std::move(create);
  return 0;
}
std::unique_ptr<Node> CreateNode(const std::string& name) {
  std::lock_guard<std::mutex> l(registry_mu());
  auto it = registry().find(name);
  if (it == registry().end()) return nullptr;
  return it->second();
}
std::vector<std::string> RegisteredNames() {
  std::lock_guard<std::mutex> l(registry_mu());
  std::vector<std::string> v;
  for (auto& kv : registry()) v.push_back(kv.first);
  return v;
}

// the builtin nodes register themselves like the reference's init() functions
// (polymesh.go:120-122, std.go:318-324, triangle.go:350-356, camera.go:325-335, parser.go:76)
namespace {
struct RegisterBuiltins {
  RegisterBuiltins() {
    Register("Globals", [] { return std::unique_ptr<Node>(new Globals()); });
    Register("Camera", [] { return std::unique_ptr<Node>(new Camera()); });
    Register("PolyMesh", [] { return std::unique_ptr<Node>(new PolyMesh()); });
    Register("ShaderStd", [] { return std::unique_ptr<Node>(new ShaderStd()); });
    Register("Include", [] { return std::unique_ptr<Node>(new Include()); });  // misc/include.go:30-36
    Register("DebugShader", [] { return std::unique_ptr<Node>(new DebugShader()); });  // debug.go:51-57
    Register("TriLight", [] { return std::unique_ptr<Node>(new TriLight()); });
    Register("DiskLight", [] { return std::unique_ptr<Node>(new DiskLight()); });      // disk.go:263-269
    Register("SphereLight", [] { return std::unique_ptr<Node>(new SphereLight()); });  // sphere.go:297-303
    Register("GeomInstance", [] { return std::unique_ptr<Node>(new GeomInstance()); });  // instance.go:163-165
    Register("Sphere", [] { return std::unique_ptr<Node>(new SphereGeom()); });        // geom/sphere/sphere.go:78-86
    // builtin/driver/outputfloat.go:44-50, outputhdr.go:59-65
    Register("OutputFloat", [] { OutputNode* o = new OutputNode(); o->Filename = "out.float"; return std::unique_ptr<Node>(o); });
    Register("OutputHDR", [] { OutputNode* o = new OutputNode(); o->Filename = "out.hdr"; o->hdr = true; return std::unique_ptr<Node>(o); });
    // builtin/filter/filter.go:14-26
    Register("AiryFilter", [] { PixelFilter* f = new PixelFilter(); f->kind = 1; f->Res = 49; f->Width = 6; f->Peak = 4; return std::unique_ptr<Node>(f); });
    Register("GaussianFilter", [] { PixelFilter* f = new PixelFilter(); f->kind = 2; f->Res = 17; f->Width = 2; f->Peak = 0; return std::unique_ptr<Node>(f); });
  }
} g_register_builtins;
}  // namespace

// ---- matrices -----------------------------------------------------------------------------------
void m4_cofactors(const float* m, float* inv) {  // math/matrix4.go:113-132, term order kept
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
}
bool m4_inverse(const M4& a, M4* out) {
  float inv[16];
  m4_cofactors(a.m, inv);
  float det = a.m[0] * inv[0] + a.m[1] * inv[4] + a.m[2] * inv[8] + a.m[3] * inv[12];
  if (det == 0.0f) return false;
  det = 1.0f / det;
  for (int i = 0; i < 16; i++) out->m[i] = inv[i] * det;
  return true;
}
float m4_det(const M4& a) {
  float inv[16];
  m4_cofactors(a.m, inv);
  return a.m[0] * inv[0] + a.m[1] * inv[4] + a.m[2] * inv[8] + a.m[3] * inv[12];
}

namespace {
struct Quat { float X, Y, Z, W; };

// math/matrix4.go:345-363 — Higham iteration Q <- (Q + Q^-T)/2, at most 10 rounds, one-sided tolerance test (:331-339)
bool polar_factor(const M4& m, M4* out) {
  M4 Q = m;
  for (int it = 0; it < 10; it++) {
    M4 Qinv;
    if (!m4_inverse(Q, &Qinv)) return false;
    M4 Qn = m4_scale(.5f, m4_add(Q, m4_transpose(Qinv)));
    bool same = true;
    for (int i = 0; i < 16; i++)
      if (Qn.m[i] - Q.m[i] > 0.000001f) { same = false; break; }
    if (same) { *out = Qn; return true; }
    Q = Qn;
  }
  return false;
}
Quat to_quat(const M4& m) {  // math/quat.go:124-154
  Quat q;
  const float tr = m.at(0, 0) + m.at(1, 1) + m.at(2, 2);
  if (tr > 0.0f) {
    const float S = sqrtf(tr + 1.0f) * 2;
    q.W = 0.25f * S; q.X = (m.at(2, 1) - m.at(1, 2)) / S; q.Y = (m.at(0, 2) - m.at(2, 0)) / S; q.Z = (m.at(1, 0) - m.at(0, 1)) / S;
  } else if ((m.at(0, 0) > m.at(1, 1)) && (m.at(0, 0) > m.at(2, 2))) {
    const float S = sqrtf(1.0f + m.at(0, 0) - m.at(1, 1) - m.at(2, 2)) * 2;
    q.W = (m.at(2, 1) - m.at(1, 2)) / S; q.X = 0.25f * S; q.Y = (m.at(0, 1) + m.at(1, 0)) / S; q.Z = (m.at(0, 2) + m.at(2, 0)) / S;
  } else if (m.at(1, 1) > m.at(2, 2)) {
    const float S = sqrtf(1.0f + m.at(1, 1) - m.at(0, 0) - m.at(2, 2)) * 2;
    q.W = (m.at(0, 2) - m.at(2, 0)) / S; q.X = (m.at(0, 1) + m.at(1, 0)) / S; q.Y = 0.25f * S; q.Z = (m.at(1, 2) + m.at(2, 1)) / S;
  } else {
    const float S = sqrtf(1.0f + m.at(2, 2) - m.at(0, 0) - m.at(1, 1)) * 2;
    q.W = (m.at(1, 0) - m.at(0, 1)) / S; q.X = (m.at(0, 2) + m.at(2, 0)) / S; q.Y = (m.at(1, 2) + m.at(2, 1)) / S; q.Z = 0.25f * S;
  }
  return q;
}
M4 from_quat(Quat q) {  // math/quat.go:73-96
  M4 m{};
  const float n = 1.0f / sqrtf(q.X * q.X + q.Y * q.Y + q.Z * q.Z + q.W * q.W);
  const float x = q.X * n, y = q.Y * n, z = q.Z * n, w = q.W * n;
  m.set(0, 0, 1 - 2 * y * y - 2 * z * z); m.set(0, 1, 2 * x * y - 2 * w * z); m.set(0, 2, 2 * x * z + 2 * w * y);
  m.set(1, 0, 2 * x * y + 2 * w * z); m.set(1, 1, 1 - 2 * x * x - 2 * z * z); m.set(1, 2, 2 * y * z - 2 * w * x);
  m.set(2, 0, 2 * x * z - 2 * w * y); m.set(2, 1, 2 * y * z + 2 * w * x); m.set(2, 2, 1 - 2 * x * x - 2 * y * y);
  m.set(3, 3, 1.0f);
  return m;
}
Quat slerp(Quat a, Quat b, float t) {  // math/quat.go:30-66
  const float c = a.W * b.W + a.X * b.X + a.Y * b.Y + a.Z * b.Z;
  if (fabs32(c) >= 1.0f) return a;
  const float half = (float)std::acos((double)c);
  const float s = sqrtf(1.0f - c * c);
  if (fabs32(s) < 0.001f) return Quat{a.X * 0.5f + b.X * 0.5f, a.Y * 0.5f + b.Y * 0.5f, a.Z * 0.5f + b.Z * 0.5f, a.W * 0.5f + b.W * 0.5f};
  const float ra = (float)std::sin((double)((1 - t) * half)) / s;
  const float rb = (float)std::sin((double)(t * half)) / s;
  return Quat{a.X * ra + b.X * rb, a.Y * ra + b.Y * rb, a.Z * ra + b.Z * rb, a.W * ra + b.W * rb};
}
// math/animdecomp.go:21-63: M = T * R * S via polar decomposition
VgTransformSRT transform_decomp(M4 mtx) {
  VgTransformSRT d;
  const float sign = m4_det(mtx) >= 0.0f ? 1.0f : -1.0f;
  d.T[0] = mtx.at(0, 3); d.T[1] = mtx.at(1, 3); d.T[2] = mtx.at(2, 3);
  mtx.set(0, 3, 0); mtx.set(1, 3, 0); mtx.set(2, 3, 0);
  if (sign < 0.0f) mtx = m4_mul(m4_scale(-1, m4_identity()), mtx);
  Quat R{0, 0, 0, 1};
  M4 S = m4_identity();
  M4 Q;
  if (polar_factor(mtx, &Q)) {
    S = m4_mul(m4_transpose(Q), mtx);
    R = to_quat(Q);
    if (sign < 0.0f) { S = m4_mul(m4_scale(-1, m4_identity()), S); S.m[15] = 1; }
  }
  d.R[0] = R.X; d.R[1] = R.Y; d.R[2] = R.Z; d.R[3] = R.W;
  std::memcpy(d.S, S.m, sizeof(d.S));
  return d;
}
}  // namespace

// ---- Camera (builtin/camera/camera.go:80-98,109-193; ComputeRay's matrix :225-236) ----------------
int Camera::PreRender(Core& core, std::string* err) {
  if (Aspect == 0.0f) Aspect = core.FrameAspect();
  const float deg = Fov / 2;
  const float tan_theta_focal = (float)std::tan((double)(deg * kPi32 / 180.0f)) * Focal;
  if (FromKeys.empty()) FromKeys.push_back(From);
  if (ToKeys.empty()) ToKeys.push_back(To);
  if (RollKeys.empty()) RollKeys.push_back(Roll);
  std::vector<M4> l2w;
  if (Type == "LookAt") {
    // calcLookatMatrices (camera.go:109-193): one matrix per key of whichever of From / To has more keys; the other one and Roll
    // are interpolated at time = i / keys
    const int nF = (int)FromKeys.size(), nT = (int)ToKeys.size(), nR = (int)RollKeys.size();
    const bool byTarget = nT > nF;
    const int n = byTarget ? nT : nF;
    for (int i = 0; i < n; i++) {
      const float time = (float)i / (float)n;
      const std::vector<V3>& other = byTarget ? FromKeys : ToKeys;
      const float k = time * (float)((int)other.size() - 1);
      const float t = k - floorf(k);
      const V3 P = lerp(other[(int)floorf(k)], other[(int)ceilf(k)], t);
      const V3 eye = byTarget ? P : FromKeys[i], tgt = byTarget ? ToKeys[i] : P;
      const V3 W = normalize(eye - tgt);  // W points away from the target
      const V3 u = normalize(cross(Up, W));
      const V3 v = normalize(cross(u, W));
      const float kr = time * (float)(nR - 1);
      const float tr = kr - floorf(kr);
      const float roll = (1 - tr) * RollKeys[(int)floorf(kr)] + tr * RollKeys[(int)ceilf(kr)];
      const float cr = (float)std::cos((double)roll), sr = (float)std::sin((double)roll);
      const V3 U = scale(cr, u) + scale(sr, v);
      const V3 V = scale(-sr, u) + scale(cr, v);
      M4 basis{};
      basis.m[0] = U.x; basis.m[1] = U.y; basis.m[2] = U.z;
      basis.m[4] = V.x; basis.m[5] = V.y; basis.m[6] = V.z;
      basis.m[8] = W.x; basis.m[9] = W.y; basis.m[10] = W.z;
      basis.m[15] = 1.0f;
      M4 trans = m4_identity();
      trans.m[12] = eye.x; trans.m[13] = eye.y; trans.m[14] = eye.z;
      l2w.push_back(m4_mul(trans, basis));
    }
  } else {
    // matrixCalc (camera.go:205-216): every Type other than "LookAt" is a matrix camera (camera.go:91-95); a singular
    // WorldToLocal inverts to the null matrix (matrix4.go:134-147)
    for (const M4& w2l : WorldToLocal) {
      M4 inv{};
      m4_inverse(w2l, &inv);
      l2w.push_back(inv);
    }
  }
  if (l2w.size() > 255) { *err = "Camera: more than 255 motion keys"; return -1; }
  decomp.clear();
  for (const M4& m : l2w) decomp.push_back(transform_decomp(m));

  // ComputeRay's matrix (camera.go:225-236) at Time 0 — what every ray of a single-key camera uses: Lerp(d0, d0, 0) recomposed;
  // no keys at all (a matrix camera without WorldToLocal): c.decomp == nil and M stays the identity
  M4 M = m4_identity();
  if (!decomp.empty()) {
    const VgTransformSRT& d0 = decomp[0];
    const V3 Tl = lerp(V3{d0.T[0], d0.T[1], d0.T[2]}, V3{d0.T[0], d0.T[1], d0.T[2]}, 0.0f);
    const Quat R0{d0.R[0], d0.R[1], d0.R[2], d0.R[3]};
    const Quat Rl = slerp(R0, R0, 0.0f);
    M4 S0;
    std::memcpy(S0.m, d0.S, sizeof(S0.m));
    const M4 Sl = m4_lerp(S0, S0, 0.0f);
    M4 tl = m4_identity();
    tl.m[12] = Tl.x; tl.m[13] = Tl.y; tl.m[14] = Tl.z;
    M = m4_mul(tl, m4_mul(from_quat(Rl), Sl));
  }

  std::memcpy(out.local_to_world, M.m, sizeof(M.m));
  out.tan_theta_focal = tan_theta_focal;
  out.aspect = Aspect;
  out.focal = Focal;
  out.radius = Radius;
  return 0;
}

// ---- PolyMesh ------------------------------------------------------------------------------------
// builtin/geom/polymesh/init.go:12-134: fan-triangulate polygons into idxp (+ normal / shader indices)
// builtin/maps/texture.go:48-83 (net/url.Parse: path up to '?', then '&'-separated key=value pairs; first value wins)
TextureMap TextureMap::Parse(const std::string& value, bool read_channel) {
  TextureMap m;
  m.set = true;
  const size_t q = value.find('?');
  m.path = value.substr(0, q);
  if (q == std::string::npos) return m;
  const size_t frag = value.find('#', q);
  std::string query = value.substr(q + 1, frag == std::string::npos ? std::string::npos : frag - q - 1);
  bool got_ch = false, got_filter = false;
  size_t pos = 0;
  while (pos <= query.size()) {
    size_t e = query.find('&', pos);
    if (e == std::string::npos) e = query.size();
    const std::string kv = query.substr(pos, e - pos);
    const size_t eq = kv.find('=');
    const std::string k = kv.substr(0, eq), v = eq == std::string::npos ? "" : kv.substr(eq + 1);
    if (k == "ch" && !got_ch) {
      got_ch = true;
      char* end = nullptr;
      const long c = std::strtol(v.c_str(), &end, 10);
      if (read_channel && end && *end == 0 && !v.empty()) m.chan = (int)c;  // strconv.Atoi; its error is dropped (texture.go:57)
    } else if (k == "filter" && !got_filter) {
      got_filter = true;
      if (v == "trilinear") m.filter = VG_TEXFILTER_TRILINEAR;
    }
    pos = e + 1;
  }
  return m;
}

void PolyMesh::triangulate() {
  const bool hasN = !Normals.Elems.empty();
  const bool hasUV = !UV.empty();  // init.go:38-48,77-83,100-106
  auto emit = [&](uint32_t a, uint32_t b, uint32_t c) {
    idxp.push_back((uint32_t)FaceIdx[a]); idxp.push_back((uint32_t)FaceIdx[b]); idxp.push_back((uint32_t)FaceIdx[c]);
    if (hasUV) {
      const std::vector<int32_t>& src = hasUVIdx ? UVIdx : FaceIdx;
      uvtriidx.push_back((uint32_t)src[a]); uvtriidx.push_back((uint32_t)src[b]); uvtriidx.push_back((uint32_t)src[c]);
    }
    if (hasN) {
      const std::vector<int32_t>& src = hasNormalIdx ? NormalIdx : FaceIdx;
      normalidx.push_back((uint32_t)src[a]); normalidx.push_back((uint32_t)src[b]); normalidx.push_back((uint32_t)src[c]);
    }
  };
  if (hasPolyCount) {
    uint32_t base = 0;
    for (size_t k = 0; k < PolyCount.size(); k++) {
      for (int j = 1; j <= PolyCount[k] - 2; j++) {
        emit(base, base + j, base + j + 1);
        if (!ShaderIdx.empty()) shaderidx.push_back((uint8_t)ShaderIdx[k]);
      }
      base += (uint32_t)PolyCount[k];
    }
  } else {
    if (hasFaceIdx) {
      for (size_t j = 0; j < FaceIdx.size(); j++) {
        idxp.push_back((uint32_t)FaceIdx[j]);
        if (hasN) normalidx.push_back(hasNormalIdx ? (uint32_t)NormalIdx[j] : (uint32_t)FaceIdx[j]);
        if (hasUV) uvtriidx.push_back(hasUVIdx ? (uint32_t)UVIdx[j] : (uint32_t)FaceIdx[j]);
      }
    } else {
      for (int j = 0; j < Verts.ElemsPerKey; j++) {
        idxp.push_back((uint32_t)j);
        if (hasN) normalidx.push_back(hasNormalIdx ? (uint32_t)NormalIdx[j] : (uint32_t)j);
        if (hasUV) uvtriidx.push_back(hasUVIdx ? (uint32_t)UVIdx[j] : (uint32_t)j);
      }
    }
    for (int32_t s : ShaderIdx) shaderidx.push_back((uint8_t)s);
  }
  FaceIdx.clear(); PolyCount.clear(); NormalIdx.clear(); ShaderIdx.clear(); UVIdx.clear();
}

// polymesh.go:73-97
int PolyMesh::PreRender(Core& core, std::string* err) {
  if (Verts.MotionKeys < 1 || Verts.ElemsPerKey < 1) { *err = "PolyMesh " + NodeName + ": no Verts"; return -1; }
  {
    // init.go:12-134 indexes NormalIdx / UVIdx by FaceIdx position and ShaderIdx by polygon; Go panics on a short slice
    // (bounds check), here the mesh is refused before anything is indexed
    size_t nidx = hasFaceIdx ? FaceIdx.size() : (size_t)Verts.ElemsPerKey;
    if (hasPolyCount) {
      size_t sum = 0;
      for (int32_t c : PolyCount) {
        if (c < 0) { *err = "PolyMesh " + NodeName + ": negative PolyCount"; return -1; }
        sum += (size_t)c;
      }
      if (!hasFaceIdx || sum > FaceIdx.size()) { *err = "PolyMesh " + NodeName + ": PolyCount sums to more indices than FaceIdx holds"; return -1; }
      if (!ShaderIdx.empty() && ShaderIdx.size() < PolyCount.size()) { *err = "PolyMesh " + NodeName + ": ShaderIdx is shorter than PolyCount"; return -1; }
      nidx = sum;
    }
    if (!Normals.Elems.empty() && hasNormalIdx && NormalIdx.size() < nidx) { *err = "PolyMesh " + NodeName + ": NormalIdx is shorter than FaceIdx"; return -1; }
    if (!UV.empty() && hasUVIdx && UVIdx.size() < nidx) { *err = "PolyMesh " + NodeName + ": UVIdx is shorter than FaceIdx"; return -1; }
  }
  triangulate();
  if (idxp.size() % 3 != 0) { *err = "PolyMesh " + NodeName + ": index count is not a multiple of 3"; return -1; }
  for (uint32_t i : idxp)
    if (i >= (uint32_t)Verts.ElemsPerKey) { *err = "PolyMesh " + NodeName + ": vertex index out of range"; return -1; }
  if (uvtriidx.size() != 0 && uvtriidx.size() != idxp.size()) { *err = "PolyMesh " + NodeName + ": UVIdx does not match FaceIdx"; return -1; }
  for (uint32_t i : uvtriidx)
    if ((size_t)i * 2 + 1 >= UV.size()) { *err = "PolyMesh " + NodeName + ": UV index out of range"; return -1; }
  facecount = (int)idxp.size() / 3;
  if (!normalidx.empty() && normalidx.size() != idxp.size()) { *err = "PolyMesh " + NodeName + ": NormalIdx does not match FaceIdx"; return -1; }
  for (uint32_t i : normalidx)
    if ((size_t)i >= Normals.Elems.size()) { *err = "PolyMesh " + NodeName + ": normal index out of range"; return -1; }
  if (!shaderidx.empty() && shaderidx.size() != (size_t)facecount) { *err = "PolyMesh " + NodeName + ": ShaderIdx does not give one shader per face"; return -1; }
  for (uint8_t i : shaderidx)
    if ((size_t)i >= Shader.size()) { *err = "PolyMesh " + NodeName + ": ShaderIdx value beyond the Shader list"; return -1; }
  for (const std::string& s : Shader) {
    Node* n = core.FindNode(s);
    if (!n) { *err = "Unable to find node (shader " + s + ")"; return -1; }
    ShaderStd* sh = dynamic_cast<ShaderStd*>(n);
    if (!sh) { *err = "Unable to find shader " + s; return -1; }
    shader.push_back(sh);
  }
  return initAccel(err, core.build_ctx, core.leaf_max);
}

// builtin/geom/polymesh/buildqbvh.go:14-144
// below this a mesh is built faster by the host builder (a device build is ~16 rounds of 13 launches and one sync each,
// ~1 ms however small the mesh; scenes of many small meshes also keep the host's threads busy in parallel)
static const int kDeviceBuildMinFaces = 32768;

int PolyMesh::initAccel(std::string* err, vg_ctx* build_ctx, int leaf_max) {
  std::vector<Box> boxes(facecount);
  std::vector<V3> cent(facecount);
  std::vector<int32_t> idxs(facecount);
  const int E = Verts.ElemsPerKey;

  if (Verts.MotionKeys > 1) {
    // topology from the mid-time snapshot (:19-52)
    const float k = 0.5f * (float)(Verts.MotionKeys - 1);
    const float time = k - floorf(k);
    const int key = (int)floorf(k), key2 = (int)ceilf(k);
    for (int i = 0; i < facecount; i++) {
      V3 p[3];
      for (int j = 0; j < 3; j++) p[j] = lerp(Verts.Elems[(int)idxp[i * 3 + j] + E * key], Verts.Elems[(int)idxp[i * 3 + j] + E * key2], time);
      boxes[i].reset();
      for (int j = 0; j < 3; j++) boxes[i].grow_point(p[j].x, p[j].y, p[j].z);
      cent[i] = V3{(p[0].x + p[1].x + p[2].x) / 3, (p[0].y + p[1].y + p[2].y) / 3, (p[0].z + p[1].z + p[2].z) / 3};
      idxs[i] = i;
    }
    if (build_ctx && facecount >= kDeviceBuildMinFaces) {
      // BuildAccelMotion (motionbuild.go:108-122) is BuildAccel's recursion on the mid-time snapshot without the boxes: the
      // device builder's topology (axes, child links, leaf ranges) is copied, the per-key boxes are filled below as before
      int n_nodes = 0;
      float b6[6];
      static std::mutex device_build_mu;
      std::lock_guard<std::mutex> build_lock(device_build_mu);
      std::vector<VgNode> tmp;
      if (vg_build_qbvh(build_ctx, &boxes[0].lo[0], &cent[0].x, facecount, leaf_max, idxs.data(), b6, &n_nodes) != VG_OK) {
        *err = std::string("device MQBVH build: ") + vg_last_error(build_ctx);
        return -1;
      }
      tmp.resize((size_t)n_nodes);
      if (vg_build_qbvh_nodes(build_ctx, tmp.data(), n_nodes) != VG_OK) {
        *err = std::string("device MQBVH build: ") + vg_last_error(build_ctx);
        return -1;
      }
      mtopo.assign((size_t)n_nodes, VgMotionNode{});
      for (int i = 0; i < n_nodes; i++) {
        mtopo[(size_t)i].axis0 = (int32_t)tmp[(size_t)i].axis0;
        mtopo[(size_t)i].axis1 = (int32_t)tmp[(size_t)i].axis1;
        mtopo[(size_t)i].axis2 = (int32_t)tmp[(size_t)i].axis2;
        for (int k = 0; k < 4; k++) mtopo[(size_t)i].children[k] = tmp[(size_t)i].children[k];
      }
    } else if (build_mqbvh(boxes.data(), cent.data(), idxs.data(), facecount, leaf_max, mtopo, err) != 0) {
      return -1;
    }
    accel_idx = idxs;  // NOTE: idxp is NOT reordered on this path (the function returns at :56) — quirk (b)
    // per-key boxes (:148-212)
    mboxes.assign((size_t)Verts.MotionKeys * mtopo.size() * 24, 0.0f);
    Box full;
    full.reset();
    for (int kk = 0; kk < Verts.MotionKeys; kk++) {
      Box b = initMotionBoxesRec(kk, 0);
      motionBounds.push_back(b);
      full.grow_box(b);
    }
    bounds = full;
    return 0;
  }

  for (int i = 0; i < facecount; i++) {
    const V3 p0 = Verts.Elems[idxp[i * 3 + 0]], p1 = Verts.Elems[idxp[i * 3 + 1]], p2 = Verts.Elems[idxp[i * 3 + 2]];
    boxes[i].reset();
    boxes[i].grow_point(p0.x, p0.y, p0.z);
    boxes[i].grow_point(p1.x, p1.y, p1.z);
    boxes[i].grow_point(p2.x, p2.y, p2.z);
    cent[i] = V3{(p0.x + p1.x + p2.x) / 3, (p0.y + p1.y + p2.y) / 3, (p0.z + p1.z + p2.z) / 3};
    idxs[i] = i;
  }
  // the same tree from either builder (node for node, box for box; only the order inside a leaf differs): on the device
  // (build_bvh.cu) for meshes large enough to pay for the round trips, else on the host
  if (build_ctx && facecount >= kDeviceBuildMinFaces) {
    static_assert(sizeof(Box) == 6 * sizeof(float) && sizeof(V3) == 3 * sizeof(float), "Box / V3 are passed as flat floats");
    int n_nodes = 0;
    float b6[6];
    // build + fetch is a two-call protocol on the context; the meshes of a round are pre-rendered concurrently
    static std::mutex device_build_mu;
    std::lock_guard<std::mutex> build_lock(device_build_mu);
    if (vg_build_qbvh(build_ctx, &boxes[0].lo[0], &cent[0].x, facecount, leaf_max, idxs.data(), b6, &n_nodes) != VG_OK) {
      *err = std::string("device QBVH build: ") + vg_last_error(build_ctx);
      return -1;
    }
    qbvh.resize((size_t)n_nodes);
    if (vg_build_qbvh_nodes(build_ctx, qbvh.data(), n_nodes) != VG_OK) {
      *err = std::string("device QBVH build: ") + vg_last_error(build_ctx);
      return -1;
    }
    for (int a = 0; a < 3; a++) { bounds.lo[a] = b6[a]; bounds.hi[a] = b6[3 + a]; }
  } else if (build_qbvh(boxes.data(), cent.data(), idxs.data(), facecount, leaf_max, qbvh, &bounds, err) != 0) {
    return -1;
  }
  accel_idx = idxs;
  // reorder per-face arrays into leaf order (:90-129)
  std::vector<uint32_t> nidx(idxp.size());
  for (int i = 0; i < facecount; i++)
    for (int j = 0; j < 3; j++) nidx[i * 3 + j] = idxp[idxs[i] * 3 + j];
  idxp.swap(nidx);
  if (!shaderidx.empty()) {
    std::vector<uint8_t> ns(shaderidx.size());
    for (int i = 0; i < facecount; i++) ns[i] = shaderidx[idxs[i]];
    shaderidx.swap(ns);
  }
  if (!normalidx.empty()) {
    std::vector<uint32_t> nn(normalidx.size());
    for (int i = 0; i < facecount; i++)
      for (int j = 0; j < 3; j++) nn[i * 3 + j] = normalidx[idxs[i] * 3 + j];
    normalidx.swap(nn);
  }
  if (!uvtriidx.empty()) {  // buildqbvh.go:106-116
    std::vector<uint32_t> uu(uvtriidx.size());
    for (int i = 0; i < facecount; i++)
      for (int j = 0; j < 3; j++) uu[i * 3 + j] = uvtriidx[idxs[i] * 3 + j];
    uvtriidx.swap(uu);
  }
  return 0;
}

static inline int ref_leaf_count(int32_t l) { return (int)((l & 0xf) + 1); }
static inline int ref_leaf_base(int32_t l) { return (int)((l & 0x7ffffff) >> 4); }  // 23-bit decode, quirk (c)
static inline void store_box(float* dst24, int k, const Box& b) {
  for (int a = 0; a < 3; a++) { dst24[k + a * 4] = b.lo[a]; dst24[k + 12 + a * 4] = b.hi[a]; }
}

// buildqbvh.go:171-212 — leaf boxes bound faces accel_idx[i] at key `key`
Box PolyMesh::initMotionBoxesRec(int key, int32_t node) {
  Box nodebox;
  nodebox.reset();
  const int E = Verts.ElemsPerKey;
  float* dst = &mboxes[((size_t)key * mtopo.size() + node) * 24];
  for (int k = 0; k < 4; k++) {
    const int32_t ch = mtopo[node].children[k];
    if (ch == -1) continue;
    Box b;
    if (ch < 0) {
      b.reset();
      const int base = ref_leaf_base(ch), cnt = ref_leaf_count(ch);
      for (int i = base; i < base + cnt; i++) {
        const int f = accel_idx[i];
        for (int j = 0; j < 3; j++) {
          const V3 p = Verts.Elems[(int)idxp[f * 3 + j] + E * key];
          b.grow_point(p.x, p.y, p.z);
        }
      }
    } else {
      b = initMotionBoxesRec(key, ch);
    }
    store_box(dst, k, b);
    nodebox.grow_box(b);
  }
  return nodebox;
}

// builtin/geom/polymesh/bounds.go:25-53
Box PolyMesh::Bounds(float time) const {
  if (!qbvh.empty()) return bounds;
  const float k = time * (float)((int)motionBounds.size() - 1);
  const float t = k - floorf(k);
  return box_lerp(motionBounds[(int)floorf(k)], motionBounds[(int)ceilf(k)], t);
}

// ---- TriLight (builtin/light/triangle.go:37-58,537-565) -----------------------------------------
int TriLight::PreRender(Core& core, std::string* err) {
  Node* n = core.FindNode(Shader);
  if (!n) { *err = "Unable to find node (shader " + Shader + ")"; return -1; }
  shader = dynamic_cast<ShaderStd*>(n);
  if (!shader) { *err = "Unable to find shader " + Shader; return -1; }
  // createMesh: a one-triangle PolyMesh with three identical vertex normals, added as a new node
  std::unique_ptr<PolyMesh> m(new PolyMesh());
  m->NodeName = NodeName + ":<mesh>";
  m->Shader = {Shader};
  m->Verts.MotionKeys = 1;
  m->Verts.ElemsPerKey = 3;
  m->Verts.Elems = {P0, P1, P2};
  const V3 N = normalize(cross(P1 - P0, P2 - P0));
  m->Normals.MotionKeys = 1;
  m->Normals.ElemsPerKey = 3;
  m->Normals.Elems = {N, N, N};
  geom = m.get();
  core.AddNode(std::move(m));
  return 0;
}

static void put3(float* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
void TriLight::Describe(VgLight* o) const {
  std::memset(o, 0, sizeof(*o));
  o->type = VG_LIGHT_TRI;
  o->samples = Samples;
  o->material = shader ? shader->material_id : -1;
  o->geom = geom ? geom->id : -1;
  put3(o->p0, P0); put3(o->p1, P1); put3(o->p2, P2);
}

// ---- DiskLight (builtin/light/disk.go:76-106,224-262) ----------------------------------------------
static inline float cos32f(float x) { return (float)std::cos((double)x); }  // math/sincos.go: float32 trig goes through float64
static inline float sin32f(float x) { return (float)std::sin((double)x); }
int DiskLight::PreRender(Core& core, std::string* err) {
  Node* n = core.FindNode(Shader);
  if (!n) { *err = "Unable to find node (shader " + Shader + ")"; return -1; }
  shader = dynamic_cast<ShaderStd*>(n);
  if (!shader) { *err = "Unable to find shader " + Shader; return -1; }
  N = normalize(LookAt - P);
  T = normalize(cross(N, Up));
  B = cross(N, T);
  // createMesh: a fan of `Segments` unindexed triangles with the normal N on every vertex (UVs omitted)
  std::unique_ptr<PolyMesh> m(new PolyMesh());
  m->NodeName = NodeName + ":<mesh>";
  m->Shader = {Shader};
  m->Verts.MotionKeys = 1;
  m->Normals.MotionKeys = 1;
  const int nv = Segments;
  const float dv = 2 * kPi32 / (float)nv;
  float ang = 0;
  for (int i = 0; i < nv; i++) {
    m->Verts.Elems.push_back(P);
    m->Verts.Elems.push_back(P + (scale(Radius * cos32f(ang), B) + scale(Radius * sin32f(ang), T)));
    m->Verts.Elems.push_back(P + (scale(Radius * cos32f(ang + dv), B) + scale(Radius * sin32f(ang + dv), T)));
    m->Verts.ElemsPerKey += 3;
    m->Normals.Elems.push_back(N);
    m->Normals.Elems.push_back(N);
    m->Normals.Elems.push_back(N);
    m->Normals.ElemsPerKey += 3;
    ang += dv;
  }
  geom = m.get();
  core.AddNode(std::move(m));
  return 0;
}
void DiskLight::Describe(VgLight* o) const {
  std::memset(o, 0, sizeof(*o));
  o->type = VG_LIGHT_DISK;
  o->samples = Samples;
  o->material = shader ? shader->material_id : -1;
  o->geom = geom ? geom->id : -1;
  put3(o->p0, P); put3(o->p1, T); put3(o->p2, B); put3(o->n, N);
  o->radius = Radius;
}

// ---- Sphere geom + SphereLight (builtin/geom/sphere/sphere.go:41-76, builtin/light/sphere.go:38-61) ----
int SphereGeom::PreRender(Core& core, std::string* err) {
  Node* n = core.FindNode(Shader);
  if (!n) { *err = "Unable to find node (shader " + Shader + ")"; return -1; }
  shader = dynamic_cast<ShaderStd*>(n);
  if (!shader) { *err = "Unable to find shader " + Shader; return -1; }
  return 0;
}
Box SphereGeom::Bounds(float) const {
  Box b;
  b.lo[0] = P.x - Radius; b.lo[1] = P.y - Radius; b.lo[2] = P.z - Radius;
  b.hi[0] = P.x + Radius; b.hi[1] = P.y + Radius; b.hi[2] = P.z + Radius;
  return b;
}
int SphereLight::PreRender(Core& core, std::string* err) {
  Node* n = core.FindNode(Shader);
  if (!n) { *err = "Unable to find node (shader " + Shader + ")"; return -1; }
  shader = dynamic_cast<ShaderStd*>(n);
  if (!shader) { *err = "Unable to find shader " + Shader; return -1; }
  std::unique_ptr<SphereGeom> g(new SphereGeom());
  g->NodeName = NodeName + ":<sphere>";  // the reference leaves the geom unnamed; a name keeps the node map unambiguous
  g->P = P;
  g->Radius = Radius;
  g->Shader = Shader;
  geom = g.get();
  core.AddNode(std::move(g));
  return 0;
}
void SphereLight::Describe(VgLight* o) const {
  std::memset(o, 0, sizeof(*o));
  o->type = VG_LIGHT_SPHERE;
  o->samples = Samples;
  o->material = shader ? shader->material_id : -1;
  o->geom = geom ? geom->id : -1;
  put3(o->p0, P);
  o->radius = Radius;
}

// ---- GeomInstance (builtin/geom/instance/instance.go:117-146) -----------------------------------------
int GeomInstance::PreRender(Core& core, std::string* err) {
  for (const M4& m : Transform) transformSRT.push_back(transform_decomp(m));
  for (size_t i = 0; i < BMin.size(); i++) {
    Box b;  // the zero value, not Reset(): the bounds always contain the origin (instance.go:124)
    for (int k = 0; k < 3; k++) b.lo[k] = b.hi[k] = 0.0f;
    b.grow_point(BMin[i].x, BMin[i].y, BMin[i].z);
    b.grow_point(BMax[i].x, BMax[i].y, BMax[i].z);
    bounds.push_back(b);
  }
  Node* n = core.FindNode(GeomName);
  if (!n) { *err = "Instance " + NodeName + ": Unable to find node " + GeomName; return -1; }
  geom = dynamic_cast<Geom*>(n);
  if (!geom) { *err = "Instance " + NodeName + ": Unable to find geom " + GeomName; return -1; }
  // a PolyMesh, or another GeomInstance (Instance.Trace simply calls ins.geom.Trace, instance.go:95: the transforms chain)
  if (!dynamic_cast<PolyMesh*>(n) && !dynamic_cast<GeomInstance*>(n)) { *err = "Instance " + NodeName + ": only PolyMesh and GeomInstance targets are supported on this path"; return -1; }
  if (n == this) { *err = "Instance " + NodeName + ": an instance of itself"; return -1; }
  if (transformSRT.empty() || bounds.empty()) { *err = "Instance " + NodeName + ": Transform and BMin/BMax need at least one element"; return -1; }
  return 0;
}

// ---- pixel filters (builtin/filter) -----------------------------------------------------------------
namespace {
double bessel_j1(double x) {  // airy.go:34-72 (Numerical-Recipes rational approximations)
  const double ax = std::fabs(x);
  if (ax < 8.0) {
    const double y = x * x;
    const double a1 = x * (72362614232.0 + y * (-7895059235.0 + y * (242396853.1 + y * (-2972611.439 + y * (15704.48260 + y * (-30.16036606))))));
    const double a2 = 144725228442.0 + y * (2300535178.0 + y * (18583304.74 + y * (99447.43394 + y * (376.9991397 + y * 1.0))));
    return a1 / a2;
  }
  const double z = 8.0 / ax, y = z * z, xx = ax - 2.356194491;
  const double a1 = 1.0 + y * (0.183105e-2 + y * (-0.3516396496e-4 + y * (0.2457520174e-5 + y * (-0.240337019e-6))));
  const double a2 = 0.04687499995 + y * (-0.2002690873e-3 + y * (0.8449199096e-5 + y * (-0.88228987e-6 + y * 0.105787412e-6)));
  double ans = std::sqrt(0.636619772 / ax) * (std::cos(xx) * a1 - z * std::sin(xx) * a2);
  return x < 0.0 ? -ans : ans;
}
}  // namespace

int PixelFilter::PreRender(Core&, std::string* err) {
  const int n = Res;
  if (n < 2 || n > 1024 || !(Width > 0)) { *err = "PixelFilter " + NodeName + ": bad Res/Width"; return -1; }
  const double w = (double)Width;
  auto f = [&](double x, double y) -> double {
    const double q = std::sqrt(x * x + y * y);
    if (q > (double)(Width / 2)) return 0;  // cut-off
    if (kind == 1) {                        // airy.go:76-97
      const double v = (20000 / (double)Width) * (M_PI * q) / (550.0 * 5.6);
      const double b = 2 * bessel_j1(v) / v;
      return (double)Peak * (b * b);
    }
    const double sigma = (double)(1.0 / std::sqrt((double)Width));  // gauss.go:33-45
    return (1 / (2 * M_PI * (sigma * sigma))) * std::exp((double)(-(x * x + y * y) / 2 * (sigma * sigma)));
  };
  // filter.go:88-173: tabulate (u outer, v inner, both advanced by repeated addition), normalise, marginal and conditional CDFs
  std::vector<double> tab((size_t)n * n);
  const double du = w / (double)(n - 1);
  double u = -w / 2, F = 0;
  for (int j = 0; j < n; j++) {
    double v = -w / 2;
    for (int i = 0; i < n; i++) {
      const double fuv = f(u, v);
      tab[j + (size_t)i * n] = fuv;
      F += fuv;
      v += du;
    }
    u += du;
  }
  std::vector<double> pV(n, 0.0), pdf((size_t)n * n);
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) {
      pdf[(size_t)j * n + i] = tab[j + (size_t)i * n] / F;
      pV[j] += pdf[(size_t)j * n + i];
    }
  cdfV.assign(n, 0.0);
  cdfVU.assign((size_t)n * n, 0.0);
  double p = 0;
  for (int j = 0; j < n; j++) {
    p += pV[j];
    cdfV[j] = p;
    double q = 0;
    for (int i = 0; i < n; i++) {
      q += pdf[(size_t)j * n + i] / pV[j];
      cdfVU[(size_t)j * n + i] = q;
    }
  }
  return 0;
}

// ---- Scene (builtin/scene/scene.go:135-268) -------------------------------------------------------
int Scene::PreRender(std::string* err) {
  const int n = (int)geoms.size();
  if (n == 0) { *err = "scene has no geoms"; return -1; }
  std::vector<Box> boxes(n);
  std::vector<V3> cent(n);
  std::vector<int32_t> idx(n);
  int maxKeys = 0;
  for (Geom* g : geoms) if (g->MotionKeys() > maxKeys) maxKeys = g->MotionKeys();
  const float t = maxKeys == 1 ? 0.0f : 0.5f;
  for (int i = 0; i < n; i++) {
    boxes[i] = geoms[i]->Bounds(t);
    idx[i] = i;
    cent[i] = boxes[i].centroid();
  }
  int rc;
  if (maxKeys == 1) rc = build_qbvh(boxes.data(), cent.data(), idx.data(), n, 1, qbvh, &bounds, err);
  else rc = build_mqbvh(boxes.data(), cent.data(), idx.data(), n, 1, mtopo, err);
  if (rc != 0) return -1;
  std::vector<Geom*> ng(n);
  for (int i = 0; i < n; i++) ng[i] = geoms[idx[i]];
  geoms.swap(ng);
  keys = maxKeys;
  if (maxKeys > 1) {
    mboxes.assign((size_t)maxKeys * mtopo.size() * 24, 0.0f);
    Box full;
    full.reset();
    for (int k = 0; k < maxKeys; k++) full.grow_box(initMotionBoxesRec(k, 0, maxKeys));
    bounds = full;
  }
  return 0;
}

Box Scene::initMotionBoxesRec(int key, int32_t node, int nkeys) {
  Box nodebox;
  nodebox.reset();
  float* dst = &mboxes[((size_t)key * mtopo.size() + node) * 24];
  for (int k = 0; k < 4; k++) {
    const int32_t ch = mtopo[node].children[k];
    if (ch == -1) continue;
    Box b;
    if (ch < 0) {
      b.reset();
      const int base = ref_leaf_base(ch), cnt = ref_leaf_count(ch);
      for (int i = base; i < base + cnt; i++) {
        const float time = (float)key / (float)(nkeys - 1);
        b.grow_box(geoms[i]->Bounds(time));
      }
    } else {
      b = initMotionBoxesRec(key, ch, nkeys);
    }
    store_box(dst, k, b);
    nodebox.grow_box(b);
  }
  return nodebox;
}

// ---- Core (core/core.go) --------------------------------------------------------------------------
Core::Core() {
  std::unique_ptr<Globals> g(new Globals());
  globals = g.get();
  owned.push_back(std::move(g));
}

void Core::AddNode(std::unique_ptr<Node> node) {
  Node* n = node.get();
  owned.push_back(std::move(node));
  pending.push_back(n);
  nodeMap[n->Name()] = n;
  if (PolyMesh* pm = dynamic_cast<PolyMesh*>(n)) {
    pm->id = next_geom_id++;
    scene.AddGeom(pm);
  } else if (GeomInstance* gi = dynamic_cast<GeomInstance*>(n)) {
    gi->id = next_geom_id++;
    scene.AddGeom(gi);
  } else if (SphereGeom* sg = dynamic_cast<SphereGeom*>(n)) {
    sg->id = next_geom_id++;
    scene.AddGeom(sg);
  } else if (Light* lt = dynamic_cast<Light*>(n)) {
    scene.AddLight(lt);
  } else if (ShaderStd* sh = dynamic_cast<ShaderStd*>(n)) {
    sh->material_id = (int)materials.size();
    materials.push_back(sh);
  } else if (Globals* g = dynamic_cast<Globals*>(n)) {
    globals = g;
  } else if (PixelFilter* pf = dynamic_cast<PixelFilter*>(n)) {
    filter = pf;
  }
}

Node* Core::FindNode(const std::string& name) const {
  auto it = nodeMap.find(name);
  return it == nodeMap.end() ? nullptr : it->second;
}

// core.go:36-61: PreRender every node; nodes added during a round are processed in the next one
int Core::PreRender() {
  if (prerendered) return 0;
  while (!pending.empty()) {
    std::vector<Node*> round;
    round.swap(pending);
    all.insert(all.end(), round.begin(), round.end());
    // PolyMesh.PreRender only reads the core (FindNode) and writes its own mesh, so the meshes of a round are pre-rendered
    // concurrently (the reference does them one after the other in this loop: C3's 1024 meshes took 0.94 s that way); every other
    // node type runs in order below, and a mesh's error is still reported at its place in that order.
    std::vector<PolyMesh*> meshes;
    for (Node* n : round)
      if (PolyMesh* pm = dynamic_cast<PolyMesh*>(n)) meshes.push_back(pm);
    std::vector<int> mesh_rc(meshes.size(), 0);
    std::vector<std::string> mesh_err(meshes.size());
    const unsigned hw = std::thread::hardware_concurrency();
    const int nthreads = (int)std::min<size_t>(std::min<unsigned>(hw ? hw : 1, 16), meshes.size());
    if (nthreads >= 2) {
      std::atomic<size_t> next(0);
      auto work = [&] {
        for (;;) {
          const size_t i = next.fetch_add(1);
          if (i >= meshes.size()) return;
          mesh_rc[i] = meshes[i]->PreRender(*this, &mesh_err[i]);
        }
      };
      std::vector<std::thread> th;
      for (int t = 1; t < nthreads; t++) th.emplace_back(work);
      work();
      for (auto& t : th) t.join();
    }
    size_t mi = 0;
    for (Node* n : round) {
      if (nthreads >= 2 && dynamic_cast<PolyMesh*>(n)) {
        const size_t i = mi++;
        if (mesh_rc[i] != 0) { err = mesh_err[i]; return -1; }
        continue;
      }
      if (n->PreRender(*this, &err) != 0) return -1;
    }
  }
  if (scene.PreRender(&err) != 0) return -1;
  prerendered = true;
  return 0;
}

}  // namespace vh
