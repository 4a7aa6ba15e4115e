// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// Flat C entry points over the C++ restatement, for tests/ (ctypes), __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs. Nothing in vermeer_b200/ may link or load this.
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <algorithm>

#include "ldseq.h"
#include "scene.h"
#include "shading.h"

using namespace orc;

namespace {
thread_local std::string g_err;
struct Handle {
  Renderer r;
};
template <class F>
int guard(F f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
Vec3 v3(const float* p) { return V3(p[0], p[1], p[2]); }
}  // namespace

extern "C" {

struct OrcRay { float o[3]; float d[3]; float tmax; float time; };
struct OrcHit { float t, u, v, w; int32_t prim, geom; int32_t nodesT, trisT; };

const char* orc_last_error() { return g_err.c_str(); }

void* orc_create() { return new Handle(); }
void orc_destroy(void* h) { delete (Handle*)h; }

int orc_set_globals(void* h, int xres, int yres) {
  Handle* H = (Handle*)h;
  H->r.XRes = xres;
  H->r.YRes = yres;
  return 0;
}

// mask bits / p[] slots: 0 EmissionColour(p0..2) 1 EmissionStrength(p3) 2 DiffuseColour(p4..6) 3 DiffuseStrength(p7)
// 4 DiffuseRoughness(p8) 5 Spec1Colour(p9..11) 6 Spec1Strength(p12) 7 Spec1Roughness(p13) 8 IOR(p14)
// 9 Spec1FresnelModel(p15: 0 "Dielectric", 1 "Metal") 10 Spec1FresnelRefl(p16..18) 11 Spec1FresnelEdge(p19..21)
int orc_add_shader(void* h, const char* name, uint32_t mask, const float* p) {
  Handle* H = (Handle*)h;
  if (mask & 4096) {  // DebugShader: Colour in the DiffuseColour slot
    auto d = std::make_unique<DebugShader>();
    d->Name = name;
    d->Colour = MakeRGB(p[4], p[5], p[6]);
    H->r.shaders.push_back(std::move(d));
    return 0;
  }
  auto s = std::make_unique<ShaderStd>();
  s->Name = name;
  if (mask & 1) { s->hasEmissionColour = true; s->EmissionColour = MakeRGB(p[0], p[1], p[2]); }
  if (mask & 2) { s->hasEmissionStrength = true; s->EmissionStrength = p[3]; }
  if (mask & 4) { s->hasDiffuseColour = true; s->DiffuseColour = MakeRGB(p[4], p[5], p[6]); }
  if (mask & 8) { s->hasDiffuseStrength = true; s->DiffuseStrength = p[7]; }
  if (mask & 16) { s->hasDiffuseRoughness = true; s->DiffuseRoughness = p[8]; }
  if (mask & 32) { s->hasSpec1Colour = true; s->Spec1Colour = MakeRGB(p[9], p[10], p[11]); }
  if (mask & 64) { s->hasSpec1Strength = true; s->Spec1Strength = p[12]; }
  if (mask & 128) { s->hasSpec1Roughness = true; s->Spec1Roughness = p[13]; }
  if (mask & 256) { s->hasIOR = true; s->IOR = p[14]; }
  if (mask & 512) s->spec1FresnelModel = p[15] != 0.0f ? 1 : 0;
  if (mask & 1024) { s->hasSpec1FresnelRefl = true; s->Spec1FresnelRefl = MakeRGB(p[16], p[17], p[18]); }
  if (mask & 2048) { s->hasSpec1FresnelEdge = true; s->Spec1FresnelEdge = MakeRGB(p[19], p[20], p[21]); }
  H->r.shaders.push_back(std::move(s));
  return 0;
}

// verts: keys*nverts*3 floats (key-major, core/param/array.go:26). polycount/faceidx/shaderidx/normals may be null.
// shaders: '\n'-separated shader names.
int orc_add_polymesh(void* h, const char* name, const float* verts, int nverts, int keys, const int32_t* polycount, int npoly,
                     const int32_t* faceidx, int nfaceidx, const char* shaders, const int32_t* shaderidx, int nshaderidx,
                     const float* normals, int nnormals, const int32_t* normalidx, int nnormalidx, float raybias) {
  Handle* H = (Handle*)h;
  return guard([&] {
    auto m = std::make_unique<PolyMesh>();
    m->Name = name;
    m->RayBias = raybias;
    m->Verts.MotionKeys = keys;
    m->Verts.ElemsPerKey = nverts;
    m->Verts.Elems.resize((size_t)keys * nverts);
    std::memcpy(m->Verts.Elems.data(), verts, sizeof(float) * 3 * (size_t)keys * nverts);
    if (polycount) { m->hasPolyCount = true; m->PolyCount.assign(polycount, polycount + npoly); }
    if (faceidx) { m->hasFaceIdx = true; m->FaceIdx.assign(faceidx, faceidx + nfaceidx); }
    if (shaderidx) m->ShaderIdx.assign(shaderidx, shaderidx + nshaderidx);
    if (normals) {
      m->Normals.MotionKeys = 1;
      m->Normals.ElemsPerKey = nnormals;
      m->Normals.Elems.resize(nnormals);
      std::memcpy(m->Normals.Elems.data(), normals, sizeof(float) * 3 * (size_t)nnormals);
      if (normalidx) { m->hasNormalIdx = true; m->NormalIdx.assign(normalidx, normalidx + nnormalidx); }
    }
    std::string s(shaders ? shaders : "");
    size_t pos = 0;
    while (pos <= s.size() && !s.empty()) {
      size_t e = s.find('\n', pos);
      std::string nm = s.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
      ShaderStd* sh = H->r.findShader(nm);
      if (!sh) throw std::runtime_error("Unable to find node (shader " + nm + ")");
      m->shader.push_back(sh);
      if (e == std::string::npos) break;
      pos = e + 1;
    }
    H->r.meshes.push_back(std::move(m));
  });
}

// PolyMesh.Transform (polymesh.go:32; init.go:14-18 decomposes every key): keys x 16 floats, column major. Documentation of
// reference quirk q only — the GPU path has no counterpart.
int orc_mesh_set_transform(void* h, const char* mesh_name, const float* transforms, int keys) {
  Handle* H = (Handle*)h;
  return guard([&] {
    for (auto& m : H->r.meshes)
      if (m->Name == mesh_name) {
        m->Transform.clear();
        m->transformSRT.clear();
        for (int k = 0; k < keys; k++) {
          Matrix4 mm;
          std::memcpy(&mm, transforms + (size_t)k * 16, sizeof(float) * 16);
          m->Transform.push_back(mm);
          m->transformSRT.push_back(TransformDecompMatrix4(mm));
        }
        return;
      }
    throw std::runtime_error(std::string("orc_mesh_set_transform: no mesh ") + mesh_name);
  });
}

// GeomInstance (builtin/geom/instance/instance.go:36-51): transforms = keys x 16 floats, column major (math.Matrix4 layout,
// i.e. what the parser stores after its transpose); bmin/bmax = nb points each.
int orc_add_instance(void* h, const char* name, const char* geom_name, const float* bmin, const float* bmax, int nb, const float* transforms, int keys) {
  Handle* H = (Handle*)h;
  return guard([&] {
    auto in = std::make_unique<Instance>();
    in->Name = name;
    for (auto& m : H->r.meshes)
      if (m->Name == geom_name) in->geom = m.get();
    for (auto& o : H->r.instances)  // core.FindNode finds any Geom: an Instance may duplicate another Instance (instance.go:131-143)
      if (o->Name == geom_name) in->geom = o.get();
    if (!in->geom) throw std::runtime_error(std::string("Instance ") + name + ": Unable to find node " + geom_name);
    for (int i = 0; i < nb; i++) { in->BMin.push_back(v3(bmin + 3 * i)); in->BMax.push_back(v3(bmax + 3 * i)); }
    for (int k = 0; k < keys; k++) {
      Matrix4 m;
      std::memcpy(m.m, transforms + 16 * k, 64);
      in->Transform.push_back(m);
    }
    if (in->Transform.empty() || in->BMin.empty()) throw std::runtime_error("Instance: Transform and BMin/BMax need at least one element");
    H->r.instances.push_back(std::move(in));
  });
}

int orc_add_trilight(void* h, const char* name, const float* p0, const float* p1, const float* p2, const char* shader, int samples) {
  Handle* H = (Handle*)h;
  return guard([&] {
    auto t = std::make_unique<Tri>();
    t->Name = name;
    t->P0 = v3(p0);
    t->P1 = v3(p1);
    t->P2 = v3(p2);
    t->Samples = samples;
    t->shader = H->r.findShader(shader);
    if (!t->shader) throw std::runtime_error(std::string("Unable to find node (shader ") + shader + ")");
    H->r.lightOrder.push_back(t.get());
    H->r.tris.push_back(std::move(t));
  });
}

// DiskLight (builtin/light/disk.go:21-34; registered defaults Segments 20, Samples 1, :263-269)
int orc_add_disklight(void* h, const char* name, const float* P, const float* lookat, const float* up, float radius, const char* shader,
                      int segments, int samples) {
  Handle* H = (Handle*)h;
  return guard([&] {
    auto d = std::make_unique<Disk>();
    d->Name = name;
    d->P = v3(P);
    d->LookAt = v3(lookat);
    d->Up = v3(up);
    d->Radius = radius;
    d->Segments = segments;
    d->Samples = samples;
    d->shader = H->r.findShader(shader);
    if (!d->shader) throw std::runtime_error(std::string("Unable to find node (shader ") + shader + ")");
    H->r.lightOrder.push_back(d.get());
    H->r.disks.push_back(std::move(d));
  });
}

// SphereLight (builtin/light/sphere.go:17-28; registered defaults Radius 1, Samples 1, :297-303)
int orc_add_spherelight(void* h, const char* name, const float* P, float radius, const char* shader, int samples) {
  Handle* H = (Handle*)h;
  return guard([&] {
    auto d = std::make_unique<SphereLight>();
    d->Name = name;
    d->P = v3(P);
    d->Radius = radius;
    d->Samples = samples;
    d->shader = H->r.findShader(shader);
    if (!d->shader) throw std::runtime_error(std::string("Unable to find node (shader ") + shader + ")");
    H->r.lightOrder.push_back(d.get());
    H->r.sphereLights.push_back(std::move(d));
  });
}

int orc_set_camera(void* h, const float* from, const float* to, const float* up, float roll, float fov, float focal, float aspect, float radius) {
  Handle* H = (Handle*)h;
  H->r.camera.From = v3(from);
  H->r.camera.To = v3(to);
  H->r.camera.Up = v3(up);
  H->r.camera.Roll = roll;
  H->r.camera.Fov = fov;
  H->r.camera.Focal = focal;
  H->r.camera.Aspect = aspect;
  H->r.camera.Radius = radius;
  return 0;
}

// Camera with motion keys (camera.go:48-73): from/to = n points each, roll = n floats (nroll 0 = Roll unset), or
// world_to_local = nmat column-major matrices for Type "Matrix".
int orc_set_camera_keys(void* h, const char* type, const float* from, int nfrom, const float* to, int nto, const float* roll, int nroll,
                        const float* up, const float* world_to_local, int nmat, float fov, float focal, float aspect, float radius) {
  Handle* H = (Handle*)h;
  Camera& c = H->r.camera;
  c.Type = type;
  c.FromKeys.clear(); c.ToKeys.clear(); c.RollKeys.clear(); c.WorldToLocal.clear();
  for (int i = 0; i < nfrom; i++) c.FromKeys.push_back(v3(from + 3 * i));
  for (int i = 0; i < nto; i++) c.ToKeys.push_back(v3(to + 3 * i));
  for (int i = 0; i < nroll; i++) c.RollKeys.push_back(roll[i]);
  for (int i = 0; i < nmat; i++) { Matrix4 m; std::memcpy(m.m, world_to_local + 16 * i, 64); c.WorldToLocal.push_back(m); }
  if (nfrom > 0) c.From = c.FromKeys[0];
  if (nto > 0) c.To = c.ToKeys[0];
  c.Up = v3(up);
  c.Fov = fov; c.Focal = focal; c.Aspect = aspect; c.Radius = radius;
  return 0;
}
int orc_camera_decomp(void* h, float* out23 /* keys x 23 */) {
  Handle* H = (Handle*)h;
  const Camera& c = H->r.camera;
  for (size_t i = 0; i < c.decomp.size(); i++) {
    float* o = out23 + 23 * i;
    for (int k = 0; k < 3; k++) o[k] = c.decomp[i].T[k];
    o[3] = c.decomp[i].R.X; o[4] = c.decomp[i].R.Y; o[5] = c.decomp[i].R.Z; o[6] = c.decomp[i].R.W;
    std::memcpy(o + 7, c.decomp[i].S.m, 64);
  }
  return (int)c.decomp.size();
}

// kind: 1 = AiryFilter (defaults Res 49, Width 6, Peak 4), 2 = GaussianFilter (Res 17, Width 2)  — builtin/filter/filter.go:14-26
int orc_set_filter(void* h, int kind, float width, int res, float peak) {
  Handle* H = (Handle*)h;
  if (kind == 0) { H->r.filter.reset(); return 0; }
  H->r.filter.reset(new PixelFilter());
  H->r.filter->kind = kind;
  H->r.filter->Width = width;
  H->r.filter->Res = res;
  H->r.filter->Peak = peak;
  return 0;
}
int orc_filter_tables(void* h, double* cdfV, double* cdfVU) {
  Handle* H = (Handle*)h;
  if (!H->r.filter) return -1;
  const FilterSampler& s = H->r.filter->sampler;
  for (int i = 0; i < s.n; i++) cdfV[i] = s.cdfV[i];
  for (int j = 0; j < s.n; j++)
    for (int i = 0; i < s.n; i++) cdfVU[j * s.n + i] = s.cdfVU[j][i];
  return 0;
}
void orc_filter_warp(void* h, double r0, double r1, double* u, double* v) {
  Handle* H = (Handle*)h;
  H->r.filter->sampler.WarpSample(r0, r1, u, v);
}
double orc_bessel_j1(double x) { return BesselJ1(x); }

int orc_set_motion_ref_compat(void* h, int on) {
  Handle* H = (Handle*)h;
  for (auto& m : H->r.meshes) m->ref_compat_motion = on != 0;
  return 0;
}

int orc_prerender(void* h) {
  Handle* H = (Handle*)h;
  return guard([&] { H->r.PreRender(); });
}

int orc_set_scramble(void* h, const uint64_t* table, int64_t npix) {
  Handle* H = (Handle*)h;
  H->r.framescramble.resize(npix);
  std::memcpy(H->r.framescramble.data(), table, sizeof(pixelscramble) * (size_t)npix);
  return 0;
}

int orc_clear_framebuffer(void* h) {
  Handle* H = (Handle*)h;
  std::fill(H->r.framebuffer.begin(), H->r.framebuffer.end(), 0.0f);
  return 0;
}

// stats: [rays, shadow rays, nanoseconds]
int orc_render(void* h, int iter_begin, int iter_end, int nthreads, int trace_last_level, float* fb_out, uint64_t* stats) {
  Handle* H = (Handle*)h;
  return guard([&] {
    H->r.trace_last_level = trace_last_level != 0;
    RenderStats s = H->r.Render(iter_begin, iter_end, nthreads);
    if (fb_out) std::memcpy(fb_out, H->r.framebuffer.data(), sizeof(float) * H->r.framebuffer.size());
    if (stats) {
      stats[0] = s.rayCount;
      stats[1] = s.shadowRayCount;
      stats[2] = (uint64_t)(s.seconds * 1e9);
    }
  });
}

// ---- textures (texture/texture.go, mipmap.go, feline.go; builtin/maps/texture.go) --------------------------------------
// rgb8: w*h*3 bytes, row 0 = BOTTOM row of the image (loadTexture flips while copying, texture.go:139).
int orc_add_texture(void* h, const char* name, int w, int hgt, const uint8_t* rgb8) {
  Handle* H = (Handle*)h;
  return guard([&] {
    auto t = std::make_unique<Texture>();
    t->url = name;
    t->w = w;
    t->h = hgt;
    t->data.assign(rgb8, rgb8 + (size_t)w * hgt * 3);
    t->mipmap = stdfilter(w, hgt, t->data, 3);
    if (t->mipmap.mipmap.empty()) throw std::runtime_error("texture " + t->url + ": a 1x1 image has no mip level (the reference panics, mipmap.go:127)");
    H->r.textures.push_back(std::move(t));
  });
}
static Texture* find_texture(Handle* H, const char* name) {
  for (auto& t : H->r.textures) if (t->url == name) return t.get();
  throw std::runtime_error(std::string("texture not loaded: ") + name);
}
// slot: ShaderStd::k* (node order). filter 1 = "?filter=trilinear" (maps.TextureTrilinear), 0 = Feline (maps.Texture).
int orc_shader_set_texture(void* h, const char* shader, int slot, const char* texname, int chan, int trilinear) {
  Handle* H = (Handle*)h;
  return guard([&] {
    ShaderStd* sh = H->r.findShader(shader);
    if (!sh) throw std::runtime_error(std::string("Unable to find node (shader ") + shader + ")");
    if (slot < 0 || slot >= ShaderStd::kNumSlots || slot == ShaderStd::kSpec1FresnelModel) throw std::runtime_error("bad parameter slot");
    sh->tex[slot].tex = find_texture(H, texname);
    sh->tex[slot].Chan = chan;
    sh->tex[slot].trilinear = trilinear != 0;
    switch (slot) {  // the parameter is present (a non-nil map)
      case ShaderStd::kEmissionColour: sh->hasEmissionColour = true; break;
      case ShaderStd::kEmissionStrength: sh->hasEmissionStrength = true; break;
      case ShaderStd::kDiffuseColour: sh->hasDiffuseColour = true; break;
      case ShaderStd::kDiffuseStrength: sh->hasDiffuseStrength = true; break;
      case ShaderStd::kDiffuseRoughness: sh->hasDiffuseRoughness = true; break;
      case ShaderStd::kSpec1Colour: sh->hasSpec1Colour = true; break;
      case ShaderStd::kSpec1Strength: sh->hasSpec1Strength = true; break;
      case ShaderStd::kSpec1Roughness: sh->hasSpec1Roughness = true; break;
      case ShaderStd::kIOR: sh->hasIOR = true; break;
      case ShaderStd::kSpec1FresnelRefl: sh->hasSpec1FresnelRefl = true; break;
      case ShaderStd::kSpec1FresnelEdge: sh->hasSpec1FresnelEdge = true; break;
    }
  });
}
// PolyMesh.UV / UVIdx (polymesh.go:36-37); before prerender.
int orc_mesh_set_uv(void* h, const char* mesh, const float* uv, int nuv, const int32_t* uvidx, int nuvidx) {
  Handle* H = (Handle*)h;
  return guard([&] {
    for (auto& m : H->r.meshes)
      if (m->Name == mesh) {
        m->UV.resize(nuv);
        std::memcpy(m->UV.data(), uv, sizeof(float) * 2 * (size_t)nuv);
        if (uvidx) { m->hasUVIdx = true; m->UVIdx.assign(uvidx, uvidx + nuvidx); }
        return;
      }
    throw std::runtime_error(std::string("Unable to find node (mesh ") + mesh + ")");
  });
}
int orc_texture_num_levels(void* h, const char* name) {
  Handle* H = (Handle*)h;
  int n = -1;
  guard([&] { n = (int)find_texture(H, name)->mipmap.mipmap.size(); });
  return n;
}
int orc_texture_level(void* h, const char* name, int level, int* w, int* hgt, uint8_t* out) {
  Handle* H = (Handle*)h;
  return guard([&] {
    const MipLevel& L = find_texture(H, name)->mipmap.mipmap.at(level);
    *w = L.w;
    *hgt = L.h;
    if (out) std::memcpy(out, L.mipmap.data(), L.mipmap.size());
  });
}
// coords: n x 8 floats {U, V, Dduvdx[2], Dduvdy[2], PixelDelta[2]}; out: n x 3
int orc_texture_sample(void* h, const char* name, int trilinear, int64_t n, const float* coords, float* out) {
  Handle* H = (Handle*)h;
  return guard([&] {
    TextureMap m;
    m.tex = find_texture(H, name);
    m.trilinear = trilinear != 0;
    for (int64_t i = 0; i < n; i++) {
      const float* c = coords + i * 8;
      TexCoord t;
      t.U = c[0]; t.V = c[1];
      t.Dduvdx[0] = c[2]; t.Dduvdx[1] = c[3];
      t.Dduvdy[0] = c[4]; t.Dduvdy[1] = c[5];
      t.PixelDelta[0] = c[6]; t.PixelDelta[1] = c[7];
      m.Sample(t, out + i * 3);
    }
  });
}
// What the texture maps see at the first hit of camera sample (iter1, x, y): out n x 8 like orc_texture_sample's coords
// (NaN rows for misses), through GenerateCameraRay -> TraceProbe -> DifferentialTransfer (core/trace.go:61-67).
int orc_camera_texcoords(void* h, int iter1, int x0, int y0, int w, int hh, float* out) {
  Handle* H = (Handle*)h;
  return guard([&] {
    H->r.PreRender();
    RenderTask task;
    task.scene = &H->r.scene;
    float pd[2];
    H->r.camera.PixelDelta(H->r.XRes, H->r.YRes, pd);
    Ray ray;
    ray.Task = &task;
    size_t n = 0;
    for (int y = y0; y < y0 + hh; y++)
      for (int x = x0; x < x0 + w; x++) {
        ShaderContext sc;
        sc.task = &task;
        H->r.GenerateCameraRay(iter1, x, y, &sc, &ray);
        ShaderContext sg;
        sg.task = &task;
        sg.Time = ray.Time;
        float* o = out + (n++) * 8;
        if (TraceProbe(&ray, &sg)) {
          ray.DifferentialTransfer(&sg);
          o[0] = sg.U; o[1] = sg.V; o[2] = sg.Dduvdx[0]; o[3] = sg.Dduvdx[1]; o[4] = sg.Dduvdy[0]; o[5] = sg.Dduvdy[1];
          o[6] = pd[0]; o[7] = pd[1];
        } else {
          for (int k = 0; k < 8; k++) o[k] = NAN;
        }
      }
  });
}

// camera rays of iteration `iter1` (1-based, as render() receives it) for the pixel rectangle
int orc_camera_rays(void* h, int iter1, int x0, int y0, int w, int hh, OrcRay* out) {
  Handle* H = (Handle*)h;
  return guard([&] {
    RenderTask task;
    task.scene = &H->r.scene;
    Ray ray;
    ray.Task = &task;
    ShaderContext sc;
    sc.task = &task;
    size_t n = 0;
    for (int y = y0; y < y0 + hh; y++)
      for (int x = x0; x < x0 + w; x++) {
        H->r.GenerateCameraRay(iter1, x, y, &sc, &ray);
        OrcRay& o = out[n++];
        for (int k = 0; k < 3; k++) { o.o[k] = ray.P[k]; o.d[k] = ray.D[k]; }
        o.tmax = ray.Tclosest;
        o.time = ray.Time;
      }
  });
}

// flags: bit0 = any-hit (RayTypeShadow). bit1 = brute force over all triangles of every mesh (anchor).
int orc_trace_batch(void* h, const OrcRay* rays, int64_t n, uint32_t flags, int nthreads, OrcHit* hits) {
  Handle* H = (Handle*)h;
  return guard([&] {
    H->r.PreRender();
    auto work = [&](int64_t b, int64_t e) {
      RenderTask task;
      task.scene = &H->r.scene;
      for (int64_t i = b; i < e; i++) {
        Ray ray;
        ray.Task = &task;
        ShaderContext sc;
        sc.task = &task;
        sc.Time = rays[i].time;
        ray.Init((flags & 1) ? RayTypeShadow : RayTypeCamera, v3(rays[i].o), v3(rays[i].d), rays[i].tmax, 0, &sc);
        bool hit;
        if (flags & 2) {
          hit = false;
          for (Geom* g : H->r.scene.geoms) {
            PolyMesh* pm = dynamic_cast<PolyMesh*>(g);
            if (pm ? pm->TraceBrute(&ray, &sc) : g->Trace(&ray, &sc)) { hit = true; sc.geom = g; }
          }
        } else {
          hit = TraceProbe(&ray, &sc);
        }
        OrcHit& o = hits[i];
        o.nodesT = (int32_t)ray.NodesT;
        o.trisT = (int32_t)ray.TrisT;
        if (hit) {
          o.t = ray.Tclosest; o.u = sc.Bu; o.v = sc.Bv; o.w = sc.Bw;
          o.prim = (int32_t)sc.ElemID;
          o.geom = sc.geom ? sc.geom->id : -1;
        } else {
          o.t = ray.Tclosest; o.u = o.v = o.w = 0;
          o.prim = -1; o.geom = -1;
        }
      }
    };
    if (nthreads <= 1) work(0, n);
    else {
      std::vector<std::thread> th;
      int64_t per = (n + nthreads - 1) / nthreads;
      for (int t = 0; t < nthreads; t++) {
        int64_t b = t * per, e = std::min<int64_t>(n, b + per);
        if (b < e) th.emplace_back(work, b, e);
      }
      for (auto& t : th) t.join();
    }
  });
}

// ---- structure export (so tests can compare the product's host builder and upload oracle-built trees) ----
int orc_num_geoms(void* h) { return (int)((Handle*)h)->r.scene.geoms.size(); }
// geom id (creation order) of the i-th geom in scene leaf order
int orc_scene_geom_id(void* h, int i) { return ((Handle*)h)->r.scene.geoms[i]->id; }
int orc_scene_num_nodes(void* h) {
  Scene& s = ((Handle*)h)->r.scene;
  return s.qbvh.empty() ? (int)s.mqbvh.Nodes.size() : (int)s.qbvh.size();
}
int orc_scene_is_motion(void* h) { return ((Handle*)h)->r.scene.qbvh.empty() ? 1 : 0; }
int orc_scene_keys(void* h) { return (int)((Handle*)h)->r.scene.mqbvh.Boxes.size(); }
int orc_scene_nodes(void* h, void* out) {
  Scene& s = ((Handle*)h)->r.scene;
  std::memcpy(out, s.qbvh.data(), s.qbvh.size() * sizeof(Node));
  return 0;
}
int orc_scene_motion_nodes(void* h, void* topo /*40 B each*/, float* boxes /*[key][node][24]*/) {
  Scene& s = ((Handle*)h)->r.scene;
  std::memcpy(topo, s.mqbvh.Nodes.data(), s.mqbvh.Nodes.size() * sizeof(MotionNode));
  size_t nn = s.mqbvh.Nodes.size();
  for (size_t k = 0; k < s.mqbvh.Boxes.size(); k++) std::memcpy(boxes + k * nn * 24, s.mqbvh.Boxes[k].data(), nn * 96);
  return 0;
}
static PolyMesh* meshById(Handle* H, int id) {
  for (auto& m : H->r.meshes) if (m->id == id) return m.get();
  return nullptr;
}
int orc_mesh_info(void* h, int id, int32_t* out /* nodes, tris, keys, nverts, is_motion, has_normals */) {
  PolyMesh* m = meshById((Handle*)h, id);
  if (!m) return -1;
  out[0] = m->accel.qbvh.empty() ? (int)m->accel.mqbvh.Nodes.size() : (int)m->accel.qbvh.size();
  out[1] = m->facecount;
  out[2] = m->Verts.MotionKeys;
  out[3] = m->Verts.ElemsPerKey;
  out[4] = m->accel.qbvh.empty() ? 1 : 0;
  out[5] = m->Normals.Elems.empty() ? 0 : 1;
  return 0;
}
int orc_mesh_nodes(void* h, int id, void* out) {
  PolyMesh* m = meshById((Handle*)h, id);
  std::memcpy(out, m->accel.qbvh.data(), m->accel.qbvh.size() * sizeof(Node));
  return 0;
}
int orc_mesh_motion_nodes(void* h, int id, void* topo, float* boxes) {
  PolyMesh* m = meshById((Handle*)h, id);
  std::memcpy(topo, m->accel.mqbvh.Nodes.data(), m->accel.mqbvh.Nodes.size() * sizeof(MotionNode));
  size_t nn = m->accel.mqbvh.Nodes.size();
  for (size_t k = 0; k < m->accel.mqbvh.Boxes.size(); k++) std::memcpy(boxes + k * nn * 24, m->accel.mqbvh.Boxes[k].data(), nn * 96);
  return 0;
}
int orc_mesh_idxp(void* h, int id, uint32_t* idxp, int32_t* accel_idx) {
  PolyMesh* m = meshById((Handle*)h, id);
  if (idxp) std::memcpy(idxp, m->idxp.data(), m->idxp.size() * 4);
  if (accel_idx) std::memcpy(accel_idx, m->accel.idx.data(), m->accel.idx.size() * 4);
  return 0;
}
int orc_mesh_bounds(void* h, int id, float* out6) {
  PolyMesh* m = meshById((Handle*)h, id);
  std::memcpy(out6, m->bounds.b, 24);
  return 0;
}
int orc_camera_matrix(void* h, float* m16, float* tan_theta_focal, float* aspect) {
  Handle* H = (Handle*)h;
  std::memcpy(m16, H->r.camera.M.m, 64);
  *tan_theta_focal = H->r.camera.TanThetaFocal;
  *aspect = H->r.camera.Aspect;
  return 0;
}

// ---- scalar function probes for unit tests ----
uint64_t orc_vdc_u(uint64_t i, uint64_t s) { return vanDerCorput_u(i, s); }
uint64_t orc_sobol_u(uint64_t i, uint64_t s) { return sobol_u(i, s); }
double orc_vdc(uint64_t i, uint64_t s) { return VanDerCorput(i, s); }
double orc_sobol(uint64_t i, uint64_t s) { return Sobol(i, s); }
uint64_t orc_raster_xy(uint32_t frame, uint32_t px, uint32_t py, uint64_t sx, uint64_t sy, double* rx, double* ry) {
  return RasterXY12(frame, px, py, sx, sy, rx, ry);
}
// box test probes: ray = P(3),D(3) ; boxes 24 floats (16-B aligned copy made inside)
void orc_box_test(const float* P, const float* D, const float* boxes, int which, int32_t* hits, float* t) {
  Ray ray;
  ray.P = v3(P);
  ray.D = v3(D);
  ray.Setup();
  alignas(16) float b[24];
  alignas(16) float tt[4];
  alignas(16) int32_t hh[4];
  std::memcpy(b, boxes, 96);
  if (which == 0) intersectBoxes(&ray, b, hh, tt);
  else intersectBoxesSlow2(&ray, b, hh, tt);
  std::memcpy(hits, hh, 16);
  std::memcpy(t, tt, 16);
}
// the three box tests of qbvh/intersect.go side by side for one ray with a given Tclosest: which = 0 asm (intersect_amd64.s),
// 1 intersectBoxesSlow2 (:52), 2 intersectBoxesSlow (:17, the Tclosest-clamped one)
void orc_box_test3(const float* P, const float* D, float tclosest, const float* boxes, int which, int32_t* hits, float* t) {
  Ray ray;
  ray.P = v3(P);
  ray.D = v3(D);
  ray.Setup();
  ray.Tclosest = tclosest;
  alignas(16) float b[24];
  alignas(16) float tt[4];
  alignas(16) int32_t hh[4];
  std::memcpy(b, boxes, 96);
  if (which == 0) intersectBoxes(&ray, b, hh, tt);
  else if (which == 1) intersectBoxesSlow2(&ray, b, hh, tt);
  else intersectBoxesSlow(&ray, b, hh, tt);
  std::memcpy(hits, hh, 16);
  std::memcpy(t, tt, 16);
}
void orc_ray_setup(const float* P, const float* D, float* out /* Dinv3 S3 */, int32_t* k /*Kx Ky Kz*/) {
  Ray ray;
  ray.P = v3(P);
  ray.D = v3(D);
  ray.Setup();
  for (int i = 0; i < 3; i++) { out[i] = ray.Dinv[i]; out[3 + i] = ray.S[i]; }
  k[0] = ray.Kx; k[1] = ray.Ky; k[2] = ray.Kz;
}
void orc_normalize(const float* a, float* out) {
  Vec3 r = Vec3Normalize(v3(a));
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
void orc_spectrum_roundtrip(const float* rgb, float lambda, float* spec4, float* rgb_out) {
  Spectrum s;
  s.Lambda = lambda;
  s.FromRGB(MakeRGB(rgb[0], rgb[1], rgb[2]));
  for (int k = 0; k < 4; k++) spec4[k] = s.C[k];
  RGB o = s.ToRGB();
  rgb_out[0] = o[0]; rgb_out[1] = o[1]; rgb_out[2] = o[2];
}

}  // extern "C"
