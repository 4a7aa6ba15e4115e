// Host layer of the C ABI (include/vermeer_gpu.h, vh_*): node creation through the registry, PreRender,
// upload of the pre-rendered scene into a device context through the vg_* layer.
#include <cstdio>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

#include "nodes.h"

using namespace vh;

struct vh_scene {
  Core core;
  std::string err;
};

namespace {
std::string g_err;
int fail(vh_scene* s, int code, const std::string& msg) {
  if (s) s->err = msg; else g_err = msg;
  return code;
}
V3 v3(const float* p) { return V3{p[0], p[1], p[2]}; }
template <class T>
T* make(vh_scene* s, const char* type, std::unique_ptr<Node>* holder) {
  *holder = CreateNode(type);
  if (!*holder) { fail(s, VG_ERR_INVALID, std::string("node type not registered: ") + type); return nullptr; }
  T* t = dynamic_cast<T*>(holder->get());
  if (!t) fail(s, VG_ERR_INVALID, std::string("registered node has the wrong type: ") + type);
  return t;
}
}  // namespace

extern "C" {

int vh_scene_create(vh_scene** out) {
  if (!out) return VG_ERR_INVALID;
  *out = new vh_scene();
  return VG_OK;
}
void vh_scene_destroy(vh_scene* s) { delete s; }
const char* vh_last_error(vh_scene* s) { return s ? s->err.c_str() : g_err.c_str(); }

int vh_registered_nodes(const char** names, int cap) {
  static std::vector<std::string> keep;
  keep = RegisteredNames();
  if (names)
    for (int i = 0; i < cap && i < (int)keep.size(); i++) names[i] = keep[i].c_str();
  return (int)keep.size();
}

int vh_set_option(vh_scene* s, const char* name, int value) {
  if (!s || !name) return VG_ERR_INVALID;
  if (!std::strcmp(name, "leaf_max")) {
    if (value < 1 || value > 16) return fail(s, VG_ERR_INVALID, "leaf_max must be in [1,16] (qbvh/build.go:296-302 clamps to the same range)");
    s->core.leaf_max = value;
    return VG_OK;
  }
  return fail(s, VG_ERR_INVALID, std::string("unknown option ") + name);
}

int vh_set_globals(vh_scene* s, int xres, int yres, int max_iter) {
  if (!s) return VG_ERR_INVALID;
  std::unique_ptr<Node> h;
  Globals* g = make<Globals>(s, "Globals", &h);
  if (!g) return VG_ERR_INVALID;
  if (xres <= 0 || yres <= 0) return fail(s, VG_ERR_INVALID, "Globals: XRes/YRes must be positive");
  g->XRes = xres;
  g->YRes = yres;
  g->MaxIter = max_iter;
  s->core.AddNode(std::move(h));
  return VG_OK;
}

int vh_add_shader_std(vh_scene* s, const char* name, const VgMaterial* params) {
  if (!s || !name || !params) return fail(s, VG_ERR_INVALID, "vh_add_shader_std: null argument");
  std::unique_ptr<Node> h;
  ShaderStd* sh = make<ShaderStd>(s, "ShaderStd", &h);
  if (!sh) return VG_ERR_INVALID;
  sh->MtlName = name;
  sh->params = *params;
  s->core.AddNode(std::move(h));
  return VG_OK;
}

int vh_add_texture(vh_scene* s, const char* name, int w, int h, const uint8_t* rgb8_bottom_up) {
  if (!s || !name || !rgb8_bottom_up || w <= 0 || h <= 0) return fail(s, VG_ERR_INVALID, "vh_add_texture: null/empty argument");
  for (TextureImage& t : s->core.textures)
    if (t.name == name) return VG_OK;  // cacheMiss: the first load of a file name wins (texture.go:178-186)
  TextureImage t;
  t.name = name;
  t.w = w;
  t.h = h;
  t.rgb.assign(rgb8_bottom_up, rgb8_bottom_up + (size_t)w * h * 3);
  s->core.textures.push_back(std::move(t));
  return VG_OK;
}

int vh_shader_set_texture(vh_scene* s, const char* shader, int slot, const char* value) {
  if (!s || !shader || !value) return fail(s, VG_ERR_INVALID, "vh_shader_set_texture: null argument");
  if (slot < 0 || slot > 11 || slot == 9) return fail(s, VG_ERR_INVALID, "vh_shader_set_texture: bad parameter slot");
  ShaderStd* sh = dynamic_cast<ShaderStd*>(s->core.FindNode(shader));
  if (!sh || (sh->params.mask & VG_MAT_DEBUG)) return fail(s, VG_ERR_INVALID, std::string("Unable to find node (shader ") + shader + ")");
  // float parameters go through CreateFloat32TextureMap (reads "ch"), colours through CreateRGBTextureMap (channel 0)
  const bool is_float = slot == 1 || slot == 3 || slot == 4 || slot == 6 || slot == 7 || slot == 8;
  sh->texmaps[slot] = TextureMap::Parse(value, is_float);
  sh->params.mask |= (1u << slot);
  return VG_OK;
}

int vh_polymesh_set_uv(vh_scene* s, const char* mesh, const float* uv, int n_uv, const int32_t* uvidx, int n_uvidx) {
  if (!s || !mesh || !uv || n_uv <= 0) return fail(s, VG_ERR_INVALID, "vh_polymesh_set_uv: null/empty argument");
  if (s->core.prerendered) return fail(s, VG_ERR_INVALID, "vh_polymesh_set_uv: after vh_prerender");
  PolyMesh* m = dynamic_cast<PolyMesh*>(s->core.FindNode(mesh));
  if (!m) return fail(s, VG_ERR_INVALID, std::string("Unable to find node (mesh ") + mesh + ")");
  m->UV.assign(uv, uv + (size_t)n_uv * 2);
  m->hasUVIdx = uvidx != nullptr;
  if (uvidx) m->UVIdx.assign(uvidx, uvidx + n_uvidx);
  return VG_OK;
}

int vh_add_shader_debug(vh_scene* s, const char* name, const float* colour) {
  if (!s || !name || !colour) return fail(s, VG_ERR_INVALID, "vh_add_shader_debug: null argument");
  std::unique_ptr<Node> h;
  DebugShader* sh = make<DebugShader>(s, "DebugShader", &h);
  if (!sh) return VG_ERR_INVALID;
  sh->MtlName = name;
  std::memset(&sh->params, 0, sizeof(sh->params));
  sh->params.mask = VG_MAT_DEBUG | VG_MAT_DIFFUSE_COLOUR;
  for (int i = 0; i < 3; i++) sh->params.diffuse_colour[i] = colour[i];
  s->core.AddNode(std::move(h));
  return VG_OK;
}

int vh_add_polymesh(vh_scene* s, const char* name, const float* verts, int n_verts, int keys, const int32_t* polycount, int n_poly,
                    const int32_t* faceidx, int n_faceidx, const char* shaders_nl, const int32_t* shaderidx, int n_shaderidx,
                    const float* normals, int n_normals, const int32_t* normalidx, int n_normalidx, float raybias) {
  if (!s || !name || !verts || n_verts <= 0 || keys <= 0) return fail(s, VG_ERR_INVALID, "vh_add_polymesh: bad argument");
  std::unique_ptr<Node> h;
  PolyMesh* m = make<PolyMesh>(s, "PolyMesh", &h);
  if (!m) return VG_ERR_INVALID;
  m->NodeName = name;
  m->RayBias = raybias;
  m->Verts.MotionKeys = keys;
  m->Verts.ElemsPerKey = n_verts;
  m->Verts.Elems.resize((size_t)keys * n_verts);
  std::memcpy(m->Verts.Elems.data(), verts, sizeof(float) * 3 * (size_t)keys * n_verts);
  if (polycount) { m->hasPolyCount = true; m->PolyCount.assign(polycount, polycount + n_poly); }
  if (faceidx) { m->hasFaceIdx = true; m->FaceIdx.assign(faceidx, faceidx + n_faceidx); }
  if (m->hasPolyCount) {
    if (!m->hasFaceIdx) return fail(s, VG_ERR_INVALID, "PolyMesh: PolyCount without FaceIdx");
    int64_t tot = 0;
    for (int32_t c : m->PolyCount) { if (c < 3) return fail(s, VG_ERR_INVALID, "PolyMesh: polygon with < 3 vertices"); tot += c; }
    if (tot != n_faceidx) return fail(s, VG_ERR_INVALID, "PolyMesh: sum(PolyCount) != len(FaceIdx)");
  }
  if (shaderidx) m->ShaderIdx.assign(shaderidx, shaderidx + n_shaderidx);
  if (normals) {
    m->Normals.MotionKeys = 1;
    m->Normals.ElemsPerKey = n_normals;
    m->Normals.Elems.resize(n_normals);
    std::memcpy(m->Normals.Elems.data(), normals, sizeof(float) * 3 * (size_t)n_normals);
    if (normalidx) { m->hasNormalIdx = true; m->NormalIdx.assign(normalidx, normalidx + n_normalidx); }
  }
  std::string sn(shaders_nl ? shaders_nl : "");
  size_t pos = 0;
  while (!sn.empty()) {
    size_t e = sn.find('\n', pos);
    m->Shader.push_back(sn.substr(pos, e == std::string::npos ? std::string::npos : e - pos));
    if (e == std::string::npos) break;
    pos = e + 1;
  }
  s->core.AddNode(std::move(h));
  return VG_OK;
}

int vh_add_filter(vh_scene* s, const char* type, const char* name, float width, int res, float peak) {
  if (!s || !type || !name) return fail(s, VG_ERR_INVALID, "vh_add_filter: null argument");
  std::unique_ptr<Node> h;
  PixelFilter* f = make<PixelFilter>(s, type, &h);
  if (!f) return VG_ERR_INVALID;
  f->NodeName = name;
  if (width > 0) f->Width = width;
  if (res > 0) f->Res = res;
  if (peak > 0) f->Peak = peak;
  s->core.AddNode(std::move(h));
  return VG_OK;
}

int vh_add_trilight(vh_scene* s, const char* name, const float* p0, const float* p1, const float* p2, const char* shader, int samples) {
  if (!s || !name || !p0 || !p1 || !p2 || !shader) return fail(s, VG_ERR_INVALID, "vh_add_trilight: null argument");
  if (samples < 0 || samples > 8) return fail(s, VG_ERR_INVALID, "TriLight: Samples must be in [0,8]");
  std::unique_ptr<Node> h;
  TriLight* t = make<TriLight>(s, "TriLight", &h);
  if (!t) return VG_ERR_INVALID;
  t->NodeName = name;
  t->P0 = v3(p0);
  t->P1 = v3(p1);
  t->P2 = v3(p2);
  t->Shader = shader;
  t->Samples = samples;
  s->core.AddNode(std::move(h));
  return VG_OK;
}

int vh_add_instance(vh_scene* s, const char* name, const char* geom, const float* bmin, const float* bmax, int n_bounds, const float* transforms,
                    int keys) {
  if (!s || !name || !geom || !bmin || !bmax || !transforms) return fail(s, VG_ERR_INVALID, "vh_add_instance: null argument");
  if (n_bounds < 1 || keys < 1 || keys > 255) return fail(s, VG_ERR_INVALID, "GeomInstance: need BMin/BMax and 1..255 Transform keys");
  std::unique_ptr<Node> h;
  GeomInstance* g = make<GeomInstance>(s, "GeomInstance", &h);
  if (!g) return VG_ERR_INVALID;
  g->NodeName = name;
  g->GeomName = geom;
  for (int i = 0; i < n_bounds; i++) { g->BMin.push_back(v3(bmin + 3 * i)); g->BMax.push_back(v3(bmax + 3 * i)); }
  for (int k = 0; k < keys; k++) {
    M4 m;
    std::memcpy(m.m, transforms + 16 * k, sizeof(m.m));
    g->Transform.push_back(m);
  }
  s->core.AddNode(std::move(h));
  return VG_OK;
}

int vh_add_disklight(vh_scene* s, const char* name, const float* P, const float* lookat, const float* up, float radius, const char* shader,
                     int segments, int samples) {
  if (!s || !name || !P || !lookat || !up || !shader) return fail(s, VG_ERR_INVALID, "vh_add_disklight: null argument");
  if (samples < 0 || samples > 8) return fail(s, VG_ERR_INVALID, "DiskLight: Samples must be in [0,8]");
  if (segments < 3 || segments > 4096) return fail(s, VG_ERR_INVALID, "DiskLight: Segments must be in [3,4096]");
  std::unique_ptr<Node> h;
  DiskLight* d = make<DiskLight>(s, "DiskLight", &h);
  if (!d) return VG_ERR_INVALID;
  d->NodeName = name;
  d->P = v3(P);
  d->LookAt = v3(lookat);
  d->Up = v3(up);
  d->Radius = radius;
  d->Shader = shader;
  d->Segments = segments;
  d->Samples = samples;
  s->core.AddNode(std::move(h));
  return VG_OK;
}

int vh_add_spherelight(vh_scene* s, const char* name, const float* P, float radius, const char* shader, int samples) {
  if (!s || !name || !P || !shader) return fail(s, VG_ERR_INVALID, "vh_add_spherelight: null argument");
  if (samples < 0 || samples > 8) return fail(s, VG_ERR_INVALID, "SphereLight: Samples must be in [0,8]");
  std::unique_ptr<Node> h;
  SphereLight* d = make<SphereLight>(s, "SphereLight", &h);
  if (!d) return VG_ERR_INVALID;
  d->NodeName = name;
  d->P = v3(P);
  d->Radius = radius;
  d->Shader = shader;
  d->Samples = samples;
  s->core.AddNode(std::move(h));
  return VG_OK;
}

int vh_set_camera_lookat(vh_scene* s, const float* from, const float* to, const float* up, float roll, float fov, float focal, float aspect,
                         float radius) {
  if (!s || !from || !to || !up) return fail(s, VG_ERR_INVALID, "vh_set_camera_lookat: null argument");
  std::unique_ptr<Node> h;
  Camera* c = make<Camera>(s, "Camera", &h);
  if (!c) return VG_ERR_INVALID;
  c->From = v3(from);
  c->To = v3(to);
  c->Up = v3(up);
  c->Roll = roll;
  c->Fov = fov;
  c->Focal = focal;
  c->Aspect = aspect;
  c->Radius = radius;
  s->core.AddNode(std::move(h));
  return VG_OK;
}

int vh_set_camera_keys(vh_scene* s, const char* type, const float* from, int n_from, const float* to, int n_to, const float* roll, int n_roll,
                       const float* up, const float* world_to_local, int n_mat, float fov, float focal, float aspect, float radius) {
  if (!s || !type || !up) return fail(s, VG_ERR_INVALID, "vh_set_camera_keys: null argument");
  if (n_from < 0 || n_to < 0 || n_roll < 0 || n_mat < 0 || (n_from > 0 && !from) || (n_to > 0 && !to) || (n_roll > 0 && !roll) ||
      (n_mat > 0 && !world_to_local))
    return fail(s, VG_ERR_INVALID, "vh_set_camera_keys: bad key arrays");
  if (std::string(type) == "LookAt" && (n_from < 1 || n_to < 1)) return fail(s, VG_ERR_INVALID, "vh_set_camera_keys: a LookAt camera needs From and To");
  std::unique_ptr<Node> h;
  Camera* c = make<Camera>(s, "Camera", &h);
  if (!c) return VG_ERR_INVALID;
  c->Type = type;
  for (int i = 0; i < n_from; i++) c->FromKeys.push_back(v3(from + 3 * i));
  for (int i = 0; i < n_to; i++) c->ToKeys.push_back(v3(to + 3 * i));
  for (int i = 0; i < n_roll; i++) c->RollKeys.push_back(roll[i]);
  for (int i = 0; i < n_mat; i++) {
    M4 m;
    std::memcpy(m.m, world_to_local + 16 * i, sizeof(m.m));
    c->WorldToLocal.push_back(m);
  }
  if (n_from > 0) c->From = c->FromKeys[0];
  if (n_to > 0) c->To = c->ToKeys[0];
  c->Up = v3(up);
  c->Fov = fov;
  c->Focal = focal;
  c->Aspect = aspect;
  c->Radius = radius;
  s->core.AddNode(std::move(h));
  return VG_OK;
}

int vh_parse_vnf(vh_scene* s, const char* text, size_t len, const char* filename) {
  if (!s || (!text && len > 0)) return fail(s, VG_ERR_INVALID, "vh_parse_vnf: null argument");
  std::string msgs;
  const int n = ParseVnf(s->core, text ? text : "", len, filename ? filename : "<memory>", &msgs);
  s->err = msgs;
  return n;
}

int vh_load_vnf(vh_scene* s, const char* path) {
  if (!s || !path) return fail(s, VG_ERR_INVALID, "vh_load_vnf: null argument");
  FILE* fp = std::fopen(path, "rb");
  if (!fp) return fail(s, VG_ERR_INVALID, std::string("open ") + path + ": no such file or directory");
  std::string text;
  char buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof(buf), fp)) > 0) text.append(buf, n);
  std::fclose(fp);
  return vh_parse_vnf(s, text.data(), text.size(), path);
}

int vh_globals(vh_scene* s, int32_t* out3) {
  if (!s || !out3) return VG_ERR_INVALID;
  out3[0] = s->core.globals->XRes;
  out3[1] = s->core.globals->YRes;
  out3[2] = s->core.globals->MaxIter;
  return VG_OK;
}

int vh_postrender(vh_scene* s, const float* framebuffer, int xres, int yres) {
  if (!s || !framebuffer || xres <= 0 || yres <= 0) return fail(s, VG_ERR_INVALID, "vh_postrender: bad argument");
  // core.PostRender (core/core.go:63-73): every node in creation order; only the output drivers do anything
  for (Node* n : s->core.all) {
    if (OutputNode* o = dynamic_cast<OutputNode*>(n)) {
      std::string err;
      if (o->Write(framebuffer, xres, yres, &err) != 0) return fail(s, VG_ERR_INVALID, err);
    }
  }
  return VG_OK;
}

int vh_rgbe(float r, float g, float b, uint8_t* out4) {
  if (!out4) return VG_ERR_INVALID;
  RgbToRgbe(r, g, b, out4);
  return VG_OK;
}

int vh_prerender(vh_scene* s) {
  if (!s) return VG_ERR_INVALID;
  if (s->core.PreRender() != 0) return fail(s, VG_ERR_BUILD, s->core.err);
  if (!s->core.include_log.empty()) s->err = s->core.include_log;  // what nodes.Parse printed for the Include'd files
  return VG_OK;
}

int vh_prerender_device(vh_scene* s, vg_ctx* ctx) {
  if (!s || !ctx) return VG_ERR_INVALID;
  s->core.build_ctx = ctx;
  const int rc = vh_prerender(s);
  s->core.build_ctx = nullptr;
  return rc;
}

static Geom* geom_by_id(vh_scene* s, int id) {
  for (Geom* g : s->core.scene.geoms)
    if (g->id == id) return g;
  return nullptr;
}
static PolyMesh* mesh_by_id(vh_scene* s, int id) { return dynamic_cast<PolyMesh*>(geom_by_id(s, id)); }

int vh_upload(vh_scene* s, vg_ctx* ctx, int motion_ref_compat) {
  if (!s || !ctx) return fail(s, VG_ERR_INVALID, "vh_upload: null argument");
  if (!s->core.prerendered) return fail(s, VG_ERR_INVALID, "vh_upload: call vh_prerender first");
  Core& c = s->core;
  auto chk = [&](int rc) { if (rc != VG_OK) s->err = std::string("device layer: ") + vg_last_error(ctx); return rc; };
  int rc;
  const int G = (int)c.scene.geoms.size();
  if ((rc = chk(vg_set_frame(ctx, c.globals->XRes, c.globals->YRes))) != VG_OK) return rc;
  std::vector<VgMaterial> mats;
  for (ShaderStd* sh : c.materials) mats.push_back(sh->params);
  if ((rc = chk(vg_set_materials(ctx, mats.data(), (int)mats.size()))) != VG_OK) return rc;
  if ((rc = chk(vg_scene_begin(ctx, G))) != VG_OK) return rc;
  const bool timing = std::getenv("VG_TIMING") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  for (int id = 0; id < G; id++) {
    Geom* gm = geom_by_id(s, id);
    if (!gm) return fail(s, VG_ERR_INVALID, "geom ids are not dense");
    if (SphereGeom* sp = dynamic_cast<SphereGeom*>(gm)) {
      const float ctr[3] = {sp->P.x, sp->P.y, sp->P.z};
      if (chk(vg_sphere_upload(ctx, id, ctr, sp->Radius, sp->shader ? sp->shader->material_id : -1)) != VG_OK) return VG_ERR_INVALID;
      continue;
    }
    if (GeomInstance* gi = dynamic_cast<GeomInstance*>(gm)) {
      if (chk(vg_instance_upload(ctx, id, gi->geom->id, gi->transformSRT.data(), (int)gi->transformSRT.size())) != VG_OK) return VG_ERR_INVALID;
      continue;
    }
    PolyMesh* m = dynamic_cast<PolyMesh*>(gm);
    if (!m) return fail(s, VG_ERR_INVALID, "unknown geom type");
    std::vector<int32_t> mids;
    for (ShaderStd* sh : m->shader) mids.push_back(sh->material_id);
    if (!m->qbvh.empty()) {
      rc = vg_mesh_upload(ctx, id, m->qbvh.data(), (int)m->qbvh.size(), m->idxp.data(), m->facecount, &m->Verts.Elems[0].x,
                          m->Verts.ElemsPerKey, m->shaderidx.empty() ? nullptr : m->shaderidx.data(), mids.data(), (int)mids.size(),
                          m->Normals.Elems.empty() ? nullptr : &m->Normals.Elems[0].x, (int)m->Normals.Elems.size(),
                          m->normalidx.empty() ? nullptr : m->normalidx.data(), m->RayBias);
    } else {
      rc = vg_mesh_upload_motion(ctx, id, m->mtopo.data(), (int)m->mtopo.size(), m->mboxes.data(), m->Verts.MotionKeys, m->idxp.data(),
                                 m->accel_idx.data(), m->facecount, &m->Verts.Elems[0].x, m->Verts.ElemsPerKey,
                                 m->shaderidx.empty() ? nullptr : m->shaderidx.data(), mids.data(), (int)mids.size(), m->RayBias,
                                 motion_ref_compat);
    }
    if (chk(rc) != VG_OK) return rc;
    if (!m->uvtriidx.empty() && !m->qbvh.empty())
      if ((rc = chk(vg_mesh_set_uv(ctx, id, m->UV.data(), (int)(m->UV.size() / 2), m->uvtriidx.data()))) != VG_OK) return rc;
  }
  if (timing) std::fprintf(stderr, "[vh_upload] mesh staging           %8.1f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
  // texture maps: the files the shaders name, each uploaded once (texture.TexStore), then the parameter bindings
  {
    if ((rc = chk(vg_textures_clear(ctx))) != VG_OK) return rc;
    std::map<std::string, int> ids;
    for (size_t mi = 0; mi < c.materials.size(); mi++)
      for (int slot = 0; slot < 12; slot++) {
        const TextureMap& tm = c.materials[mi]->texmaps[slot];
        if (!tm.set) continue;
        auto it = ids.find(tm.path);
        if (it == ids.end()) {
          const TextureImage* img = nullptr;
          for (const TextureImage& t : c.textures)
            if (t.name == tm.path) img = &t;
          if (!img)
            return fail(s, VG_ERR_INVALID, "texture \"" + tm.path + "\" (shader " + c.materials[mi]->MtlName +
                                               ") was not registered: decode the file and pass it to vh_add_texture before vh_upload");
          int id = -1;
          if ((rc = chk(vg_texture_upload(ctx, img->rgb.data(), img->w, img->h, &id))) != VG_OK) return rc;
          it = ids.emplace(tm.path, id).first;
        }
        if ((rc = chk(vg_material_set_texture(ctx, (int)mi, slot, it->second, tm.chan, tm.filter))) != VG_OK) return rc;
      }
  }
  std::vector<int32_t> order;
  for (Geom* g : c.scene.geoms) order.push_back(g->id);
  if (c.scene.keys == 1)
    rc = vg_scene_upload(ctx, c.scene.qbvh.data(), (int)c.scene.qbvh.size(), order.data(), (int)order.size());
  else
    rc = vg_scene_upload_motion(ctx, c.scene.mtopo.data(), (int)c.scene.mtopo.size(), c.scene.mboxes.data(), c.scene.keys, order.data(), (int)order.size());
  if (chk(rc) != VG_OK) return rc;
  if ((rc = chk(vg_scene_commit(ctx))) != VG_OK) return rc;

  std::vector<VgLight> lights;
  for (Light* t : c.scene.lights) {
    VgLight l;
    t->Describe(&l);
    lights.push_back(l);
  }
  if ((rc = chk(vg_set_area_lights(ctx, lights.data(), (int)lights.size()))) != VG_OK) return rc;

  if (c.filter) rc = vg_set_filter(ctx, c.filter->Res, (double)c.filter->Width, c.filter->cdfV.data(), c.filter->cdfVU.data());
  else rc = vg_set_filter(ctx, 0, 0.0, nullptr, nullptr);
  if (chk(rc) != VG_OK) return rc;

  // core.Render finds the node named "camera" unless Globals.Camera overrides it (core/render.go:148-164)
  std::string camName = c.globals->Camera.empty() ? "camera" : c.globals->Camera;
  Camera* cam = dynamic_cast<Camera*>(c.FindNode(camName));
  if (cam) {
    if ((rc = chk(vg_set_camera_motion(ctx, &cam->out, cam->decomp.data(), (int)cam->decomp.size()))) != VG_OK) return rc;
  }
  return VG_OK;
}

int vh_num_geoms(vh_scene* s) { return s ? (int)s->core.scene.geoms.size() : 0; }

int vh_scene_info(vh_scene* s, int32_t* out4) {
  if (!s || !out4) return VG_ERR_INVALID;
  Scene& sc = s->core.scene;
  out4[0] = sc.keys == 1 ? (int)sc.qbvh.size() : (int)sc.mtopo.size();
  out4[1] = sc.keys == 1 ? 0 : 1;
  out4[2] = sc.keys;
  out4[3] = (int)sc.geoms.size();
  return VG_OK;
}
int vh_scene_nodes(vh_scene* s, VgNode* out) {
  if (!s || !out) return VG_ERR_INVALID;
  std::memcpy(out, s->core.scene.qbvh.data(), s->core.scene.qbvh.size() * sizeof(VgNode));
  return VG_OK;
}
int vh_scene_motion_nodes(vh_scene* s, VgMotionNode* topo, float* boxes) {
  if (!s || !topo || !boxes) return VG_ERR_INVALID;
  Scene& sc = s->core.scene;
  std::memcpy(topo, sc.mtopo.data(), sc.mtopo.size() * sizeof(VgMotionNode));
  std::memcpy(boxes, sc.mboxes.data(), sc.mboxes.size() * sizeof(float));
  return VG_OK;
}
int vh_scene_geom_order(vh_scene* s, int32_t* out) {
  if (!s || !out) return VG_ERR_INVALID;
  int i = 0;
  for (Geom* g : s->core.scene.geoms) out[i++] = g->id;
  return VG_OK;
}
int vh_mesh_info(vh_scene* s, int geom_id, int32_t* o) {
  if (!s || !o) return VG_ERR_INVALID;
  PolyMesh* m = mesh_by_id(s, geom_id);
  if (!m) return fail(s, VG_ERR_INVALID, "no such geom");
  o[0] = m->qbvh.empty() ? (int)m->mtopo.size() : (int)m->qbvh.size();
  o[1] = m->facecount;
  o[2] = m->Verts.MotionKeys;
  o[3] = m->Verts.ElemsPerKey;
  o[4] = m->qbvh.empty() ? 1 : 0;
  o[5] = m->Normals.Elems.empty() ? 0 : 1;
  return VG_OK;
}
int vh_mesh_nodes(vh_scene* s, int geom_id, VgNode* out) {
  PolyMesh* m = s ? mesh_by_id(s, geom_id) : nullptr;
  if (!m || !out) return VG_ERR_INVALID;
  std::memcpy(out, m->qbvh.data(), m->qbvh.size() * sizeof(VgNode));
  return VG_OK;
}
int vh_mesh_motion_nodes(vh_scene* s, int geom_id, VgMotionNode* topo, float* boxes) {
  PolyMesh* m = s ? mesh_by_id(s, geom_id) : nullptr;
  if (!m || !topo || !boxes) return VG_ERR_INVALID;
  std::memcpy(topo, m->mtopo.data(), m->mtopo.size() * sizeof(VgMotionNode));
  std::memcpy(boxes, m->mboxes.data(), m->mboxes.size() * sizeof(float));
  return VG_OK;
}
int vh_mesh_idxp(vh_scene* s, int geom_id, uint32_t* idxp, int32_t* accel_idx) {
  PolyMesh* m = s ? mesh_by_id(s, geom_id) : nullptr;
  if (!m) return VG_ERR_INVALID;
  if (idxp) std::memcpy(idxp, m->idxp.data(), m->idxp.size() * 4);
  if (accel_idx) std::memcpy(accel_idx, m->accel_idx.data(), m->accel_idx.size() * 4);
  return VG_OK;
}
int vh_camera(vh_scene* s, VgCamera* out) {
  if (!s || !out) return VG_ERR_INVALID;
  Core& c = s->core;
  std::string camName = c.globals->Camera.empty() ? "camera" : c.globals->Camera;
  Camera* cam = dynamic_cast<Camera*>(c.FindNode(camName));
  if (!cam) return fail(s, VG_ERR_INVALID, "no camera node");
  *out = cam->out;
  return VG_OK;
}
int vh_camera_decomp(vh_scene* s, VgTransformSRT* out) {
  if (!s) return VG_ERR_INVALID;
  Core& c = s->core;
  std::string camName = c.globals->Camera.empty() ? "camera" : c.globals->Camera;
  Camera* cam = dynamic_cast<Camera*>(c.FindNode(camName));
  if (!cam) return fail(s, VG_ERR_INVALID, "no camera node");
  if (out) std::memcpy(out, cam->decomp.data(), cam->decomp.size() * sizeof(VgTransformSRT));
  return (int)cam->decomp.size();
}

}  // extern "C"
