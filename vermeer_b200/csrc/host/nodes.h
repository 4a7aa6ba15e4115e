// Host mirror of the reference's node registry and PreRender pipeline (the vh_* layer).
//
//   nodes.Register / createNode          nodes/register.go:14-33
//   core.Node                            core/node.go:14-28      (Name, PreRender, PostRender)
//   core.AddNode / FindNode / PreRender  core/core.go:36-105
//   core.Geom                            core/geom.go:8-18       (MotionKeys, Bounds; Trace runs on the device)
//   core.Scene (AddGeom, AddLight, PreRender) core/scene.go:8-24, builtin/scene/scene.go:119-268
//   PolyMesh, Sphere, ShaderStd, TriLight, DiskLight, SphereLight, Camera, Globals: see nodes.cpp
//
// Same names and argument meaning as the reference; error behaviour: the reference returns `error` from
// PreRender and panics on invariant violations — here both become an int status + message, never an abort.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/vermeer_gpu.h"
#include "hmath.h"

namespace vh {

struct Core;

struct Node {
  virtual ~Node() {}
  virtual std::string Name() const = 0;
  virtual int PreRender(Core& core, std::string* err) { (void)core; (void)err; return 0; }
  virtual int PostRender() { return 0; }
};

using CreateFn = std::function<std::unique_ptr<Node>()>;
// nodes.Register: error (-1) if `name` is already registered.
int Register(const std::string& name, CreateFn create);
std::unique_ptr<Node> CreateNode(const std::string& name);
std::vector<std::string> RegisteredNames();

struct Geom {
  virtual ~Geom() {}
  virtual int MotionKeys() const = 0;
  virtual Box Bounds(float time) const = 0;
  int id = -1;  // creation order (core.AddNode order)
};

struct Globals : Node {
  int XRes = 1024, YRes = 1024, MaxGoRoutines = 5, MaxIter = 16;  // core/core.go:17-23
  std::string Camera;
  std::string Name() const override { return "<globals>"; }
};

// builtin/maps/texture.go:17-83: a shader parameter read from a texture file. `Parse` is CreateRGBTextureMap /
// CreateFloat32TextureMap: url.Parse the value, path = EscapedPath, "?filter=trilinear" picks TextureTrilinear, "ch=N" the
// channel of a float parameter (only CreateFloat32TextureMap reads it; the .vnf `rgbtex` form always goes through
// CreateRGBTextureMap, nodes/parser.go:259, so its channel is 0).
struct TextureMap {
  bool set = false;
  std::string path;
  int chan = 0;
  int filter = VG_TEXFILTER_FELINE;
  static TextureMap Parse(const std::string& value, bool read_channel);
};

struct ShaderStd : Node {
  std::string MtlName;
  VgMaterial params{};
  TextureMap texmaps[12];  // by parameter slot (bit index of the VG_MAT_* flag)
  int material_id = -1;
  std::string Name() const override { return MtlName; }
};

// builtin/shader/debug.go:15-22 ("DebugShader"): a core.Shader whose Eval is OutRGB = Colour. It shares ShaderStd's slot in the
// material table; params.mask = VG_MAT_DEBUG | VG_MAT_DIFFUSE_COLOUR, Colour in params.diffuse_colour.
struct DebugShader : ShaderStd {};

struct PointArray {  // core/param/array.go:19-28
  int MotionKeys = 0, ElemsPerKey = 0;
  std::vector<V3> Elems;
};

struct PolyMesh : Node, Geom {
  std::string NodeName;
  float RayBias = 0;
  PointArray Verts;
  std::vector<int32_t> PolyCount, FaceIdx, ShaderIdx, NormalIdx;
  bool hasPolyCount = false, hasFaceIdx = false, hasNormalIdx = false;
  PointArray Normals;
  std::vector<float> UV;         // param.Vec2Array, one key (polymesh.go:36)
  std::vector<int32_t> UVIdx;    // polymesh.go:37
  bool hasUVIdx = false;
  std::vector<uint32_t> uvtriidx;  // triangulated UV indexes (polymesh.go:45), leaf order after initAccel
  std::vector<std::string> Shader;
  bool IsVisible = true;

  // products of PreRender (polymesh.go:41-57)
  int facecount = 0;
  std::vector<uint32_t> idxp, normalidx;
  std::vector<uint8_t> shaderidx;
  std::vector<VgNode> qbvh;
  std::vector<VgMotionNode> mtopo;
  std::vector<float> mboxes;  // [key][node][24]
  std::vector<int32_t> accel_idx;
  std::vector<ShaderStd*> shader;
  Box bounds;
  std::vector<Box> motionBounds;

  std::string Name() const override { return NodeName; }
  int PreRender(Core& core, std::string* err) override;
  int MotionKeys() const override { return !qbvh.empty() ? 1 : Verts.MotionKeys; }
  Box Bounds(float time) const override;

 private:
  void triangulate();
  int initAccel(std::string* err, vg_ctx* build_ctx, int leaf_max = 16);
  Box initMotionBoxesRec(int key, int32_t node);
};

// core.Light (core/light.go:21-42) as far as the host needs it: which geom the light made for itself and the record the
// device consumes. Sampling (SampleArea / ValidSample) runs on the device.
struct Light {
  virtual ~Light() {}
  virtual Geom* LightGeom() const = 0;
  virtual void Describe(VgLight* out) const = 0;
};

// sphere.Sphere (builtin/geom/sphere/sphere.go:15-86): "only used for spherical light sources"
struct SphereGeom : Node, Geom {
  std::string NodeName;
  V3 P{};
  float Radius = 1;
  std::string Shader;
  ShaderStd* shader = nullptr;
  std::string Name() const override { return NodeName; }
  int PreRender(Core& core, std::string* err) override;
  int MotionKeys() const override { return 1; }
  Box Bounds(float time) const override;
};

// instance.Instance (builtin/geom/instance/instance.go:36-160)
struct GeomInstance : Node, Geom {
  std::string NodeName;
  std::string GeomName;           // `Geom`
  std::vector<V3> BMin, BMax;     // param.PointArray elements
  std::vector<M4> Transform;      // param.MatrixArray elements (column major)
  std::vector<VgTransformSRT> transformSRT;
  std::vector<Box> bounds;
  Geom* geom = nullptr;
  std::string Name() const override { return NodeName; }
  int PreRender(Core& core, std::string* err) override;
  int MotionKeys() const override { return geom ? geom->MotionKeys() : 1; }
  Box Bounds(float) const override { return bounds[0]; }
};

struct TriLight : Node, Light {
  std::string NodeName;
  V3 P0{}, P1{}, P2{};
  std::string Shader;
  int Samples = 1;
  ShaderStd* shader = nullptr;
  PolyMesh* geom = nullptr;
  std::string Name() const override { return NodeName; }
  int PreRender(Core& core, std::string* err) override;
  Geom* LightGeom() const override { return geom; }
  void Describe(VgLight* out) const override;
};

// light.Disk (builtin/light/disk.go:21-34,76-106,224-262)
struct DiskLight : Node, Light {
  std::string NodeName;
  V3 P{}, Up{}, LookAt{};
  V3 T{}, B{}, N{};
  float Radius = 0;
  std::string Shader;
  int Segments = 20, Samples = 1;  // registered defaults, disk.go:263-269
  ShaderStd* shader = nullptr;
  PolyMesh* geom = nullptr;
  std::string Name() const override { return NodeName; }
  int PreRender(Core& core, std::string* err) override;
  Geom* LightGeom() const override { return geom; }
  void Describe(VgLight* out) const override;
};

// light.Sphere (builtin/light/sphere.go:17-28,38-61)
struct SphereLight : Node, Light {
  std::string NodeName;
  V3 P{};
  float Radius = 1;  // registered defaults Radius 1, Samples 1 (sphere.go:297-303)
  std::string Shader;
  int Samples = 1;
  ShaderStd* shader = nullptr;
  SphereGeom* geom = nullptr;
  std::string Name() const override { return NodeName; }
  int PreRender(Core& core, std::string* err) override;
  Geom* LightGeom() const override { return geom; }
  void Describe(VgLight* out) const override;
};

struct Camera : Node {
  std::string NodeName = "camera";
  std::string Type = "LookAt";
  V3 From{}, To{}, Up{0, 1, 0};
  float Roll = 0;
  std::vector<V3> FromKeys, ToKeys;   // motion keys of From / To (empty: the single From / To above)
  std::vector<float> RollKeys;
  std::vector<M4> WorldToLocal;       // Type "Matrix"
  float Aspect = 0, Fov = 90, Focal = 12, Radius = 0;  // camera.go:325-335 defaults
  VgCamera out{};                     // single-key matrix (Time 0)
  std::vector<VgTransformSRT> decomp; // c.decomp: one per LocalToWorld key (camera.go:188-192)
  std::string Name() const override { return NodeName; }
  int PreRender(Core& core, std::string* err) override;
};

// builtin/filter: core.PixelFilter nodes. PreRender tabulates the filter and builds the FIS CDF tables
// (filter.CreateSampler, filter.go:88-173); WarpSample itself runs on the device.
struct PixelFilter : Node {
  std::string NodeName;
  int kind = 1;  // 1 Airy, 2 Gaussian
  float Width = 6, Peak = 4;
  int Res = 49;
  std::vector<double> cdfV, cdfVU;  // n, n*n
  std::string Name() const override { return NodeName; }
  int PreRender(Core& core, std::string* err) override;
};

// driver.OutputFloat / driver.OutputHDR (builtin/driver/outputfloat.go, outputhdr.go): PostRender writes the finished frame.
struct OutputNode : Node {
  std::string Filename;
  bool hdr = false;
  std::string Name() const override { return hdr ? "OutputHDR<>" : "OutputFloat<>"; }
  int Write(const float* fb, int w, int h, std::string* err) const;  // vnf.cpp
};
void RgbToRgbe(float r, float g, float b, uint8_t out[4]);  // image/hdr/hdr.go:26-50

// misc.Include (builtin/misc/include.go:9-36): PreRender parses another .vnf file into the core; the nodes it adds are
// pre-rendered in the next round of core.PreRender (core/core.go:46-57). Only a file that cannot be opened is an error
// (nodes.Parse returns nil after printing its parse errors, parser.go:862-899); those land in Core::include_log.
struct Include : Node {
  std::string NodeName, Filename;
  std::string Name() const override { return NodeName; }
  int PreRender(Core& core, std::string* err) override;  // vnf.cpp
};

// nodes.Parse (nodes/parser.go): adds the file's nodes to `core` in file order; returns the number of parse errors.
int ParseVnf(Core& core, const char* text, size_t len, const std::string& filename, std::string* messages);

// builtin/scene/scene.go
struct Scene {
  std::vector<Geom*> geoms;
  std::vector<Light*> lights;
  std::vector<VgNode> qbvh;
  std::vector<VgMotionNode> mtopo;
  std::vector<float> mboxes;
  int keys = 1;
  Box bounds;
  void AddGeom(Geom* g) { geoms.push_back(g); }
  void AddLight(Light* l) { lights.push_back(l); }
  int PreRender(std::string* err);

 private:
  Box initMotionBoxesRec(int key, int32_t node, int nkeys);
};

// One entry of texture.TexStore (texture/texture.go:36-47): the decoded image, rows bottom-up (loadTexture flips, :139).
// Image decoding is the caller's (vh_add_texture); a map whose file was never registered is an error at vh_upload, where the
// reference would substitute its embedded checker image (texture.go:103-106,187-190).
struct TextureImage {
  std::string name;
  int w = 0, h = 0;
  std::vector<uint8_t> rgb;
};

struct Core {
  vg_ctx* build_ctx = nullptr;  // vh_prerender_device: static-mesh QBVHs are built on this device (vg_build_qbvh)
  // leafMax of the per-mesh builds. 16 is what the reference passes (buildqbvh.go:85, motionbuild.go callers); any other value is the
  // opt-in NON-PARITY mode of vh_set_option("leaf_max"): another tree, same triangles, same intersection routine.
  int leaf_max = 16;
  std::vector<TextureImage> textures;
  Globals* globals = nullptr;
  Scene scene;
  std::vector<std::unique_ptr<Node>> owned;
  std::vector<Node*> pending;  // nodes added since the last PreRender round (core.go:46-58)
  std::vector<Node*> all;
  std::map<std::string, Node*> nodeMap;
  std::vector<ShaderStd*> materials;
  PixelFilter* filter = nullptr;  // core.filter (core/core.go:15,89)
  int next_geom_id = 0;
  bool prerendered = false;
  std::string err;
  std::string include_log;  // parse messages of files read by Include nodes
  int include_errors = 0;

  Core();
  void AddNode(std::unique_ptr<Node> node);  // core.go:77-93 type-switch wiring
  Node* FindNode(const std::string& name) const;
  int PreRender();
  float FrameAspect() const { return (float)globals->XRes / (float)globals->YRes; }
};

}  // namespace vh
