// .vnf scene reader and output drivers of the host layer (vh_parse_vnf / vh_load_vnf / vh_postrender).
//
// Mirrors, for the node types on this path:
//   nodes.Lex                      nodes/lex.go:60-261     (tokens, numbers, strings, '#' comments)
//   nodes.Parse / parser.node      nodes/parser.go:110-131,757-940 (NodeType { Field value ... }, required/optional fields,
//                                  unknown fields skipped up to the next non-keyword token)
//   parser.parseParam and friends  nodes/parser.go:133-640 (scalars, "<n> int|string" slices, "rgb r g b" / "float f" constant
//                                  maps, vec3, "<keys> <n> point|vec3|vec2|float" arrays, "<keys> matrix" arrays)
//   driver.OutputFloat / OutputHDR builtin/driver/outputfloat.go:30-42, outputhdr.go:30-57, image/hdr/hdr.go:26-50,
//                                  image/hdr/writer.go:58-91
// Error behaviour: the reference prints "<file>:<line>:<col>: message", keeps going and exits after more than 10 errors;
// here the messages are collected (vh_last_error), the count is returned, and parsing stops after more than 10.
// Node types that are registered in the reference but out of scope here (QuadLight, Proc, Include,
// DebugShader) are reported like an unregistered type. `rgbtex "file?filter=trilinear"` maps are kept as TextureMap bindings.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

#include "nodes.h"

namespace vh {

namespace {

enum Tok { TEof = 0, TToken, TString, TFloat, TInt, TOpenBrace, TCloseBrace, TOpenCurly, TCloseCurly, TComma };
struct Sym {
  double numFloat = 0;
  int64_t numInt = 0;
  std::string str;
};

// nodes/lex.go. Bytes are treated as Latin letters when >= 0x80 (the reference decodes UTF-8 runes and asks unicode.IsLetter;
// identifiers in scene files are ASCII).
class Lex {
 public:
  Lex(const char* text, size_t len) : p_(text), end_(text + len) {}
  int LineNumber = 0, ColNumber = 1, BeginColNumber = 0;

  int lex(Sym* v) {
    if (peekToken_) {
      peekToken_ = false;
      *v = psym_;
      return ptoken_;
    }
    return lex1(v);
  }
  int peek(Sym* v) {
    const int t = lex1(v);
    peekToken_ = true;
    psym_ = *v;
    ptoken_ = t;
    return t;
  }
  void skip() {
    Sym s;
    lex(&s);
  }

 private:
  static const int kEof = -1;
  const char* p_;
  const char* end_;
  std::string line_;
  size_t lpos_ = 0;
  int peekc_ = kEof;
  bool peekToken_ = false;
  Sym psym_;
  int ptoken_ = 0;

  static bool isAlpha(int c) { return c == '_' || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c >= 0x80; }

  bool readLine() {  // lex.go:207-219
    ColNumber = 0;
    LineNumber++;
    if (p_ >= end_) return false;
    const char* nl = (const char*)memchr(p_, '\n', (size_t)(end_ - p_));
    const char* e = nl ? nl + 1 : end_;
    line_.assign(p_, e);
    lpos_ = 0;
    p_ = e;
    return true;
  }
  int next() {  // lex.go:222-261
    if (peekc_ != kEof) {
      const int r = peekc_;
      peekc_ = kEof;
      return r;
    }
    for (;;) {
      ColNumber++;
      if (lpos_ >= line_.size()) {
        if (!readLine()) return kEof;
        continue;
      }
      const int c = (unsigned char)line_[lpos_++];
      if (c == '#') {
        if (!readLine()) return kEof;
        return '\n';
      }
      return c;
    }
  }
  int lex1(Sym* v) {  // lex.go:72-108
    for (;;) {
      const int c = next();
      BeginColNumber = ColNumber;
      if (c != kEof && isAlpha(c)) return token(c, v);
      switch (c) {
        case kEof: return TEof;
        case '-': case '0': case '1': case '2': case '3': case '4': case '5': case '6': case '7': case '8': case '9':
          return num(c, v);
        case '[': return TOpenBrace;
        case ']': return TCloseBrace;
        case '{': return TOpenCurly;
        case '}': return TCloseCurly;
        case ',': return TComma;
        case '"': return str(v);
        default: break;  // whitespace and anything else is skipped
      }
    }
  }
  int str(Sym* v) {  // lex.go:110-137: no escapes
    std::string b;
    for (;;) {
      const int c = next();
      if (c == kEof) return TEof;
      if (c == '"') break;
      b.push_back((char)c);
    }
    v->str = b;
    return TString;
  }
  int token(int c, Sym* v) {  // lex.go:139-166
    std::string b(1, (char)c);
    for (;;) {
      c = next();
      if (c == kEof) return TEof;
      if (c == '"' || c == '[' || c == ']' || c == '{' || c == '}' || c == ',') {
        peekc_ = c;
        break;
      }
      if (c == ' ' || c == '\t' || c == '\r' || c == '\n') break;
      b.push_back((char)c);
    }
    v->str = b;
    return TToken;
  }
  int num(int c, Sym* v) {  // lex.go:168-205
    std::string b(1, (char)c);
    bool isFloat = false;
    for (;;) {
      c = next();
      if ((c >= '0' && c <= '9') || c == '-') {
        b.push_back((char)c);
      } else if (c == '.' || c == 'e' || c == 'E') {
        isFloat = true;
        b.push_back((char)c);
      } else {
        break;
      }
    }
    if (c != kEof) peekc_ = c;
    char* endp = nullptr;
    if (isFloat) {
      const double f = std::strtod(b.c_str(), &endp);  // strconv.ParseFloat(s, 64): both round to nearest even
      if (endp == b.c_str() || *endp != 0) return TEof;
      v->numFloat = f;
      return TFloat;
    }
    const long long i = std::strtoll(b.c_str(), &endp, 10);
    if (endp == b.c_str() || *endp != 0) return TEof;
    v->numInt = i;
    return TInt;
  }
};

// ---- parsed values -------------------------------------------------------------------------------
enum Kind { KInt, KFloat, KBool, KString, KInt32Slice, KStringSlice, KMap, KVec3, KPointArray, KVec3Array, KVec2Array, KFloat32Array, KMatrixArray };
struct Value {
  int64_t i = 0;
  double f = 0;
  std::string s;
  std::vector<int32_t> ints;
  std::vector<std::string> strs;
  float c[3] = {0, 0, 0};          // constant map / vec3
  bool is_float_map = false;
  bool is_texture = false;         // rgbtex "file?query": v.s
  int MotionKeys = 0, ElemsPerKey = 0;
  std::vector<float> elems;        // arrays, flattened
  bool set = false;
};
struct FieldDef {
  const char* name;
  Kind kind;
  bool required;
};
struct NodeDefn {
  const char* type;
  std::vector<FieldDef> fields;
  bool in_scope;
};

const std::vector<NodeDefn>& node_table() {
  static const std::vector<NodeDefn> t = {
      // core/globals.go:8-16 (all optional)
      {"Globals", {{"XRes", KInt, false}, {"YRes", KInt, false}, {"UseProgress", KBool, false}, {"MaxGoRoutines", KInt, false},
                   {"Camera", KString, false}, {"MaxIter", KInt, false}, {"Output", KString, false}}, true},
      // builtin/camera/camera.go:48-73
      {"Camera", {{"Name", KString, true}, {"Type", KString, true}, {"From", KPointArray, true}, {"To", KPointArray, true},
                  {"Roll", KFloat32Array, true}, {"Up", KVec3, true}, {"Aspect", KFloat, false}, {"Fov", KFloat, false}, {"Focal", KFloat, false},
                  {"WorldToLocal", KMatrixArray, false}, {"LocalToWorld", KMatrixArray, false}, {"L", KFloat, false}, {"R", KFloat, false},
                  {"T", KFloat, false}, {"B", KFloat, false}, {"Radius", KFloat, false}}, true},
      // builtin/shader/std.go:25-47
      {"ShaderStd", {{"Name", KString, true}, {"EmissionColour", KMap, false}, {"EmissionStrength", KMap, false}, {"Sides", KInt, false},
                     {"DiffuseColour", KMap, false}, {"DiffuseStrength", KMap, false}, {"DiffuseRoughness", KMap, false},
                     {"Spec1Colour", KMap, false}, {"Spec1Strength", KMap, false}, {"Spec1Roughness", KMap, false},
                     {"Spec1FresnelModel", KString, false}, {"Spec1FresnelRefl", KMap, false}, {"Spec1FresnelEdge", KMap, false},
                     {"IOR", KMap, false}}, true},
      // builtin/shader/debug.go:15-22
      {"DebugShader", {{"Name", KString, true}, {"Sides", KInt, false}, {"Colour", KMap, true}}, true},
      // builtin/geom/polymesh/polymesh.go:17-40
      {"PolyMesh", {{"Name", KString, true}, {"RayBias", KFloat, false}, {"Verts", KPointArray, true}, {"PolyCount", KInt32Slice, false},
                    {"FaceIdx", KInt32Slice, false}, {"Shader", KStringSlice, true}, {"ShaderIdx", KInt32Slice, false},
                    {"CalcNormals", KBool, false}, {"IsVisible", KBool, false}, {"Transform", KMatrixArray, false},
                    {"UV", KVec2Array, false}, {"UVIdx", KInt32Slice, false}, {"Normals", KVec3Array, false}, {"NormalIdx", KInt32Slice, false}}, true},
      // builtin/light/triangle.go:18-28, disk.go:21-34, sphere.go:17-28
      {"TriLight", {{"Name", KString, true}, {"P0", KVec3, true}, {"P1", KVec3, true}, {"P2", KVec3, true}, {"Shader", KString, true},
                    {"Samples", KInt, true}}, true},
      {"DiskLight", {{"Name", KString, true}, {"P", KVec3, true}, {"Up", KVec3, true}, {"LookAt", KVec3, true}, {"Radius", KFloat, true},
                     {"Shader", KString, true}, {"Segments", KInt, false}, {"Samples", KInt, false}}, true},
      {"SphereLight", {{"Name", KString, true}, {"P", KVec3, true}, {"Radius", KFloat, true}, {"Shader", KString, true}, {"Samples", KInt, false}}, true},
      // builtin/geom/sphere/sphere.go:15-28
      {"Sphere", {{"Name", KString, true}, {"RayBias", KFloat, false}, {"P", KVec3, true}, {"Radius", KFloat, true}, {"Shader", KString, true}}, true},
      // builtin/geom/instance/instance.go:36-51
      {"GeomInstance", {{"Name", KString, true}, {"Geom", KString, true}, {"BMin", KPointArray, true}, {"BMax", KPointArray, true},
                        {"Transform", KMatrixArray, true}}, true},
      // builtin/filter/airy.go:13-22 (GaussianFilter's untagged NodeDef field makes it unparseable in the reference, gauss.go:14;
      // accepted here with its two fields)
      {"AiryFilter", {{"Name", KString, true}, {"Width", KFloat, false}, {"Res", KInt, false}, {"Peak", KFloat, false}}, true},
      {"GaussianFilter", {{"Name", KString, true}, {"Width", KFloat, true}, {"Res", KInt, true}}, true},
      // builtin/driver/outputfloat.go:14-17, outputhdr.go:15-18
      {"OutputFloat", {{"Filename", KString, true}}, true},
      {"OutputHDR", {{"Filename", KString, true}}, true},
      // registered in the reference, not on this path
      // builtin/misc/include.go:9-13
      {"Include", {{"Name", KString, false}, {"Filename", KString, true}}, true},
      {"QuadLight", {}, false}, {"Proc", {}, false},
  };
  return t;
}

const char* const kKeywords[] = {"int", "float", "vec2", "vec3", "point", "rgb", "rgbtex", "matrix"};  // parser.go:48
bool is_keyword(const std::string& s) {
  for (const char* k : kKeywords)
    if (s == k) return true;
  return false;
}

struct Parser {
  Lex lex;
  std::string filename;
  std::ostringstream log;
  int nerrors = 0;
  bool aborted = false;
  Parser(const char* text, size_t len, const std::string& fn) : lex(text, len), filename(fn) {}

  void errorf(const std::string& msg) {  // parser.go:862-874
    log << filename << ":" << lex.LineNumber << ":" << lex.BeginColNumber << ": " << msg << "\n";
    nerrors++;
    if (nerrors > 10) {
      log << "Too many errors, stopping.\n";
      aborted = true;
    }
  }

  bool number(float* out) {
    Sym s;
    const int t = lex.lex(&s);
    if (t == TInt) { *out = (float)s.numInt; return true; }
    if (t == TFloat) { *out = (float)s.numFloat; return true; }
    return false;
  }
  // "<keys> <n> <type> values..." (parser.go:293-343 and siblings); comps = floats per element
  const char* array(Value& v, int comps, bool has_count) {
    Sym s;
    if (lex.lex(&s) != TInt) return "Expected number of motion keys.";
    v.MotionKeys = (int)s.numInt;
    v.ElemsPerKey = 1;
    if (has_count) {
      if (lex.lex(&s) != TInt) return "Expected number of elements.";
      v.ElemsPerKey = (int)s.numInt;
    }
    lex.lex(&s);  // the element type keyword; the reference's check (`t != TokToken && str != ...`) never fires for a token
    const int k = v.MotionKeys == 0 ? 1 : v.MotionKeys;
    const long long n = (long long)k * v.ElemsPerKey * comps;
    if (n < 0 || n > (1ll << 31)) return "Array too large.";
    v.elems.resize((size_t)n);
    for (long long j = 0; j < n; j++)
      if (!number(&v.elems[(size_t)j])) return "Expected array component.";
    return nullptr;
  }

  // parser.go:566-690. Returns an error text for the "Error parsing field" message, or null.
  const char* param(Kind kind, Value& v) {
    Sym s;
    v.set = true;
    switch (kind) {
      case KInt: {
        const int t = lex.lex(&s);
        if (t == TFloat) v.i = (int64_t)s.numFloat;
        else if (t == TInt) v.i = s.numInt;
        return nullptr;
      }
      case KFloat: {
        const int t = lex.lex(&s);
        if (t == TFloat) v.f = s.numFloat;
        else if (t == TInt) v.f = (double)s.numInt;
        return nullptr;
      }
      case KBool: {
        if (lex.lex(&s) == TInt) v.i = s.numInt != 0;
        return nullptr;
      }
      case KString: {
        if (lex.lex(&s) == TString) v.s = s.str;
        return nullptr;
      }
      case KInt32Slice:
      case KStringSlice: {
        if (lex.peek(&s) != TInt) {
          errorf("Invalid token for param (expecting length of slice)");
          lex.skip();
          return nullptr;
        }
        lex.lex(&s);
        const long long count = s.numInt;
        lex.lex(&s);  // "int" / "string"
        for (long long i = 0; i < count; i++) {
          const int t = lex.lex(&s);
          if (kind == KInt32Slice) {
            if (t != TInt) return nullptr;  // int32slice's error is dropped by parseParam (parser.go:613)
            v.ints.push_back((int32_t)s.numInt);
          } else {
            if (t != TString) return nullptr;
            v.strs.push_back(s.str);
          }
        }
        return nullptr;
      }
      case KMap: {  // parser.go:637-649: rgb | float | rgbtex
        if (lex.peek(&s) != TToken) return nullptr;
        if (s.str == "rgb") {
          lex.lex(&s);
          for (int i = 0; i < 3; i++)
            if (!number(&v.c[i])) return nullptr;
        } else if (s.str == "float") {
          lex.lex(&s);
          float f;
          if (!number(&f)) return nullptr;
          v.c[0] = v.c[1] = v.c[2] = f;
          v.is_float_map = true;
        } else if (s.str == "rgbtex") {
          lex.lex(&s);
          if (lex.lex(&s) != TString) {  // parser.go:255-257; the error value is dropped by parseParam (:646)
            v.set = false;
            return nullptr;
          }
          v.s = s.str;
          v.is_texture = true;
        } else {
          v.set = false;  // the reference leaves the interface nil and the stray token is reported as an unknown field
        }
        return nullptr;
      }
      case KVec3:
        for (int i = 0; i < 3; i++)
          if (!number(&v.c[i])) return "Expected vector component.";
        return nullptr;
      case KPointArray:
      case KVec3Array: return array(v, 3, true);
      case KVec2Array: return array(v, 2, true);
      case KFloat32Array: return array(v, 1, true);
      case KMatrixArray: {
        const char* e = array(v, 16, false);
        if (e) return e;
        // stored transposed (parser.go:487): file order is row major, Matrix4 is column major
        for (size_t m = 0; m + 16 <= v.elems.size(); m += 16) {
          float t[16];
          for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) t[i * 4 + j] = v.elems[m + (size_t)j * 4 + i];
          std::memcpy(&v.elems[m], t, sizeof(t));
        }
        return nullptr;
      }
    }
    return nullptr;
  }

  void skipToToken() {  // parser.go:741-755
    Sym s;
    for (;;) {
      const int t = lex.peek(&s);
      if (t == TEof || t == TCloseCurly || (t == TToken && !is_keyword(s.str))) return;
      lex.skip();
    }
  }

  // parser.go:757-860. Returns false when the node is dropped.
  bool node(const std::string& type, const NodeDefn** defn_out, std::map<std::string, Value>* vals) {
    const NodeDefn* defn = nullptr;
    for (const NodeDefn& d : node_table())
      if (type == d.type) defn = &d;
    *defn_out = defn;
    if (!defn || !defn->in_scope) return false;  // createNode fails (register.go:14-23)
    Sym v;
    for (;;) {
      if (aborted) return false;
      const int t = lex.lex(&v);
      if (t == TToken) {
        const std::string pname = v.str;
        const FieldDef* fd = nullptr;
        for (const FieldDef& f : defn->fields)
          if (pname == f.name) fd = &f;
        if (fd && (*vals)[pname].set) {
          errorf("Field " + pname + " already found in " + type);
          return false;
        }
        if (!fd) {
          errorf("Field \"" + pname + "\" not found in node " + type);
          skipToToken();
          continue;
        }
        Value& val = (*vals)[pname];
        if (const char* e = param(fd->kind, val)) errorf(std::string("Error parsing field: ") + e);
        val.set = true;  // field.present = true whatever the outcome (parser.go:812-813)
      } else if (t == TCloseCurly) {
        for (const FieldDef& f : defn->fields)
          if (f.required && !(*vals)[f.name].set) {
            errorf(std::string("node: required field ") + f.name + " not found in " + type);
            return false;
          }
        return true;
      } else if (t == TEof) {
        errorf("node: unexpected end of file in object " + type);
        return false;
      } else {
        errorf("node: Parse error, invalid token in object");
      }
    }
  }
};

V3 v3of(const float* c) { return V3{c[0], c[1], c[2]}; }

void set_map(VgMaterial& m, uint32_t bit, const Value& v, float* dst3, float* dst1, TextureMap* texmaps = nullptr) {
  if (!v.set) return;
  m.mask |= bit;
  if (v.is_texture && texmaps) {
    // parser.rgbtex always builds the map with CreateRGBTextureMap (parser.go:259): channel 0 even on a float parameter
    int slot = 0;
    while ((1u << slot) != bit) slot++;
    texmaps[slot] = TextureMap::Parse(v.s, false);
    return;
  }
  if (dst3) { dst3[0] = v.c[0]; dst3[1] = v.c[1]; dst3[2] = v.c[2]; }
  if (dst1) *dst1 = v.c[0];  // param.Float32Uniform of a Constant map returns C[0] (builtin/maps/constant.go)
}

// Build the vh node from the parsed fields. Returns an error text or "".
std::string build_node(Core& core, const std::string& type, std::map<std::string, Value>& f) {
  auto has = [&](const char* k) { auto it = f.find(k); return it != f.end() && it->second.set; };
  std::unique_ptr<Node> h = CreateNode(type);
  if (!h) return "node type not registered: " + type;
  if (type == "Globals") {
    Globals* g = static_cast<Globals*>(h.get());
    g->XRes = 256; g->YRes = 256; g->MaxGoRoutines = 5; g->MaxIter = 0;  // parser.go:76-80
    if (has("XRes")) g->XRes = (int)f["XRes"].i;
    if (has("YRes")) g->YRes = (int)f["YRes"].i;
    if (has("MaxGoRoutines")) g->MaxGoRoutines = (int)f["MaxGoRoutines"].i;
    if (has("MaxIter")) g->MaxIter = (int)f["MaxIter"].i;
    if (has("Camera")) g->Camera = f["Camera"].s;
    if (g->XRes <= 0 || g->YRes <= 0) return "Globals: XRes/YRes must be positive";
  } else if (type == "Camera") {
    Camera* c = static_cast<Camera*>(h.get());
    c->NodeName = f["Name"].s;
    c->Type = f["Type"].s;
    // From / To / Roll keep their motion keys; calcLookatMatrices indexes Elems[key] (camera.go:109-193), i.e. one element per key
    const Value &from = f["From"], &to = f["To"], &roll = f["Roll"];
    const int nF = from.MotionKeys == 0 ? 1 : from.MotionKeys, nT = to.MotionKeys == 0 ? 1 : to.MotionKeys;
    const int nR = roll.MotionKeys == 0 ? 1 : roll.MotionKeys;
    if (c->Type == "LookAt" && (from.elems.size() < (size_t)3 * nF || to.elems.size() < (size_t)3 * nT)) return "Camera: From/To need one point per motion key";
    for (int i = 0; i < nF && (size_t)3 * (i + 1) <= from.elems.size(); i++) c->FromKeys.push_back(v3of(from.elems.data() + 3 * i));
    for (int i = 0; i < nT && (size_t)3 * (i + 1) <= to.elems.size(); i++) c->ToKeys.push_back(v3of(to.elems.data() + 3 * i));
    for (int i = 0; i < nR && (size_t)(i + 1) <= roll.elems.size(); i++) c->RollKeys.push_back(roll.elems[i]);
    if (!c->FromKeys.empty()) c->From = c->FromKeys[0];
    if (!c->ToKeys.empty()) c->To = c->ToKeys[0];
    c->Roll = c->RollKeys.empty() ? 0.0f : c->RollKeys[0];
    if (has("WorldToLocal")) {
      const Value& w = f["WorldToLocal"];
      for (size_t m = 0; m + 16 <= w.elems.size(); m += 16) {
        M4 mm;
        std::memcpy(mm.m, &w.elems[m], sizeof(mm.m));
        c->WorldToLocal.push_back(mm);
      }
    }
    if (has("LocalToWorld")) return "Camera: a LocalToWorld given in the file (PreRender appends to it) is outside this path";
    c->Up = v3of(f["Up"].c);
    if (has("Aspect")) c->Aspect = (float)f["Aspect"].f;
    if (has("Fov")) c->Fov = (float)f["Fov"].f;
    if (has("Focal")) c->Focal = (float)f["Focal"].f;
    if (has("Radius")) c->Radius = (float)f["Radius"].f;
  } else if (type == "ShaderStd") {
    ShaderStd* s = static_cast<ShaderStd*>(h.get());
    s->MtlName = f["Name"].s;
    VgMaterial& m = s->params;
    std::memset(&m, 0, sizeof(m));
    set_map(m, VG_MAT_EMISSION_COLOUR, f["EmissionColour"], m.emission_colour, nullptr, s->texmaps);
    set_map(m, VG_MAT_EMISSION_STRENGTH, f["EmissionStrength"], nullptr, &m.emission_strength, s->texmaps);
    set_map(m, VG_MAT_DIFFUSE_COLOUR, f["DiffuseColour"], m.diffuse_colour, nullptr, s->texmaps);
    set_map(m, VG_MAT_DIFFUSE_STRENGTH, f["DiffuseStrength"], nullptr, &m.diffuse_strength, s->texmaps);
    set_map(m, VG_MAT_DIFFUSE_ROUGHNESS, f["DiffuseRoughness"], nullptr, &m.diffuse_roughness, s->texmaps);
    set_map(m, VG_MAT_SPEC1_COLOUR, f["Spec1Colour"], m.spec1_colour, nullptr, s->texmaps);
    set_map(m, VG_MAT_SPEC1_STRENGTH, f["Spec1Strength"], nullptr, &m.spec1_strength, s->texmaps);
    set_map(m, VG_MAT_SPEC1_ROUGHNESS, f["Spec1Roughness"], nullptr, &m.spec1_roughness, s->texmaps);
    set_map(m, VG_MAT_IOR, f["IOR"], nullptr, &m.ior, s->texmaps);
    set_map(m, VG_MAT_SPEC1_FRESNEL_REFL, f["Spec1FresnelRefl"], m.spec1_fresnel_refl, nullptr, s->texmaps);
    set_map(m, VG_MAT_SPEC1_FRESNEL_EDGE, f["Spec1FresnelEdge"], m.spec1_fresnel_edge, nullptr, s->texmaps);
    if (has("Spec1FresnelModel")) {  // std.go:65-73: anything but "Metal" leaves the zero value (Dielectric)
      m.mask |= VG_MAT_SPEC1_FRESNEL_MODEL;
      m.spec1_fresnel_model = f["Spec1FresnelModel"].s == "Metal" ? VG_FRESNEL_CONDUCTOR : VG_FRESNEL_DIELECTRIC;
    }
  } else if (type == "Include") {
    Include* inc = static_cast<Include*>(h.get());
    if (has("Name")) inc->NodeName = f["Name"].s;
    inc->Filename = f["Filename"].s;
  } else if (type == "DebugShader") {
    DebugShader* s = static_cast<DebugShader*>(h.get());
    s->MtlName = f["Name"].s;
    VgMaterial& m = s->params;
    std::memset(&m, 0, sizeof(m));
    m.mask = VG_MAT_DEBUG;
    set_map(m, VG_MAT_DIFFUSE_COLOUR, f["Colour"], m.diffuse_colour, nullptr);
  } else if (type == "PolyMesh") {
    PolyMesh* m = static_cast<PolyMesh*>(h.get());
    m->NodeName = f["Name"].s;
    if (has("RayBias")) m->RayBias = (float)f["RayBias"].f;
    const Value& verts = f["Verts"];
    m->Verts.MotionKeys = verts.MotionKeys == 0 ? 1 : verts.MotionKeys;
    m->Verts.ElemsPerKey = verts.ElemsPerKey;
    if (verts.ElemsPerKey <= 0) return "PolyMesh: no vertices";
    m->Verts.Elems.resize(verts.elems.size() / 3);
    std::memcpy(m->Verts.Elems.data(), verts.elems.data(), m->Verts.Elems.size() * sizeof(V3));
    if (has("PolyCount")) { m->hasPolyCount = true; m->PolyCount = f["PolyCount"].ints; }
    if (has("FaceIdx")) { m->hasFaceIdx = true; m->FaceIdx = f["FaceIdx"].ints; }
    if (m->hasPolyCount) {
      if (!m->hasFaceIdx) return "PolyMesh: PolyCount without FaceIdx";
      int64_t tot = 0;
      for (int32_t c : m->PolyCount) { if (c < 3) return "PolyMesh: polygon with < 3 vertices"; tot += c; }
      if (tot != (int64_t)m->FaceIdx.size()) return "PolyMesh: sum(PolyCount) != len(FaceIdx)";
    }
    m->Shader = f["Shader"].strs;
    if (has("ShaderIdx")) m->ShaderIdx = f["ShaderIdx"].ints;
    if (has("Normals")) {
      const Value& nv = f["Normals"];
      if (nv.MotionKeys > 1) return "PolyMesh: Normals with motion keys are outside this path";
      m->Normals.MotionKeys = 1;
      m->Normals.ElemsPerKey = nv.ElemsPerKey;
      m->Normals.Elems.resize(nv.elems.size() / 3);
      std::memcpy(m->Normals.Elems.data(), nv.elems.data(), m->Normals.Elems.size() * sizeof(V3));
      if (has("NormalIdx")) { m->hasNormalIdx = true; m->NormalIdx = f["NormalIdx"].ints; }
    }
    if (has("Transform")) {
      // object transforms (polymesh/trace.go:22-55) are SURVEY.md 8(f).3; an identity single-key transform is a no-op
      const Value& t = f["Transform"];
      static const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
      if (t.MotionKeys > 1 || t.elems.size() != 16 || std::memcmp(t.elems.data(), ident, sizeof(ident)) != 0)
        return "PolyMesh: Transform other than a single identity matrix is outside this path";
    }
    if (has("UV")) {  // polymesh.go:36-37
      const Value& uv = f["UV"];
      if (uv.MotionKeys > 1) return "PolyMesh: UV with motion keys is outside this path";
      m->UV = uv.elems;
      if (has("UVIdx")) { m->hasUVIdx = true; m->UVIdx = f["UVIdx"].ints; }
    }
    // CalcNormals and IsVisible are never read by the reference
  } else if (type == "TriLight") {
    TriLight* t = static_cast<TriLight*>(h.get());
    t->NodeName = f["Name"].s;
    t->P0 = v3of(f["P0"].c); t->P1 = v3of(f["P1"].c); t->P2 = v3of(f["P2"].c);
    t->Shader = f["Shader"].s;
    t->Samples = (int)f["Samples"].i;
    if (t->Samples < 0 || t->Samples > 8) return "TriLight: Samples must be in [0,8]";
  } else if (type == "DiskLight") {
    DiskLight* d = static_cast<DiskLight*>(h.get());
    d->NodeName = f["Name"].s;
    d->P = v3of(f["P"].c); d->Up = v3of(f["Up"].c); d->LookAt = v3of(f["LookAt"].c);
    d->Radius = (float)f["Radius"].f;
    d->Shader = f["Shader"].s;
    if (has("Segments")) d->Segments = (int)f["Segments"].i;
    if (has("Samples")) d->Samples = (int)f["Samples"].i;
    if (d->Samples < 0 || d->Samples > 8) return "DiskLight: Samples must be in [0,8]";
    if (d->Segments < 3 || d->Segments > 4096) return "DiskLight: Segments must be in [3,4096]";
  } else if (type == "SphereLight") {
    SphereLight* d = static_cast<SphereLight*>(h.get());
    d->NodeName = f["Name"].s;
    d->P = v3of(f["P"].c);
    d->Radius = (float)f["Radius"].f;
    d->Shader = f["Shader"].s;
    if (has("Samples")) d->Samples = (int)f["Samples"].i;
    if (d->Samples < 0 || d->Samples > 8) return "SphereLight: Samples must be in [0,8]";
  } else if (type == "Sphere") {
    SphereGeom* g = static_cast<SphereGeom*>(h.get());
    g->NodeName = f["Name"].s;
    g->P = v3of(f["P"].c);
    g->Radius = (float)f["Radius"].f;
    g->Shader = f["Shader"].s;
  } else if (type == "GeomInstance") {
    GeomInstance* g = static_cast<GeomInstance*>(h.get());
    g->NodeName = f["Name"].s;
    g->GeomName = f["Geom"].s;
    const Value &bmin = f["BMin"], &bmax = f["BMax"], &tr = f["Transform"];
    if (bmin.elems.size() < 3 || bmax.elems.size() != bmin.elems.size()) return "GeomInstance: BMin/BMax need the same number (>= 1) of points";
    for (size_t i = 0; i + 3 <= bmin.elems.size(); i += 3) { g->BMin.push_back(v3of(&bmin.elems[i])); g->BMax.push_back(v3of(&bmax.elems[i])); }
    if (tr.elems.size() < 16 || tr.elems.size() / 16 > 255) return "GeomInstance: Transform needs 1..255 matrices";
    for (size_t i = 0; i + 16 <= tr.elems.size(); i += 16) {
      M4 m;
      std::memcpy(m.m, &tr.elems[i], sizeof(m.m));
      g->Transform.push_back(m);
    }
  } else if (type == "AiryFilter" || type == "GaussianFilter") {
    PixelFilter* p = static_cast<PixelFilter*>(h.get());
    p->NodeName = f["Name"].s;
    if (has("Width")) p->Width = (float)f["Width"].f;
    if (has("Res")) p->Res = (int)f["Res"].i;
    if (has("Peak")) p->Peak = (float)f["Peak"].f;
  } else if (type == "OutputFloat" || type == "OutputHDR") {
    static_cast<OutputNode*>(h.get())->Filename = f["Filename"].s;
  }
  core.AddNode(std::move(h));
  return "";
}

}  // namespace

// nodes.Parse (nodes/parser.go:110-131,876-907)
int ParseVnf(Core& core, const char* text, size_t len, const std::string& filename, std::string* messages) {
  Parser p(text, len, filename);
  Sym v;
  for (;;) {
    if (p.aborted) break;
    const int t = p.lex.lex(&v);
    if (t != TToken) break;
    const std::string type = v.str;
    if (p.lex.lex(&v) != TOpenCurly) p.errorf("Invalid token in node preamble");
    const NodeDefn* defn = nullptr;
    std::map<std::string, Value> vals;
    bool ok = p.node(type, &defn, &vals);
    std::string why;
    if (ok) {
      why = build_node(core, type, vals);
      if (!why.empty()) {
        p.errorf(why);
        continue;  // the node's own '}' has been consumed; nothing to skip
      }
      continue;
    }
    // createNode's error text (nodes/register.go:23) for an unknown type; a node that p.node() dropped comes back as (nil, nil)
    // and Go's %v prints that error as "<nil>" (parser.go:889)
    if (!defn) why = "Node type " + type + " not registered.";
    else if (!defn->in_scope) why = "node type \"" + type + "\" is outside this path (SURVEY.md 8, out of scope)";
    p.errorf("Node is nil: " + (why.empty() ? std::string("<nil>") : why));
    // parser.go:893-899: skip to the next '}' — for a node dropped at its own '}' this swallows the FOLLOWING node, as in
    // the reference
    for (;;) {
      const int t2 = p.lex.lex(&v);
      if (t2 == TCloseCurly || t2 == TEof) break;
    }
  }
  if (messages) *messages = p.log.str();
  return p.nerrors;
}

// misc.Include.PreRender (include.go:24-26) = nodes.Parse(Filename): os.Open resolves the name against the working directory
int Include::PreRender(Core& core, std::string* err) {
  FILE* fp = std::fopen(Filename.c_str(), "rb");
  if (!fp) { *err = "open " + Filename + ": no such file or directory"; return -1; }
  std::string text;
  char buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof(buf), fp)) > 0) text.append(buf, n);
  std::fclose(fp);
  std::string msgs;
  core.include_errors += ParseVnf(core, text.data(), text.size(), Filename, &msgs);
  core.include_log += msgs;
  return 0;
}

// ---- output drivers ----------------------------------------------------------------------------------
// image/hdr/hdr.go:26-50. Go's float32 -> byte conversion on amd64 is CVTTSS2SL + low byte: negative values wrap and NaN /
// out-of-range become 0x80000000 -> 0.
static inline uint8_t go_byte(float v) {
  int32_t i;
  if (!(v > -2147483904.0f && v < 2147483648.0f)) i = INT32_MIN;  // also NaN
  else i = (int32_t)v;
  return (uint8_t)(uint32_t)i;
}
void RgbToRgbe(float r, float g, float b, uint8_t out[4]) {
  float d = r;
  if (g > d) d = g;
  if (b > d) d = b;
  if (d < 0.000001f) {
    out[0] = out[1] = out[2] = out[3] = 0;
    return;
  }
  int e = 0;
  const double nd = std::frexp((double)d, &e);
  const float n = (float)nd;
  const float df = n * 255.999f / d;
  out[0] = go_byte(r * df);
  out[1] = go_byte(g * df);
  out[2] = go_byte(b * df);
  out[3] = (uint8_t)(e + 128);
}

int OutputNode::Write(const float* fb, int w, int h, std::string* err) const {
  std::ofstream out(Filename, std::ios::binary | std::ios::trunc);
  if (!out) { *err = "cannot create " + Filename; return -1; }
  if (!hdr) {
    // driver/outputfloat.go:30-42: binary.Write(LittleEndian, []float32) of the whole framebuffer, rows top to bottom
    out.write(reinterpret_cast<const char*>(fb), (std::streamsize)((size_t)w * h * 3 * sizeof(float)));
  } else {
    // image/hdr/writer.go:58-91: flat (not run-length encoded) RGBE scanlines, written bottom row first under a "+Y" header
    std::ostringstream hd;
    hd << "#?RADIANCE\n# Created by Vermeer Light Tools (http://www.vermeerlt.com)\nFORMAT=32-bit_rle_rgbe\n\n+Y " << h << " +X " << w << "\n";
    const std::string hs = hd.str();
    out.write(hs.data(), (std::streamsize)hs.size());
    std::vector<uint8_t> scan((size_t)w * 4);
    for (int j = 0; j < h; j++) {
      const int k = h - j - 1;
      for (int i = 0; i < w; i++) {
        const float* px = fb + ((size_t)i + (size_t)k * w) * 3;
        RgbToRgbe(px[0], px[1], px[2], &scan[(size_t)i * 4]);
      }
      out.write(reinterpret_cast<const char*>(scan.data()), (std::streamsize)scan.size());
    }
  }
  out.flush();
  if (!out) { *err = "write failed: " + Filename; return -1; }
  return 0;
}

}  // namespace vh
