"""Scene descriptions and seeded synthetic scene generators (numpy only, no GPU, no oracle).

A `SceneDesc` is the neutral, in-memory equivalent of a reference `.vnf` file: a list of nodes
(`Globals`, `Camera`, `ShaderStd`, `PolyMesh`, `TriLight`) with the reference's field names
(builtin/geom/polymesh/polymesh.go:17-40, builtin/shader/std.go:25-47,
builtin/light/triangle.go:18-28, builtin/camera/camera.go:48-73, core/globals.go:8-16).  The same
description is fed to the GPU host library (`vermeer_b200.host`) and, in tests and the CPU baseline,
to the oracle.

Generators follow SURVEY.md §8(d):
  C1 `cornell_box`      5 diffuse quads + 2 boxes, light = 2 TriLights (QuadLight panics in the reference)
  C2 `heightfield_scene` one PolyMesh, 2*nq^2 triangles (nq=708 -> 1 002 528), TriLight pair above
  C3 `sphere_field_scene` n meshes x 2*slices*(stacks-1) triangles, mirror+diffuse ShaderStd
  C4 `heightfield_scene(motion=True)` same mesh with Verts.MotionKeys=2
All geometry is wound so the geometric normal (V1-V0)x(V2-V0) faces outward/up: the reference never
face-forwards.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

__all__ = [
    "ShaderStd", "DebugShader", "Texture", "parse_texture_url", "SHADER_SLOTS", "textured_room", "PolyMesh", "GeomInstance", "matrix4", "srt_matrix", "TriLight", "DiskLight", "SphereLight", "Camera", "PixelFilter", "SceneDesc", "splitmix64_table", "camera_motion_variants", "debug_shader_box",
    "heightfield_mesh", "heightfield_scene", "sphere_field_scene", "cornell_box", "glossy_box", "instanced_scene", "incoherent_rays", "to_vnf",
]


@dataclass
class Texture:
    """One entry of the reference's texture store (texture/texture.go:36-47), keyed by the file name the shader parameters
    use.  Pixels: (h, w, 3) uint8 in IMAGE order (row 0 = top), i.e. what image.Decode hands to loadTexture; `rows_bottom_up()`
    is the texture's own storage order (texture.go:139 flips while copying).  Image decoding is outside the path."""
    Name: str
    Pixels: np.ndarray

    def __post_init__(self):
        self.Pixels = np.ascontiguousarray(self.Pixels, np.uint8)
        assert self.Pixels.ndim == 3 and self.Pixels.shape[2] == 3

    def rows_bottom_up(self) -> np.ndarray:
        return np.ascontiguousarray(self.Pixels[::-1])


def parse_texture_url(value: str):
    """builtin/maps/texture.go:48-83: `path?filter=trilinear&ch=N` -> (path, chan, trilinear)."""
    from urllib.parse import urlsplit, parse_qs
    u = urlsplit(value)
    q = parse_qs(u.query)
    ch = q.get("ch", ["0"])[0]
    try:
        chan = int(ch)
    except ValueError:
        chan = 0                      # strconv.Atoi error is dropped (texture.go:57)
    return u.path, chan, q.get("filter", [""])[0] == "trilinear"


# ShaderStd parameter slots in node order (std.go:28-46): (name, offset in the packed float block, floats)
SHADER_SLOTS = [("EmissionColour", 0, 3), ("EmissionStrength", 3, 1), ("DiffuseColour", 4, 3), ("DiffuseStrength", 7, 1),
                ("DiffuseRoughness", 8, 1), ("Spec1Colour", 9, 3), ("Spec1Strength", 12, 1), ("Spec1Roughness", 13, 1), ("IOR", 14, 1),
                ("Spec1FresnelModel", 15, 1), ("Spec1FresnelRefl", 16, 3), ("Spec1FresnelEdge", 19, 3)]


@dataclass
class ShaderStd:
    """A parameter given as a `str` is a texture map (the .vnf `rgbtex "file?filter=trilinear"` form, nodes/parser.go:247-268;
    the `ch` query selects the channel of a float parameter, builtin/maps/texture.go:48-67)."""
    Name: str
    EmissionColour: Optional[tuple] = None
    EmissionStrength: Optional[float] = None
    DiffuseColour: Optional[tuple] = None
    DiffuseStrength: Optional[float] = None
    DiffuseRoughness: Optional[float] = None
    Spec1Colour: Optional[tuple] = None
    Spec1Strength: Optional[float] = None
    Spec1Roughness: Optional[float] = None
    IOR: Optional[float] = None
    Spec1FresnelModel: Optional[str] = None   # "Dielectric" | "Metal" (std.go:65-73); unset = Dielectric (the zero value)
    Spec1FresnelRefl: Optional[tuple] = None
    Spec1FresnelEdge: Optional[tuple] = None

    def packed(self):
        """(mask, 22 floats) in the slot order shared by the host C-ABI and the oracle."""
        p = np.zeros(22, np.float32)
        mask = 0
        for bit, (name, off, n) in enumerate(SHADER_SLOTS):
            v = getattr(self, name)
            if isinstance(v, str) and name != "Spec1FresnelModel":
                continue   # texture map: bound separately (texture_bindings)
            if v is not None:
                mask |= 1 << bit
                if name == "Spec1FresnelModel":
                    v = {"Dielectric": 0.0, "Metal": 1.0}[v]
                p[off:off + n] = np.asarray(v, np.float32).reshape(-1)
        return mask, p

    def texture_bindings(self):
        """[(slot, path, chan, trilinear)] for the parameters that are texture maps."""
        out = []
        for slot, (name, _, _) in enumerate(SHADER_SLOTS):
            v = getattr(self, name)
            if isinstance(v, str) and name != "Spec1FresnelModel":
                out.append((slot,) + parse_texture_url(v))
        return out


@dataclass
class DebugShader:
    """builtin/shader/debug.go:15-22: OutRGB = Colour."""
    Name: str
    Colour: tuple = (0.0, 0.0, 0.0)

    def packed(self):
        p = np.zeros(22, np.float32)
        p[4:7] = np.asarray(self.Colour, np.float32)
        return 4096 | 4, p


@dataclass
class PolyMesh:
    Name: str
    Verts: np.ndarray                      # (keys, nverts, 3) float32
    Shader: List[str]
    PolyCount: Optional[np.ndarray] = None  # int32
    FaceIdx: Optional[np.ndarray] = None    # int32
    ShaderIdx: Optional[np.ndarray] = None  # int32, per polygon
    Normals: Optional[np.ndarray] = None    # (n,3) float32
    NormalIdx: Optional[np.ndarray] = None
    RayBias: float = 0.0
    UV: Optional[np.ndarray] = None         # (n,2) float32 (param.Vec2Array, one key)
    UVIdx: Optional[np.ndarray] = None

    def __post_init__(self):
        v = np.ascontiguousarray(self.Verts, np.float32)
        if v.ndim == 2:
            v = v[None]
        self.Verts = v
        if self.UV is not None:
            self.UV = np.ascontiguousarray(self.UV, np.float32).reshape(-1, 2)
        for k in ("PolyCount", "FaceIdx", "ShaderIdx", "NormalIdx", "UVIdx"):
            a = getattr(self, k)
            if a is not None:
                setattr(self, k, np.ascontiguousarray(a, np.int32))
        if self.Normals is not None:
            self.Normals = np.ascontiguousarray(self.Normals, np.float32)

    @property
    def num_tris(self) -> int:
        if self.PolyCount is not None:
            return int((self.PolyCount - 2).sum())
        if self.FaceIdx is not None:
            return len(self.FaceIdx) // 3
        return self.Verts.shape[1] // 3


@dataclass
class GeomInstance:
    """builtin/geom/instance/instance.go:36-51. Transform: (keys, 16) float32 in math.Matrix4 layout (column major: element
    (row i, col j) at [j*4 + i]); a .vnf file lists each matrix row by row and the parser transposes (nodes/parser.go:487).
    BMin/BMax: world-space bounds given by the scene author (the reference never derives them; Bounds() also always contains
    the origin, instance.go:124)."""
    Name: str
    Geom: str
    Transform: np.ndarray
    BMin: tuple
    BMax: tuple

    def __post_init__(self):
        self.Transform = np.ascontiguousarray(self.Transform, np.float32).reshape(-1, 16)


def matrix4(rows) -> np.ndarray:
    """A 4x4 given row by row (as in a .vnf file) -> math.Matrix4 storage (column major), float32."""
    return np.ascontiguousarray(np.asarray(rows, np.float32).reshape(4, 4).T).reshape(16)


def srt_matrix(translate=(0, 0, 0), rotate_y_deg=0.0, scale=1.0) -> np.ndarray:
    """translate * rotateY * uniform scale as math.Matrix4 storage; every factor rounded to float32 first."""
    a = np.float32(np.deg2rad(rotate_y_deg))
    c, s_ = np.float32(np.cos(a)), np.float32(np.sin(a))
    k = np.float32(scale)
    rows = [[c * k, 0, s_ * k, translate[0]], [0, k, 0, translate[1]], [-s_ * k, 0, c * k, translate[2]], [0, 0, 0, 1]]
    return matrix4(rows)


@dataclass
class TriLight:
    Name: str
    P0: tuple
    P1: tuple
    P2: tuple
    Shader: str
    Samples: int = 1


@dataclass
class DiskLight:
    """builtin/light/disk.go:21-34 (registered defaults Segments 20, Samples 1)."""
    Name: str
    P: tuple
    LookAt: tuple
    Up: tuple
    Radius: float
    Shader: str
    Segments: int = 20
    Samples: int = 1


@dataclass
class SphereLight:
    """builtin/light/sphere.go:17-28 (registered defaults Radius 1, Samples 1)."""
    Name: str
    P: tuple
    Shader: str
    Radius: float = 1.0
    Samples: int = 1


@dataclass
class Camera:
    From: tuple
    To: tuple
    Up: tuple = (0.0, 1.0, 0.0)
    Roll: float = 0.0
    Fov: float = 90.0
    Focal: float = 12.0
    Aspect: float = 0.0
    Radius: float = 0.0
    Name: str = "camera"
    Type: str = "LookAt"
    # motion keys (camera.go:48-73): lists of points / floats / 4x4 column-major matrices; None = the single From / To / Roll above
    FromKeys: Optional[list] = None
    ToKeys: Optional[list] = None
    RollKeys: Optional[list] = None
    WorldToLocal: Optional[list] = None   # Type "Matrix"

    def keys(self):
        """(from points, to points, rolls, matrices) as float32 arrays, one row per motion key."""
        fr = np.asarray(self.FromKeys if self.FromKeys else [self.From], np.float32).reshape(-1, 3)
        to = np.asarray(self.ToKeys if self.ToKeys else [self.To], np.float32).reshape(-1, 3)
        ro = np.asarray(self.RollKeys if self.RollKeys else [self.Roll], np.float32).reshape(-1)
        w2l = np.asarray(self.WorldToLocal if self.WorldToLocal else [], np.float32).reshape(-1, 16)
        return fr, to, ro, w2l

    @property
    def has_keys(self):
        return bool(self.FromKeys or self.ToKeys or self.RollKeys or self.WorldToLocal) or self.Type != "LookAt"


@dataclass
class PixelFilter:
    """AiryFilter / GaussianFilter node (builtin/filter/airy.go:13-22, gauss.go:13-21). None = the registered default."""
    Type: str = "AiryFilter"
    Name: str = "filter"
    Width: Optional[float] = None
    Res: Optional[int] = None
    Peak: Optional[float] = None


@dataclass
class SceneDesc:
    XRes: int
    YRes: int
    camera: Camera
    shaders: List[ShaderStd] = field(default_factory=list)
    meshes: List[PolyMesh] = field(default_factory=list)
    lights: list = field(default_factory=list)   # TriLight | DiskLight | SphereLight, in node order
    instances: list = field(default_factory=list)   # GeomInstance nodes (created after the meshes)
    MaxIter: int = 16
    name: str = "scene"
    filter: Optional[PixelFilter] = None
    textures: list = field(default_factory=list)   # Texture entries: the decoded files the shaders' texture maps name

    @property
    def num_tris(self) -> int:
        per_light = {"TriLight": lambda l: 1, "DiskLight": lambda l: l.Segments, "SphereLight": lambda l: 0}
        return sum(m.num_tris for m in self.meshes) + sum(per_light[type(l).__name__](l) for l in self.lights)


# ------------------------------------------------------------------------------------------------
def _num(v) -> str:
    """Shortest text that nodes.Lex + float32(strconv.ParseFloat(..., 64)) reads back to exactly this float32
    (nodes/lex.go:168-205 accepts [0-9-.eE] only: no '+' in exponents)."""
    f = np.float32(v)
    if float(f).is_integer() and abs(float(f)) < 1e15:
        return str(int(f))
    return ("%.9g" % float(f)).replace("e+", "e")


def _vec(v) -> str:
    return " ".join(_num(x) for x in v)


def to_vnf(sc: "SceneDesc", outputs=()) -> str:
    """The scene as reference-valid .vnf text (syntax: nodes/parser.go). `outputs` = [("OutputFloat" |
    "OutputHDR", filename), ...]. Nodes are written in the order the in-memory path adds them (Globals, shaders, meshes,
    lights, filter, camera), so both paths build identical scenes."""
    o = []
    o.append("Globals { XRes %d YRes %d MaxIter %d }" % (sc.XRes, sc.YRes, sc.MaxIter))
    for s in sc.shaders:
        if isinstance(s, DebugShader):
            o.append('DebugShader { Name "%s" Colour rgb %s }' % (s.Name, _vec(s.Colour)))
            continue
        parts = ['Name "%s"' % s.Name]
        for name in ("EmissionColour", "DiffuseColour", "Spec1Colour", "Spec1FresnelRefl", "Spec1FresnelEdge"):
            v = getattr(s, name)
            if isinstance(v, str):
                parts.append('%s rgbtex "%s"' % (name, v))      # nodes/parser.go:247-268
            elif v is not None:
                parts.append("%s rgb %s" % (name, _vec(v)))
        for name in ("EmissionStrength", "DiffuseStrength", "DiffuseRoughness", "Spec1Strength", "Spec1Roughness", "IOR"):
            v = getattr(s, name)
            if isinstance(v, str):
                # the only texture form the parser has; it builds an RGB map, so a float parameter reads channel 0
                parts.append('%s rgbtex "%s"' % (name, v))
            elif v is not None:
                parts.append("%s float %s" % (name, _num(v)))
        if s.Spec1FresnelModel is not None:
            parts.append('Spec1FresnelModel "%s"' % s.Spec1FresnelModel)
        o.append("ShaderStd { %s }" % " ".join(parts))
    for m in sc.meshes:
        keys, nv, _ = m.Verts.shape
        parts = ['Name "%s"' % m.Name]
        if m.RayBias:
            parts.append("RayBias %s" % _num(m.RayBias))
        parts.append("Verts %d %d point %s" % (keys, nv, _vec(m.Verts.reshape(-1))))
        if m.PolyCount is not None:
            parts.append("PolyCount %d int %s" % (len(m.PolyCount), " ".join(str(int(x)) for x in m.PolyCount)))
        if m.FaceIdx is not None:
            parts.append("FaceIdx %d int %s" % (len(m.FaceIdx), " ".join(str(int(x)) for x in m.FaceIdx)))
        parts.append("Shader %d string %s" % (len(m.Shader), " ".join('"%s"' % x for x in m.Shader)))
        if m.ShaderIdx is not None:
            parts.append("ShaderIdx %d int %s" % (len(m.ShaderIdx), " ".join(str(int(x)) for x in m.ShaderIdx)))
        if m.Normals is not None:
            parts.append("Normals 1 %d vec3 %s" % (len(m.Normals), _vec(m.Normals.reshape(-1))))
            if m.NormalIdx is not None:
                parts.append("NormalIdx %d int %s" % (len(m.NormalIdx), " ".join(str(int(x)) for x in m.NormalIdx)))
        if m.UV is not None:
            parts.append("UV 1 %d vec2 %s" % (len(m.UV), _vec(m.UV.reshape(-1))))
            if m.UVIdx is not None:
                parts.append("UVIdx %d int %s" % (len(m.UVIdx), " ".join(str(int(x)) for x in m.UVIdx)))
        o.append("PolyMesh { %s }" % "\n  ".join(parts))
    for ins in sc.instances:
        rows = ins.Transform.reshape(-1, 4, 4).transpose(0, 2, 1).reshape(-1)     # file order is row major (parser.go:487)
        bmin, bmax = np.asarray(ins.BMin, np.float32).reshape(-1), np.asarray(ins.BMax, np.float32).reshape(-1)
        o.append('GeomInstance { Name "%s" Geom "%s" BMin 1 %d point %s BMax 1 %d point %s Transform %d matrix %s }' % (
            ins.Name, ins.Geom, len(bmin) // 3, _vec(bmin), len(bmax) // 3, _vec(bmax), len(ins.Transform), _vec(rows)))
    for l in sc.lights:
        kind = type(l).__name__
        if kind == "TriLight":
            o.append('TriLight { Name "%s" Shader "%s" P0 %s P1 %s P2 %s Samples %d }' % (l.Name, l.Shader, _vec(l.P0), _vec(l.P1), _vec(l.P2), l.Samples))
        elif kind == "DiskLight":
            o.append('DiskLight { Name "%s" Shader "%s" P %s LookAt %s Up %s Radius %s Segments %d Samples %d }' % (
                l.Name, l.Shader, _vec(l.P), _vec(l.LookAt), _vec(l.Up), _num(l.Radius), l.Segments, l.Samples))
        elif kind == "SphereLight":
            o.append('SphereLight { Name "%s" Shader "%s" P %s Radius %s Samples %d }' % (l.Name, l.Shader, _vec(l.P), _num(l.Radius), l.Samples))
    if sc.filter is not None:
        f = sc.filter
        parts = ['Name "%s"' % f.Name]
        if f.Width is not None:
            parts.append("Width %s" % _num(f.Width))
        if f.Res is not None:
            parts.append("Res %d" % f.Res)
        if f.Peak is not None and f.Type == "AiryFilter":
            parts.append("Peak %s" % _num(f.Peak))
        o.append("%s { %s }" % (f.Type, " ".join(parts)))
    c = sc.camera
    fr, to, ro, w2l = c.keys()
    cam = ['Name "%s"' % c.Name, 'Type "%s"' % c.Type, "From %d 1 point %s" % (len(fr), " ".join(_vec(v) for v in fr)),
           "To %d 1 point %s" % (len(to), " ".join(_vec(v) for v in to)), "Roll %d 1 float %s" % (len(ro), " ".join(_num(v) for v in ro)),
           "Up %s" % _vec(c.Up), "Fov %s" % _num(c.Fov), "Focal %s" % _num(c.Focal)]
    if len(w2l):
        # the file holds matrices row major (parser.go:487 transposes on read)
        cam.append("WorldToLocal %d matrix %s" % (len(w2l), _vec(w2l.reshape(-1, 4, 4).transpose(0, 2, 1).reshape(-1))))
    if c.Aspect:
        cam.append("Aspect %s" % _num(c.Aspect))
    if c.Radius:
        cam.append("Radius %s" % _num(c.Radius))
    o.append("Camera { %s }" % " ".join(cam))
    for kind, fn in outputs:
        o.append('%s { Filename "%s" }' % (kind, fn))
    return "\n".join(o) + "\n"


def splitmix64_table(seed: int, npix: int) -> np.ndarray:
    """Per-pixel scramble table, (npix, 6) uint64 = {lensU, lensV, time, lambda, scramble[0], scramble[1]}
    (core/render.go:18-23).  The reference draws these from unseeded math/rand (render.go:169-174); any
    u64s are valid, so both sides share this seeded splitmix64 stream."""
    n = npix * 6
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z.reshape(npix, 6)


def _hash01(ix, iz, seed):
    """Cheap integer hash -> [0,1) float32, vectorised."""
    with np.errstate(over="ignore"):
        h = (ix.astype(np.uint32) * np.uint32(73856093)) ^ (iz.astype(np.uint32) * np.uint32(19349663)) ^ np.uint32(seed * 83492791 & 0xFFFFFFFF)
        h ^= h >> np.uint32(13)
        h *= np.uint32(0x5bd1e995)
        h ^= h >> np.uint32(15)
    return (h & np.uint32(0xFFFFFF)).astype(np.float32) / np.float32(1 << 24)


def heightfield_mesh(nq: int = 708, amp: float = 0.08, seed: int = 2, motion: bool = False, name: str = "heightfield",
                     shader: str = "ground", extent: float = 1.0) -> PolyMesh:
    """(nq+1)^2-vertex grid over [-extent,extent]^2 in xz, y = amp*sin(fx)*sin(fz) + hash noise; 2*nq^2 triangles
    given as an indexed triangle list (FaceIdx only), wound so Ng = +Y."""
    n = nq + 1
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    x = (-extent + 2.0 * extent * ii / nq).astype(np.float32)
    z = (-extent + 2.0 * extent * jj / nq).astype(np.float32)
    y = (amp * np.sin(7.0 * x) * np.sin(5.0 * z)).astype(np.float32)
    y += np.float32(0.25 * 2.0 * extent / max(nq, 1)) * (_hash01(ii, jj, seed) - np.float32(0.5))
    v0 = np.stack([x, y.astype(np.float32), z], -1).reshape(-1, 3).astype(np.float32)
    keys = [v0]
    if motion:
        d = np.zeros_like(v0)
        d[:, 1] = (0.5 * amp * np.sin(3.0 * v0[:, 0] + 1.0) * np.cos(4.0 * v0[:, 2])).astype(np.float32)
        d[:, 0] = np.float32(0.01) * np.sin(2.0 * v0[:, 2]).astype(np.float32)
        keys.append((v0 + d).astype(np.float32))
    qi, qj = np.meshgrid(np.arange(nq), np.arange(nq), indexing="xy")
    a = (qj * n + qi).reshape(-1)
    b = a + 1
    c = a + n + 1
    dd = a + n
    tris = np.stack([a, c, b, a, dd, c], -1).reshape(-1).astype(np.int32)
    return PolyMesh(Name=name, Verts=np.stack(keys, 0), Shader=[shader], FaceIdx=tris)


def _light_pair(y: float, half: float, shader: str, samples: int = 1, cx: float = 0.0, cz: float = 0.0, dy: float = 0.1):
    """Two TriLights forming a 45-degree-rotated square (diamond) whose normals (P1-P0)x(P2-P0) face -Y (into the scene),
    the second one lowered by `dy`.

    * A diamond, not an axis-aligned quad: two triangles sharing an axis-aligned rectangle's diagonal have identical
      bounding-box centroids, and the reference's scene-level builder (leafMax=1) then recurses forever
      (qbvh/build.go:35-43 flat-axis case; DESIGN.md "reference quirks", f).
    * Not coplanar: a ray that hits light A (directly or through a mirror) evaluates light B from the shading point
      (builtin/scene/scene.go:100-116 only excludes the hit light itself). If B lies in A's plane, B's spherical-triangle
      sampling runs from inside its own plane and is decided by rounding noise — NaN or finite, black or bright
      (DESIGN.md "reference quirks", g). The offset keeps the reference's own result well defined."""
    pw = (cx - half, y, cz)
    pe = (cx + half, y, cz)
    pn = (cx, y, cz + half)
    lw = (cx - half, y - dy, cz)
    le = (cx + half, y - dy, cz)
    ls = (cx, y - dy, cz - half)
    return [TriLight("lightA", pw, pe, pn, shader, samples), TriLight("lightB", le, lw, ls, shader, samples)]


def heightfield_scene(xres: int = 1920, yres: int = 1080, nq: int = 708, motion: bool = False, seed: int = 2) -> SceneDesc:
    """Configs C2 / C4: one 2*nq^2-triangle PolyMesh, one TriLight pair above, camera at ~30 deg elevation."""
    shaders = [
        ShaderStd("ground", DiffuseColour=(0.7, 0.6, 0.5), DiffuseStrength=1.0),
        ShaderStd("lightmtl", EmissionColour=(1.0, 0.9, 0.8), EmissionStrength=15.0, DiffuseColour=(0.0, 0.0, 0.0), DiffuseStrength=1.0),
    ]
    mesh = heightfield_mesh(nq=nq, seed=seed, motion=motion)
    cam = Camera(From=(0.0, 1.25, 2.1), To=(0.0, 0.0, 0.0), Fov=50.0, Focal=1.0)
    return SceneDesc(XRes=xres, YRes=yres, camera=cam, shaders=shaders, meshes=[mesh],
                     lights=_light_pair(1.5, 0.4, "lightmtl"), MaxIter=64,
                     name=("C4-motion-heightfield" if motion else "C2-heightfield") + "-%dtri" % mesh.num_tris)


def _uv_sphere(slices: int, stacks: int):
    """Unit UV sphere: 2 + (stacks-1)*slices vertices, 2*slices*(stacks-1) triangles, outward winding."""
    verts = [(0.0, 1.0, 0.0)]
    for s in range(1, stacks):
        th = np.pi * s / stacks
        for l in range(slices):
            ph = 2.0 * np.pi * l / slices
            verts.append((np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)))
    verts.append((0.0, -1.0, 0.0))
    v = np.asarray(verts, np.float32)
    tris = []
    ring = lambda s, l: 1 + (s - 1) * slices + (l % slices)
    south = len(verts) - 1
    for l in range(slices):
        tris += [0, ring(1, l + 1), ring(1, l)]
    for s in range(1, stacks - 1):
        for l in range(slices):
            a, b, c, d = ring(s, l), ring(s, l + 1), ring(s + 1, l + 1), ring(s + 1, l)
            tris += [a, b, c, a, c, d]
    for l in range(slices):
        tris += [south, ring(stacks - 1, l), ring(stacks - 1, l + 1)]
    t = np.asarray(tris, np.int32).reshape(-1, 3)
    # make every triangle face outward
    p0, p1, p2 = v[t[:, 0]], v[t[:, 1]], v[t[:, 2]]
    n = np.cross(p1 - p0, p2 - p0)
    flip = (n * (p0 + p1 + p2)).sum(-1) < 0
    t[flip] = t[flip][:, [0, 2, 1]]
    return v, t.reshape(-1)


def sphere_field_scene(xres: int = 1920, yres: int = 1080, nmesh: int = 1024, slices: int = 70, stacks: int = 71,
                       seed: int = 3, mirror: bool = True) -> SceneDesc:
    """Config C3 / C5: nmesh displaced spheres on a jittered grid over a ground quad; mirror+diffuse shader so
    Level 0..3 chains (the reference's "4 bounces": builtin/shader/std.go:95,219-261)."""
    rng = np.random.default_rng(seed)
    base_v, base_t = _uv_sphere(slices, stacks)
    g = int(np.ceil(np.sqrt(nmesh)))
    cell = 2.0 / g
    shaders = [
        ShaderStd("ground", DiffuseColour=(0.6, 0.6, 0.6), DiffuseStrength=1.0),
        ShaderStd("lightmtl", EmissionColour=(1.0, 0.95, 0.9), EmissionStrength=20.0, DiffuseColour=(0.0, 0.0, 0.0), DiffuseStrength=1.0),
    ]
    if mirror:
        shaders.append(ShaderStd("chrome", DiffuseColour=(0.5, 0.45, 0.4), DiffuseStrength=0.3,
                                 Spec1Colour=(0.9, 0.9, 0.9), Spec1Strength=0.7, Spec1Roughness=0.0))
    else:
        shaders.append(ShaderStd("chrome", DiffuseColour=(0.5, 0.45, 0.4), DiffuseStrength=1.0))
    meshes = []
    gv = np.asarray([[-1.2, 0, -1.2], [-1.2, 0, 1.2], [1.2, 0, 1.2], [1.2, 0, -1.2]], np.float32)
    meshes.append(PolyMesh("floor", gv, ["ground"], PolyCount=np.asarray([4]), FaceIdx=np.asarray([0, 1, 2, 3])))
    for m in range(nmesh):
        gx, gz = m % g, m // g
        r = cell * (0.28 + 0.12 * rng.random())
        cx = -1.0 + cell * (gx + 0.5) + cell * 0.15 * (rng.random() - 0.5)
        cz = -1.0 + cell * (gz + 0.5) + cell * 0.15 * (rng.random() - 0.5)
        cy = r * 1.05 + cell * 0.3 * rng.random()
        disp = 1.0 + 0.06 * np.sin(base_v[:, 0] * 9.0 + m) * np.sin(base_v[:, 1] * 7.0 + 2 * m) * np.sin(base_v[:, 2] * 8.0)
        v = (base_v * disp[:, None].astype(np.float32) * np.float32(r) + np.asarray([cx, cy, cz], np.float32)).astype(np.float32)
        meshes.append(PolyMesh("sphere%d" % m, v, ["chrome"], FaceIdx=base_t.copy()))
    cam = Camera(From=(0.0, 1.4, 2.3), To=(0.0, 0.1, 0.0), Fov=45.0, Focal=1.0)
    sc = SceneDesc(XRes=xres, YRes=yres, camera=cam, shaders=shaders, meshes=meshes,
                   lights=_light_pair(1.8, 0.5, "lightmtl"), MaxIter=256)
    sc.name = "C3-spherefield-%dtri" % sc.num_tris
    return sc


def _quad(name, p, shader):
    return PolyMesh(name, np.asarray(p, np.float32), [shader], PolyCount=np.asarray([4]), FaceIdx=np.asarray([0, 1, 2, 3]))


def _box(name, lo, hi, shader):
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    v = np.asarray([[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0], [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]], np.float32)
    faces = [[0, 3, 2, 1], [4, 5, 6, 7], [0, 1, 5, 4], [3, 7, 6, 2], [0, 4, 7, 3], [1, 2, 6, 5]]  # outward
    return PolyMesh(name, v, [shader], PolyCount=np.full(6, 4), FaceIdx=np.asarray(faces).reshape(-1))


def cornell_box(xres: int = 512, yres: int = 512, boxes: bool = True) -> SceneDesc:
    """Config C1: room [-1,1]x[0,2]x[-1,1], camera on +z looking at -z."""
    shaders = [
        ShaderStd("white", DiffuseColour=(0.73, 0.73, 0.73), DiffuseStrength=1.0),
        ShaderStd("red", DiffuseColour=(0.65, 0.05, 0.05), DiffuseStrength=1.0),
        ShaderStd("green", DiffuseColour=(0.12, 0.45, 0.15), DiffuseStrength=1.0),
        ShaderStd("lightmtl", EmissionColour=(1.0, 0.9, 0.8), EmissionStrength=15.0, DiffuseColour=(0.0, 0.0, 0.0), DiffuseStrength=1.0),
    ]
    meshes = [
        _quad("floor", [[-1, 0, -1], [-1, 0, 1], [1, 0, 1], [1, 0, -1]], "white"),      # Ng=+Y
        _quad("ceiling", [[-1, 2, -1], [1, 2, -1], [1, 2, 1], [-1, 2, 1]], "white"),    # Ng=-Y
        _quad("back", [[-1, 0, -1], [1, 0, -1], [1, 2, -1], [-1, 2, -1]], "white"),     # Ng=+Z
        _quad("left", [[-1, 0, -1], [-1, 2, -1], [-1, 2, 1], [-1, 0, 1]], "red"),       # Ng=+X
        _quad("right", [[1, 0, -1], [1, 0, 1], [1, 2, 1], [1, 2, -1]], "green"),        # Ng=-X
    ]
    if boxes:
        meshes.append(_box("shortbox", (0.1, 0.0, 0.0), (0.7, 0.6, 0.6), "white"))
        meshes.append(_box("tallbox", (-0.7, 0.0, -0.6), (-0.1, 1.2, 0.0), "white"))
    lights = _light_pair(1.99, 0.35, "lightmtl", dy=0.03)
    cam = Camera(From=(0.0, 1.0, 3.4), To=(0.0, 1.0, 0.0), Fov=40.0, Focal=1.0)
    return SceneDesc(XRes=xres, YRes=yres, camera=cam, shaders=shaders, meshes=meshes, lights=lights, MaxIter=16, name="C1-cornell")


def _test_texture(w: int, h: int, seed: int) -> np.ndarray:
    """A seeded image with structure at every scale (checker + stripes + noise), (h, w, 3) uint8, row 0 = top."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    chk = (((x * 8) // w + (y * 8) // h) & 1).astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    img[..., 0] = 40 + 180 * chk
    img[..., 1] = 30 + 200 * (0.5 + 0.5 * np.sin(x * (2 * np.pi * 5 / w)))
    img[..., 2] = 50 + 150 * (0.5 + 0.5 * np.cos(y * (2 * np.pi * 3 / h)))
    img += rng.integers(-25, 26, size=(h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def textured_room(xres: int = 160, yres: int = 120, mirror: bool = True, smooth: bool = True, float_maps: bool = False) -> SceneDesc:
    """SURVEY.md 8(f).4 test scene: the Cornell room with texture maps (builtin/maps/texture.go) on ShaderStd parameters.
      floor     DiffuseColour = Feline map, UVs tiled 3x -> grazing, strongly anisotropic footprints (many probes)
      back      EmissionColour = trilinear map on a black diffuse wall: the image shows the filtered texture itself
      left      DiffuseColour = Feline map on a mesh WITHOUT UVs (the barycentric fallback of trace.go:355-358,495-501)
      right     odd-sized (37x23) texture: the NP2 branches of the pyramid
      shortbox  mirror (mirror=True): the floor texture seen through Ray.Init's reflected differentials
      ball      a smooth-shaded sphere mesh (smooth=True) with a mirror lobe: DdNdx/DdNdy of interpolated normals
      tallbox   float_maps=True: DiffuseStrength / DiffuseRoughness read from texture channels
    """
    sc = cornell_box(xres, yres)
    sc.textures = [Texture("floor.png", _test_texture(64, 64, 11)), Texture("picture.png", _test_texture(128, 64, 12)),
                   Texture("wall.png", _test_texture(32, 32, 13)), Texture("odd.png", _test_texture(37, 23, 14))]
    sh = {s.Name: s for s in sc.shaders}
    sc.shaders.append(ShaderStd("floor_tex", DiffuseColour="floor.png", DiffuseStrength=1.0))
    sc.shaders.append(ShaderStd("picture", EmissionColour="picture.png?filter=trilinear", EmissionStrength=0.8,
                                DiffuseColour=(0.0, 0.0, 0.0), DiffuseStrength=1.0))
    sc.shaders.append(ShaderStd("wall_tex", DiffuseColour="wall.png", DiffuseStrength=1.0))
    sc.shaders.append(ShaderStd("odd_tex", DiffuseColour="odd.png?filter=trilinear", DiffuseStrength=1.0))
    by = {m.Name: m for m in sc.meshes}
    by["floor"].Shader = ["floor_tex"]
    by["floor"].UV = np.asarray([[0, 0], [0, 3], [3, 3], [3, 0]], np.float32)
    by["back"].Shader = ["picture"]
    by["back"].UV = np.asarray([[0.1, 0.2], [1.7, 0.2], [1.7, 1.3], [0.1, 1.3]], np.float32)
    by["back"].UVIdx = np.asarray([0, 1, 2, 3], np.int32)
    by["left"].Shader = ["wall_tex"]
    by["right"].Shader = ["odd_tex"]
    by["right"].UV = np.asarray([[0, 0], [2, 0], [2, 2], [0, 2]], np.float32)
    if mirror:
        sc.shaders.append(ShaderStd("mirror", DiffuseColour=(0.4, 0.4, 0.4), DiffuseStrength=0.3,
                                    Spec1Colour=(0.9, 0.9, 0.9), Spec1Strength=0.7, Spec1Roughness=0.0))
        by["shortbox"].Shader = ["mirror"]
    if smooth:
        v, t = _uv_sphere(12, 8)
        if not mirror:
            sc.shaders.append(ShaderStd("mirror", DiffuseColour=(0.4, 0.4, 0.4), DiffuseStrength=0.3,
                                        Spec1Colour=(0.9, 0.9, 0.9), Spec1Strength=0.7, Spec1Roughness=0.0))
        pos = (v * np.float32(0.28) + np.asarray([0.45, 0.95, 0.25], np.float32)).astype(np.float32)
        sc.meshes.append(PolyMesh("ball", pos, ["mirror"], FaceIdx=t.copy(), Normals=v.copy()))
    if float_maps:
        sc.shaders.append(ShaderStd("tall_tex", DiffuseColour=(0.7, 0.7, 0.6), DiffuseStrength="wall.png?ch=1", DiffuseRoughness="floor.png?ch=2&filter=trilinear",
                                    Spec1Colour=(0.9, 0.9, 0.9), Spec1Strength=0.25, Spec1Roughness=0.0))
        by["tallbox"].Shader = ["tall_tex"]
    # looking down into the room: the floor is seen at a grazing angle (long, thin texture footprints)
    sc.camera = Camera(From=(0.03, 1.6, 3.4), To=(0.0, 0.4, 0.0), Fov=50.0, Focal=1.0)
    sc.name = "textured-room"
    return sc


def debug_shader_box(xres: int = 128, yres: int = 128, mirrors: bool = True) -> SceneDesc:
    """The Cornell room with DebugShader walls (builtin/shader/debug.go: OutRGB = Colour). With `mirrors` the floor and the ceiling
    are facing mirrors, so chains run to Level 4, where a DebugShader still answers while ShaderStd does not (std.go:95)."""
    sc = cornell_box(xres, yres)
    sc.shaders.append(DebugShader("dbg_blue", Colour=(0.2, 0.6, 0.9)))
    sc.shaders.append(DebugShader("dbg_orange", Colour=(0.9, 0.5, 0.1)))
    sc.meshes[3].Shader = ["dbg_blue"]      # left wall
    sc.meshes[2].Shader = ["dbg_orange"]    # back wall
    if mirrors:
        sc.shaders.append(ShaderStd("mirror", DiffuseColour=(0.5, 0.5, 0.5), DiffuseStrength=0.3,
                                    Spec1Colour=(0.9, 0.9, 0.9), Spec1Strength=0.7, Spec1Roughness=0.0))
        sc.meshes[0].Shader = ["mirror"]    # floor
        sc.meshes[1].Shader = ["mirror"]    # ceiling
    # off the room's axis: from (0, 1, z) the wall/ceiling edges project exactly onto the image diagonals, and the one sample per
    # diagonal pixel that sits on the sub-pixel diagonal then runs precisely along the crack between two meshes, where the last bit
    # of the direction's normalize (RSQRTSS in the reference) decides between a constant colour and a lit wall
    sc.camera = Camera(From=(0.03, 1.02, 3.4), To=(0.0, 1.0, 0.0), Fov=40.0, Focal=1.0)
    sc.name = "debug-shader-box"
    return sc


def glossy_box(xres: int = 256, yres: int = 256, lights: str = "tri,disk,sphere") -> SceneDesc:
    """SURVEY.md 8(f).1 test scene: the Cornell room with a GGX-glossy dielectric floor, a GGX "Metal" (conductor Fresnel)
    tall box, a mirror short box, and one light of each in-scope type (TriLight, DiskLight, SphereLight) in node order."""
    shaders = [
        ShaderStd("white", DiffuseColour=(0.73, 0.73, 0.73), DiffuseStrength=1.0),
        ShaderStd("red", DiffuseColour=(0.65, 0.05, 0.05), DiffuseStrength=1.0),
        ShaderStd("green", DiffuseColour=(0.12, 0.45, 0.15), DiffuseStrength=1.0),
        ShaderStd("lightmtl", EmissionColour=(1.0, 0.9, 0.8), EmissionStrength=12.0, DiffuseColour=(0.0, 0.0, 0.0), DiffuseStrength=1.0),
        ShaderStd("glossfloor", DiffuseColour=(0.5, 0.5, 0.55), DiffuseStrength=0.6, Spec1Colour=(1.0, 1.0, 1.0), Spec1Strength=0.4,
                  Spec1Roughness=0.45, IOR=1.5),
        ShaderStd("gold", DiffuseColour=(0.3, 0.2, 0.05), DiffuseStrength=0.2, Spec1Colour=(1.0, 0.85, 0.5), Spec1Strength=0.8,
                  Spec1Roughness=0.6, Spec1FresnelModel="Metal", Spec1FresnelRefl=(0.95, 0.75, 0.35), Spec1FresnelEdge=(0.9, 0.8, 0.6)),
        ShaderStd("mirror", DiffuseColour=(0.2, 0.2, 0.2), DiffuseStrength=0.2, Spec1Colour=(0.9, 0.9, 0.9), Spec1Strength=0.8,
                  Spec1Roughness=0.0, Spec1FresnelModel="Metal", Spec1FresnelRefl=(0.8, 0.8, 0.85)),
    ]
    meshes = [
        _quad("floor", [[-1, 0, -1], [-1, 0, 1], [1, 0, 1], [1, 0, -1]], "glossfloor"),
        _quad("ceiling", [[-1, 2, -1], [1, 2, -1], [1, 2, 1], [-1, 2, 1]], "white"),
        _quad("back", [[-1, 0, -1], [1, 0, -1], [1, 2, -1], [-1, 2, -1]], "white"),
        _quad("left", [[-1, 0, -1], [-1, 2, -1], [-1, 2, 1], [-1, 0, 1]], "red"),
        _quad("right", [[1, 0, -1], [1, 0, 1], [1, 2, 1], [1, 2, -1]], "green"),
        _box("shortbox", (0.1, 0.0, 0.0), (0.7, 0.6, 0.6), "mirror"),
        _box("tallbox", (-0.7, 0.0, -0.6), (-0.1, 1.2, 0.0), "gold"),
    ]
    ls = []
    for kind in [k for k in lights.split(",") if k]:
        if kind == "tri":
            ls.append(TriLight("lightT", (-0.9, 1.95, -0.2), (-0.3, 1.95, -0.2), (-0.6, 1.95, 0.4), "lightmtl", 1))
        elif kind == "disk":
            ls.append(DiskLight("lightD", P=(0.45, 1.97, 0.1), LookAt=(0.45, 0.0, 0.15), Up=(0.0, 0.0, 1.0), Radius=0.3, Shader="lightmtl",
                                Segments=12, Samples=1))
        elif kind == "sphere":
            ls.append(SphereLight("lightS", P=(0.0, 1.3, 0.55), Shader="lightmtl", Radius=0.12, Samples=1))
        else:
            raise ValueError(kind)
    cam = Camera(From=(0.0, 1.0, 3.4), To=(0.0, 1.0, 0.0), Fov=40.0, Focal=1.0)
    return SceneDesc(XRes=xres, YRes=yres, camera=cam, shaders=shaders, meshes=meshes, lights=ls, MaxIter=16, name="F1-glossy-box")


def instanced_scene(xres: int = 128, yres: int = 96, moving: bool = True, motion_base: bool = False, nested: bool = False) -> SceneDesc:
    """SURVEY.md 8(f).3 test scene: one displaced-sphere mesh at the origin, three GeomInstances of it (translated, rotated,
    uniformly scaled; the last one with two transform keys when `moving`), a floor and a TriLight pair. Bounds are the
    transformed mesh bounds with a margin, over both keys."""
    base_v, base_t = _uv_sphere(20, 21)
    disp = 1.0 + 0.15 * np.sin(base_v[:, 0] * 5.0) * np.sin(base_v[:, 1] * 4.0 + 1.0)
    v0 = (base_v * disp[:, None] * 0.22 + np.asarray([0.0, 0.25, 0.0])).astype(np.float32)
    verts = v0[None]
    if motion_base:
        v1 = (v0 + np.asarray([0.0, 0.05, 0.0], np.float32) * np.sin(v0[:, :1] * 9.0)).astype(np.float32)
        verts = np.stack([v0, v1], 0)
    shaders = [
        ShaderStd("ground", DiffuseColour=(0.6, 0.6, 0.6), DiffuseStrength=1.0),
        ShaderStd("lightmtl", EmissionColour=(1.0, 0.95, 0.9), EmissionStrength=20.0, DiffuseColour=(0.0, 0.0, 0.0), DiffuseStrength=1.0),
        ShaderStd("clay", DiffuseColour=(0.7, 0.35, 0.25), DiffuseStrength=1.0),
    ]
    gv = np.asarray([[-1.5, 0, -1.5], [-1.5, 0, 1.5], [1.5, 0, 1.5], [1.5, 0, -1.5]], np.float32)
    meshes = [PolyMesh("floor", gv, ["ground"], PolyCount=np.asarray([4]), FaceIdx=np.asarray([0, 1, 2, 3])),
              PolyMesh("blob", verts, ["clay"], FaceIdx=base_t.copy())]

    def bounds(mats):
        lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
        for m in mats:
            M = m.reshape(4, 4).T.astype(np.float64)
            for vv in verts:
                w = vv.astype(np.float64) @ M[:3, :3].T + M[:3, 3]
                lo, hi = np.minimum(lo, w.min(0)), np.maximum(hi, w.max(0))
        return tuple((lo - 0.02).astype(np.float32)), tuple((hi + 0.02).astype(np.float32))

    specs = [[srt_matrix((-0.75, 0.0, 0.2), 30.0, 1.0)],
             [srt_matrix((0.7, 0.1, -0.3), -50.0, 1.4)],
             [srt_matrix((0.1, 0.0, 0.8), 10.0, 0.8)] + ([srt_matrix((0.3, 0.15, 0.8), 70.0, 0.8)] if moving else [])]
    instances = []
    for i, mats in enumerate(specs):
        lo, hi = bounds(mats)
        instances.append(GeomInstance("inst%d" % i, "blob", np.stack(mats, 0), lo, hi))
    if nested:
        # an instance of an INSTANCE (Instance.Trace just calls ins.geom.Trace, instance.go:95): "inst0" placed once more, the
        # outer transform applied on top of inst0's own; and an instance of that one (three transforms deep)
        def compose(outer, inner):
            return matrix4((outer.reshape(4, 4).T.astype(np.float64) @ inner.reshape(4, 4).T.astype(np.float64)))
        outer = srt_matrix((0.9, 0.35, 0.9), 20.0, 0.7)
        lo, hi = bounds([compose(outer, specs[0][0])])
        instances.append(GeomInstance("nest1", "inst0", np.stack([outer], 0), lo, hi))
        outer2 = srt_matrix((-1.6, 0.0, -0.2), -35.0, 1.0)
        lo, hi = bounds([compose(outer2, compose(outer, specs[0][0]))])
        instances.append(GeomInstance("nest2", "nest1", np.stack([outer2], 0), lo, hi))
    cam = Camera(From=(0.0, 1.3, 2.6), To=(0.0, 0.2, 0.0), Fov=42.0, Focal=1.0)
    sc = SceneDesc(XRes=xres, YRes=yres, camera=cam, shaders=shaders, meshes=meshes, lights=_light_pair(1.9, 0.5, "lightmtl"),
                   instances=instances, MaxIter=16, name="F3-instances")
    return sc


def incoherent_rays(rays: np.ndarray, hits: np.ndarray, seed: int = 7) -> np.ndarray:
    """Level-1 "incoherent closest-hit" batch: cosine-hemisphere directions about the
    direction-facing axis-aligned approximation of the surface normal at the primary hit points.
    `rays` is the structured VgRay array, `hits` the VgHit array of the primary batch.  Only hit rays spawn
    a bounce ray.  The normal is estimated from the incoming direction (flipped +Y), which is enough to make the
    batch incoherent; exact normals are not needed because both sides trace these identical rays."""
    rng = np.random.default_rng(seed)
    mask = hits["prim"] >= 0
    o = rays["o"][mask] + rays["d"][mask] * hits["t"][mask][:, None]
    n = len(o)
    u0, u1 = rng.random(n), rng.random(n)
    r = np.sqrt(1.0 - u0)
    th = 2.0 * np.pi * u1
    d = np.stack([r * np.cos(th), np.sqrt(u0), r * np.sin(th)], -1)  # hemisphere about +Y
    out = np.zeros(n, dtype=rays.dtype)
    out["o"] = (o + np.asarray([0, 1e-3, 0])).astype(np.float32)
    dn = d / np.linalg.norm(d, axis=1, keepdims=True)
    out["d"] = dn.astype(np.float32)
    out["tmax"] = np.float32(np.inf)
    out["time"] = rays["time"][mask]
    return out


def diffuse_bounce_rays(scene: SceneDesc, rays: np.ndarray, hits: np.ndarray, seed: int = 7) -> np.ndarray:
    """One diffuse bounce off the sphere field (bench.py: the 4-bounce incoherent wavefront of the 10 M-triangle scene, the workload
    BASELINE.json's third config names): from every hit point a cosine-hemisphere direction about the surface normal, taken as the
    direction from the hit mesh's centroid (the displaced spheres) or +Y (floor, lights), flipped to face the incoming ray. Both sides
    trace these identical rays, so the normal only has to make the batch look like a diffuse bounce, not match the shading normal."""
    rng = np.random.default_rng(seed)
    mask = hits["prim"] >= 0
    o = rays["o"][mask].astype(np.float64) + rays["d"][mask].astype(np.float64) * hits["t"][mask][:, None].astype(np.float64)
    geom = hits["geom"][mask]
    cent = np.zeros((len(scene.meshes), 3))
    is_sphere = np.zeros(len(scene.meshes), bool)
    for gi, m in enumerate(scene.meshes):
        cent[gi] = np.asarray(m.Verts[0], np.float64).mean(0)   # Verts is (keys, nverts, 3)
        is_sphere[gi] = m.Name.startswith("sphere")
    gi = np.clip(geom, 0, len(scene.meshes) - 1)
    nrm = np.where((is_sphere[gi] & (geom < len(scene.meshes)))[:, None], o - cent[gi], np.asarray([0.0, 1.0, 0.0]))
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
    nrm = np.where((nrm * rays["d"][mask]).sum(1, keepdims=True) > 0, -nrm, nrm)
    n = len(o)
    u0, u1 = rng.random(n), rng.random(n)
    r, th = np.sqrt(1.0 - u0), 2.0 * np.pi * u1
    a = np.where(np.abs(nrm[:, :1]) > 0.9, np.asarray([0.0, 1.0, 0.0]), np.asarray([1.0, 0.0, 0.0]))
    t1 = np.cross(nrm, a)
    t1 /= np.linalg.norm(t1, axis=1, keepdims=True)
    t2 = np.cross(nrm, t1)
    d = t1 * (r * np.cos(th))[:, None] + t2 * (r * np.sin(th))[:, None] + nrm * np.sqrt(u0)[:, None]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    out = np.zeros(n, dtype=rays.dtype)
    out["o"] = (o + nrm * 1e-4).astype(np.float32)
    out["d"] = d.astype(np.float32)
    out["tmax"] = np.float32(np.inf)
    out["time"] = rays["time"][mask]
    return out


def camera_motion_variants():
    """Camera nodes with motion keys (camera.go:48-73) for the heightfield / Cornell framing: name -> Camera.
    'from3': three From keys, fixed target; 'to4_from2': more To keys than From keys (the other calcLookatMatrices branch, where the
    eye is the interpolated From); 'roll': two Roll keys and a lens radius; 'matrix2': Type "Matrix" with two WorldToLocal keys."""
    base = dict(Fov=50.0, Focal=1.0)
    out = {
        "from3": Camera(From=(0.0, 1.25, 2.1), To=(0.0, 0.0, 0.0), FromKeys=[(-0.25, 1.25, 2.1), (0.0, 1.3, 2.0), (0.3, 1.2, 2.15)], **base),
        "to4_from2": Camera(From=(0.0, 1.25, 2.1), To=(0.0, 0.0, 0.0), FromKeys=[(-0.1, 1.25, 2.1), (0.1, 1.25, 2.1)],
                            ToKeys=[(0.0, 0.0, 0.0), (0.05, 0.02, 0.0), (0.1, 0.0, 0.05), (0.12, 0.0, 0.1)], **base),
        "roll": Camera(From=(0.0, 1.25, 2.1), To=(0.0, 0.0, 0.0), RollKeys=[-0.1, 0.15], Radius=0.01, **base),
    }
    # two rigid world-to-camera matrices: local-to-world = T(eye) * Ry(yaw) * Rx(-pitch) (the camera looks down its local -z),
    # yaw -4 and +4 degrees; WorldToLocal is the inverse, rounded to float32
    pitch = np.arctan2(1.25, 2.1)
    cp, sp = np.cos(-pitch), np.sin(-pitch)
    rx = np.array([[1, 0, 0, 0], [0, cp, -sp, 0], [0, sp, cp, 0], [0, 0, 0, 1]], np.float64)
    t = np.eye(4)
    t[:3, 3] = (0.0, 1.25, 2.1)
    mats = []
    for deg in (-4.0, 4.0):
        a = np.deg2rad(deg)
        ry = np.array([[np.cos(a), 0, np.sin(a), 0], [0, 1, 0, 0], [-np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1]], np.float64)
        mats.append(matrix4(np.linalg.inv(t @ ry @ rx)))
    out["matrix2"] = Camera(From=(0.0, 0.0, 0.0), To=(0.0, 0.0, -1.0), Type="Matrix", WorldToLocal=mats, **base)
    return out
