"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding over oracle/liboracle.so.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product (vermeer_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3), ("tmax", np.float32), ("time", np.float32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("w", np.float32),
                      ("prim", np.int32), ("geom", np.int32), ("nodesT", np.int32), ("trisT", np.int32)])
NODE_DTYPE = np.dtype([("Boxes", np.float32, 24), ("Axis0", np.uint32), ("Axis1", np.uint32), ("Axis2", np.uint32),
                       ("Children", np.int32, 4), ("Parent", np.int32)])
MNODE_DTYPE = np.dtype([("Axis0", np.int32), ("Axis1", np.int32), ("Axis2", np.int32), ("Children", np.int32, 4),
                        ("Parent", np.int32), ("pad", np.uint32, 2)])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 32 and NODE_DTYPE.itemsize == 128 and MNODE_DTYPE.itemsize == 40


def build(force: bool = False) -> str:
    """Compile the C++ restatement (make -C oracle)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_create.restype = C.c_void_p
        L.orc_last_error.restype = C.c_char_p
        L.orc_vdc_u.restype = C.c_uint64
        L.orc_sobol_u.restype = C.c_uint64
        L.orc_vdc.restype = C.c_double
        L.orc_sobol.restype = C.c_double
        L.orc_raster_xy.restype = C.c_uint64
        L.orc_bessel_j1.restype = C.c_double
        L.orc_bessel_j1.argtypes = [C.c_double]
        L.orc_filter_warp.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_filter_warp.restype = None
        for f in ("orc_vdc_u", "orc_sobol_u", "orc_vdc", "orc_sobol"):
            getattr(L, f).argtypes = [C.c_uint64, C.c_uint64]
        L.orc_raster_xy.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class Oracle:
    """One reference-semantics renderer instance built from a vermeer_b200.scenes.SceneDesc."""

    def __init__(self, scene, motion_ref_compat: bool = True):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create())
        self.scene = scene
        L, h = self.L, self.h
        self._chk(L.orc_set_globals(h, scene.XRes, scene.YRes))
        for s in scene.shaders:
            mask, p = s.packed()
            self._chk(L.orc_add_shader(h, s.Name.encode(), C.c_uint32(mask), _p(p)))
        for t in getattr(scene, "textures", []):
            px = t.rows_bottom_up()
            self._chk(L.orc_add_texture(h, t.Name.encode(), px.shape[1], px.shape[0], _p(px)))
        for s in scene.shaders:
            for slot, path, chan, tri in (s.texture_bindings() if hasattr(s, "texture_bindings") else []):
                self._chk(L.orc_shader_set_texture(h, s.Name.encode(), slot, path.encode(), chan, 1 if tri else 0))
        for m in scene.meshes:
            keys, nverts, _ = m.Verts.shape
            self._chk(L.orc_add_polymesh(
                h, m.Name.encode(), _p(m.Verts), nverts, keys,
                _p(m.PolyCount), 0 if m.PolyCount is None else len(m.PolyCount),
                _p(m.FaceIdx), 0 if m.FaceIdx is None else len(m.FaceIdx),
                "\n".join(m.Shader).encode(),
                _p(m.ShaderIdx), 0 if m.ShaderIdx is None else len(m.ShaderIdx),
                _p(m.Normals), 0 if m.Normals is None else len(m.Normals),
                _p(m.NormalIdx), 0 if m.NormalIdx is None else len(m.NormalIdx),
                C.c_float(m.RayBias)))
            if getattr(m, "UV", None) is not None:
                self._chk(L.orc_mesh_set_uv(h, m.Name.encode(), _p(m.UV), len(m.UV), _p(m.UVIdx), 0 if m.UVIdx is None else len(m.UVIdx)))
        for ins in getattr(scene, "instances", []):
            bmin = np.ascontiguousarray(ins.BMin, np.float32).reshape(-1, 3)
            bmax = np.ascontiguousarray(ins.BMax, np.float32).reshape(-1, 3)
            tr = np.ascontiguousarray(ins.Transform, np.float32).reshape(-1, 16)
            self._chk(L.orc_add_instance(h, ins.Name.encode(), ins.Geom.encode(), _p(bmin), _p(bmax), len(bmin), _p(tr), len(tr)))
        for l in scene.lights:
            kind = type(l).__name__
            if kind == "TriLight":
                self._chk(L.orc_add_trilight(h, l.Name.encode(), _f3(l.P0), _f3(l.P1), _f3(l.P2), l.Shader.encode(), l.Samples))
            elif kind == "DiskLight":
                self._chk(L.orc_add_disklight(h, l.Name.encode(), _f3(l.P), _f3(l.LookAt), _f3(l.Up), C.c_float(l.Radius), l.Shader.encode(),
                                              int(l.Segments), int(l.Samples)))
            elif kind == "SphereLight":
                self._chk(L.orc_add_spherelight(h, l.Name.encode(), _f3(l.P), C.c_float(l.Radius), l.Shader.encode(), int(l.Samples)))
            else:
                raise ValueError("unknown light node %r" % kind)
        c = scene.camera
        if getattr(c, "has_keys", False):
            fr, to, ro, w2l = c.keys()
            self._chk(L.orc_set_camera_keys(h, c.Type.encode(), _p(fr), len(fr), _p(to), len(to), _p(ro), len(ro), _f3(c.Up), _p(w2l), len(w2l),
                                            C.c_float(c.Fov), C.c_float(c.Focal), C.c_float(c.Aspect), C.c_float(c.Radius)))
        else:
            self._chk(L.orc_set_camera(h, _f3(c.From), _f3(c.To), _f3(c.Up), C.c_float(c.Roll), C.c_float(c.Fov), C.c_float(c.Focal),
                                       C.c_float(c.Aspect), C.c_float(c.Radius)))
        if getattr(scene, "filter", None) is not None:
            f = scene.filter
            airy = f.Type == "AiryFilter"
            L.orc_set_filter(h, 1 if airy else 2, C.c_float(f.Width or (6 if airy else 2)), int(f.Res or (49 if airy else 17)), C.c_float(f.Peak or (4 if airy else 0)))
        L.orc_set_motion_ref_compat(h, 1 if motion_ref_compat else 0)
        self._chk(L.orc_prerender(h))

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- rendering -------------------------------------------------------------------------------
    def mesh_set_transform(self, mesh_name: str, matrices):
        """PolyMesh.Transform (polymesh.go:32): keys x 16 floats in math.Matrix4 storage. Call before the first trace / render.
        Exists to document reference quirk q (tests/test_oracle_mesh_transform.py); the GPU path refuses such meshes."""
        m = np.ascontiguousarray(matrices, np.float32).reshape(-1, 16)
        self._chk(self.L.orc_mesh_set_transform(self.h, mesh_name.encode(), _p(m), len(m)))

    def set_scramble(self, table: np.ndarray):
        table = np.ascontiguousarray(table, np.uint64)
        assert table.shape == (self.scene.XRes * self.scene.YRes, 6)
        self._tab = table
        self._chk(self.L.orc_set_scramble(self.h, _p(table), C.c_int64(table.shape[0])))

    def clear(self):
        self.L.orc_clear_framebuffer(self.h)

    def render(self, iter_begin: int, iter_end: int, nthreads: int = 1, trace_last_level: bool = True):
        """Returns (framebuffer (H,W,3) float32, dict(rays, shadow_rays, seconds)). nthreads < 0 = the reference's own worker
        policy: min(-nthreads, 10) workers (core/render.go:190) and global atomic ray counters (core/stats.go:26-33)."""
        fb = np.zeros((self.scene.YRes, self.scene.XRes, 3), np.float32)
        st = np.zeros(3, np.uint64)
        self._chk(self.L.orc_render(self.h, iter_begin, iter_end, nthreads, 1 if trace_last_level else 0, _p(fb), _p(st)))
        return fb, {"rays": int(st[0]), "shadow_rays": int(st[1]), "seconds": float(st[2]) * 1e-9}

    def camera_rays(self, iter1: int, x0=0, y0=0, w=None, h=None) -> np.ndarray:
        w = self.scene.XRes if w is None else w
        h = self.scene.YRes if h is None else h
        out = np.zeros(w * h, RAY_DTYPE)
        self._chk(self.L.orc_camera_rays(self.h, iter1, x0, y0, w, h, _p(out)))
        return out

    def trace(self, rays: np.ndarray, any_hit: bool = False, brute: bool = False, nthreads: int = 1) -> np.ndarray:
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.zeros(len(rays), HIT_DTYPE)
        flags = (1 if any_hit else 0) | (2 if brute else 0)
        self._chk(self.L.orc_trace_batch(self.h, _p(rays), C.c_int64(len(rays)), C.c_uint32(flags), nthreads, _p(hits)))
        return hits

    # -- textures ------------------------------------------------------------------------------------
    def texture_levels(self, name: str):
        """The mip pyramid stdfilter built: list of (h, w, 3) uint8 arrays, rows bottom-up."""
        out = []
        for l in range(self.L.orc_texture_num_levels(self.h, name.encode())):
            w, hh = C.c_int(), C.c_int()
            self._chk(self.L.orc_texture_level(self.h, name.encode(), l, C.byref(w), C.byref(hh), None))
            a = np.zeros((hh.value, w.value, 3), np.uint8)
            self._chk(self.L.orc_texture_level(self.h, name.encode(), l, C.byref(w), C.byref(hh), _p(a)))
            out.append(a)
        return out

    def texture_sample(self, name: str, coords: np.ndarray, trilinear: bool = False) -> np.ndarray:
        """coords: (n, 8) float32 {U, V, Dduvdx[2], Dduvdy[2], PixelDelta[2]} -> (n, 3)."""
        coords = np.ascontiguousarray(coords, np.float32).reshape(-1, 8)
        out = np.zeros((len(coords), 3), np.float32)
        self._chk(self.L.orc_texture_sample(self.h, name.encode(), 1 if trilinear else 0, C.c_int64(len(coords)), _p(coords), _p(out)))
        return out

    def camera_texcoords(self, iter1: int, x0=0, y0=0, w=None, h=None) -> np.ndarray:
        """(n, 8): what a texture map sees at the first hit of each camera sample (NaN rows for misses)."""
        w = self.scene.XRes if w is None else w
        h = self.scene.YRes if h is None else h
        out = np.zeros((w * h, 8), np.float32)
        self._chk(self.L.orc_camera_texcoords(self.h, iter1, x0, y0, w, h, _p(out)))
        return out

    # -- structure export --------------------------------------------------------------------------
    def num_geoms(self):
        return self.L.orc_num_geoms(self.h)

    def scene_geom_order(self):
        return np.asarray([self.L.orc_scene_geom_id(self.h, i) for i in range(self.num_geoms())], np.int32)

    def scene_is_motion(self):
        return bool(self.L.orc_scene_is_motion(self.h))

    def scene_nodes(self):
        n = self.L.orc_scene_num_nodes(self.h)
        if self.scene_is_motion():
            keys = self.L.orc_scene_keys(self.h)
            topo = np.zeros(n, MNODE_DTYPE)
            boxes = np.zeros((keys, n, 24), np.float32)
            self.L.orc_scene_motion_nodes(self.h, _p(topo), _p(boxes))
            return topo, boxes
        out = np.zeros(n, NODE_DTYPE)
        self.L.orc_scene_nodes(self.h, _p(out))
        return out

    def mesh_info(self, gid):
        o = np.zeros(6, np.int32)
        self._chk(self.L.orc_mesh_info(self.h, gid, _p(o)))
        return dict(nodes=int(o[0]), tris=int(o[1]), keys=int(o[2]), nverts=int(o[3]), motion=bool(o[4]), normals=bool(o[5]))

    def mesh_nodes(self, gid):
        info = self.mesh_info(gid)
        if info["motion"]:
            topo = np.zeros(info["nodes"], MNODE_DTYPE)
            boxes = np.zeros((info["keys"], info["nodes"], 24), np.float32)
            self.L.orc_mesh_motion_nodes(self.h, gid, _p(topo), _p(boxes))
            return topo, boxes
        out = np.zeros(info["nodes"], NODE_DTYPE)
        self.L.orc_mesh_nodes(self.h, gid, _p(out))
        return out

    def mesh_idxp(self, gid):
        info = self.mesh_info(gid)
        idxp = np.zeros(info["tris"] * 3, np.uint32)
        aidx = np.zeros(info["tris"], np.int32)
        self.L.orc_mesh_idxp(self.h, gid, _p(idxp), _p(aidx))
        return idxp, aidx

    def camera_decomp(self):
        """Camera.decomp: [keys, 23] float32."""
        out = np.zeros((256, 23), np.float32)
        n = self.L.orc_camera_decomp(self.h, _p(out))
        return out[:n].copy()

    def camera_matrix(self):
        m = np.zeros(16, np.float32)
        ttf = C.c_float()
        asp = C.c_float()
        self.L.orc_camera_matrix(self.h, _p(m), C.byref(ttf), C.byref(asp))
        return m, ttf.value, asp.value
