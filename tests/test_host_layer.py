"""Host layer (vh_*) and C-ABI surface, no GPU needed: the library loads, exports every declared symbol, the registry
mirrors nodes.Register, and the host PreRender produces bit-identical structures to the oracle's restatement."""
import ctypes as C
import re

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(built_library):
    import os
    from vermeer_b200.host import DECLARED_SYMBOLS, load_library
    lib = load_library()
    hdr = open(os.path.join(os.path.dirname(built_library), "..", "include", "vermeer_gpu.h")).read()
    declared = sorted(set(re.findall(r"\b(v[gh]_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(DECLARED_SYMBOLS), "host.py DECLARED_SYMBOLS is out of sync with include/vermeer_gpu.h"
    for s in declared:
        assert hasattr(lib, s), "missing export: " + s


def test_no_cpu_fallback_without_device(built_library):
    """On a box without a GPU, creating a device context must fail loudly (VG_ERR_NO_DEVICE)."""
    from vermeer_b200.host import Device, device_count
    if device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        Device(0)


def test_registry_mirrors_nodes_register(built_library):
    from vermeer_b200.host import load_library
    lib = load_library()
    n = lib.vh_registered_nodes(None, 0)
    arr = (C.c_char_p * n)()
    lib.vh_registered_nodes(arr, n)
    names = sorted(a.decode() for a in arr)
    assert names == ["AiryFilter", "Camera", "DebugShader", "DiskLight", "GaussianFilter", "GeomInstance", "Globals", "Include", "OutputFloat", "OutputHDR", "PolyMesh", "ShaderStd", "Sphere", "SphereLight", "TriLight"]


def _equal_nodes(a, b):
    if isinstance(a, tuple):
        return a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()
    return a.tobytes() == b.tobytes()


@pytest.mark.parametrize("name", ["cornell", "heightfield", "motion", "spheres", "glossy", "instances", "nested", "debug"])
def test_host_prerender_matches_oracle_bit_for_bit(built_library, name):
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    from vermeer_b200.host import HostScene
    sc = {"cornell": lambda: scenes.cornell_box(64, 48),
          "heightfield": lambda: scenes.heightfield_scene(64, 48, nq=300),   # > 65536 triangles: exercises the parallel builder
          "motion": lambda: scenes.heightfield_scene(64, 48, nq=90, motion=True),
          "spheres": lambda: scenes.sphere_field_scene(64, 48, nmesh=9, slices=12, stacks=13),
          # TriLight + DiskLight (fan mesh) + SphereLight (analytic Sphere geom in the scene-level tree)
          "glossy": lambda: scenes.glossy_box(64, 48),
          # GeomInstances (one with two transform keys) of a two-key motion mesh: scene-level MQBVH over user-given bounds
          "instances": lambda: scenes.instanced_scene(64, 48, moving=True, motion_base=True),
          "nested": lambda: scenes.instanced_scene(64, 48, moving=False, nested=True),   # instances of instances
          "debug": lambda: scenes.debug_shader_box(64, 48)}[name]()
    o = Oracle(sc)
    h = HostScene(sc).prerender()
    assert np.array_equal(o.scene_geom_order(), h.scene_geom_order())
    assert _equal_nodes(o.scene_nodes(), h.scene_nodes())
    for g in range(o.num_geoms()):
        try:
            info = o.mesh_info(g)
        except RuntimeError:   # an analytic Sphere geom or a GeomInstance: no mesh of its own on either side
            with pytest.raises(RuntimeError):
                h.mesh_info(g)
            continue
        assert info == h.mesh_info(g)
        assert _equal_nodes(o.mesh_nodes(g), h.mesh_nodes(g))
        io, ih = o.mesh_idxp(g), h.mesh_idxp(g)
        assert np.array_equal(io[0], ih[0]) and np.array_equal(io[1], ih[1])
    mo, ttf, asp = o.camera_matrix()
    mh, ttf2, asp2 = h.camera()
    assert mo.tobytes() == mh.tobytes() and ttf == ttf2 and asp == asp2


def test_host_errors_are_statuses_not_aborts(built_library):
    from vermeer_b200 import scenes
    from vermeer_b200.host import HostScene
    sc = scenes.cornell_box(16, 16)
    sc.meshes[0].Shader = ["nonexistent"]
    with pytest.raises(RuntimeError, match="Unable to find node"):
        HostScene(sc).prerender()
    sc = scenes.cornell_box(16, 16)
    sc.meshes[0].FaceIdx = np.asarray([0, 1, 2, 9], np.int32)
    with pytest.raises(RuntimeError, match="vertex index out of range"):
        HostScene(sc).prerender()


def test_tile_partition_is_a_partition():
    """Every pixel is owned by exactly one rank for world = 1, 2, 4, 8 (same formula as render.cu: prepare())."""
    from vermeer_b200.partition import owned_pixels
    for (w, h) in [(1920, 1080), (512, 512), (100, 70)]:
        for world in (1, 2, 4, 8):
            seen = np.zeros(w * h, np.int32)
            sizes = []
            for r in range(world):
                p = owned_pixels(w, h, r, world)
                seen[p] += 1
                sizes.append(len(p))
            assert np.all(seen == 1)
            assert max(sizes) - min(sizes) <= 2 * 32 * 32 * (1 if world > 1 else 0) + 0 or world == 1


@pytest.mark.parametrize("variant", ["from3", "to4_from2", "roll", "matrix2"])
def test_camera_motion_keys_decompose_like_the_oracle(built_library, variant):
    """Camera.PreRender with motion keys (camera.go:109-216): the host's LocalToWorld decompositions are bit-identical to the
    oracle's, through the vh_* calls and through the .vnf reader."""
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    from vermeer_b200.host import HostScene
    sc = scenes.heightfield_scene(48, 32, nq=8)
    sc.camera = scenes.camera_motion_variants()[variant]
    want_keys = {"from3": 3, "to4_from2": 4, "roll": 1, "matrix2": 2}[variant]
    o = Oracle(sc)
    h = HostScene(sc).prerender()
    do, dh = o.camera_decomp(), h.camera_decomp()
    assert do.shape == (want_keys, 23) and np.isfinite(do).all()
    assert do.tobytes() == dh.tobytes()
    mo, ttf, asp = o.camera_matrix()
    mh, ttf2, asp2 = h.camera()
    assert mo.tobytes() == mh.tobytes() and ttf == ttf2 and asp == asp2
    h2 = HostScene.from_vnf(scenes.to_vnf(sc)).prerender()
    assert h2.camera_decomp().tobytes() == dh.tobytes()
    assert h2.camera()[0].tobytes() == mh.tobytes()
    if want_keys > 1:   # the keys differ: the camera really moves
        assert not np.array_equal(do[0], do[-1])


def test_matrix_camera_without_keys_is_the_identity(built_library):
    """A non-LookAt camera with no WorldToLocal: c.decomp is nil and ComputeRay keeps Matrix4Identity (camera.go:223-236)."""
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    from vermeer_b200.host import HostScene
    sc = scenes.heightfield_scene(48, 32, nq=8)
    sc.camera = scenes.Camera(From=(0, 0, 0), To=(0, 0, -1), Type="Matrix", Fov=50.0, Focal=1.0)
    h = HostScene(sc).prerender()
    o = Oracle(sc)
    assert h.camera_decomp().shape[0] == 0 and o.camera_decomp().shape[0] == 0
    assert np.array_equal(h.camera()[0], np.eye(4, dtype=np.float32).reshape(-1))
    assert np.array_equal(o.camera_matrix()[0], np.eye(4, dtype=np.float32).reshape(-1))
