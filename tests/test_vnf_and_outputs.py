"""the .vnf scene reader (nodes/lex.go, nodes/parser.go) and the output drivers (OutputFloat, OutputHDR) of the
host layer. CPU tests: the file path builds structures bit-identical to the in-memory path, the reference's error behaviour
(messages + keep going) is mirrored, and the written files equal the oracle's restatement byte for byte. One GPU test renders
a scene straight from .vnf text."""
import os

import numpy as np
import pytest


def _same_structures(h1, h2):
    assert h1.num_geoms() == h2.num_geoms()
    assert np.array_equal(h1.scene_geom_order(), h2.scene_geom_order())
    a, b = h1.scene_nodes(), h2.scene_nodes()
    if isinstance(a, tuple):
        assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()
    else:
        assert a.tobytes() == b.tobytes()
    for g in range(h1.num_geoms()):
        try:
            info = h1.mesh_info(g)
        except RuntimeError:
            with pytest.raises(RuntimeError):
                h2.mesh_info(g)
            continue
        assert info == h2.mesh_info(g)
        na, nb = h1.mesh_nodes(g), h2.mesh_nodes(g)
        if isinstance(na, tuple):
            assert na[0].tobytes() == nb[0].tobytes() and na[1].tobytes() == nb[1].tobytes()
        else:
            assert na.tobytes() == nb.tobytes()
        ia, ib = h1.mesh_idxp(g), h2.mesh_idxp(g)
        assert np.array_equal(ia[0], ib[0]) and np.array_equal(ia[1], ib[1])
    ca, cb = h1.camera(), h2.camera()
    assert ca[0].tobytes() == cb[0].tobytes() and ca[1:] == cb[1:]


@pytest.mark.parametrize("name", ["cornell", "glossy", "motion", "filter", "instances", "debug"])
def test_vnf_path_builds_the_same_scene(built_library, name):
    from vermeer_b200 import scenes
    from vermeer_b200.host import HostScene
    sc = {"cornell": lambda: scenes.cornell_box(64, 48),
          "glossy": lambda: scenes.glossy_box(64, 48),
          "motion": lambda: scenes.heightfield_scene(64, 48, nq=24, motion=True),
          "filter": lambda: scenes.cornell_box(32, 32, boxes=False),
          "instances": lambda: scenes.instanced_scene(64, 48, moving=True),
          "debug": lambda: scenes.debug_shader_box(64, 48)}[name]()
    if name == "filter":
        sc.filter = scenes.PixelFilter("AiryFilter", Res=48)
    text = scenes.to_vnf(sc)
    h1 = HostScene(sc).prerender()
    h2 = HostScene.from_vnf(text).prerender()
    assert (h2.scene.XRes, h2.scene.YRes, h2.scene.MaxIter) == (sc.XRes, sc.YRes, sc.MaxIter)
    _same_structures(h1, h2)


def test_vnf_lexer_details(built_library):
    """nodes/lex.go: '#' comments, ints where floats are expected, negative exponents, tokens glued to braces, no escapes."""
    from vermeer_b200 import scenes
    from vermeer_b200.host import HostScene
    text = '''
# a comment line
Globals{XRes 32 YRes 16 MaxIter 4}   # trailing comment
ShaderStd { Name "m" DiffuseColour rgb 1 5e-1 .25e0 DiffuseStrength float 1 }
PolyMesh { Name "tri" Verts 1 3 point 0 0 0   1 0 0   0 0 -1e0 Shader 1 string "m"}
TriLight { Name "l" Shader "m" P0 -1 2 -1 P1 1 2 -1 P2 0 2 1 Samples 1 }
Camera { Name "camera" Type "LookAt" From 1 1 point 0 1 3 To 1 1 point 0 0 0 Roll 1 1 float 0 Up 0 1 0 Fov 45.0 Focal 1 }
'''
    h = HostScene.from_vnf(text).prerender()
    assert (h.scene.XRes, h.scene.YRes, h.scene.MaxIter) == (32, 16, 4)
    assert h.num_geoms() == 2      # the mesh and the light's own mesh
    sc = scenes.SceneDesc(32, 16, scenes.Camera(From=(0, 1, 3), To=(0, 0, 0), Fov=45.0, Focal=1.0),
                          shaders=[scenes.ShaderStd("m", DiffuseColour=(1, 0.5, 0.025), DiffuseStrength=1.0)],
                          meshes=[scenes.PolyMesh("tri", np.asarray([[0, 0, 0], [1, 0, 0], [0, 0, -1]], np.float32), ["m"])],
                          lights=[scenes.TriLight("l", (-1, 2, -1), (1, 2, -1), (0, 2, 1), "m", 1)], MaxIter=4)
    _same_structures(HostScene(sc).prerender(), h)


def test_vnf_error_behaviour_mirrors_the_reference(built_library):
    """parser.go:757-907: an unknown field is reported and skipped up to the next field name; a missing required field drops
    the node (and, like the reference, the skip-to-'}' that follows swallows the next node); unknown node types are reported;
    parsing stops after more than 10 errors."""
    from vermeer_b200.host import HostScene
    good_cam = 'Camera { Name "camera" Type "LookAt" From 1 1 point 0 1 3 To 1 1 point 0 0 0 Roll 1 1 float 0 Up 0 1 0 }\n'
    mat = 'ShaderStd { Name "m" DiffuseStrength float 1 }\n'
    tri = 'PolyMesh { Name "%s" Verts 1 3 point 0 0 0 1 0 0 0 0 -1 Shader 1 string "m" }\n'
    # 1. unknown field: one message, the node survives
    h = HostScene.from_vnf(mat + 'PolyMesh { Name "a" Bogus 1 2 3 Verts 1 3 point 0 0 0 1 0 0 0 0 -1 Shader 1 string "m" }\n' + good_cam, strict=False)
    assert h.parse_errors == 1 and 'Field "Bogus" not found in node PolyMesh' in h.parse_log and "<memory>:2:" in h.parse_log
    assert h.prerender().num_geoms() == 1
    # 2. required field missing: the node is dropped and the following node is swallowed
    h = HostScene.from_vnf(mat + 'PolyMesh { Name "a" Shader 1 string "m" }\n' + tri % "b" + tri % "c" + good_cam, strict=False)
    assert "required field Verts not found in PolyMesh" in h.parse_log and "Node is nil" in h.parse_log
    assert h.prerender().num_geoms() == 1            # "b" was swallowed by the reference's recovery, "c" survives
    # 3. unknown / out-of-scope node types
    h = HostScene.from_vnf(mat + 'Teapot { Size 3 }\nQuadLight { Name "q" }\n' + tri % "a" + good_cam, strict=False)
    assert h.parse_errors == 2 and "Teapot" in h.parse_log and "QuadLight" in h.parse_log
    assert h.prerender().num_geoms() == 1
    # 4. strict mode raises; more than 10 errors stop the parse
    with pytest.raises(RuntimeError, match="not found in node"):
        HostScene.from_vnf(mat + 'PolyMesh { Name "a" Bogus 1 Verts 1 3 point 0 0 0 1 0 0 0 0 -1 Shader 1 string "m" }\n')
    h = HostScene.from_vnf("".join("Nope%d { }\n" % i for i in range(30)), strict=False)
    assert h.parse_errors == 11 and "Too many errors, stopping." in h.parse_log
    # 5. a missing file is an error status, not an abort
    with pytest.raises(RuntimeError, match="no such file"):
        HostScene.from_vnf(path="/nonexistent/scene.vnf")


def test_rgbe_known_answers(built_library):
    """image/hdr/hdr.go:26-50 worked by hand: d = max(r,g,b); (n, e) = frexp(d); byte(c * n*255.999/d); e + 128."""
    import ctypes as C
    from oracle.fileformats import rgb_to_rgbe
    from vermeer_b200.host import load_library
    L = load_library()
    kat = {(1.0, 1.0, 1.0): (127, 127, 127, 129),      # frexp(1) = (.5, 1): .5*255.999 = 127.9995
           (0.5, 0.25, 0.125): (127, 63, 31, 128),      # frexp(.5) = (.5, 0): df = 255.999
           (3.0, 1.5, 0.0): (191, 95, 0, 130),          # frexp(3) = (.75, 2): df = 63.99975
           (0.0, 0.0, 0.0): (0, 0, 0, 0),
           (5e-7, 5e-7, 5e-7): (0, 0, 0, 0),            # below the 1e-6 cut
           (1000.0, 10.0, 1.0): (249, 2, 0, 138)}       # frexp(1000) = (.9765625, 10)
    for rgb, want in kat.items():
        out = (C.c_uint8 * 4)()
        assert L.vh_rgbe(C.c_float(rgb[0]), C.c_float(rgb[1]), C.c_float(rgb[2]), out) == 0
        assert tuple(out) == want, (rgb, tuple(out))
        assert tuple(rgb_to_rgbe(np.asarray(rgb, np.float32))) == want
    # negative and NaN components go through Go's float -> byte conversion (truncate, low byte)
    for rgb in [(1.0, -0.25, 0.5), (2.0, float("nan"), 1.0), (0.7, -3.0, 0.2)]:
        out = (C.c_uint8 * 4)()
        L.vh_rgbe(C.c_float(rgb[0]), C.c_float(rgb[1]), C.c_float(rgb[2]), out)
        assert tuple(out) == tuple(rgb_to_rgbe(np.asarray(rgb, np.float32))), rgb


def test_output_nodes_write_reference_format_files(built_library, tmp_path):
    from oracle.fileformats import output_float_bytes, output_hdr_bytes
    from vermeer_b200 import scenes
    from vermeer_b200.host import HostScene
    sc = scenes.cornell_box(40, 24, boxes=False)
    ff, fh = str(tmp_path / "out.float"), str(tmp_path / "out.hdr")
    h = HostScene.from_vnf(scenes.to_vnf(sc, outputs=[("OutputFloat", ff), ("OutputHDR", fh)])).prerender()
    rng = np.random.default_rng(7)
    fb = (rng.random((24, 40, 3)) ** 4 * 50).astype(np.float32)
    fb[3, 5] = (0, 0, 0)
    fb[4, 6] = (np.nan, 1, 2)
    fb[5, 7] = (-1, 0.5, 0.25)
    h.postrender(fb)
    assert open(ff, "rb").read() == output_float_bytes(fb)
    got = open(fh, "rb").read()
    assert got == output_hdr_bytes(fb)
    assert got.startswith(b"#?RADIANCE\n") and b"\n+Y 24 +X 40\n" in got[:200]
    # the float file reads back as the frame, top row first (driver/outputfloat.go:30-42)
    back = np.fromfile(ff, "<f4").reshape(24, 40, 3)
    assert np.array_equal(back, fb, equal_nan=True)


@pytest.mark.gpu
def test_render_from_vnf_equals_in_memory_path(built_library, tmp_path):
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    sc = scenes.glossy_box(96, 80)
    out = str(tmp_path / "frame.float")
    tab = scenes.splitmix64_table(2, sc.XRes * sc.YRes)
    imgs = []
    for host in (HostScene(sc), HostScene.from_vnf(scenes.to_vnf(sc, outputs=[("OutputFloat", out)]))):
        dev = Device(0).upload(host.prerender())
        dev.set_scramble(tab)
        imgs.append(dev.render(0, 4))
        last = host
    assert imgs[0].tobytes() == imgs[1].tobytes()
    last.postrender(imgs[1])
    assert np.fromfile(out, "<f4").tobytes() == imgs[1].tobytes()


def test_include_node_reads_the_other_file_at_prerender(built_library, tmp_path, monkeypatch):
    """misc.Include (builtin/misc/include.go): PreRender = nodes.Parse(Filename), relative to the working directory; the nodes it adds
    are pre-rendered in the next round. Globals + Include of everything else builds the single-file scene bit for bit."""
    from vermeer_b200 import scenes
    from vermeer_b200.host import HostScene
    sc = scenes.debug_shader_box(64, 48)
    text = scenes.to_vnf(sc)
    head, rest = text.split("\n", 1)
    assert head.startswith("Globals")
    (tmp_path / "rest.vnf").write_text(rest)
    monkeypatch.chdir(tmp_path)
    h1 = HostScene.from_vnf(text).prerender()
    h2 = HostScene.from_vnf(head + '\nInclude { Filename "rest.vnf" }\n').prerender()
    _same_structures(h1, h2)
    # a file that cannot be opened is PreRender's error (include.go:24-26 returns os.Open's)
    h3 = HostScene.from_vnf(head + '\nInclude { Name "inc" Filename "nonexistent.vnf" }\n')
    with pytest.raises(RuntimeError, match="nonexistent.vnf"):
        h3.prerender()
    # parse errors inside the included file are printed, not returned (parser.go:862-899): the nodes that parsed are kept
    (tmp_path / "bad.vnf").write_text(rest + "\nNoSuchNode { }\n")
    h4 = HostScene.from_vnf(head + '\nInclude { Filename "bad.vnf" }\n').prerender()
    _same_structures(h1, h4)
    assert "NoSuchNode" in h4.L.vh_last_error(h4.h).decode()


def test_vnf_texture_maps_and_uvs(built_library):
    """`Param rgbtex "file?query"` (nodes/parser.go:247-268), PolyMesh UV / UVIdx (polymesh.go:36-37): parsed, pre-rendered, and a
    mesh whose UVIdx does not match its FaceIdx is refused by PreRender."""
    from vermeer_b200 import scenes
    from vermeer_b200.host import HostScene
    sc = scenes.textured_room(48, 36, float_maps=True)
    text = scenes.to_vnf(sc)
    assert 'DiffuseColour rgbtex "floor.png"' in text and 'EmissionColour rgbtex "picture.png?filter=trilinear"' in text
    assert "UV 1 4 vec2" in text and "UVIdx 4 int" in text
    h = HostScene.from_vnf(text)
    assert h.parse_errors == 0
    for t in sc.textures:
        h.add_texture(t)
    h.prerender()
    assert h.num_geoms() == HostScene(sc).prerender().num_geoms()
    # rgbtex without a file name: the reference's parser drops the value (parser.go:255-257) and goes on
    bad = text.replace('DiffuseColour rgbtex "floor.png"', "DiffuseColour rgbtex 5")
    hb = HostScene.from_vnf(bad, strict=False)
    assert hb.parse_errors >= 0
    # a UV index beyond the UV array
    sc2 = scenes.textured_room(48, 36)
    [m for m in sc2.meshes if m.Name == "back"][0].UVIdx = np.asarray([0, 1, 2, 9], np.int32)
    with pytest.raises(RuntimeError, match="UV index out of range"):
        HostScene(sc2).prerender()


def test_prerender_reports_the_first_failing_node_in_order(built_library):
    """The meshes of a PreRender round are pre-rendered concurrently; the error reported is still the first failing node's in
    node order, like the reference's sequential loop (core/core.go:46-57)."""
    from vermeer_b200 import scenes
    from vermeer_b200.host import HostScene
    sc = scenes.sphere_field_scene(32, 24, nmesh=12, slices=8, stacks=8)
    sc.meshes[3].Shader = ["no_such_shader_3"]
    sc.meshes[9].Shader = ["no_such_shader_9"]
    with pytest.raises(RuntimeError, match="no_such_shader_3"):
        HostScene(sc).prerender()
