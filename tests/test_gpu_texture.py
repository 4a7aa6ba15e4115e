"""texture maps on the device (vermeer_b200/csrc/texture.cu, texture.cuh, k_surface in render.cu) against the
oracle (oracle/texture.h): the mip pyramid byte for byte, both filters on batches of lookups, and rendered images with
textured ShaderStd parameters, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_with(textures):
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    sc = scenes.cornell_box(8, 8, boxes=False)
    sc.textures = [scenes.Texture(n, p) for n, p in textures]
    return Oracle(sc)


SIZES = [(2, 2), (3, 3), (4, 4), (5, 5), (6, 6), (7, 5), (5, 9), (16, 8), (8, 16), (37, 23), (23, 37), (64, 64), (100, 60), (129, 127), (256, 256), (2, 7)]


def test_pyramid_is_byte_identical(built_library):
    from vermeer_b200.host import Device
    rng = np.random.default_rng(3)
    imgs = [("t%d" % i, rng.integers(0, 256, (h, w, 3), dtype=np.uint8)) for i, (w, h) in enumerate(SIZES)]
    ora = _oracle_with(imgs)
    dev = Device(0)
    for name, img in imgs:
        tid = dev.texture_upload(img[::-1])
        lo, lg = ora.texture_levels(name), dev.texture_levels(tid)
        assert [a.shape for a in lo] == [a.shape for a in lg], name
        for l, (a, b) in enumerate(zip(lo, lg)):
            assert np.array_equal(a, b), "%s %s level %d differs at %d texels" % (name, img.shape, l, (a != b).any(-1).sum())


def test_upload_refuses_what_the_reference_cannot_load(built_library):
    from vermeer_b200.host import Device
    dev = Device(0)
    for shape in ((1, 1, 3), (2, 16, 3), (4, 64, 3)):
        with pytest.raises(RuntimeError):
            dev.texture_upload(np.zeros(shape, np.uint8))
    dev.texture_upload(np.zeros((64, 4, 3), np.uint8))


def _random_coords(n, seed):
    rng = np.random.default_rng(seed)
    c = np.zeros((n, 8), np.float32)
    c[:, 0:2] = rng.uniform(-1.5, 2.5, (n, 2))
    # footprints from far below a texel to larger than the image, isotropic to 40:1, any orientation
    r_major = 10.0 ** rng.uniform(-3.5, 0.3, n)
    ratio = 10.0 ** rng.uniform(0, 1.6, n)
    ang = rng.uniform(0, 2 * np.pi, n)
    ax = np.stack([np.cos(ang), np.sin(ang)], 1) * r_major[:, None]
    bx = np.stack([-np.sin(ang), np.cos(ang)], 1) * (r_major / ratio)[:, None]
    mix = rng.uniform(-1, 1, (n, 2, 2))
    c[:, 2:4] = ax * mix[:, 0, 0:1] + bx * mix[:, 0, 1:2]
    c[:, 4:6] = ax * mix[:, 1, 0:1] + bx * mix[:, 1, 1:2]
    c[:, 6] = 1.0
    c[:, 7] = rng.uniform(0.5, 1.5, n)
    # a few exact cases: texel centres, zero footprints, axis-aligned ellipses
    c[0] = [0.25, 0.5, 0, 0, 0, 0, 1, 1]
    c[1] = [0.25, 0.5, 4 / 64, 0, 0, 1 / 64, 1, 1]
    c[2] = [0.25, 0.5, 1 / 64, 0, 0, 1 / 64, 1, 1]
    return c


@pytest.mark.parametrize("trilinear", [False, True])
@pytest.mark.parametrize("size", [(64, 64), (37, 23), (256, 128)])
def test_filters_match_oracle(built_library, trilinear, size):
    from vermeer_b200.host import Device
    rng = np.random.default_rng(size[0])
    img = rng.integers(0, 256, (size[1], size[0], 3), dtype=np.uint8)
    ora = _oracle_with([("t", img)])
    dev = Device(0)
    tid = dev.texture_upload(img[::-1])
    co = _random_coords(20000, 17)
    o = ora.texture_sample("t", co, trilinear=trilinear)
    g = dev.texture_sample(tid, co, trilinear=trilinear)
    both_nan = np.isnan(o) & np.isnan(g)
    assert (np.isnan(o) == np.isnan(g)).all()        # degenerate footprints (parallel derivatives): NaN in both, like the reference
    d = np.abs(np.where(both_nan, 0, o - g))
    # the only non-IEEE steps are log2 / atan / cos / sin / exp through two different double-precision libms
    assert d.max() <= 2e-5, (d.max(), co[np.unravel_index(d.argmax(), d.shape)[0]])
    same = ((o.view(np.uint32) == g.view(np.uint32)) | both_nan).all(1).mean()
    assert same >= 0.98, same


def _render_pair(sc, iters, seed=1, nthreads=8, **opts):
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    tab = scenes.splitmix64_table(seed, sc.XRes * sc.YRes)
    ora = Oracle(sc)
    ora.set_scramble(tab)
    fo, so = ora.render(0, iters, nthreads=nthreads)
    dev = Device(0).upload(HostScene(sc).prerender())
    dev.set_scramble(tab)
    for k, v in opts.items():
        dev.set_option(k, v)
    fg = dev.render(0, iters)
    return fo, so, fg, dev.stats(), dev


def _rmse(fo, fg):
    ok = np.isfinite(fo).all(-1) & np.isfinite(fg).all(-1)
    return float(np.sqrt(((fo[ok] - fg[ok]) ** 2).mean())), ok


def test_emissive_textures_show_the_filtered_texels(built_library):
    """Every textured surface emits its map and nothing else lights the room: a pixel is the filter's output times N/(N+1), so this
    compares U, V, the transferred differentials and both filters per pixel, with no light sampling in between."""
    from vermeer_b200 import scenes
    sc = scenes.textured_room(160, 120, mirror=False, smooth=False)
    for s in sc.shaders:
        if isinstance(s.DiffuseColour, str):
            s.EmissionColour, s.EmissionStrength, s.DiffuseColour = s.DiffuseColour, 1.0, (0.0, 0.0, 0.0)
    sc.lights = []
    fo, so, fg, st, _ = _render_pair(sc, 3)
    rmse, ok = _rmse(fo, fg)
    assert ok.all()
    assert fo.max() > 0.3
    assert rmse <= 1e-4, rmse
    # pixels whose sample straddles a mesh edge can land on the other mesh (normalize of the camera direction is not
    # bit-reproducible); everywhere else the lookups agree to the last bits
    assert np.quantile(np.abs(fo - fg), 0.99) <= 5e-6


@pytest.mark.parametrize("variant", ["plain", "mirror", "smooth+mirror", "float_maps"])
def test_textured_room_image(built_library, variant):
    from vermeer_b200 import scenes
    sc = scenes.textured_room(160, 120, mirror="mirror" in variant, smooth="smooth" in variant, float_maps=variant == "float_maps")
    fo, so, fg, st, _ = _render_pair(sc, 6)
    rmse, ok = _rmse(fo, fg)
    assert ok.mean() > 0.99
    assert rmse <= 1e-3, rmse
    assert np.median(np.abs(fo - fg)[ok]) <= 2e-6
    assert abs(st["rays"] - so["rays"]) <= 2e-3 * so["rays"]


def test_generic_kernel_and_precise_trig_agree(built_library):
    from vermeer_b200 import scenes
    sc = scenes.textured_room(128, 96)
    fo, so, fg, st, _ = _render_pair(sc, 4, generic_shade=1, precise_trig=1)
    rmse, ok = _rmse(fo, fg)
    assert rmse <= 1e-3, rmse


def test_vnf_path_renders_the_same_image(built_library):
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    sc = scenes.textured_room(96, 72)
    tab = scenes.splitmix64_table(2, sc.XRes * sc.YRes)
    d1 = Device(0).upload(HostScene(sc).prerender())
    d1.set_scramble(tab)
    f1 = d1.render(0, 3)
    h2 = HostScene.from_vnf(scenes.to_vnf(sc))
    for t in sc.textures:
        h2.add_texture(t)
    d2 = Device(0).upload(h2.prerender())
    d2.set_scramble(tab)
    f2 = d2.render(0, 3)
    assert np.array_equal(f1.view(np.uint32), f2.view(np.uint32))


def test_errors(built_library):
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    # a map whose file was never registered
    sc = scenes.textured_room(32, 24)
    sc.textures = sc.textures[:1]
    with pytest.raises(RuntimeError, match="was not registered"):
        Device(0).upload(HostScene(sc).prerender())
    # texture maps next to a motion mesh
    sc = scenes.heightfield_scene(32, 24, nq=8, motion=True)
    sc.textures = [scenes.Texture("t.png", np.zeros((4, 4, 3), np.uint8))]
    sc.shaders[0].DiffuseColour = "t.png"
    dev = Device(0).upload(HostScene(sc).prerender())
    dev.set_scramble(scenes.splitmix64_table(1, 32 * 24))
    with pytest.raises(RuntimeError, match="static PolyMeshes"):
        dev.render(0, 1)
    # a texture on a light's emission
    sc = scenes.textured_room(32, 24)
    [s for s in sc.shaders if s.Name == "lightmtl"][0].EmissionColour = "wall.png"
    dev = Device(0).upload(HostScene(sc).prerender())
    dev.set_scramble(scenes.splitmix64_table(1, 32 * 24))
    with pytest.raises(RuntimeError, match="light"):
        dev.render(0, 1)


def test_cooperative_lookups_are_bit_identical(built_library):
    """k_surface<COOP>: the probes of a warp's lookups shared between its lanes against one lane per lookup."""
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    sc = scenes.textured_room(160, 120, float_maps=True)
    tab = scenes.splitmix64_table(3, sc.XRes * sc.YRes)
    imgs = []
    for coop in (0, 1):
        dev = Device(0).upload(HostScene(sc).prerender())
        dev.set_scramble(tab)
        dev.set_option("texture_coop", coop)
        imgs.append(dev.render(0, 4))
    assert np.array_equal(imgs[0].view(np.uint32), imgs[1].view(np.uint32))


def test_tile_partitions_of_a_textured_scene_tile_the_frame(built_library):
    """Two contexts rendering the interleaved tiles of ranks 0 and 1 (what two GPUs do) reproduce the single-context frame bit
    for bit: texture lookups and ray differentials do not depend on which pixels a context owns."""
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    sc = scenes.textured_room(160, 128, float_maps=True)
    tab = scenes.splitmix64_table(5, sc.XRes * sc.YRes)
    host = HostScene(sc).prerender()
    one = Device(0).upload(host)
    one.set_scramble(tab)
    full = one.render(0, 3)
    acc = np.zeros_like(full)
    for rank in range(2):
        d = Device(0).upload(host)
        d.set_partition(rank, 2)
        d.set_scramble(tab)
        acc += d.render(0, 3)
    assert np.array_equal(np.nan_to_num(acc).view(np.uint32), np.nan_to_num(full).view(np.uint32))


def _instanced_cards(xres=160, yres=120, emissive=True, moving=False):
    """A textured card (quad with UVs) and a textured, smooth-shaded mirror ball, each placed twice more through GeomInstances
    (rotated / scaled; one instance of an instance), over the textured floor of the room."""
    from vermeer_b200 import scenes
    sc = scenes.textured_room(xres, yres, mirror=False, smooth=False)
    sc.meshes = [m for m in sc.meshes if m.Name in ("floor", "back", "left", "right", "ceiling")]
    card = scenes._quad("card", [[-0.25, 0.05, 0.0], [0.25, 0.05, 0.0], [0.25, 0.55, 0.0], [-0.25, 0.55, 0.0]], "odd_tex")
    card.UV = np.asarray([[0, 0], [2, 0], [2, 2], [0, 2]], np.float32)
    sc.meshes.append(card)
    v, t = scenes._uv_sphere(12, 8)
    sc.shaders.append(scenes.ShaderStd("ballmtl", DiffuseColour="wall.png", DiffuseStrength=0.5, Spec1Colour=(0.9, 0.9, 0.9), Spec1Strength=0.5, Spec1Roughness=0.0))
    ball = scenes.PolyMesh("ball", (v * np.float32(0.16) + np.asarray([0.0, 0.2, 0.0], np.float32)).astype(np.float32), ["ballmtl"], FaceIdx=t.copy(), Normals=v.copy())
    sc.meshes.append(ball)

    def inst(name, geom, mats, verts):
        lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
        for m in mats:
            M = m.reshape(4, 4).T.astype(np.float64)
            w = verts.astype(np.float64) @ M[:3, :3].T + M[:3, 3]
            lo, hi = np.minimum(lo, w.min(0)), np.maximum(hi, w.max(0))
        return scenes.GeomInstance(name, geom, np.stack(mats, 0), tuple((lo - 0.02).astype(np.float32)), tuple((hi + 0.02).astype(np.float32)))
    cv, bv = card.Verts[0], ball.Verts[0]
    m1 = [scenes.srt_matrix((-0.55, 0.1, 0.3), 40.0, 1.3)] + ([scenes.srt_matrix((-0.45, 0.2, 0.3), 75.0, 1.3)] if moving else [])
    m2 = [scenes.srt_matrix((0.55, 0.0, 0.45), -30.0, 0.8)]
    m3 = [scenes.srt_matrix((0.5, 0.05, -0.35), 15.0, 1.5)]
    outer = scenes.srt_matrix((-0.1, 0.45, 0.1), 25.0, 0.9)
    comp = scenes.matrix4(outer.reshape(4, 4).T.astype(np.float64) @ m2[0].reshape(4, 4).T.astype(np.float64))
    sc.instances = [inst("card1", "card", m1, cv), inst("card2", "card", m2, cv), inst("ball1", "ball", m3, bv),
                    inst("card3", "card2", [outer], (cv.astype(np.float64) @ m2[0].reshape(4, 4).T[:3, :3].T.astype(np.float64) + m2[0].reshape(4, 4).T[:3, 3]).astype(np.float32))]
    del comp
    if emissive:
        for s in sc.shaders:
            if isinstance(s.DiffuseColour, str) and s.Name != "ballmtl":
                s.EmissionColour, s.EmissionStrength, s.DiffuseColour = s.DiffuseColour, 1.0, (0.0, 0.0, 0.0)
        sc.lights = []
        sc.meshes = [m for m in sc.meshes if m.Name != "ball"]
        sc.instances = [i for i in sc.instances if i.Geom != "ball"]
    return sc


def test_texture_footprints_through_instances(built_library):
    """a textured mesh seen through GeomInstances. PolyMesh.TraceElems runs with the ray in OBJECT space
    (instance.go:86-95) and transfers the untouched world-space ray differentials with the object-space direction
    (trace.go:360); k_surface mirrors that. All-emissive: a pixel is the filter's output, so U, V and the footprint are
    compared with no light sampling in between."""
    sc = _instanced_cards(emissive=True)
    fo, so, fg, st, _ = _render_pair(sc, 3)
    rmse, ok = _rmse(fo, fg)
    assert ok.all() and fo.max() > 0.3
    assert rmse <= 1e-4, rmse
    assert np.quantile(np.abs(fo - fg), 0.99) <= 5e-6
    # the instanced cards are in the picture: the frame differs from the same room without them
    sc0 = _instanced_cards(emissive=True)
    sc0.instances = []
    f0 = _render_pair(sc0, 3)[2]
    assert (np.abs(f0 - fg).max(-1) > 0.05).mean() > 0.03


@pytest.mark.parametrize("moving", [False, True])
def test_textured_instances_lit_with_mirror_ball(built_library, moving):
    """The same placements lit by the TriLights, with an instanced smooth-shaded mirror ball: the reflected ray's differentials
    (Ray.Init, core/ray.go:72-87) use the context's normal AFTER ApplyTransform with the instance's matrix, and the texture the
    mirror shows is filtered with them."""
    sc = _instanced_cards(emissive=False, moving=moving)
    fo, so, fg, st, _ = _render_pair(sc, 6)
    rmse, ok = _rmse(fo, fg)
    assert ok.mean() > 0.99
    assert rmse <= 1e-3, rmse
    assert np.median(np.abs(fo - fg)[ok]) <= 2e-6
    assert abs(st["rays"] - so["rays"]) <= 2e-3 * so["rays"]
