"""Oracle of the texture subsystem (oracle/texture.h: texture/mipmap.go, texture.go, feline.go) against known answers worked by
hand / re-derived independently in numpy.  The reference has no tests under texture/, so these pin the restatement."""
import numpy as np
import pytest


def _oracle_with(textures):
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    sc = scenes.cornell_box(8, 8, boxes=False)
    sc.textures = [scenes.Texture(n, p) for n, p in textures]
    return Oracle(sc)


def _coords(U, V, dudx=0.0, dvdx=0.0, dudy=0.0, dvdy=0.0, pd=(1.0, 1.0)):
    return np.asarray([[U, V, dudx, dvdx, dudy, dvdy, pd[0], pd[1]]], np.float32)


def test_pyramid_level_count_and_box_filter(oracle_lib):
    rng = np.random.default_rng(5)
    img4 = rng.integers(0, 256, (4, 4, 3), dtype=np.uint8)
    img2 = rng.integers(0, 256, (2, 2, 3), dtype=np.uint8)
    img16 = rng.integers(0, 256, (8, 16, 3), dtype=np.uint8)
    o = _oracle_with([("t4", img4), ("t2", img2), ("t16", img16)])
    # maxlevel = ceil(log2(max(w,h))) levels, level 0 included: the pyramid stops at 2 texels, not 1 (mipmap.go:122-127)
    assert [l.shape for l in o.texture_levels("t2")] == [(2, 2, 3)]
    assert [l.shape for l in o.texture_levels("t4")] == [(4, 4, 3), (2, 2, 3)]
    assert [l.shape for l in o.texture_levels("t16")] == [(8, 16, 3), (4, 8, 3), (2, 4, 3), (1, 2, 3)]
    # stored bottom-up (texture.go:139)
    L = o.texture_levels("t4")
    assert np.array_equal(L[0], img4[::-1])
    a = L[0].astype(np.float32)
    exp = (0.25 * (a[0::2, 0::2] + a[1::2, 0::2] + a[0::2, 1::2] + a[1::2, 1::2])).astype(np.uint8)   # byte() truncates
    assert np.array_equal(L[1], exp)
    # worked by hand: texels 10, 20, 30, 41 -> 0.25 * 101 = 25.25 -> 25
    hand = np.zeros((2, 2, 3), np.uint8)
    hand[0, 0], hand[0, 1], hand[1, 0], hand[1, 1] = 10, 20, 30, 41
    big = np.zeros((4, 4, 3), np.uint8)
    big[0:2, 0:2] = hand
    o2 = _oracle_with([("h", big)])
    assert int(o2.texture_levels("h")[1][1, 0, 0]) == 25     # image rows 0-1 are the TOP: bottom-up row 1


def test_pyramid_odd_sizes_polyphase_weights(oracle_lib):
    # 3x3 -> 2x2 takes the even/even (box) branch because the NEW size is even (mipmap.go:147-148): x1 = min(2x+1, 2)
    img = np.arange(27, dtype=np.uint8).reshape(3, 3, 3) * 9
    o = _oracle_with([("t3", img), ("t5", np.full((5, 5, 3), 200, np.uint8)), ("t6", np.full((6, 6, 3), 77, np.uint8))])
    a = o.texture_levels("t3")[0].astype(np.float32)
    L1 = o.texture_levels("t3")[1]
    for y in range(2):
        for x in range(2):
            x0, x1, y0, y1 = 2 * x, min(2 * x + 1, 2), 2 * y, min(2 * y + 1, 2)
            exp = (np.float32(0.25) * (a[y0, x0] + a[y1, x0] + a[y0, x1] + a[y1, x1])).astype(np.uint8)
            assert np.array_equal(L1[y, x], exp)
    # 5x5 -> 3x3 (odd/odd, 9 taps) and 6x6 -> 3x3 (odd new size from an even one): the three weights sum to 1 on each axis, so a
    # constant image stays constant up to the truncation of a product that rounds just below the integer
    for name, c in (("t5", 200), ("t6", 77)):
        for lvl in o.texture_levels(name)[1:]:
            assert np.abs(lvl.astype(int) - c).max() <= 1
    # one 9-tap value by the formula of mipmap.go:300-302, float32 in the same order: 5x5 ramp, output (1,1)
    ramp = np.zeros((5, 5, 3), np.uint8)
    ramp[..., 0] = np.arange(5, dtype=np.uint8)[None, :] * 10 + np.arange(5, dtype=np.uint8)[:, None] * 3
    o3 = _oracle_with([("r", ramp)])
    src = o3.texture_levels("r")[0][..., 0].astype(np.float32)
    f = np.float32
    w0, w1, w2 = f(3 - 1 - 1) / f(5), f(3) / f(5), f(1) / f(5)
    x0, x1, x2, y0, y1, y2 = 1, 2, 3, 1, 2, 3
    row = lambda yy: w0 * src[yy, x0] + w1 * src[yy, x1] + w2 * src[yy, x2]
    exp = int(np.uint8(w0 * row(y0) + w1 * row(y1) + w2 * row(y2)))
    assert int(o3.texture_levels("r")[1][1, 1, 0]) == exp


def test_reference_cannot_load_wide_or_single_texel_images(oracle_lib):
    # 1x1: stdfilter indexes mipmap[0] of an empty slice (mipmap.go:127); 16x2: a level filtered from a 1-row level reads row 1
    for shape in ((1, 1, 3), (2, 16, 3)):
        with pytest.raises(RuntimeError):
            _oracle_with([("bad", np.zeros(shape, np.uint8))])
    _oracle_with([("ok", np.zeros((16, 2, 3), np.uint8))])   # tall is fine: the x axis has a wrap rule


def test_bilinear_taps_and_wrap(oracle_lib):
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (4, 4, 3), dtype=np.uint8)
    o = _oracle_with([("t", img)])
    tex = img[::-1].astype(np.float32)     # storage order
    pd = (1.0, 1.0)
    # integer texel coordinates ARE texel centres: floor == ceil, dx = 0
    for x in range(4):
        for y in range(4):
            c = o.texture_sample("t", _coords(x / 4, y / 4, 0.25, 0, 0, 0.25, pd), trilinear=True) * 255
            assert np.allclose(c[0], tex[y, x], atol=1e-4), (x, y)
    # half way between texel 3 and the wrap-around texel 0
    c = o.texture_sample("t", _coords(3.5 / 4, 0, 0.25, 0, 0, 0.25, pd), trilinear=True) * 255
    assert np.allclose(c[0], 0.5 * tex[0, 3] + 0.5 * tex[0, 0], atol=1e-4)
    # negative / > 1 coordinates wrap (s - floor(s))
    a = o.texture_sample("t", _coords(-0.25, 1.5, 0.25, 0, 0, 0.25, pd), trilinear=True)
    b = o.texture_sample("t", _coords(0.75, 0.5, 0.25, 0, 0, 0.25, pd), trilinear=True)
    assert np.array_equal(a, b)


def test_trilinear_lod_selection(oracle_lib):
    rng = np.random.default_rng(10)
    img = rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)
    o = _oracle_with([("t", img)])
    L = [l.astype(np.float32) for l in o.texture_levels("t")]
    assert len(L) == 4
    # footprint of 4 texels along x (Dduvdx = (4/16, 0)) -> lod = log2(4) = 2 -> level 2 only, texel (1, 2) at its centre
    c = o.texture_sample("t", _coords(1 / 4, 2 / 4, 4 / 16, 0, 0, 1 / 16), trilinear=True) * 255
    assert np.allclose(c[0], L[2][2, 1], atol=1e-4)
    # lod = log2(2*sqrt(2)) = 1.5: dl*level2 + (1-dl)*level1 with dl = .5 (mipmap.go:98-101)
    c = o.texture_sample("t", _coords(0, 0, 2 / 16, 2 / 16, 0, 0), trilinear=True) * 255
    assert np.allclose(c[0], 0.5 * L[2][0, 0] + 0.5 * L[1][0, 0], atol=1e-3)
    # beyond the last level: clamped to MaxLevelOfDetail (3); a zero footprint (log2 0 = -inf) clamps to 0
    c = o.texture_sample("t", _coords(0, 0, 100.0, 0, 0, 0), trilinear=True) * 255
    assert np.allclose(c[0], L[3][0, 0], atol=1e-4)
    c = o.texture_sample("t", _coords(3 / 16, 5 / 16), trilinear=True) * 255
    assert np.allclose(c[0], L[0][5, 3], atol=1e-4)
    # PixelDelta scales the footprint (texture.go:241-242)
    c = o.texture_sample("t", _coords(1 / 4, 2 / 4, 1.0, 0, 0, 0, pd=(4 / 16, 1.0)), trilinear=True) * 255
    assert np.allclose(c[0], L[2][2, 1], atol=1e-4)


def test_feline_isotropic_is_one_trilinear_probe(oracle_lib):
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)
    o = _oracle_with([("t", img)])
    for r in (1.0, 2.0, 3.0, 0.25):
        co = _coords(0.3, 0.6, r / 16, 0, 0, r / 16)
        f = o.texture_sample("t", co, trilinear=False)
        t = o.texture_sample("t", co, trilinear=True)
        assert np.allclose(f, t, atol=1e-6), r


def test_feline_four_to_one_footprint(oracle_lib):
    """Ellipse 4 x 1 texels along u: fProbes = 2*4/1 - 1 = 7 probes one texel apart on level 0, Gaussian weights
    exp(-0.6 * (n/2)^2 / 16) (feline.go:86-137).  Re-derived here in float64 from the paper's formulas."""
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, (16, 16, 3), dtype=np.uint8)
    o = _oracle_with([("t", img)])
    tex = img[::-1].astype(np.float64)
    U, V = 7 / 16, 5 / 16
    got = o.texture_sample("t", _coords(U, V, 4 / 16, 0, 0, 1 / 16), trilinear=False)[0] * 255
    acc, wsum = np.zeros(3), 0.0
    for k in range(-3, 4):
        wgt = np.exp(-0.6 * (k * k) / 16.0)
        acc += wgt * tex[5, (7 + k) % 16]
        wsum += wgt
    assert np.allclose(got, acc / wsum, atol=2e-3), (got, acc / wsum)
    # the same ellipse along v
    got = o.texture_sample("t", _coords(U, V, 1 / 16, 0, 0, 4 / 16), trilinear=False)[0] * 255
    acc, wsum = np.zeros(3), 0.0
    for k in range(-3, 4):
        wgt = np.exp(-0.6 * (k * k) / 16.0)
        acc += wgt * tex[(5 + k) % 16, 7]
        wsum += wgt
    assert np.allclose(got, acc / wsum, atol=2e-3), (got, acc / wsum)
    # more than 16 probes wanted: capped, the minor radius widens (feline.go:92-96) -> lod = log2(2*40/17)
    got = o.texture_sample("t", _coords(U, V, 40 / 16, 0, 0, 1 / 16), trilinear=False)
    assert np.isfinite(got).all()


def test_texcoords_of_camera_hits(oracle_lib):
    """U, V and their screen derivatives at the first hit (trace.go:350-502), against a finite difference of the same
    camera: the floor quad of the textured room carries UVs 3x its extent."""
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    sc = scenes.textured_room(64, 48, mirror=False, smooth=False)
    o = Oracle(sc)
    o.set_scramble(scenes.splitmix64_table(1, sc.XRes * sc.YRes))
    tc = o.camera_texcoords(1).reshape(48, 64, 8)
    rays = o.camera_rays(1).reshape(48, 64)
    hits = o.trace(rays.reshape(-1)).reshape(48, 64)
    floor_id = [i for i, m in enumerate(sc.meshes) if m.Name == "floor"][0]
    on_floor = hits["geom"] == floor_id
    assert on_floor.sum() > 200
    # UV = 1.5 * (x + 1), 1.5 * (z + 1) on the floor: check against the hit point
    P = rays["o"] + rays["d"] * hits["t"][..., None]
    assert np.allclose(tc[..., 0][on_floor], 1.5 * (P[..., 0][on_floor] + 1), atol=2e-3)
    assert np.allclose(tc[..., 1][on_floor], 1.5 * (P[..., 2][on_floor] + 1), atol=2e-3)
    # PixelDelta * Dduvdx ~ the UV step between horizontally adjacent pixels (|Sx step| = 2/w; PixelDelta = 2 tan(th) f / w)
    pd0 = tc[..., 6][on_floor][0]
    du = (tc[:, 1:, 0] - tc[:, :-1, 0])
    both = on_floor[:, 1:] & on_floor[:, :-1]
    pred = pd0 * tc[:, :-1, 2]
    rel = np.abs(du[both] - pred[both]) / np.maximum(np.abs(du[both]), 1e-6)
    assert np.median(rel) < 0.2, np.median(rel)     # jittered sample positions: the step is only roughly one pixel
    # meshes without UVs: U, V are the barycentrics
    left_id = [i for i, m in enumerate(sc.meshes) if m.Name == "left"][0]
    on_left = hits["geom"] == left_id
    assert np.allclose(tc[..., 0][on_left], hits["u"][on_left]) and np.allclose(tc[..., 1][on_left], hits["v"][on_left])


def test_textured_render_differs_from_constant(oracle_lib):
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    sc = scenes.textured_room(48, 36)
    tab = scenes.splitmix64_table(1, sc.XRes * sc.YRes)
    o = Oracle(sc)
    o.set_scramble(tab)
    ft, _ = o.render(0, 2, nthreads=4)
    sc2 = scenes.textured_room(48, 36)
    for s in sc2.shaders:
        for name, _, _ in scenes.SHADER_SLOTS:
            if isinstance(getattr(s, name), str) and name != "Spec1FresnelModel":
                setattr(s, name, (0.5, 0.5, 0.5) if "Colour" in name else 0.5)
    o2 = Oracle(sc2)
    o2.set_scramble(tab)
    fc, _ = o2.render(0, 2, nthreads=4)
    ok = np.isfinite(ft).all(-1) & np.isfinite(fc).all(-1)
    assert np.abs(ft[ok] - fc[ok]).max() > 0.05


def test_texture_url_parsing():
    from vermeer_b200.scenes import parse_texture_url
    assert parse_texture_url("a/b.png") == ("a/b.png", 0, False)
    assert parse_texture_url("a.png?filter=trilinear") == ("a.png", 0, True)
    assert parse_texture_url("a.png?ch=2&filter=trilinear") == ("a.png", 2, True)
    assert parse_texture_url("a.png?ch=x") == ("a.png", 0, False)
    assert parse_texture_url("a.png?filter=feline") == ("a.png", 0, False)
