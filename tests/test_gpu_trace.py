"""Parity of the CUDA traversal (vg_trace_batch through the C ABI) against the oracle on identical rays.

Bar (BASELINE.json north_star): primitive id and hit/miss bit-exact, t bit-exact (the traversal arithmetic is
sub/mul/add/div only, so parity mode is expected to be exact, not just within 1e-5)."""
import numpy as np
import pytest

from conftest import assert_hits_equal, random_rays

pytestmark = pytest.mark.gpu


def _pair(scene, motion_ref_compat=False):
    from oracle.binding import Oracle
    from vermeer_b200.host import Device, HostScene
    ora = Oracle(scene, motion_ref_compat=motion_ref_compat)
    host = HostScene(scene).prerender()
    dev = Device(0).upload(host, motion_ref_compat=motion_ref_compat)
    return ora, dev


@pytest.mark.parametrize("any_hit", [False, True])
def test_cornell_random_rays(built_library, any_hit):
    from vermeer_b200 import scenes
    ora, dev = _pair(scenes.cornell_box(64, 64))
    rays = random_rays(20000, 11, lo=(-0.99, 0.01, -0.99), hi=(0.99, 1.98, 0.99), tmax=(1.5 if any_hit else np.inf))
    g = dev.trace(rays, any_hit=any_hit)
    o = ora.trace(rays, any_hit=any_hit)
    assert_hits_equal(g, o, what="cornell any_hit=%s" % any_hit)
    assert (g["prim"] >= 0).mean() > 0.3


def test_cornell_camera_rays(built_library):
    from vermeer_b200 import scenes
    sc = scenes.cornell_box(96, 96)
    ora, dev = _pair(sc)
    ora.set_scramble(scenes.splitmix64_table(1, 96 * 96))
    rays = ora.camera_rays(1)
    assert_hits_equal(dev.trace(rays), ora.trace(rays), what="cornell camera")


@pytest.mark.parametrize("nq", [40, 300])
def test_heightfield_primary_and_incoherent(built_library, nq):
    from vermeer_b200 import scenes
    sc = scenes.heightfield_scene(128, 72, nq=nq)
    ora, dev = _pair(sc)
    ora.set_scramble(scenes.splitmix64_table(2, 128 * 72))
    rays = ora.camera_rays(3)
    g = dev.trace(rays)
    o = ora.trace(rays)
    assert_hits_equal(g, o, what="heightfield primary")
    assert (g["prim"] >= 0).mean() > 0.5
    inc = scenes.incoherent_rays(rays, g, seed=5)
    assert_hits_equal(dev.trace(inc), ora.trace(inc), what="heightfield incoherent")
    sh = inc.copy()
    sh["tmax"] = 0.7
    assert_hits_equal(dev.trace(sh, any_hit=True), ora.trace(sh, any_hit=True), what="heightfield any-hit")


def test_sphere_field_two_level(built_library):
    from vermeer_b200 import scenes
    sc = scenes.sphere_field_scene(64, 64, nmesh=25, slices=16, stacks=17)
    ora, dev = _pair(sc)
    rays = random_rays(30000, 3, lo=(-1.1, 0.0, -1.1), hi=(1.1, 1.5, 1.1))
    assert_hits_equal(dev.trace(rays), ora.trace(rays), what="sphere field")
    sh = rays.copy()
    sh["tmax"] = 0.5
    assert_hits_equal(dev.trace(sh, any_hit=True), ora.trace(sh, any_hit=True), what="sphere field any-hit")


@pytest.mark.parametrize("ref_compat", [False, True])
def test_motion_heightfield(built_library, ref_compat):
    from vermeer_b200 import scenes
    sc = scenes.heightfield_scene(64, 64, nq=60, motion=True)
    ora, dev = _pair(sc, motion_ref_compat=ref_compat)
    rays = random_rays(20000, 4, lo=(-1.0, 0.3, -1.0), hi=(1.0, 1.4, 1.0))
    rays["d"][:, 1] = -np.abs(rays["d"][:, 1])  # look down at the field
    g = dev.trace(rays)
    o = ora.trace(rays)
    assert_hits_equal(g, o, what="motion ref_compat=%s" % ref_compat)
    if not ref_compat:
        assert (g["prim"] >= 0).mean() > 0.2
    # ref_compat reproduces reference quirk (b): leaf boxes bound face accel.idx[i] but face i is tested, so almost
    # every ray misses — parity with the oracle above is what matters there.


def test_degenerate_rays(built_library):
    """Axis-parallel directions (Dinv = +-Inf), origins on box planes (0*Inf = NaN in the slab test), zero-length and
    NaN rays: the x86 MINPS/MAXPS operand semantics decide hit/miss here (qbvh/intersect_amd64.s)."""
    from vermeer_b200 import scenes
    from vermeer_b200.host import RAY_DTYPE
    sc = scenes.cornell_box(32, 32)
    ora, dev = _pair(sc)
    rays = []
    for o in [(0, 1, 0), (0.1, 0.0, 0.0), (-1, 1, -1), (1, 2, 1), (0.7, 0.6, 0.6), (0.1, 0.3, 0.0), (-0.7, 1.2, -0.6), (0.25, 1.99, 0.0)]:
        for d in [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (1, 1, 0), (0, 1, 1), (-1, 0, 1), (1, -1, 1)]:
            rays.append((o, d, np.inf, 0.0))
    rays.append(((0, 1, 0), (0, 0, 0), np.inf, 0.0))
    rays.append(((0, 1, 0), (np.nan, 0, 1), np.inf, 0.0))
    rays.append(((0, 1, 0), (0, 0, -1), 0.0, 0.0))
    r = np.zeros(len(rays), RAY_DTYPE)
    for i, (o, d, tm, ti) in enumerate(rays):
        r[i] = (o, d, tm, ti)
    g, o = dev.trace(r), ora.trace(r)
    # t of a miss is tmax (may be inf/nan-free); compare everything bitwise
    assert_hits_equal(g, o, what="degenerate")


def test_empty_and_ragged_batches(built_library):
    from vermeer_b200 import scenes
    sc = scenes.cornell_box(32, 32)
    ora, dev = _pair(sc)
    assert len(dev.trace(random_rays(0, 1))) == 0
    for n in (1, 31, 33, 127, 129, 1000):
        rays = random_rays(n, n, lo=(-0.9, 0.1, -0.9), hi=(0.9, 1.9, 0.9))
        assert_hits_equal(dev.trace(rays), ora.trace(rays), what="n=%d" % n)


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_traversal_variants(built_library, variant):
    """The three persistent-kernel variants (0 per-lane while-while, 1 the same over the cp.async.bulk/TMA-staged ray queue,
    2 warp-cooperative leaves = default) give the same bits as the oracle, counters included."""
    from vermeer_b200 import scenes
    sc = scenes.heightfield_scene(128, 72, nq=64)
    ora, dev = _pair(sc)
    dev.set_option("traversal", variant)
    for n in (1, 33, 1000, 50001):
        rays = random_rays(n, n, lo=(-1, 0.2, -1), hi=(1, 1.2, 1))
        rays["d"][:, 1] = -np.abs(rays["d"][:, 1])
        assert_hits_equal(dev.trace(rays), ora.trace(rays), what="variant %d n=%d" % (variant, n))
        sh = rays.copy()
        sh["tmax"] = 0.6
        assert_hits_equal(dev.trace(sh, any_hit=True), ora.trace(sh, any_hit=True), what="variant %d any-hit n=%d" % (variant, n))
    sc = scenes.sphere_field_scene(64, 64, nmesh=25, slices=16, stacks=17)
    ora, dev = _pair(sc)
    dev.set_option("traversal", variant)
    rays = random_rays(30000, 3, lo=(-1.1, 0.0, -1.1), hi=(1.1, 1.5, 1.1))
    assert_hits_equal(dev.trace(rays), ora.trace(rays), what="variant %d sphere field" % variant)


def test_sphere_geom_leaf_is_bit_exact(built_library):
    """The analytic Sphere geom a SphereLight adds to the scene-level tree (builtin/geom/sphere/trace.go): closest-hit and
    any-hit batches through every traversal variant agree with the oracle bit for bit, sphere hits included."""
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    from conftest import assert_hits_equal, random_rays
    sc = scenes.glossy_box(64, 64)
    ora = Oracle(sc)
    ora.set_scramble(scenes.splitmix64_table(1, 64 * 64))
    host = HostScene(sc).prerender()
    dev = Device(0).upload(host)
    rays = np.concatenate([ora.camera_rays(1), random_rays(40000, 11, lo=(-0.9, 0.1, -0.9), hi=(0.9, 1.9, 0.9))])
    # aim a third of the random rays at the sphere light so that the analytic leaf is hit often
    n = len(rays)
    tgt = np.asarray([0.0, 1.3, 0.55], np.float32) + np.random.default_rng(5).normal(scale=0.08, size=(n, 3)).astype(np.float32)
    aim = np.arange(n) % 3 == 0
    d = tgt - rays["o"]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["d"][aim] = d[aim].astype(np.float32)
    o = ora.trace(rays)
    sphere_geom = int(o["geom"][(o["prim"] == 0) & (o["u"] == 0) & (o["v"] == 0) & (o["geom"] >= 0)].max())
    assert (o["geom"] == sphere_geom).sum() > 1000
    for variant in (0, 1, 2):
        dev.set_option("traversal", variant)
        assert_hits_equal(dev.trace(rays), o, what="sphere scene closest variant %d" % variant)
        sh = rays.copy()
        sh["tmax"] = 2.5
        og, gg = ora.trace(sh, any_hit=True), dev.trace(sh, any_hit=True)
        assert np.array_equal(og["prim"] >= 0, gg["prim"] >= 0), "any-hit occlusion differs (variant %d)" % variant


def test_nested_instances_trace_parity(built_library):
    """An instance of an instance (and one three transforms deep): Instance.Trace calls ins.geom.Trace, whatever Geom that is
    (instance.go:95), so the inverses are applied one after the other and the hit reports the OUTERMOST instance's geom id (the
    scene-level leaf, scene.go:65). Static transforms: bit-exact incl. the counters."""
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    from conftest import assert_hits_equal, random_rays
    sc = scenes.instanced_scene(96, 72, moving=False, nested=True)
    ora = Oracle(sc, motion_ref_compat=False)
    ora.set_scramble(scenes.splitmix64_table(1, 96 * 72))
    dev = Device(0).upload(HostScene(sc).prerender())
    rays = np.concatenate([ora.camera_rays(1), random_rays(40000, 5, lo=(-2.0, 0.05, -1.5), hi=(1.6, 1.2, 1.6))])
    o = ora.trace(rays)
    nest1, nest2 = 2 + 3, 2 + 4                      # geoms: floor, blob, inst0..2, nest1, nest2 (+ the lights' meshes after)
    assert (o["geom"] == nest1).sum() > 300 and (o["geom"] == nest2).sum() > 100
    assert_hits_equal(dev.trace(rays), o, what="nested instances")
    sh = rays.copy()
    sh["tmax"] = 3.0
    assert np.array_equal(ora.trace(sh, any_hit=True)["prim"] >= 0, dev.trace(sh, any_hit=True)["prim"] >= 0)


@pytest.mark.parametrize("kind", ["static", "moving", "moving_motion_base"])
def test_instances_trace_parity(built_library, kind):
    """GeomInstance (builtin/geom/instance/instance.go): the ray is taken into object space with the inverse of the
    SRT-interpolated transform, re-Setup, and traced through the target mesh's tree; a hit reports the instance's geom id.
    Single-key transforms are bit-exact (the matrices come out of the same float32 operations as the oracle's); with motion
    keys the per-ray slerp goes through acos/sin of a different libm, so t is held to the 1e-5 relative bar instead."""
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    from conftest import assert_hits_equal, random_rays
    sc = scenes.instanced_scene(96, 72, moving=kind != "static", motion_base=kind == "moving_motion_base")
    ora = Oracle(sc, motion_ref_compat=False)
    ora.set_scramble(scenes.splitmix64_table(1, 96 * 72))
    dev = Device(0).upload(HostScene(sc).prerender())
    rays = np.concatenate([ora.camera_rays(1), ora.camera_rays(2), random_rays(30000, 3, lo=(-1.2, 0.05, -1.2), hi=(1.2, 1.2, 1.2))])
    o = ora.trace(rays)
    assert (o["geom"] >= 2).sum() > 2000          # plenty of hits on the three instances (geoms 2, 3, 4)
    g = dev.trace(rays)
    if kind == "static":
        assert_hits_equal(g, o, what="instances " + kind)
    else:
        same = (g["prim"] == o["prim"]) & (g["geom"] == o["geom"])
        assert same.mean() > 0.999, same.mean()
        hit = same & (o["prim"] >= 0)
        rel = np.abs(g["t"][hit] - o["t"][hit]) / np.abs(o["t"][hit])
        assert rel.max() <= 1e-5, rel.max()
        # rays that never enter the moving instance's bounds are still bit-exact
        assert (g["t"].view(np.uint32) == o["t"].view(np.uint32)).mean() > 0.97
    sh = rays.copy()
    sh["tmax"] = 3.0
    og, gg = ora.trace(sh, any_hit=True), dev.trace(sh, any_hit=True)
    assert ((og["prim"] >= 0) == (gg["prim"] >= 0)).mean() > (0.9999 if kind == "static" else 0.999)


def test_pinned_batch_is_pipelined_and_identical(built_library):
    """vg_trace_batch with page-locked ray/hit buffers overlaps H2D, traversal and D2H chunk by chunk on three streams; the
    hits must equal those of the plain path (pageable buffers, one kernel) bit for bit, counters included."""
    import torch
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene, HIT_DTYPE, RAY_DTYPE
    from conftest import random_rays
    sc = scenes.heightfield_scene(64, 48, nq=96)
    dev = Device(0).upload(HostScene(sc).prerender())
    n = (1 << 20) + 12345          # two full chunks and a ragged third
    rays = random_rays(n, 9, lo=(-1.0, 0.05, -1.0), hi=(1.0, 0.6, 1.0))
    rays["d"][:, 1] = -np.abs(rays["d"][:, 1])      # aim down at the heightfield
    plain = dev.trace(rays)
    rp = torch.from_numpy(rays.view(np.uint8).reshape(n, 32)).pin_memory().numpy().reshape(-1).view(RAY_DTYPE)
    hp = torch.zeros((n, 32), dtype=torch.uint8, pin_memory=True).numpy().reshape(-1).view(HIT_DTYPE)
    got = dev.trace(rp, out=hp)
    assert got is hp
    assert (plain["prim"] >= 0).mean() > 0.5
    assert got.tobytes() == plain.tobytes()
    sh = dev.trace(rp, any_hit=True, out=hp).copy()
    assert np.array_equal(sh["prim"] >= 0, dev.trace(rays, any_hit=True)["prim"] >= 0)
