"""Structural anchors of the oracle's traversal: the SSE box test equals the reference's own
scalar twin intersectBoxesSlow2 on finite inputs, and BVH traversal equals brute force through the same triangle routine."""
import ctypes as C

import numpy as np

from conftest import random_rays


def _box(L, P, D, boxes, which):
    hits = (C.c_int32 * 4)()
    t = (C.c_float * 4)()
    L.orc_box_test((C.c_float * 3)(*P), (C.c_float * 3)(*D), (C.c_float * 24)(*boxes), which, hits, t)
    return list(hits), np.asarray(list(t), np.float32)


def test_sse_box_test_matches_scalar_twin(oracle_lib):
    rng = np.random.default_rng(1)
    for _ in range(3000):
        lo = rng.uniform(-2, 2, (3, 4)).astype(np.float32)
        hi = lo + rng.uniform(0, 1.5, (3, 4)).astype(np.float32)
        boxes = np.concatenate([lo.reshape(-1), hi.reshape(-1)])  # [min x(4) y(4) z(4), max x y z]
        P = rng.uniform(-3, 3, 3).astype(np.float32)
        D = rng.normal(size=3).astype(np.float32)
        h0, t0 = _box(oracle_lib, P, D, boxes, 0)
        h1, t1 = _box(oracle_lib, P, D, boxes, 1)
        assert h0 == h1
        assert np.array_equal(t0.view(np.uint32), t1.view(np.uint32))


def test_box_test_nan_semantics(oracle_lib):
    """Dinv = +-Inf with the origin exactly on a slab plane gives 0*Inf = NaN; MINPS/MAXPS return the second operand, so an
    x-axis NaN is dropped (the slab is ignored) while a z-axis NaN poisons tNear and the box is missed."""
    boxes = np.zeros(24, np.float32)
    boxes[0:4] = 0.0    # min x
    boxes[4:8] = -1.0   # min y
    boxes[8:12] = -1.0  # min z
    boxes[12:16] = 1.0  # max x
    boxes[16:20] = 1.0
    boxes[20:24] = 1.0
    h, t = _box(oracle_lib, (0.0, 0.0, -5.0), (0.0, 0.0, 1.0), boxes, 0)   # D.x = 0 -> Dinv.x = Inf, min.x - O.x = 0
    assert h == [-1, -1, -1, -1] and np.all(t == 4.0)
    boxes2 = boxes.copy()
    boxes2[0:4] = -1.0
    boxes2[8:12] = 0.0  # min z on the origin plane, D.z = 0
    h, t = _box(oracle_lib, (-5.0, 0.0, 0.0), (1.0, 0.0, 0.0), boxes2, 0)
    assert h == [0, 0, 0, 0] and np.all(np.isnan(t))


def test_ray_setup(oracle_lib):
    L = oracle_lib
    out = (C.c_float * 6)()
    k = (C.c_int32 * 3)()
    L.orc_ray_setup((C.c_float * 3)(0, 0, 0), (C.c_float * 3)(0.1, -0.9, 0.2), out, k)
    assert list(k) == [0, 2, 1]          # Kz = 1 (|y| largest); D[Kz] < 0 swaps Kx, Ky (core/ray.go:132-136)
    assert out[5] == np.float32(1.0 / np.float64(np.float32(-0.9)))
    L.orc_ray_setup((C.c_float * 3)(0, 0, 0), (C.c_float * 3)(0.5, 0.5, 0.5), out, k)
    assert list(k) == [1, 2, 0]          # ties resolve to Kz = 0 (ray.go:107-117)


def test_bvh_equals_brute_force():
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    for sc, lo, hi in [(scenes.cornell_box(32, 32), (-0.95, 0.05, -0.95), (0.95, 1.9, 0.95)),
                       (scenes.heightfield_scene(32, 32, nq=24), (-1, 0.2, -1), (1, 1.0, 1)),
                       (scenes.sphere_field_scene(32, 32, nmesh=4, slices=8, stacks=9), (-1, 0.05, -1), (1, 1, 1))]:
        ora = Oracle(sc)
        rays = random_rays(4000, 5, lo=lo, hi=hi)
        a = ora.trace(rays)
        b = ora.trace(rays, brute=True)
        assert np.array_equal((a["prim"] >= 0), (b["prim"] >= 0))
        # The visiting order differs, and the accept test compares T against Tclosest*det without dividing
        # (trace.go:182), so on shared edges a neighbour within an ulp can win in one order and lose in the other:
        # t agrees bit for bit on almost every ray and to 1e-5 relative on all; the primitive differs only on such ties.
        hit = a["prim"] >= 0
        exact = a["t"].view(np.uint32) == b["t"].view(np.uint32)
        assert exact.mean() > 0.99
        rel = np.abs(a["t"][hit] - b["t"][hit]) / np.abs(b["t"][hit])
        assert rel.max() < 1e-5
        same = (a["prim"] == b["prim"]) & (a["geom"] == b["geom"])
        assert same.mean() > 0.98


def test_motion_modes_differ_only_by_leaf_indexing():
    """ref_compat (quirk b) tests face i at leaf slot i; fixed tests face accel.idx[i]. Brute force over all faces visits the same
    set of triangles in both modes, so it must give the same closest t."""
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    sc = scenes.heightfield_scene(32, 32, nq=20, motion=True)
    rays = random_rays(3000, 6, lo=(-1, 0.3, -1), hi=(1, 1.2, 1))
    rays["d"][:, 1] = -np.abs(rays["d"][:, 1])
    fixed = Oracle(sc, motion_ref_compat=False)
    compat = Oracle(sc, motion_ref_compat=True)
    bf, bc = fixed.trace(rays, brute=True), compat.trace(rays, brute=True)
    assert np.array_equal(bf["prim"] >= 0, bc["prim"] >= 0)
    assert np.allclose(bf["t"], bc["t"], rtol=1e-5, equal_nan=True)
    # the fixed mode's BVH agrees with brute force; the reference-compatible one mostly misses (boxes bound other faces)
    tf = fixed.trace(rays)
    assert np.array_equal(tf["prim"] >= 0, bf["prim"] >= 0)
    assert np.allclose(tf["t"], bf["t"], rtol=1e-5, equal_nan=True)
    tc = compat.trace(rays)
    assert (tc["prim"] >= 0).mean() < (tf["prim"] >= 0).mean()


def test_builder_reports_reference_quirk_f():
    """Two geoms with identical bounds centroids make the reference's scene-level build (leafMax=1) recurse forever
    (qbvh/build.go:35-43). The oracle and the host builder both report it instead of overflowing the stack."""
    import pytest
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    from vermeer_b200.host import HostScene
    sc = scenes.cornell_box(16, 16, boxes=False)
    sc.lights = [scenes.TriLight("a", (-0.25, 1.99, -0.25), (0.25, 1.99, -0.25), (0.25, 1.99, 0.25), "lightmtl", 1),
                 scenes.TriLight("b", (-0.25, 1.99, -0.25), (0.25, 1.99, 0.25), (-0.25, 1.99, 0.25), "lightmtl", 1)]
    with pytest.raises(RuntimeError, match="unbounded recursion"):
        Oracle(sc)
    with pytest.raises(RuntimeError, match="unbounded recursion"):
        HostScene(sc).prerender()


def _box3(L, P, D, tclosest, boxes, which):
    hits = (C.c_int32 * 4)()
    t = (C.c_float * 4)()
    L.orc_box_test3.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_box_test3((C.c_float * 3)(*P), (C.c_float * 3)(*D), C.c_float(tclosest), (C.c_float * 24)(*boxes), which, hits, t)
    return np.asarray(list(hits), np.int32), np.asarray(list(t), np.float32)


def test_reference_self_check_asm_vs_slow_vs_slow2(oracle_lib):
    """The reference's own (commented-out) cross-check, qbvh/intersect.go:116-133: the asm box test against intersectBoxesSlow
    (:17-50) and intersectBoxesSlow2 (:52-87), here on the nodes of a real tree plus random boxes, and WHERE they may differ:
      * asm == Slow2 everywhere on finite inputs (hits and t bit for bit): Slow2 is the asm's twin;
      * Slow clamps tFar by Tclosest (:39), the asm does not: with Tclosest = +Inf the three agree; with a finite Tclosest the
        asm's extra hits are exactly the children whose tNear lies beyond Tclosest — the ones Trace discards at pop time
        (:106, Tclosest < Stack.T), which is why the reference can push them unclamped;
      * tNear itself is the same number in all three (the near plane picked by the sign of Dinv is the min of the two products).
    Rays with a zero direction component are left to test_box_test_nan_semantics: there Slow's sign-based plane choice and the
    MINPS/MAXPS operand order genuinely part ways."""
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    sc = scenes.heightfield_scene(32, 32, nq=40)
    ora = Oracle(sc)
    nodes = ora.mesh_nodes(0)
    rng = np.random.default_rng(3)
    n_extra = 0
    for i in range(4000):
        if i % 2 == 0:
            nd = nodes[rng.integers(len(nodes))]
            boxes = nd["Boxes"].copy()
            boxes[~np.isfinite(boxes)] = 0.0       # empty children hold +-Inf boxes; the finite-input claim is about real boxes
            P = np.float32([rng.uniform(-1, 1), rng.uniform(0.2, 1.5), rng.uniform(-1, 1)])
            D = np.float32([rng.normal() * 0.5, -abs(rng.normal()) - 0.05, rng.normal() * 0.5])
        else:
            lo = rng.uniform(-2, 2, (3, 4)).astype(np.float32)
            hi = lo + rng.uniform(0, 1.5, (3, 4)).astype(np.float32)
            boxes = np.concatenate([lo.reshape(-1), hi.reshape(-1)])
            P = rng.uniform(-3, 3, 3).astype(np.float32)
            D = rng.normal(size=3).astype(np.float32)
        D[D == 0] = np.float32(0.25)
        for tcl in (np.inf, float(rng.uniform(0.05, 1.0))):
            ha, ta = _box3(oracle_lib, P, D, tcl, boxes, 0)
            h2, t2 = _box3(oracle_lib, P, D, tcl, boxes, 1)
            hs, ts = _box3(oracle_lib, P, D, tcl, boxes, 2)
            assert np.array_equal(ha, h2) and np.array_equal(ta.view(np.uint32), t2.view(np.uint32))
            assert np.array_equal(ta, ts)                          # the same tNear (== : the sign of a zero may differ)
            if np.isinf(tcl):
                assert np.array_equal(ha, hs)
            else:
                assert np.all(hs[ha == 0] == 0)                    # Slow never reports a box the asm misses
                extra = (ha != 0) & (hs == 0)
                assert np.all(ta[extra] > tcl)                     # ... and misses only children beyond Tclosest
                assert np.all(hs[(ha != 0) & (ta <= tcl)] != 0)
                n_extra += int(extra.sum())
    assert n_extra > 50                                            # the difference is exercised, not vacuous
