"""Reference quirk q, demonstrated on the oracle: what the reference does with a PolyMesh's own `Transform`
(builtin/geom/polymesh/trace.go:22-55,62-74, bounds.go:26) — the reason the GPU path accepts only an identity Transform and sends
transformed placement through GeomInstance instead (round-1 verdict, item 8: "show with an oracle test that the reference's
result is unusable and keep the refusal").

Three facts, each asserted below on a floor + cube scene whose cube carries Transform = translate(tx, 0, 0):
  1. `Bounds()` ignores the transform (`if false && mesh.transformBounds != nil`), so the scene-level tree culls the mesh by its
     UNtransformed box. Rays that enter that box are taken to object space and hit the cube exactly where the moved cube is, so
     the mesh is drawn only where a ray crosses BOTH the untransformed box and the moved cube: clipped for a small translation
     (tx = 0.3: the boxes overlap), entirely invisible once the two no longer line up (tx = 0.8);
  2. where it is drawn, the hit is the intended one (same t): the ray transform itself is right, the culling is not;
  3. `Trace` rewrites sg.Transform / sg.P / sg.N after the traversal whether or not it hit (:62-74): a floor or wall point seen
     through the cube's box is shaded at P moved by the translation — the picture of everything BEHIND the box changes too.
So no placement of the mesh yields the image of the transformed mesh; the feature is unusable as it stands."""
import numpy as np

from conftest import random_rays


def _scene(translate=0.0):
    from vermeer_b200 import scenes
    sc = scenes.cornell_box(64, 64, boxes=False)
    cube = scenes._box("cube", (-0.3 + translate, 0.0, -0.3), (0.3 + translate, 0.6, 0.3), "red")
    sc.meshes.append(cube)
    return sc


def _xlate(tx):
    from vermeer_b200 import scenes
    return scenes.matrix4([[1, 0, 0, tx], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])


def test_mesh_transform_as_the_reference_executes_it(oracle_lib):
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    tab = scenes.splitmix64_table(1, 64 * 64)
    cube_geom = len(_scene().meshes) - 1
    drawn_share = {}
    for tx in (0.3, 0.8):
        intended = Oracle(_scene(translate=tx))         # the cube's vertices moved by hand: what Transform is meant to give
        xf = Oracle(_scene(translate=0.0))
        xf.mesh_set_transform("cube", _xlate(tx))       # the same placement through PolyMesh.Transform
        for o in (intended, xf):
            o.set_scramble(tab)
        cam = intended.camera_rays(1)
        hi, hx = intended.trace(cam), xf.trace(cam)
        seen = hi["geom"] == cube_geom
        drawn = hx["geom"] == cube_geom
        assert seen.sum() > 100
        # 2. whatever is drawn is the intended hit ...
        assert (seen[drawn]).all() and np.array_equal(hx["t"][drawn], hi["t"][drawn])
        # 1. ... but only a part of the cube (or none of it) is drawn
        drawn_share[tx] = drawn.sum() / seen.sum()
    assert 0.05 < drawn_share[0.3] < 0.9, drawn_share
    assert drawn_share[0.8] == 0.0, drawn_share

    # 3. the frame (tx = 0.8, cube invisible): still not the picture of the room without the cube — points seen THROUGH the
    #    untransformed box are shaded with the stale transform the missed mesh left in the context (trace.go:62-74)
    fx, _ = xf.render(0, 4, nthreads=4)
    fi, _ = intended.render(0, 4, nthreads=4)
    ok = np.isfinite(fi).all(-1) & np.isfinite(fx).all(-1)
    assert float(np.sqrt(((fi[ok] - fx[ok]) ** 2).mean())) > 0.01
    nocube = Oracle(scenes.cornell_box(64, 64, boxes=False))
    nocube.set_scramble(tab)
    fn, _ = nocube.render(0, 4, nthreads=4)
    hn = nocube.trace(cam)
    same_hit = (hx["geom"] == hn["geom"]) & (hx["prim"] == hn["prim"]) & (hx["t"] == hn["t"]) & (hx["geom"] >= 0)
    assert same_hit.mean() > 0.95                      # the traversal finds the same surfaces as in the empty room ...
    d = np.abs(fx - fn).max(-1).reshape(-1)
    fin = np.isfinite(d)
    assert (d[same_hit & fin] > 1e-3).mean() > 0.01    # ... and shades a visible share of them differently


def test_gpu_path_refuses_what_it_cannot_mirror(built_library):
    """The host layer's side of the decision: an identity Transform passes, anything else is an error at PreRender."""
    import pytest
    from vermeer_b200.host import HostScene
    base = ('ShaderStd { Name "m" DiffuseStrength float 1 }\n'
            'Camera { Name "camera" Type "LookAt" From 1 1 point 0 1 3 To 1 1 point 0 0 0 Roll 1 1 float 0 Up 0 1 0 }\n'
            'PolyMesh { Name "t" Verts 1 3 point 0 0 0 1 0 0 0 0 -1 Shader 1 string "m" Transform 1 matrix %s }\n')
    HostScene.from_vnf(text=base % "1 0 0 0  0 1 0 0  0 0 1 0  0 0 0 1").prerender()
    with pytest.raises(RuntimeError, match="Transform"):
        HostScene.from_vnf(text=base % "1 0 0 0.8  0 1 0 0  0 0 1 0  0 0 0 1").prerender()
