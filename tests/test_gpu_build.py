"""qbvh.BuildAccel on the device (vermeer_b200/csrc/build_bvh.cu) against the host builder, which is itself
bit-identical to the oracle's restatement of qbvh/build.go (tests/test_host_layer.py).  Bar: the same nodes, boxes, axes and leaf
ranges bit for bit; the same SET of faces in every leaf (the order inside a leaf is the one documented difference)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _soup(n, seed, extent=(4.0, 1.0, 2.5), size=0.08):
    """n random triangles in a flat-ish box, so that different levels split on different axes."""
    from vermeer_b200 import scenes
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, (n, 1, 3)) * np.asarray(extent)
    v = (c + rng.normal(size=(n, 3, 3)) * size).astype(np.float32).reshape(-1, 3)
    return scenes.PolyMesh("soup", v, ["white"])


def _scene_with(mesh):
    from vermeer_b200 import scenes
    sc = scenes.cornell_box(64, 48, boxes=False)
    sc.meshes.append(mesh)
    return sc


def _leaf_ranges(nodes):
    out = []
    for ch in nodes["Children"].reshape(-1):
        if ch < 0 and ch != -1:
            base, cnt = (int(ch) & 0x7ffffff) >> 4, (int(ch) & 0xf) + 1
            out.append((base, cnt))
    return out


def _compare_trees(sc, gid):
    from vermeer_b200.host import Device, HostScene
    hh = HostScene(sc).prerender()
    dev = Device(0)
    hd = HostScene(sc).prerender(device=dev)
    nh, nd = hh.mesh_nodes(gid), hd.mesh_nodes(gid)
    assert len(nh) == len(nd)
    for f in ("Axis0", "Axis1", "Axis2", "Children"):
        assert np.array_equal(nh[f], nd[f]), f
    assert np.array_equal(nh["Boxes"].view(np.uint32), nd["Boxes"].view(np.uint32))
    (ih, ah), (idd, ad) = hh.mesh_idxp(gid), hd.mesh_idxp(gid)
    assert sorted(ah.tolist()) == sorted(ad.tolist()) == list(range(len(ah)))
    for base, cnt in _leaf_ranges(nh):
        assert sorted(ah[base:base + cnt].tolist()) == sorted(ad[base:base + cnt].tolist())
    # the per-face arrays follow the permutation: slot i holds the vertices of face accel_idx[i]
    return hh, hd, dev, ah, ad


@pytest.mark.parametrize("case", ["heightfield", "soup20k", "soup100k", "sphere"])
def test_device_tree_equals_host_tree(built_library, case):
    from vermeer_b200 import scenes
    if case == "heightfield":
        sc = scenes.heightfield_scene(96, 64, nq=150)
        gid = 0
    elif case == "sphere":
        sc = scenes.sphere_field_scene(96, 64, nmesh=3, slices=130, stacks=131)
        gid = 1
    else:
        sc = _scene_with(_soup(40000 if case == "soup20k" else 100000, 5))
        gid = len(sc.meshes) - 1
    _compare_trees(sc, gid)


def test_hits_and_image_through_the_device_tree(built_library):
    from oracle.binding import Oracle
    from conftest import random_rays
    from vermeer_b200 import scenes
    sc = scenes.heightfield_scene(128, 96, nq=150)
    hh, hd, dev, ah, ad = _compare_trees(sc, 0)
    tab = scenes.splitmix64_table(1, sc.XRes * sc.YRes)
    ora = Oracle(sc)
    ora.set_scramble(tab)
    dev.upload(hd)
    dev.set_scramble(tab)
    rays = np.concatenate([ora.camera_rays(1), random_rays(30000, 3, lo=(-1, 0.0, -1), hi=(1, 1.5, 1))])
    g, o = dev.trace(rays), ora.trace(rays)
    hit = o["prim"] >= 0
    assert np.array_equal(g["prim"] >= 0, hit) and np.array_equal(g["geom"], o["geom"])
    on_mesh = hit & (o["geom"] == 0)
    # same FACE (prim is the leaf-order slot; the two builders order a leaf differently), same t/u/v/w bits, same counters
    same_face = ad[g["prim"][on_mesh]] == ah[o["prim"][on_mesh]]
    assert same_face.mean() >= 0.9999, same_face.mean()     # a ray through a shared edge can pick the other triangle of a tie
    ok = np.flatnonzero(on_mesh)[same_face]
    for f in ("t", "u", "v", "w"):
        assert np.array_equal(g[f][ok].view(np.uint32), o[f][ok].view(np.uint32)), f
    assert np.array_equal(g["nodesT"], o["nodesT"]) and np.array_equal(g["trisT"], o["trisT"])
    # images: device tree vs host tree, both rendered on the GPU
    f_dev = dev.render(0, 4)
    d2 = type(dev)(0).upload(hh)
    d2.set_scramble(tab)
    f_host = d2.render(0, 4)
    assert np.sqrt(((f_dev - f_host) ** 2).mean()) <= 1e-6
    assert (f_dev.view(np.uint32) == f_host.view(np.uint32)).mean() >= 0.999


def test_raw_entry_point_and_degenerate_inputs(built_library):
    from vermeer_b200.host import Device
    dev = Device(0)
    rng = np.random.default_rng(1)
    # fewer primitives than leafMax: one node, one leaf, three empty children (qbvh.go:67-79)
    c = rng.uniform(-1, 1, (5, 3)).astype(np.float32)
    b = np.concatenate([c - 0.1, c + 0.1], 1).astype(np.float32)
    nodes, idx, bounds = dev.build_qbvh(b, c)
    assert len(nodes) == 1 and sorted(idx.tolist()) == [0, 1, 2, 3, 4]
    ch = nodes["Children"][0]
    assert ch[0] == np.int32(np.uint32((1 << 31) | (0 << 4) | 4).view(np.int32)) and list(ch[1:]) == [-1, -1, -1]
    assert np.allclose(bounds, np.concatenate([b[:, :3].min(0), b[:, 3:].max(0)]))
    assert np.isinf(nodes["Boxes"][0][[1, 2, 3]]).all()
    # leafMax = 1 (the scene-level tree's setting): every leaf holds exactly one primitive
    c = rng.uniform(-1, 1, (1000, 3)).astype(np.float32)
    b = np.concatenate([c - 0.01, c + 0.01], 1).astype(np.float32)
    nodes, idx, _ = dev.build_qbvh(b, c, leaf_max=1)
    leaves = _leaf_ranges(nodes)
    assert len(leaves) == 1000 and all(cnt == 1 for _, cnt in leaves)
    # 40 primitives on one centroid, leafMax 16: the flat branch halves the range without partitioning (build.go:35-43) and ends
    c = np.zeros((40, 3), np.float32)
    b = np.concatenate([c - 0.1, c + 0.1], 1).astype(np.float32)
    nodes, idx, _ = dev.build_qbvh(b, c)
    assert sorted(cnt for _, cnt in _leaf_ranges(nodes)) == [9, 10, 10, 11] and idx.tolist() == list(range(40))
    # two primitives on one centroid with leafMax 1: len/2+1 == len, the reference recurses until its stack overflows (quirk f)
    with pytest.raises(RuntimeError, match="unbounded recursion"):
        dev.build_qbvh(b[:2], c[:2], leaf_max=1)
    c[0, 0] = np.nan
    with pytest.raises(RuntimeError):
        dev.build_qbvh(b, c)


def test_motion_topology_through_the_device_builder(built_library):
    """BuildAccelMotion (qbvh/motionbuild.go:108-122): the same recursion on the mid-time snapshot; per-key boxes follow from the
    leaf sets, so topology AND boxes must equal the host builder's, and traversal must stay bit-identical to the oracle."""
    from oracle.binding import Oracle
    from conftest import random_rays, assert_hits_equal
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    sc = scenes.heightfield_scene(96, 64, nq=150, motion=True)
    hh = HostScene(sc).prerender()
    dev = Device(0)
    hd = HostScene(sc).prerender(device=dev)
    (th, bh), (td, bd) = hh.mesh_nodes(0), hd.mesh_nodes(0)
    for f in ("Axis0", "Axis1", "Axis2", "Children"):
        assert np.array_equal(th[f], td[f]), f
    assert np.array_equal(bh.view(np.uint32), bd.view(np.uint32))
    tab = scenes.splitmix64_table(1, sc.XRes * sc.YRes)
    ora = Oracle(sc, motion_ref_compat=False)
    ora.set_scramble(tab)
    dev.upload(hd)
    rays = np.concatenate([ora.camera_rays(2), random_rays(20000, 4, lo=(-1, 0.0, -1), hi=(1, 1.5, 1))])
    g, o = dev.trace(rays), ora.trace(rays)
    # motion meshes report the ORIGINAL face index (trace.go:624), so prim ids compare directly
    same = g["prim"] == o["prim"]
    assert same.mean() >= 0.9999
    for f in ("t", "u", "v", "w"):
        assert np.array_equal(g[f][same].view(np.uint32), o[f][same].view(np.uint32)), f
    assert np.array_equal(g["nodesT"], o["nodesT"]) and np.array_equal(g["trisT"], o["trisT"])
