"""Parity properties at BASELINE.json's full C2 size (1 002 528 triangles, 1920x1080), where the oracle would take minutes:
size-independent properties the domain offers, each checked bit for bit through the C ABI.
  * a progressive render equals a one-shot render ([0,2)+[2,4) == [0,4)): the running mean continues exactly;
  * the union of two tile partitions equals the single-context frame (what the multi-GPU gather relies on);
  * closest-hit and any-hit agree on occlusion for every camera ray, and a ray shortened to just before its hit is unoccluded;
  * the device-built QBVH equals the host-built one (nodes, boxes, leaf ranges) at 1 M triangles;
  * ray counts follow the reference's definition (every camera sample is one TraceProbe, every shadow sample another)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2(built_library):
    from vermeer_b200 import scenes
    from vermeer_b200.host import Device, HostScene
    sc = scenes.heightfield_scene(1920, 1080, nq=708)
    assert sc.num_tris == 1002530
    host = HostScene(sc).prerender()
    tab = scenes.splitmix64_table(1, sc.XRes * sc.YRes)
    return sc, host, tab


def test_progressive_equals_one_shot_and_partitions_tile_the_frame(c2):
    from vermeer_b200.host import Device
    sc, host, tab = c2
    dev = Device(0).upload(host)
    dev.set_scramble(tab)
    one = dev.render(0, 4).copy()
    st = dev.stats()
    npix = sc.XRes * sc.YRes
    assert st["rays"] - st["shadow_rays"] == 4 * npix           # one camera TraceProbe per pixel and iteration
    dev.clear()
    dev.render(0, 2)
    two = dev.render(2, 4)
    assert np.array_equal(one.view(np.uint32), two.view(np.uint32))
    parts = []
    for rank in range(2):
        d = Device(0).upload(host)
        d.set_partition(rank, 2)
        d.set_scramble(tab)
        parts.append(d.render(0, 4))
    owned0, owned1 = (parts[0] != 0).any(-1), (parts[1] != 0).any(-1)
    assert not (owned0 & owned1).any()                           # non-owned pixels stay 0
    union = parts[0] + parts[1]
    assert np.array_equal(union.view(np.uint32), one.view(np.uint32))


def test_any_hit_agrees_with_closest_hit(c2):
    from oracle.binding import Oracle  # camera rays only (ray generation is cheap on the CPU); traversal is all on the GPU
    from vermeer_b200.host import Device
    sc, host, tab = c2
    ora = Oracle(sc)
    ora.set_scramble(tab)
    rays = ora.camera_rays(1)
    dev = Device(0).upload(host)
    closest = dev.trace(rays)
    anyhit = dev.trace(rays, any_hit=True)
    assert np.array_equal(closest["prim"] >= 0, anyhit["prim"] >= 0)
    hit = closest["prim"] >= 0
    assert 0.5 < hit.mean() < 1.0
    short = rays[hit].copy()
    short["tmax"] = closest["t"][hit] * np.float32(0.999)
    assert (dev.trace(short, any_hit=True)["prim"] < 0).all()    # nothing in front of the closest hit
    longer = rays[hit].copy()
    longer["tmax"] = closest["t"][hit] * np.float32(1.001)
    assert (dev.trace(longer, any_hit=True)["prim"] >= 0).all()


def test_device_built_tree_at_full_size(c2):
    from vermeer_b200.host import Device, HostScene
    sc, host, tab = c2
    dev = Device(0)
    hd = HostScene(sc).prerender(device=dev)
    nh, nd = host.mesh_nodes(0), hd.mesh_nodes(0)
    assert len(nh) == len(nd)
    for f in ("Axis0", "Axis1", "Axis2", "Children"):
        assert np.array_equal(nh[f], nd[f]), f
    assert np.array_equal(nh["Boxes"].view(np.uint32), nd["Boxes"].view(np.uint32))
    (_, ah), (_, ad) = host.mesh_idxp(0), hd.mesh_idxp(0)
    assert np.array_equal(np.sort(ad), np.arange(len(ad)))
    # same face set per leaf: sort both permutations inside every leaf range and compare once
    ch = nh["Children"].reshape(-1)
    leaf = ch[(ch < 0) & (ch != -1)]
    base, cnt = (leaf & 0x7ffffff) >> 4, (leaf & 0xf) + 1
    key = np.zeros(len(ah), np.int64)
    key[base] = 1
    key = np.cumsum(key)                                          # leaf number of every slot (leaves tile the slots)
    assert int(cnt.sum()) == len(ah)
    assert np.array_equal(ah[np.lexsort((ah, key))], ad[np.lexsort((ad, key))])
