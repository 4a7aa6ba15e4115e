"""PolyMesh.PreRender refuses index arrays that do not fit together (the reference panics on a Go bounds check; the round-1
advisor found the host mirror reading and writing out of bounds instead). Host code only: no GPU."""
import numpy as np
import pytest


def _mesh(**over):
    from vermeer_b200 import scenes
    v = np.asarray([[0, 0, 0], [1, 0, 0], [1, 0, 1], [0, 0, 1], [2, 0, 0], [2, 0, 1]], np.float32)
    kw = dict(PolyCount=np.asarray([4, 4]), FaceIdx=np.asarray([0, 3, 2, 1, 1, 2, 5, 4]))
    kw.update(over)
    return scenes.PolyMesh("m", v, kw.pop("Shader", ["a", "b"]), **kw)


def _prerender(mesh):
    from vermeer_b200 import scenes
    from vermeer_b200.host import HostScene
    sc = scenes.cornell_box(32, 32, boxes=False)
    sc.shaders += [scenes.ShaderStd("a", DiffuseStrength=1.0), scenes.ShaderStd("b", DiffuseStrength=1.0)]
    sc.meshes.append(mesh)
    return HostScene(sc).prerender()


def test_well_formed_mesh_prerenders(built_library):
    h = _prerender(_mesh(ShaderIdx=np.asarray([0, 1])))
    assert h.mesh_info(h.num_geoms() - 3)["tris"] >= 0


@pytest.mark.parametrize("over,msg", [
    (dict(ShaderIdx=np.asarray([0])), "ShaderIdx"),                                  # shorter than PolyCount
    (dict(ShaderIdx=np.asarray([0, 2])), "Shader list"),                             # value beyond the Shader list
    (dict(PolyCount=np.asarray([4, 5])), "PolyCount"),                               # sums beyond FaceIdx
    (dict(Normals=np.tile(np.float32([0, 1, 0]), (6, 1)), NormalIdx=np.asarray([0, 1, 2])), "NormalIdx"),
    (dict(Normals=np.tile(np.float32([0, 1, 0]), (2, 1)), NormalIdx=np.asarray([0, 1, 2, 3, 1, 2, 5, 4])), "normal index"),
])
def test_malformed_index_arrays_are_refused(built_library, over, msg):
    with pytest.raises(RuntimeError) as e:
        _prerender(_mesh(**over))
    assert msg in str(e.value), str(e.value)
