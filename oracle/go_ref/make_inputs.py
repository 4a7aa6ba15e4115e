"""Write the scenes (.vnf through the repo's writer) and ray batches the Go dumper runs on: oracle/_ref/go/<name>/{scene.vnf, rays.bin}."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import random_rays  # noqa: E402
from vermeer_b200 import scenes  # noqa: E402

SCENES = {
    "cornell": lambda: scenes.cornell_box(96, 96),
    "heightfield": lambda: scenes.heightfield_scene(128, 96, nq=60),
    "motion": lambda: scenes.heightfield_scene(96, 64, nq=40, motion=True),
    "spheres": lambda: scenes.sphere_field_scene(96, 96, nmesh=16, slices=16, stacks=17),
}


def main():
    out = os.path.join(ROOT, "oracle", "_ref", "go")
    for name, make in SCENES.items():
        d = os.path.join(out, name)
        os.makedirs(d, exist_ok=True)
        sc = make()
        open(os.path.join(d, "scene.vnf"), "w").write(scenes.to_vnf(sc))
        random_rays(20000, 11, lo=(-1.0, 0.05, -1.0), hi=(1.0, 1.8, 1.0)).tofile(os.path.join(d, "rays.bin"))
        print("wrote", d)


if __name__ == "__main__":
    main()
