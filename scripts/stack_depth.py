"""Deepest traversal stack per config, measured with a -DVG_STACK_STATS build (VG_SO_PATH=build_variants/lib_stackstats.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vermeer_b200 import scenes
from vermeer_b200.host import Device, HostScene
for cfg in ("c1", "c2", "c4", "c3"):
    sc = bench.build_scene(cfg)
    dev = Device(0).upload(HostScene(sc).prerender())
    dev.set_scramble(scenes.splitmix64_table(1, sc.XRes * sc.YRes))
    dev.set_option("iters_per_batch", 8)
    dev.render(0, 8, fetch=False)
    print(cfg, "max traversal stack depth over %d rays: %d entries" % (dev.stats()["rays"], dev.stats()["max_stack_depth"]))
