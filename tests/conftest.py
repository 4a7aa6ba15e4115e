import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_library():
    """The CUDA/C-ABI library, built in-tree (nvcc cross-compiles without a GPU)."""
    from vermeer_b200.build import build
    return build()


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import binding
    binding.build()
    return binding.lib()


def random_rays(n, seed, lo=(-1.2, -0.3, -1.2), hi=(1.2, 2.2, 1.2), tmax=np.inf):
    """Uniform origins in a box, uniform directions on the sphere (not normalised exactly: the traversal does not need it)."""
    from vermeer_b200.host import RAY_DTYPE
    rng = np.random.default_rng(seed)
    r = np.zeros(n, RAY_DTYPE)
    r["o"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r["d"] = d.astype(np.float32)
    r["tmax"] = np.float32(tmax)
    r["time"] = rng.random(n).astype(np.float32)
    return r


def assert_hits_equal(gpu, ora, check_counters=True, what=""):
    """Bit-exact comparison of two VgHit arrays (the bar for integer/index work and for t)."""
    for f in ("prim", "geom"):
        bad = np.nonzero(gpu[f] != ora[f])[0]
        assert bad.size == 0, "%s %s differs at %d rays, first %s: gpu=%s oracle=%s" % (what, f, bad.size, bad[:5], gpu[bad[:5]], ora[bad[:5]])
    for f in ("t", "u", "v", "w"):
        a, b = gpu[f].view(np.uint32), ora[f].view(np.uint32)
        # NaN payload/sign is not specified (x86 produces the negative "indefinite" NaN, the GPU the canonical positive one)
        bad = np.nonzero((a != b) & ~(np.isnan(gpu[f]) & np.isnan(ora[f])))[0]
        assert bad.size == 0, "%s %s not bit-identical at %d rays, first %s: gpu=%s oracle=%s" % (what, f, bad.size, bad[:5], gpu[f][bad[:5]], ora[f][bad[:5]])
    if check_counters:
        for f in ("nodesT", "trisT"):
            bad = np.nonzero(gpu[f] != ora[f])[0]
            assert bad.size == 0, "%s %s differs at %d rays" % (what, f, bad.size)
