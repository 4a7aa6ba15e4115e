"""vermeer_b200 — B200-native ray-traversal and path-integration engine behind Vermeer's
core.Trace/TraceProbe, core.Geom, core.Shader and nodes-registry boundary.

The product is the C-ABI shared library built from `vermeer_b200/csrc` (declared in
`include/vermeer_gpu.h`); this package is the thin Python harness over it (ctypes, numpy) used by
tests and bench.py.  There is no CPU fallback: every compute entry point fails loudly when the CUDA
library or a GPU is missing.
"""
__version__ = "0.1.0"
