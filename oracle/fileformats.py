"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.

numpy restatement of the reference's output drivers, used by tests/ to check the host layer's vh_postrender byte for byte.
parity unpinned: the reference has no tests or golden files for these writers; anchored by hand-computed known answers
(tests/test_vnf_and_outputs.py).

Follows:
  driver.OutputFloat.PostRender   builtin/driver/outputfloat.go:30-42  (binary.Write LittleEndian of the []float32 framebuffer)
  hdr.convertRGBToRGBE            image/hdr/hdr.go:26-50
  hdr.Writer.WriteImage           image/hdr/writer.go:58-91            (header, flat RGBE scanlines, bottom row first)
"""
from __future__ import annotations

import numpy as np


def output_float_bytes(fb: np.ndarray) -> bytes:
    return np.ascontiguousarray(fb, "<f4").tobytes()


def _go_byte(v: np.ndarray) -> np.ndarray:
    """Go/amd64 float32 -> byte: CVTTSS2SL (truncate toward zero; NaN/out of range -> 0x80000000) then the low byte."""
    v = np.asarray(v, np.float32)
    bad = ~np.isfinite(v) | (v >= np.float32(2147483648.0)) | (v <= np.float32(-2147483904.0))
    i = np.where(bad, np.int64(-2147483648), np.trunc(np.where(bad, 0, v)).astype(np.int64))
    return (i & 0xFF).astype(np.uint8)


def rgb_to_rgbe(rgb: np.ndarray) -> np.ndarray:
    """(..., 3) float32 -> (..., 4) uint8, image/hdr/hdr.go:26-50 in float32 arithmetic."""
    rgb = np.asarray(rgb, np.float32)
    r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    d = r.copy()
    d = np.where(g > d, g, d)
    d = np.where(b > d, b, d)
    zero = d < np.float32(0.000001)          # NaN compares false: a NaN maximum goes through the arithmetic like in Go
    dd = np.where(zero, np.float32(1), d)
    with np.errstate(all="ignore"):
        nd, e = np.frexp(dd.astype(np.float64))
        n = nd.astype(np.float32)
        df = (n * np.float32(255.999)) / dd
        out = np.stack([_go_byte(r * df), _go_byte(g * df), _go_byte(b * df), ((e + 128) & 0xFF).astype(np.uint8)], -1)
    out[zero] = 0
    return out


def output_hdr_bytes(fb: np.ndarray) -> bytes:
    h, w, _ = fb.shape
    head = "#?RADIANCE\n# Created by Vermeer Light Tools (http://www.vermeerlt.com)\nFORMAT=32-bit_rle_rgbe\n\n+Y %d +X %d\n" % (h, w)
    return head.encode() + rgb_to_rgbe(fb[::-1]).tobytes()
