"""Oracle pins for the QMC sequences (math/ldseq). The reference's own tests hold no vectors for this path
(core/qmc_test.go does not compile and asserts nothing), so the anchors are the published sequence values, the
elementary-interval property of RasterXY, and the known-answer triples recorded in SURVEY.md section 4."""
import ctypes as C

import numpy as np


def test_vdc_sobol_first_points(oracle_lib):
    L = oracle_lib
    assert [L.orc_vdc(i, 0) for i in range(1, 5)] == [0.5, 0.25, 0.75, 0.125]
    assert [L.orc_sobol(i, 0) for i in range(1, 5)] == [0.5, 0.75, 0.25, 0.625]
    assert L.orc_vdc_u(5, 0x123456789abcdef0) == 0xb23456789abcd
    assert L.orc_sobol_u(5, 0x123456789abcdef0) == 0x323456789abcd


def test_raster_xy_known_answers(oracle_lib):
    L = oracle_lib
    kat = [((1, 0, 0), 17895424, 0.5333251953125, 0.5333251953125),
           ((1, 1, 0), 26843136, 1.5999755859375, 0.4705810546875),
           ((2, 0, 0), 35790848, 0.26666259765625, 0.79998779296875),
           ((16, 1919, 1079), 271089646, 1919.8798904418945, 1079.419075012207),
           ((64, 3839, 2159), 1089421303, 3839.8103046417236, 2159.883707046509)]
    rx, ry = C.c_double(), C.c_double()
    for (f, px, py), idx, ex, ey in kat:
        got = L.orc_raster_xy(f, px, py, 0, 0, C.byref(rx), C.byref(ry))
        assert (got, rx.value, ry.value) == (idx, ex, ey)


def test_raster_xy_lands_in_its_pixel(oracle_lib):
    """floor(rx)==px and floor(ry)==py for every (frame, px, py): exercises both GF(2) tables."""
    L = oracle_lib
    rng = np.random.default_rng(0)
    rx, ry = C.c_double(), C.c_double()
    for _ in range(20000):
        f = int(rng.integers(1, 1 << 12))
        px, py = int(rng.integers(0, 4096)), int(rng.integers(0, 4096))
        L.orc_raster_xy(f, px, py, 0, 0, C.byref(rx), C.byref(ry))
        assert int(rx.value) == px and int(ry.value) == py


def test_sequence_is_a_0_2_net_per_pixel(oracle_lib):
    """Within one pixel the first 16 frames stratify the pixel into a 4x4 grid in (rx, ry) — (0,2)-sequence property."""
    L = oracle_lib
    rx, ry = C.c_double(), C.c_double()
    cells = set()
    for f in range(16, 32):
        L.orc_raster_xy(f, 7, 9, 0, 0, C.byref(rx), C.byref(ry))
        cells.add((int((rx.value - 7) * 4), int((ry.value - 9) * 4)))
    assert len(cells) == 16


def test_pixel_filter_tables(oracle_lib):
    """FIS pixel filters (builtin/filter): Bessel J1 against published values, and the warp stays inside the filter support
    and is monotone in r0 within a CDF bin."""
    from oracle.binding import Oracle
    from vermeer_b200 import scenes
    L = oracle_lib
    assert abs(L.orc_bessel_j1(1.0) - 0.4400505857449335) < 1e-8
    assert abs(L.orc_bessel_j1(10.0) - 0.04347274616886144) < 1e-8
    assert abs(L.orc_bessel_j1(-2.5) + 0.4970941024642741) < 1e-8
    # Reference quirk (h): the default AiryFilter (Res 49, Width 6) tabulates the point (0,0) exactly, where
    # 2*J1(v)/v is 0/0 = NaN (airy.go:90-92); the NaN poisons the normalisation and every CDF, so WarpSample never finds a
    # bin and returns (0,0): all samples land on the pixel centre. An even Res avoids the origin and behaves as intended.
    sc = scenes.cornell_box(16, 16)
    sc.filter = scenes.PixelFilter(Type="AiryFilter")
    o = Oracle(sc)
    u, v = C.c_double(), C.c_double()
    for r0, r1 in ((0.1, 0.9), (0.5, 0.5), (0.99, 0.01)):
        L.orc_filter_warp(o.h, r0, r1, C.byref(u), C.byref(v))
        assert (u.value, v.value) == (0.0, 0.0)
    for ftype, w, res in (("AiryFilter", 6.0, 48), ("GaussianFilter", 2.0, None)):
        sc = scenes.cornell_box(16, 16)
        sc.filter = scenes.PixelFilter(Type=ftype, Res=res)
        o = Oracle(sc)
        u, v = C.c_double(), C.c_double()
        rng = np.random.default_rng(3)
        us = []
        for _ in range(2000):
            r0, r1 = rng.random(), rng.random()
            L.orc_filter_warp(o.h, r0, r1, C.byref(u), C.byref(v))
            hi = w / 2 + w / 15 + 1e-9   # bin i maps to -w/2 + w*(i+du)/(n-1), so the last bin reaches w/2 + w/(n-1)
            assert -w / 2 - 1e-9 <= u.value <= hi and -w / 2 - 1e-9 <= v.value <= hi
            us.append(u.value)
        assert np.std(us) > 0.05 * w   # it actually spreads samples
