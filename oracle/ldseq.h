// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// CPU restatement of the reference's deterministic QMC sequences (integer exact).
// Follows:
//   math/ldseq/ldseq.go:50-96   (VanDerCorput / Sobol, 52-bit radical inverses)
//   math/ldseq/raster.go:10-57  (lookup / RasterXY: global (0,2)-sequence stratified per pixel)
//   math/ldseq/vdc_sobol_matrices52.go:470-499, 865-890 (rows m=12, the only rows the renderer
//   uses: core/render.go:89 calls RasterXY(12, ...))
// Pinned by: the sequence invariants VdC(1..4,0)=.5,.25,.75,.125, Sobol(1..4,0)=.5,.75,.25,.625 and
// the elementary-interval property floor(rx)==px && floor(ry)==py (tests/test_oracle_ldseq.py).
#pragma once
#include <cstdint>

namespace orc {

static inline uint64_t vanDerCorput_u(uint64_t i, uint64_t scramble) {
  uint64_t bits = (i << 32) | (i >> 32);
  bits = ((bits & 0x0000ffff0000ffffull) << 16) | ((bits & 0xffff0000ffff0000ull) >> 16);
  bits = ((bits & 0x00ff00ff00ff00ffull) << 8) | ((bits & 0xff00ff00ff00ff00ull) >> 8);
  bits = ((bits & 0x0f0f0f0f0f0f0f0full) << 4) | ((bits & 0xf0f0f0f0f0f0f0f0ull) >> 4);
  bits = ((bits & 0x3333333333333333ull) << 2) | ((bits & 0xccccccccccccccccull) >> 2);
  bits = ((bits & 0x5555555555555555ull) << 1) | ((bits & 0xaaaaaaaaaaaaaaaaull) >> 1);
  return (scramble ^ bits) >> (64 - 52);
}
static inline double VanDerCorput(uint64_t i, uint64_t scramble) {
  return (double)vanDerCorput_u(i, scramble) / (double)(1ull << 52);
}
static inline uint64_t sobol_u(uint64_t i, uint64_t scramble) {
  uint64_t r = scramble >> (64 - 52);
  for (uint64_t v = 1ull << (52 - 1); i != 0; i >>= 1) {
    if (i & 1) r ^= v;
    v ^= v >> 1;
  }
  return r;
}
static inline double Sobol(uint64_t i, uint64_t scramble) { return (double)sobol_u(i, scramble) / (double)(1ull << 52); }

// Rows m=12 of the two GF(2) tables.
static const uint64_t kVdcSobolM12[28] = {0x808, 0xc0c, 0xa0a, 0xf0f, 0x888, 0xccc, 0xaaa, 0xfff, 0x800, 0xc00,
                                          0xa00, 0xf00, 0x880, 0xcc0, 0xaa0, 0xff0, 0x808, 0xc0c, 0xa0a, 0xf0f,
                                          0x888, 0xccc, 0xaaa, 0xfff, 0x800, 0xc00, 0xa00, 0xf00};
static const uint64_t kVdcSobolInvM12[24] = {0xf0f000, 0x505000, 0x303000, 0x101000, 0xff0000, 0x550000, 0x330000, 0x110000,
                                             0xf0000,  0x50000,  0x30000,  0x10000,  0x888800, 0x444400, 0x222200, 0x111100,
                                             0x800080, 0x400040, 0x200020, 0x100010, 0x80008,  0x40004,  0x20002,  0x10001};

// raster.go:10-41 with m fixed to 12
static inline uint64_t lookup12(uint32_t frame, uint32_t px, uint32_t py, uint64_t scrambleX, uint64_t scrambleY) {
  const uint32_t m = 12;
  uint32_t m2 = m << 1;
  uint64_t index = (uint64_t)frame << m2;
  uint64_t delta = 0;
  for (uint32_t c = 0; frame != 0; frame >>= 1) {
    if (frame & 1) delta ^= kVdcSobolM12[c];
    c++;
  }
  px ^= (uint32_t)(scrambleX >> (64 - m));
  py ^= (uint32_t)(scrambleY >> (64 - m));
  uint64_t b = (((uint64_t)px << m) | (uint64_t)py) ^ delta;
  for (uint32_t c = 0; b != 0; b >>= 1) {
    if (b & 1) index ^= kVdcSobolInvM12[c];
    c++;
  }
  return index;
}
// raster.go:50-57
static inline uint64_t RasterXY12(uint32_t frame, uint32_t px, uint32_t py, uint64_t scrambleX, uint64_t scrambleY, double* rx, double* ry) {
  uint64_t index = lookup12(frame, px, py, scrambleX, scrambleY);
  *rx = (double)vanDerCorput_u(index, scrambleX) / (double)(1ull << (52 - 12));
  *ry = (double)sobol_u(index, scrambleY) / (double)(1ull << (52 - 12));
  return index;
}

}  // namespace orc
