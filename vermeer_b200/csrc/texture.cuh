// Texture maps on the device: the reference's mip pyramid, bilinear/trilinear taps and the Feline anisotropic filter.
//   texture/mipmap.go:15-58     miplevel.BilinearSample (wrap mode, taps at floor/ceil of the texel coordinate)
//   texture/mipmap.go:69-104    mipmap.TrilinearSample
//   texture/mipmap.go:122-315   stdfilter (pyramid: 2x2 box for even sizes, NP2 polyphase weights for odd ones)
//   texture/texture.go:219-311  SampleRGB   (maps.TextureTrilinear)
//   texture/feline.go:25-151    SampleFeline (maps.Texture, the default)
//
// Layout in HBM: every level of every texture lives in one uchar4 array (RGB + one pad byte, so a tap is one 32-bit load instead
// of the reference's three byte loads), rows bottom-up like texture.Texture.data; `DevTexLevel` says where a level starts.
// The hardware texture units are not used: their 9-bit fixed-point filter weights and texel-centre convention differ from the
// reference's float arithmetic (integer coordinates ARE texel centres here), and parity is the contract.
// The pyramid itself is built on the device (k_mip_level), one thread per output texel, with the reference's float operation
// order and no contraction (-fmad=false), byte-identical to stdfilter.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vg {

struct DevTexLevel {
  uint32_t off;  // first texel in the shared texel array
  int32_t w, h;
  int32_t pad;
};
struct DevTexture {
  int32_t w, h;
  int32_t first_level;  // into the level table
  int32_t n_levels;     // ceil(log2(max(w,h))) (mipmap.go:123-125)
};
struct DevTexStore {
  const uchar4* texels;
  const DevTexLevel* levels;
  const DevTexture* textures;
  int32_t n_textures;
};

// What a map reads of the ShaderContext (U, V, Dduvdx, Dduvdy) and of core.Image (PixelDelta).
struct TexCoord {
  float U, V;
  float dudx, dvdx, dudy, dvdy;
  float pd0, pd1;
};

// math/pow.go:12-19, sincos.go:16-64: float32 -> float64 stdlib -> float32. `fast` = the single-precision libm (1-2 ulp), the same
// switch as the shading kernels' trig ("precise_trig" option); both are far inside the image tolerance.
// `fast` also replaces the correctly rounded divisions of the footprint analysis and of the probes by the approximate
// reciprocal (2 ulp): everything here is downstream of a normalize, i.e. tolerance-only either way, and a correctly rounded
// float division is ~10 instructions against 2.
__device__ __forceinline__ float tex_div(float a, float b, bool fast) { return fast ? __fdividef(a, b) : a / b; }
__device__ __forceinline__ float tex_log2(float x, bool fast) { return fast ? log2f(x) : (float)log2((double)x); }
__device__ __forceinline__ float tex_exp(float x, bool fast) { return fast ? expf(x) : (float)exp((double)x); }
__device__ __forceinline__ float tex_atan(float x, bool fast) { return fast ? atanf(x) : (float)atan((double)x); }
__device__ __forceinline__ void tex_sincos(float x, bool fast, float* sn, float* cs) {
  if (fast) {
    sincosf(x, sn, cs);
  } else {
    *sn = (float)sin((double)x);
    *cs = (float)cos((double)x);
  }
}

// mipmap.go:15-58
__device__ inline void tex_bilinear(const DevTexStore& ts, const DevTexLevel& L, float s, float t, float c[3]) {
  const float ms = s - floorf(s);
  const float mt = t - floorf(t);
  const float fw = (float)L.w, fh = (float)L.h;
  int x0 = (int)floorf(ms * fw);
  int x1 = (int)ceilf(ms * fw);
  const float dx = ms * fw - floorf(ms * fw);
  int y0 = (int)floorf(mt * fh);
  int y1 = (int)ceilf(mt * fh);
  const float dy = mt * fh - floorf(mt * fh);
  // x0 %= w etc. (mipmap.go:30-47): ms is in [0,1), so the texel coordinates are in [0, w] and the modulo only ever folds w
  // onto 0 (ms*w can round up to w; ceil reaches w on the last texel) — a compare instead of an integer division. NaN
  // coordinates convert to 0 here where amd64's CVTTSS2SQ gives the "integer indefinite"; the colour is NaN either way.
  x0 = (x0 >= L.w || x0 < 0) ? 0 : x0;
  x1 = (x1 >= L.w || x1 < 0) ? 0 : x1;
  y0 = (y0 >= L.h || y0 < 0) ? 0 : y0;
  y1 = (y1 >= L.h || y1 < 0) ? 0 : y1;
  const uchar4* base = ts.texels + L.off;
  const uchar4 t00 = __ldg(base + x0 + y0 * L.w), t10 = __ldg(base + x1 + y0 * L.w);
  const uchar4 t01 = __ldg(base + x0 + y1 * L.w), t11 = __ldg(base + x1 + y1 * L.w);
  const float a[3] = {(float)t00.x, (float)t00.y, (float)t00.z}, b[3] = {(float)t10.x, (float)t10.y, (float)t10.z};
  const float d[3] = {(float)t01.x, (float)t01.y, (float)t01.z}, e[3] = {(float)t11.x, (float)t11.y, (float)t11.z};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float c0 = (1 - dx) * a[k] + dx * b[k];
    const float c1 = (1 - dx) * d[k] + dx * e[k];
    c[k] = (1 - dy) * c0 + dy * c1;
  }
}

// mipmap.go:69-104
__device__ inline void tex_trilinear(const DevTexStore& ts, const DevTexture& T, float s, float t, float lod, float c[3]) {
  int l0 = (int)ceilf(lod);
  int l1 = (int)floorf(lod);
  const float dl = lod - floorf(lod);
  if (l0 < 0) l0 = 0;
  if (l0 > T.n_levels - 1) l0 = T.n_levels - 1;
  if (l1 < 0) l1 = 0;
  if (l1 > T.n_levels - 1) l1 = T.n_levels - 1;
  if (l1 == l0) {
    tex_bilinear(ts, ts.levels[T.first_level + l0], s, t, c);
    return;
  }
  float c0[3], c1[3];
  tex_bilinear(ts, ts.levels[T.first_level + l0], s, t, c0);
  tex_bilinear(ts, ts.levels[T.first_level + l1], s, t, c1);
#pragma unroll
  for (int k = 0; k < 3; k++) c[k] = dl * c0[k] + (1 - dl) * c1[k];
}

// One lookup = a set-up (the footprint analysis, once) and 1..16 probes (a trilinear tap and its weight each). The split is what
// lets a warp share the probes of its 32 lookups evenly (tex_sample_warp): Feline's probe count follows the anisotropy of the
// footprint, so lane-per-lookup execution leaves most of a warp idle (ncu: 6.6 of 32 threads active in the taps).
struct TexProbeSetup {
  int32_t tex;      // texture id; < 0: no lookup
  int32_t nprobes;  // 1 for the trilinear filter
  int32_t feline;   // bit 0: Feline lookup; bits 1 / 2: fast arithmetic (Feline / trilinear lookup)
  float U, V;       // feline: w*U, h*V (texel units); trilinear: U, V
  float dU, dV, lod, mr2, n0;
};

// texture.go:219-311 up to the tap (SampleRGB)
__device__ inline TexProbeSetup tex_setup_rgb(const DevTexture& T, int tex, const TexCoord& sg, bool fast) {
  float tx0 = sg.dudx * sg.pd0, tx1 = sg.dvdx * sg.pd0;
  float ty0 = sg.dudy * sg.pd1, ty1 = sg.dvdy * sg.pd1;
  tx0 = tx0 * (float)T.w;
  ty0 = ty0 * (float)T.w;
  tx1 = tx1 * (float)T.h;
  ty1 = ty1 * (float)T.h;
  const float ds = sqrtf(tx0 * tx0 + tx1 * tx1);
  const float dt = sqrtf(ty0 * ty0 + ty1 * ty1);
  float lod = tex_log2(ds > dt ? ds : dt, fast);  // m.Max = MAXSS: the second operand when unordered
  const float maxlod = (float)(T.n_levels - 1);
  if (lod > maxlod) lod = maxlod;
  if (lod < 0) lod = 0;
  TexProbeSetup s;
  s.tex = tex; s.nprobes = 1; s.feline = fast ? 4 : 0;  // bit 2: fast arithmetic for a trilinear lookup (bit 0 stays clear)
  s.U = sg.U; s.V = sg.V; s.dU = 0; s.dV = 0; s.lod = lod; s.mr2 = 1; s.n0 = 0;
  return s;
}

// feline.go:25-116 (SampleFeline up to the probe loop)
__device__ inline TexProbeSetup tex_setup_feline(const DevTexture& T, int tex, const TexCoord& sc, bool fast) {
  float ux = sc.dudx * sc.pd0, vx = sc.dvdx * sc.pd0;
  float uy = sc.dudy * sc.pd1, vy = sc.dvdy * sc.pd1;
  const float fw = (float)T.w, fh = (float)T.h;
  ux = ux * fw;
  uy = uy * fw;
  vx = vx * fh;
  vy = vy * fh;
  const float Ann = vx * vx + vy * vy;
  const float Bnn = -2 * (ux * vx + uy * vy);
  const float Cnn = ux * ux + uy * uy;
  const float F = Ann * Cnn - (Bnn * Bnn / 4);
  const float A = tex_div(Ann, F, fast);
  const float B = tex_div(Bnn, F, fast);
  const float C = tex_div(Cnn, F, fast);
  const float amc = A - C;
  const float root = sqrtf(amc * amc + B * B);
  const float Aprm = (A + C - root) / 2;
  const float Cprm = (A + C + root) / 2;
  float majorRadius = sqrtf(tex_div(1, Aprm, fast));
  float minorRadius = sqrtf(tex_div(1, Cprm, fast));
  float theta = tex_atan(tex_div(B, amc, fast), fast) / 2;
  if (A > C) theta = theta + 3.14159265358f / 2;
  minorRadius = minorRadius > 1.0f ? minorRadius : 1.0f;
  majorRadius = majorRadius > 1.0f ? majorRadius : 1.0f;
  const float fProbes = 2 * tex_div(majorRadius, minorRadius, fast) - 1;
  float iProbes = floorf(fProbes + 0.5f);
  iProbes = iProbes < 16.0f ? iProbes : 16.0f;
  if (iProbes < fProbes) minorRadius = tex_div(2 * majorRadius, iProbes + 1, fast);
  float lod = tex_log2(minorRadius, fast);
  const float maxlod = (float)(T.n_levels - 1);
  if (lod > maxlod) {
    lod = maxlod;
    iProbes = 1;
  }
  if (lod < 0) lod = 0;
  const float lineLength = 2 * (majorRadius - minorRadius);
  float sn, cs;
  tex_sincos(theta, fast, &sn, &cs);
  float dU = tex_div(cs * lineLength, iProbes - 1, fast);
  float dV = tex_div(sn * lineLength, iProbes - 1, fast);
  // int(iProbes): a NaN footprint (parallel derivatives: F = 0) reaches here as NaN, which amd64 converts to the most negative
  // integer, i.e. no probe and 0/0 = NaN for the colour; `0 probes` gives the same NaN
  const int nProbes = iProbes == iProbes ? (int)iProbes : 0;
  if (nProbes == 1) {
    dU = 0;
    dV = 0;
  }
  TexProbeSetup s;
  s.tex = tex; s.nprobes = nProbes; s.feline = fast ? 3 : 1;
  s.U = fw * sc.U; s.V = fh * sc.V; s.dU = dU; s.dV = dV; s.lod = lod; s.mr2 = majorRadius * majorRadius;
  s.n0 = (float)(-(nProbes - 1));
  return s;
}

__device__ inline TexProbeSetup tex_setup(const DevTexStore& ts, int tex, int filter, const TexCoord& tc, bool fast = false) {
  const DevTexture T = ts.textures[tex];
  return filter == 1 ? tex_setup_rgb(T, tex, tc, fast) : tex_setup_feline(T, tex, tc, fast);
}

// probe i of a lookup: Feline: (sample / 255) * weight in .xyz and the weight in .w (feline.go:118-137); trilinear: sample / 255
// (texture.go:305-308). The divisions sit here, with the taps, so that the owner's in-order pass is additions only.
__device__ inline float4 tex_probe(const DevTexStore& ts, const TexProbeSetup& s, int i) {
  const DevTexture T = ts.textures[s.tex];
  float sample[3];
  float w = 1.0f;
  if (s.feline & 1) {
    const float fw = (float)T.w, fh = (float)T.h;
    const float n = s.n0 + (float)(2 * i);  // the reference's running n += 2: small integers, exact either way
    const float u = s.U + (n / 2) * s.dU;
    const float v = s.V + (n / 2) * s.dV;
    const bool fast = (s.feline & 2) != 0;
    const float d2 = tex_div(((n * n) / 4) * (s.dU * s.dU + s.dV * s.dV), s.mr2, fast);
    w = tex_exp(-0.6f * d2, fast);
    tex_trilinear(ts, T, tex_div(u, fw, fast), tex_div(v, fh, fast), s.lod, sample);
    if (fast) {
      const float k = w * (1.0f / 255.0f);
      return make_float4(sample[0] * k, sample[1] * k, sample[2] * k, w);
    }
    return make_float4((sample[0] / 255.0f) * w, (sample[1] / 255.0f) * w, (sample[2] / 255.0f) * w, w);
  }
  tex_trilinear(ts, T, s.U, s.V, s.lod, sample);
  if (s.feline & 4) return make_float4(sample[0] * (1.0f / 255.0f), sample[1] * (1.0f / 255.0f), sample[2] * (1.0f / 255.0f), w);
  return make_float4(sample[0] / 255.0f, sample[1] / 255.0f, sample[2] / 255.0f, w);
}

// the owner's side: probes in order (feline.go:133-147), or the single tap / 255 (texture.go:305-308)
struct TexAccum {
  float a0, a1, a2, w;
  __device__ __forceinline__ void init() { a0 = a1 = a2 = w = 0.0f; }
  __device__ __forceinline__ void add(const TexProbeSetup& s, const float4 r) {
    if (s.feline & 1) {
      a0 += r.x;
      a1 += r.y;
      a2 += r.z;
      w += r.w;
    } else {
      a0 = r.x;
      a1 = r.y;
      a2 = r.z;
    }
  }
  __device__ __forceinline__ void finish(const TexProbeSetup& s, float out[3]) const {
    if (s.feline & 2) {
      const float r = __fdividef(1.0f, w);
      out[0] = a0 * r;
      out[1] = a1 * r;
      out[2] = a2 * r;
    } else if (s.feline & 1) {
      out[0] = a0 / w;
      out[1] = a1 / w;
      out[2] = a2 / w;
    } else {
      out[0] = a0; out[1] = a1; out[2] = a2;
    }
  }
};

// one lane, one lookup (the batch entry point and the reference for the cooperative form)
__device__ inline void tex_sample(const DevTexStore& ts, int tex, int filter, const TexCoord& tc, float out[3]) {
  const TexProbeSetup s = tex_setup(ts, tex, filter, tc);
  TexAccum acc;
  acc.init();
  for (int i = 0; i < s.nprobes; i++) acc.add(s, tex_probe(ts, s, i));
  acc.finish(s, out);
}

// Warp-cooperative form: all 32 lanes call it; lanes with `mine.tex >= 0` own a lookup. The probes of the warp's lookups are
// numbered consecutively (exclusive prefix sum of the probe counts) and taken 32 at a time, one per lane; each owner then adds
// the results of its own probes in probe order, so the sums are bit-identical to tex_sample's.
struct TexWarpScratch {  // per warp, shared memory
  TexProbeSetup setup[32];
  int incl[32];
  float4 res[32];
};
__device__ inline void tex_sample_warp(const DevTexStore& ts, const TexProbeSetup& mine, TexWarpScratch* sh, float out[3]) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int n = mine.tex >= 0 ? mine.nprobes : 0;
  int incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(full, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(full, incl, 31);
  sh->setup[lane] = mine;
  sh->incl[lane] = incl;
  __syncwarp();
  const int begin = incl - n;
  TexAccum acc;
  acc.init();
  for (int base = 0; base < total; base += 32) {
    const int j = base + lane;
    if (j < total) {
      int lo = 0, hi = 31;  // owner = first lane whose inclusive count exceeds j
#pragma unroll
      for (int step = 0; step < 5; step++) {
        const int mid = (lo + hi) >> 1;
        if (sh->incl[mid] > j) hi = mid;
        else lo = mid + 1;
      }
      const TexProbeSetup& s = sh->setup[lo];
      sh->res[lane] = tex_probe(ts, s, j - (sh->incl[lo] - s.nprobes));
    }
    __syncwarp();
    const int a = begin > base ? begin : base;
    const int b = begin + n < base + 32 ? begin + n : base + 32;
    for (int q = a; q < b; q++) acc.add(mine, sh->res[q - base]);
    __syncwarp();
  }
  if (n > 0 || mine.tex >= 0) acc.finish(mine, out);
}

// ---- kernels (texture.cu) --------------------------------------------------------------------------------------------
cudaError_t launch_mip_level(uchar4* texels, DevTexLevel src, DevTexLevel dst, cudaStream_t stream);
cudaError_t launch_texture_sample(const DevTexStore& ts, int tex, int filter, const float* d_coords8, long long n, float* d_out3, cudaStream_t stream);

}  // namespace vg
