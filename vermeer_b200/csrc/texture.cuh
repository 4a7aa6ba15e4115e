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

__device__ __forceinline__ float tex_log2(float x) { return (float)log2((double)x); }  // math/pow.go:12-14
__device__ __forceinline__ float tex_exp(float x) { return (float)exp((double)x); }

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
  x0 %= L.w; x1 %= L.w;
  if (x0 < 0) x0 += L.w;
  if (x1 < 0) x1 += L.w;
  y0 %= L.h; y1 %= L.h;
  if (y0 < 0) y0 += L.h;
  if (y1 < 0) y1 += L.h;
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

// texture.go:219-311
__device__ inline void tex_sample_rgb(const DevTexStore& ts, const DevTexture& T, const TexCoord& sg, float out[3]) {
  float tx0 = sg.dudx * sg.pd0, tx1 = sg.dvdx * sg.pd0;
  float ty0 = sg.dudy * sg.pd1, ty1 = sg.dvdy * sg.pd1;
  tx0 = tx0 * (float)T.w;
  ty0 = ty0 * (float)T.w;
  tx1 = tx1 * (float)T.h;
  ty1 = ty1 * (float)T.h;
  const float ds = sqrtf(tx0 * tx0 + tx1 * tx1);
  const float dt = sqrtf(ty0 * ty0 + ty1 * ty1);
  float lod = tex_log2(ds > dt ? ds : dt);  // m.Max = MAXSS: the second operand when unordered
  const float maxlod = (float)(T.n_levels - 1);
  if (lod > maxlod) lod = maxlod;
  if (lod < 0) lod = 0;
  tex_trilinear(ts, T, sg.U, sg.V, lod, out);
  out[0] /= 255.0f;
  out[1] /= 255.0f;
  out[2] /= 255.0f;
}

// feline.go:25-151
__device__ inline void tex_sample_feline(const DevTexStore& ts, const DevTexture& T, const TexCoord& sc, float c[3]) {
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
  const float A = Ann / F;
  const float B = Bnn / F;
  const float C = Cnn / F;
  const float amc = A - C;
  const float root = sqrtf(amc * amc + B * B);
  const float Aprm = (A + C - root) / 2;
  const float Cprm = (A + C + root) / 2;
  float majorRadius = sqrtf(1 / Aprm);
  float minorRadius = sqrtf(1 / Cprm);
  float theta = (float)atan((double)(B / amc)) / 2;
  if (A > C) theta = theta + 3.14159265358f / 2;
  minorRadius = minorRadius > 1.0f ? minorRadius : 1.0f;
  majorRadius = majorRadius > 1.0f ? majorRadius : 1.0f;
  const float fProbes = 2 * (majorRadius / minorRadius) - 1;
  float iProbes = floorf(fProbes + 0.5f);
  iProbes = iProbes < 16.0f ? iProbes : 16.0f;
  if (iProbes < fProbes) minorRadius = 2 * majorRadius / (iProbes + 1);
  float lod = tex_log2(minorRadius);
  const float maxlod = (float)(T.n_levels - 1);
  if (lod > maxlod) {
    lod = maxlod;
    iProbes = 1;
  }
  if (lod < 0) lod = 0;
  const float lineLength = 2 * (majorRadius - minorRadius);
  float dU = (float)cos((double)theta) * lineLength / (iProbes - 1);
  float dV = (float)sin((double)theta) * lineLength / (iProbes - 1);
  const int nProbes = (int)iProbes;
  if (nProbes == 1) {
    dU = 0;
    dV = 0;
  }
  float n = (float)(-(nProbes - 1));
  const float alpha = 0.6f;
  float accum[3] = {0, 0, 0};
  float accumWeight = 0;
  const float mr2 = majorRadius * majorRadius;
  for (int i = 0; i < nProbes; i++) {
    const float u = fw * sc.U + (n / 2) * dU;
    const float v = fh * sc.V + (n / 2) * dV;
    const float d2 = ((n * n) / 4) * (dU * dU + dV * dV) / mr2;
    const float relativeWeight = tex_exp(-alpha * d2);
    float sample[3];
    tex_trilinear(ts, T, u / fw, v / fh, lod, sample);
#pragma unroll
    for (int k = 0; k < 3; k++) accum[k] += (sample[k] / 255.0f) * relativeWeight;
    accumWeight += relativeWeight;
    n += 2;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) c[k] = accum[k] / accumWeight;
}

#define VG_TEXFILTER_FELINE_ 0
__device__ inline void tex_sample(const DevTexStore& ts, int tex, int filter, const TexCoord& tc, float out[3]) {
  const DevTexture T = ts.textures[tex];
  if (filter == 1) tex_sample_rgb(ts, T, tc, out);
  else tex_sample_feline(ts, T, tc, out);
}

// ---- kernels (texture.cu) --------------------------------------------------------------------------------------------
cudaError_t launch_mip_level(uchar4* texels, DevTexLevel src, DevTexLevel dst, cudaStream_t stream);
cudaError_t launch_texture_sample(const DevTexStore& ts, int tex, int filter, const float* d_coords8, long long n, float* d_out3, cudaStream_t stream);

}  // namespace vg
