// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// CPU restatement of the reference's texture store, mip pyramid and the two texture filters.
// parity unpinned (the reference has no tests under texture/); anchored by hand-worked known answers in
// tests/test_oracle_texture.py (2x2 / 4x4 pyramids, texel-centre and wrap-around bilinear taps, isotropic and 4:1 footprints).
// Follows:
//   texture/texture.go:36-42,164-176        (Texture, SetRGB, CreateRGBTexture; rows are stored bottom-up: loadTexture :139 flips)
//   texture/texture.go:219-311              (SampleRGB: trilinear filter, LOD from the longer footprint axis)
//   texture/mipmap.go:10-58                 (miplevel.BilinearSample, wrap mode)
//   texture/mipmap.go:60-104                (mipmap.TrilinearSample)
//   texture/mipmap.go:122-315               (stdfilter: the pyramid; box filter for even sizes, NVIDIA NP2 polyphase weights for odd ones)
//   texture/feline.go:25-151                (SampleFeline, WRL-99-1)
//   builtin/maps/texture.go:17-83           (Texture / TextureTrilinear maps; "?filter=trilinear&ch=N" in the file name)
// Image decoding (image/png, jpeg, tiff, tga) is outside the path: textures arrive as RGB8 rows, already flipped.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "vmath.h"

namespace orc {

struct MipLevel {
  int w = 0, h = 0;
  std::vector<uint8_t> mipmap;

  // mipmap.go:15-58
  void BilinearSample(float s, float t, int cp, float c[3]) const {
    const float ms = s - Floor(s);
    const float mt = t - Floor(t);
    int x0 = (int)Floor(ms * (float)w);
    int x1 = (int)Ceil(ms * (float)w);
    const float dx = ms * (float)w - Floor(ms * (float)w);
    int y0 = (int)Floor(mt * (float)h);
    int y1 = (int)Ceil(mt * (float)h);
    const float dy = mt * (float)h - Floor(mt * (float)h);
    x0 %= w;
    x1 %= w;
    if (x0 < 0) x0 += w;
    if (x1 < 0) x1 += w;
    y0 %= h;
    y1 %= h;
    if (y0 < 0) y0 += h;
    if (y1 < 0) y1 += h;
    for (int k = 0; k < 3; k++) {
      const float c0 = (1 - dx) * (float)mipmap[(x0 + (y0 * w)) * cp + k] + dx * (float)mipmap[(x1 + (y0 * w)) * cp + k];
      const float c1 = (1 - dx) * (float)mipmap[(x0 + (y1 * w)) * cp + k] + dx * (float)mipmap[(x1 + (y1 * w)) * cp + k];
      c[k] = (1 - dy) * c0 + dy * c1;
    }
  }
};

static inline int maxi(int a, int b) { return a > b ? a : b; }
static inline int mini(int a, int b) { return a < b ? a : b; }

struct MipMap {
  int components = 3;
  std::vector<MipLevel> mipmap;
  int MaxLevelOfDetail() const { return (int)mipmap.size() - 1; }

  // mipmap.go:69-104
  void TrilinearSample(float s, float t, float lod, float c[3]) const {
    int l0 = (int)Ceil(lod);
    int l1 = (int)Floor(lod);
    const float dl = lod - Floor(lod);
    const int n = (int)mipmap.size();
    if (l0 < 0) l0 = 0;
    if (l0 > n - 1) l0 = n - 1;
    if (l1 < 0) l1 = 0;
    if (l1 > n - 1) l1 = n - 1;
    if (l1 == l0) {
      mipmap[l0].BilinearSample(s, t, components, c);
      return;
    }
    float c0[3], c1[3];
    mipmap[l0].BilinearSample(s, t, components, c0);
    mipmap[l1].BilinearSample(s, t, components, c1);
    for (int k = 0; k < 3; k++) c[k] = dl * c0[k] + (1 - dl) * c1[k];
  }
};

// mipmap.go:122-315. The branch is chosen by the parity of the NEW level's size, as the reference does.
static inline MipMap stdfilter(int w, int h, const std::vector<uint8_t>& img, int components) {
  MipMap out;
  const int maxlevel = (int)Ceil(Log2(Max((float)w, (float)h)));
  out.mipmap.resize(maxlevel > 0 ? maxlevel : 0);
  if (maxlevel <= 0) return out;  // the reference indexes mipmap[0] here and panics for a 1x1 image
  out.mipmap[0].mipmap = img;
  out.mipmap[0].w = w;
  out.mipmap[0].h = h;
  int width = w, height = h;
  for (int l = 1; l < maxlevel; l++) {
    const int nwidth = (int)Max(1, Ceil((float)width / 2));
    const int nheight = (int)Max(1, Ceil((float)height / 2));
    MipLevel& L = out.mipmap[l];
    const std::vector<uint8_t>& P = out.mipmap[l - 1].mipmap;
    L.mipmap.assign((size_t)nwidth * nheight * components, 0);
    L.w = nwidth;
    L.h = nheight;
    auto px = [&](int x, int y, int k) -> float { return (float)P.at((size_t)(x + y * width) * components + k); };
    if (nheight % 2 == 0) {
      if (nwidth % 2 == 0) {
        for (int y = 0; y < nheight; y++) {
          const int y0 = y * 2, y1 = mini(y0 + 1, maxi(1, height - 1));
          for (int x = 0; x < nwidth; x++) {
            const int x0 = x * 2, x1 = mini(x0 + 1, maxi(1, width - 1));
            for (int k = 0; k < components; k++)
              L.mipmap[(x + y * nwidth) * components + k] = (uint8_t)(0.25f * (px(x0, y0, k) + px(x0, y1, k) + px(x1, y0, k) + px(x1, y1, k)));
          }
        }
      } else {  // height even, width odd
        for (int y = 0; y < nheight; y++) {
          const int y0 = y * 2, y1 = mini(y0 + 1, maxi(1, height - 1));
          for (int x = 0; x < nwidth; x++) {
            int x0 = maxi(x * 2 - 1, -maxi(1, width - 1));
            const int x1 = x * 2;
            int x2 = mini(x * 2 + 1, maxi(1, width - 1));
            if (x0 < 0) { x0 += width; x0 = mini(x0, maxi(1, width - 1)); }
            if (x2 > width - 1) { x2 -= width; x2 = maxi(x2, -maxi(1, width - 1)); }
            const float w0 = (float)(nwidth - x - 1) / (float)(2 * nwidth - 1);
            const float w1 = (float)(nwidth) / (float)(2 * nwidth - 1);
            const float w2 = (float)(x) / (float)(2 * nwidth - 1);
            for (int k = 0; k < components; k++) {
              const float c00 = px(x0, y0, k), c10 = px(x1, y0, k), c20 = px(x2, y0, k);
              const float c01 = px(x0, y1, k), c11 = px(x1, y1, k), c21 = px(x2, y1, k);
              L.mipmap[(x + y * nwidth) * components + k] = (uint8_t)(0.5f * (w0 * c00 + w1 * c10 + w2 * c20 + w0 * c01 + w1 * c11 + w2 * c21));
            }
          }
        }
      }
    } else {
      if (nwidth % 2 == 0) {  // height odd, width even
        for (int y = 0; y < nheight; y++) {
          int y0 = maxi(y * 2 - 1, -maxi(1, height - 1));
          const int y1 = y * 2;
          const int y2 = mini(y * 2 + 1, maxi(1, height - 1));
          if (y0 < 0) { y0 += height; y0 = mini(y0, maxi(1, height - 1)); }
          const float w0 = (float)(nheight - y - 1) / (float)(2 * nheight - 1);
          const float w1 = (float)(nheight) / (float)(2 * nheight - 1);
          const float w2 = (float)(y) / (float)(2 * nheight - 1);
          for (int x = 0; x < nwidth; x++) {
            const int x0 = x * 2, x1 = mini(x0 + 1, maxi(1, width - 1));
            for (int k = 0; k < components; k++) {
              const float c00 = px(x0, y0, k), c01 = px(x0, y1, k), c02 = px(x0, y2, k);
              const float c10 = px(x1, y0, k), c11 = px(x1, y1, k), c12 = px(x1, y2, k);
              L.mipmap[(x + y * nwidth) * components + k] = (uint8_t)(0.5f * (w0 * c00 + w1 * c01 + w2 * c02 + w0 * c10 + w1 * c11 + w2 * c12));
            }
          }
        }
      } else {
        for (int y = 0; y < nheight; y++) {
          int y0 = maxi(y * 2 - 1, -maxi(1, height - 1));
          const int y1 = y * 2;
          const int y2 = mini(y * 2 + 1, maxi(1, height - 1));
          if (y0 < 0) { y0 += height; y0 = mini(y0, maxi(1, height - 1)); }
          const float wy0 = (float)(nheight - y - 1) / (float)(2 * nheight - 1);
          const float wy1 = (float)(nheight) / (float)(2 * nheight - 1);
          const float wy2 = (float)(y) / (float)(2 * nheight - 1);
          for (int x = 0; x < nwidth; x++) {
            int x0 = maxi(x * 2 - 1, -maxi(1, width - 1));
            const int x1 = x * 2;
            int x2 = mini(x * 2 + 1, maxi(1, width - 1));
            if (x0 < 0) { x0 += width; x0 = mini(x0, maxi(1, width - 1)); }
            if (x2 > width - 1) { x2 -= width; x2 = maxi(x2, -maxi(1, width - 1)); }
            const float w0 = (float)(nwidth - x - 1) / (float)(2 * nwidth - 1);
            const float w1 = (float)(nwidth) / (float)(2 * nwidth - 1);
            const float w2 = (float)(x) / (float)(2 * nwidth - 1);
            for (int k = 0; k < components; k++) {
              const float c00 = px(x0, y0, k), c01 = px(x0, y1, k), c02 = px(x0, y2, k);
              const float c10 = px(x1, y0, k), c11 = px(x1, y1, k), c12 = px(x1, y2, k);
              const float c20 = px(x2, y0, k), c21 = px(x2, y1, k), c22 = px(x2, y2, k);
              L.mipmap[(x + y * nwidth) * components + k] =
                  (uint8_t)(wy0 * (w0 * c00 + w1 * c10 + w2 * c20) + wy1 * (w0 * c01 + w1 * c11 + w2 * c21) + wy2 * (w0 * c02 + w1 * c12 + w2 * c22));
            }
          }
        }
      }
    }
    width = nwidth;
    height = nheight;
  }
  out.components = components;
  return out;
}

// texture.go:36-42
struct Texture {
  std::string url;
  int w = 0, h = 0;
  std::vector<uint8_t> data;
  MipMap mipmap;
};

// What a sampler reads of the ShaderContext (U, V, Dduvdx, Dduvdy, Image.PixelDelta).
struct TexCoord {
  float U, V;
  float Dduvdx[2], Dduvdy[2];
  float PixelDelta[2];
};

// texture.go:219-311
static inline void SampleRGB(const Texture* img, const TexCoord& sg, float out[3]) {
  float deltaTx[2] = {sg.Dduvdx[0] * sg.PixelDelta[0], sg.Dduvdx[1] * sg.PixelDelta[0]};
  float deltaTy[2] = {sg.Dduvdy[0] * sg.PixelDelta[1], sg.Dduvdy[1] * sg.PixelDelta[1]};
  deltaTx[0] = deltaTx[0] * (float)img->w;
  deltaTy[0] = deltaTy[0] * (float)img->w;
  deltaTx[1] = deltaTx[1] * (float)img->h;
  deltaTy[1] = deltaTy[1] * (float)img->h;
  const float ds = Sqrt(deltaTx[0] * deltaTx[0] + deltaTx[1] * deltaTx[1]);
  const float dt = Sqrt(deltaTy[0] * deltaTy[0] + deltaTy[1] * deltaTy[1]);
  float lod = Log2(Max(ds, dt));
  if (lod > (float)img->mipmap.MaxLevelOfDetail()) lod = (float)img->mipmap.MaxLevelOfDetail();
  if (lod < 0) lod = 0;
  img->mipmap.TrilinearSample(sg.U, sg.V, lod, out);
  out[0] /= 255.0f;
  out[1] /= 255.0f;
  out[2] /= 255.0f;
}

static inline float tex_sqr(float x) { return x * x; }

// feline.go:25-151
static inline void SampleFeline(const Texture* img, const TexCoord& sc, float c[3]) {
  const int maxProbes = 16;
  float Dduvdx[2] = {sc.Dduvdx[0] * sc.PixelDelta[0], sc.Dduvdx[1] * sc.PixelDelta[0]};
  float Dduvdy[2] = {sc.Dduvdy[0] * sc.PixelDelta[1], sc.Dduvdy[1] * sc.PixelDelta[1]};
  Dduvdx[0] = Dduvdx[0] * (float)img->w;
  Dduvdy[0] = Dduvdy[0] * (float)img->w;
  Dduvdx[1] = Dduvdx[1] * (float)img->h;
  Dduvdy[1] = Dduvdy[1] * (float)img->h;

  const float Ann = Dduvdx[1] * Dduvdx[1] + Dduvdy[1] * Dduvdy[1];
  const float Bnn = -2 * (Dduvdx[0] * Dduvdx[1] + Dduvdy[0] * Dduvdy[1]);
  const float Cnn = Dduvdx[0] * Dduvdx[0] + Dduvdy[0] * Dduvdy[0];
  const float F = Ann * Cnn - (Bnn * Bnn / 4);
  const float A = Ann / F;
  const float B = Bnn / F;
  const float C = Cnn / F;

  const float root = Sqrt(tex_sqr(A - C) + tex_sqr(B));
  const float Aprm = (A + C - root) / 2;
  const float Cprm = (A + C + root) / 2;
  float majorRadius = Sqrt(1 / Aprm);
  float minorRadius = Sqrt(1 / Cprm);
  float theta = Atan(B / (A - C)) / 2;
  if (A > C) theta = theta + kPi / 2;
  minorRadius = Max(minorRadius, 1);
  majorRadius = Max(majorRadius, 1);

  const float fProbes = 2 * (majorRadius / minorRadius) - 1;
  float iProbes = Floor(fProbes + 0.5f);
  iProbes = Min(iProbes, (float)maxProbes);
  if (iProbes < fProbes) minorRadius = 2 * majorRadius / (iProbes + 1);

  float levelOfDetail = Log2(minorRadius);
  if (levelOfDetail > (float)img->mipmap.MaxLevelOfDetail()) {
    levelOfDetail = (float)img->mipmap.MaxLevelOfDetail();
    iProbes = 1;
  }
  if (levelOfDetail < 0) levelOfDetail = 0;

  const float lineLength = 2 * (majorRadius - minorRadius);
  float dU = Cos(theta) * lineLength / (iProbes - 1);
  float dV = Sin(theta) * lineLength / (iProbes - 1);
  const int nProbes = (int)iProbes;
  if (nProbes == 1) {
    dU = 0;
    dV = 0;
  }
  float n = (float)(-(nProbes - 1));
  const float alpha = 0.6f;
  float accum[3] = {0, 0, 0};
  float accumWeight = 0;
  for (int i = 0; i < nProbes; i++) {
    const float u = (float)img->w * sc.U + (n / 2) * dU;
    const float v = (float)img->h * sc.V + (n / 2) * dV;
    const float d2 = (tex_sqr(n) / 4) * (tex_sqr(dU) + tex_sqr(dV)) / tex_sqr(majorRadius);
    const float relativeWeight = Exp(-alpha * d2);
    float sample[3];
    img->mipmap.TrilinearSample(u / (float)img->w, v / (float)img->h, levelOfDetail, sample);
    for (int k = 0; k < 3; k++) accum[k] += (sample[k] / 255.0f) * relativeWeight;
    accumWeight += relativeWeight;
    n += 2;
  }
  for (int k = 0; k < 3; k++) c[k] = accum[k] / accumWeight;
}

// builtin/maps/texture.go:17-46: one shader parameter bound to a texture file.
struct TextureMap {
  const Texture* tex = nullptr;
  int Chan = 0;
  bool trilinear = false;  // "?filter=trilinear" -> TextureTrilinear, else Feline
  void Sample(const TexCoord& sg, float out[3]) const {
    if (trilinear) SampleRGB(tex, sg, out);
    else SampleFeline(tex, sg, out);
  }
};

}  // namespace orc
