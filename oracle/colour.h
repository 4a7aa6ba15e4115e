// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// CPU restatement of the reference's hero-wavelength colour pipeline.
// Follows:
//   colour/rgb.go:11-37                 (RGB Scale/Add/Mul)
//   colour/spectrum.go:18-112           (Spectrum: 4 hero wavelengths rotated over [450,750))
//   colour/spectrum_smits9.go:9-84      (Smits'99 RGB -> spectrum, 10 bins over [380,720))
//   colour/cie1931_2deg.go:62-95        (nearest-bin CIE observer lookup, -1 outside [360,830))
//   colour/space_srgb.go:8-14, colourspace.go:13-19 (XYZ -> sRGB)
#pragma once
#include "colour_tables.h"

namespace orc {

struct RGB {
  float c[3] = {0, 0, 0};
  float& operator[](int i) { return c[i]; }
  const float& operator[](int i) const { return c[i]; }
  void Scale(float f) { for (int k = 0; k < 3; k++) c[k] *= f; }
  void Add(const RGB& o) { for (int k = 0; k < 3; k++) c[k] += o.c[k]; }
  void Mul(const RGB& o) { for (int k = 0; k < 3; k++) c[k] *= o.c[k]; }
};
static inline RGB MakeRGB(float r, float g, float b) { RGB x; x.c[0] = r; x.c[1] = g; x.c[2] = b; return x; }

static inline float smitsEval(const float* s, float lambda) {  // spectrum_smits9.go:16-25
  if (lambda < 380.0f || lambda >= 720.0f) return 0;
  int bin = (int)(((lambda - 380.0f) / (720.0f - 380.0f)) * 10.0f);
  return s[bin];
}

static inline float RGBToSpectrumSmits99(float r, float g, float b, float lambda) {  // spectrum_smits9.go:48-84
  float c = 0;
  if (r <= g && r <= b) {
    c += r * smitsEval(kSmitsWhite, lambda);
    if (g <= b) {
      c += (g - r) * smitsEval(kSmitsCyan, lambda);
      c += (b - g) * smitsEval(kSmitsBlue, lambda);
    } else {
      c += (b - r) * smitsEval(kSmitsCyan, lambda);
      c += (g - b) * smitsEval(kSmitsGreen, lambda);
    }
  } else if (g <= r && g <= b) {
    c += g * smitsEval(kSmitsWhite, lambda);
    if (r <= b) {
      c += (r - g) * smitsEval(kSmitsMagenta, lambda);
      c += (b - r) * smitsEval(kSmitsBlue, lambda);
    } else {
      c += (b - g) * smitsEval(kSmitsMagenta, lambda);
      c += (r - b) * smitsEval(kSmitsRed, lambda);
    }
  } else {
    c += b * smitsEval(kSmitsWhite, lambda);
    if (r <= g) {
      c += (r - b) * smitsEval(kSmitsYellow, lambda);
      c += (g - r) * smitsEval(kSmitsGreen, lambda);
    } else {
      c += (g - b) * smitsEval(kSmitsYellow, lambda);
      c += (r - g) * smitsEval(kSmitsRed, lambda);
    }
  }
  return c;
}

static inline float cieLookup(const float* tab, float lambda) {  // cie1931_2deg.go:62-70
  if (lambda < 360.0f || lambda >= 830.0f) return -1;
  int bin = (int)(((lambda - 360.0f) / (830.0f - 360.0f)) * (float)95);
  return tab[bin];
}

struct Spectrum {
  float C[4] = {0, 0, 0, 0};
  float Lambda = 0;

  float Wavelength(int j) const {  // spectrum.go:101-112
    float v = (Lambda - 450.0f + ((float)j / 4.0f) * 300.0f);
    if (v >= 300.0f) v -= 300.0f;
    v += 450.0f;
    return v;
  }
  void FromRGB(const RGB& rgb) {  // spectrum.go:49-54
    for (int k = 0; k < 4; k++) C[k] = RGBToSpectrumSmits99(rgb[0], rgb[1], rgb[2], Wavelength(k));
  }
  RGB ToRGB() const {  // spectrum.go:57-72
    float x = 0, y = 0, z = 0;
    for (int i = 0; i < 4; i++) {
      x += C[i] * cieLookup(kCieX, Wavelength(i));
      y += C[i] * cieLookup(kCieY, Wavelength(i));
      z += C[i] * cieLookup(kCieZ, Wavelength(i));
    }
    RGB rgb;
    rgb[0] = x * 3.2404542f + y * -1.5371385f + z * -0.4985314f;
    rgb[1] = x * -0.9692660f + y * 1.8760108f + z * 0.0415560f;
    rgb[2] = x * 0.0556434f + y * -0.2040259f + z * 1.0572252f;
    return rgb;
  }
  void Mul(const Spectrum& o) { for (int k = 0; k < 4; k++) C[k] *= o.C[k]; }
  void Scale(float s) { for (int k = 0; k < 4; k++) C[k] *= s; }
};

}  // namespace orc
