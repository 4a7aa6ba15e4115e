// Texture kernels: the mip pyramid of texture/mipmap.go:122-315 built level by level on the device, and a batch entry point
// of the two filters (for callers that want maps.Texture / maps.TextureTrilinear lookups outside a render, and for parity tests).
#include "texture.cuh"

namespace vg {

__device__ __forceinline__ int dmaxi(int a, int b) { return a > b ? a : b; }
__device__ __forceinline__ int dmini(int a, int b) { return a < b ? a : b; }

// One thread per texel of the new level. The branch is chosen by the parity of the NEW level's size, as the reference does
// (mipmap.go:147-148,211); index clamps and wrap rules are the reference's, including the ones that only matter for odd sizes.
__global__ void __launch_bounds__(256) k_mip_level(uchar4* texels, DevTexLevel S, DevTexLevel D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.w * D.h) return;
  const int x = i % D.w, y = i / D.w;
  const int width = S.w, height = S.h, nwidth = D.w, nheight = D.h;
  const uchar4* P = texels + S.off;
  auto px = [&](int xx, int yy) {
    const uchar4 t = P[xx + yy * width];
    return make_float3((float)t.x, (float)t.y, (float)t.z);
  };
  float3 r;
  if (nheight % 2 == 0) {
    const int y0 = y * 2, y1 = dmini(y0 + 1, dmaxi(1, height - 1));
    if (nwidth % 2 == 0) {
      const int x0 = x * 2, x1 = dmini(x0 + 1, dmaxi(1, width - 1));
      const float3 a = px(x0, y0), b = px(x0, y1), c = px(x1, y0), d = px(x1, y1);
      r = make_float3(0.25f * (a.x + b.x + c.x + d.x), 0.25f * (a.y + b.y + c.y + d.y), 0.25f * (a.z + b.z + c.z + d.z));
    } else {  // height even, width odd
      int x0 = dmaxi(x * 2 - 1, -dmaxi(1, width - 1));
      const int x1 = x * 2;
      int x2 = dmini(x * 2 + 1, dmaxi(1, width - 1));
      if (x0 < 0) { x0 += width; x0 = dmini(x0, dmaxi(1, width - 1)); }
      if (x2 > width - 1) { x2 -= width; x2 = dmaxi(x2, -dmaxi(1, width - 1)); }
      const float w0 = (float)(nwidth - x - 1) / (float)(2 * nwidth - 1);
      const float w1 = (float)(nwidth) / (float)(2 * nwidth - 1);
      const float w2 = (float)(x) / (float)(2 * nwidth - 1);
      const float3 c00 = px(x0, y0), c10 = px(x1, y0), c20 = px(x2, y0), c01 = px(x0, y1), c11 = px(x1, y1), c21 = px(x2, y1);
      r.x = 0.5f * (w0 * c00.x + w1 * c10.x + w2 * c20.x + w0 * c01.x + w1 * c11.x + w2 * c21.x);
      r.y = 0.5f * (w0 * c00.y + w1 * c10.y + w2 * c20.y + w0 * c01.y + w1 * c11.y + w2 * c21.y);
      r.z = 0.5f * (w0 * c00.z + w1 * c10.z + w2 * c20.z + w0 * c01.z + w1 * c11.z + w2 * c21.z);
    }
  } else {
    int y0 = dmaxi(y * 2 - 1, -dmaxi(1, height - 1));
    const int y1 = y * 2;
    const int y2 = dmini(y * 2 + 1, dmaxi(1, height - 1));
    if (y0 < 0) { y0 += height; y0 = dmini(y0, dmaxi(1, height - 1)); }
    const float wy0 = (float)(nheight - y - 1) / (float)(2 * nheight - 1);
    const float wy1 = (float)(nheight) / (float)(2 * nheight - 1);
    const float wy2 = (float)(y) / (float)(2 * nheight - 1);
    if (nwidth % 2 == 0) {  // height odd, width even
      const int x0 = x * 2, x1 = dmini(x0 + 1, dmaxi(1, width - 1));
      const float3 c00 = px(x0, y0), c01 = px(x0, y1), c02 = px(x0, y2), c10 = px(x1, y0), c11 = px(x1, y1), c12 = px(x1, y2);
      r.x = 0.5f * (wy0 * c00.x + wy1 * c01.x + wy2 * c02.x + wy0 * c10.x + wy1 * c11.x + wy2 * c12.x);
      r.y = 0.5f * (wy0 * c00.y + wy1 * c01.y + wy2 * c02.y + wy0 * c10.y + wy1 * c11.y + wy2 * c12.y);
      r.z = 0.5f * (wy0 * c00.z + wy1 * c01.z + wy2 * c02.z + wy0 * c10.z + wy1 * c11.z + wy2 * c12.z);
    } else {
      int x0 = dmaxi(x * 2 - 1, -dmaxi(1, width - 1));
      const int x1 = x * 2;
      int x2 = dmini(x * 2 + 1, dmaxi(1, width - 1));
      if (x0 < 0) { x0 += width; x0 = dmini(x0, dmaxi(1, width - 1)); }
      if (x2 > width - 1) { x2 -= width; x2 = dmaxi(x2, -dmaxi(1, width - 1)); }
      const float w0 = (float)(nwidth - x - 1) / (float)(2 * nwidth - 1);
      const float w1 = (float)(nwidth) / (float)(2 * nwidth - 1);
      const float w2 = (float)(x) / (float)(2 * nwidth - 1);
      const float3 c00 = px(x0, y0), c01 = px(x0, y1), c02 = px(x0, y2);
      const float3 c10 = px(x1, y0), c11 = px(x1, y1), c12 = px(x1, y2);
      const float3 c20 = px(x2, y0), c21 = px(x2, y1), c22 = px(x2, y2);
      r.x = wy0 * (w0 * c00.x + w1 * c10.x + w2 * c20.x) + wy1 * (w0 * c01.x + w1 * c11.x + w2 * c21.x) + wy2 * (w0 * c02.x + w1 * c12.x + w2 * c22.x);
      r.y = wy0 * (w0 * c00.y + w1 * c10.y + w2 * c20.y) + wy1 * (w0 * c01.y + w1 * c11.y + w2 * c21.y) + wy2 * (w0 * c02.y + w1 * c12.y + w2 * c22.y);
      r.z = wy0 * (w0 * c00.z + w1 * c10.z + w2 * c20.z) + wy1 * (w0 * c01.z + w1 * c11.z + w2 * c21.z) + wy2 * (w0 * c02.z + w1 * c12.z + w2 * c22.z);
    }
  }
  // byte(float32): truncation toward zero; the weights sum to <= 1 so the value stays within [0, 255]
  texels[D.off + i] = make_uchar4((unsigned char)(int)r.x, (unsigned char)(int)r.y, (unsigned char)(int)r.z, 0);
}

cudaError_t launch_mip_level(uchar4* texels, DevTexLevel src, DevTexLevel dst, cudaStream_t stream) {
  const int n = dst.w * dst.h;
  k_mip_level<<<(n + 255) / 256, 256, 0, stream>>>(texels, src, dst);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(128) k_texture_sample(DevTexStore ts, int tex, int filter, const float* __restrict__ coords, long long n,
                                                        float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = *reinterpret_cast<const float4*>(coords + i * 8), b = *reinterpret_cast<const float4*>(coords + i * 8 + 4);
  TexCoord tc;
  tc.U = a.x; tc.V = a.y; tc.dudx = a.z; tc.dvdx = a.w;
  tc.dudy = b.x; tc.dvdy = b.y; tc.pd0 = b.z; tc.pd1 = b.w;
  float c[3];
  tex_sample(ts, tex, filter, tc, c);
  out[i * 3 + 0] = c[0];
  out[i * 3 + 1] = c[1];
  out[i * 3 + 2] = c[2];
}

cudaError_t launch_texture_sample(const DevTexStore& ts, int tex, int filter, const float* d_coords8, long long n, float* d_out3, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  k_texture_sample<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(ts, tex, filter, d_coords8, n, d_out3);
  return cudaGetLastError();
}

}  // namespace vg
