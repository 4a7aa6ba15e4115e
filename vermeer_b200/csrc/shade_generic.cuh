// General shading kernel: every lobe of ShaderStd (Oren-Nayar diffuse, mirror, GGX glossy) with either Fresnel model and
// every in-scope light type (Tri, Disk, Sphere). Included by render.cu after k_shade, which stays the specialised fast
// path for scenes that only use "diffuse + dielectric mirror + TriLights" (BASELINE configs C1-C5).
//
// Reference functions mirrored (evaluation order kept; compiled with -fmad=false):
//   ShaderStd.Eval glossy branch          builtin/shader/std.go:172-192,194-217,269-284
//   bsdf.MicrofacetGGX                    builtin/shader/bsdf/microfacetggx.go:27-146
//   fresnel.Conductor                     builtin/shader/fresnel/conductor.go:33-89
//   light.Disk                            builtin/light/disk.go:38-68,123-217
//   light.Sphere                          builtin/light/sphere.go:69-127,132-270
//   sphere.Sphere hit record              builtin/geom/sphere/trace.go:13-51
#pragma once

namespace vg {

// ---- conductor Fresnel -------------------------------------------------------------------------------
__device__ __forceinline__ float cond_nmin(float r) { return (1 - r) / (1 + r); }
__device__ __forceinline__ float cond_nmax(float r) { return (1 + sqrtf(r)) / (1 - sqrtf(r)); }
__device__ inline float conductor_fresnel(float r, float g, float c) {  // conductor.go:58-75
  const float nr = maxf_x86(0.0f, minf_x86(r, 0.99f));
  const float n = cond_nmin(nr) * g + (1 - g) * cond_nmax(nr);
  const float k2 = ((n + 1) * (n + 1) * nr - (n - 1) * (n - 1)) / (1 - nr);
  const float rsNum = n * n + k2 - 2 * n * c + c * c;
  const float rsDen = n * n + k2 + 2 * n * c + c * c;
  const float rs = rsNum / rsDen;
  const float rpNum = (n * n + k2) * c * c - 2 * n * c + 1;
  const float rpDen = (n * n + k2) * c * c + 2 * n * c + 1;
  const float rp = rpNum / rpDen;
  return 0.5f * (rs + rp);
}
__device__ inline f3 fresnel_kr(const DevMat& m, float c) {  // std.go:172-192
  if (m.fresnel_model == VG_FRESNEL_CONDUCTOR)
    return mk3(conductor_fresnel(m.fres_refl.x, m.fres_edge.x, c), conductor_fresnel(m.fres_refl.y, m.fres_edge.y, c),
               conductor_fresnel(m.fres_refl.z, m.fres_edge.z, c));
  const float k = dielectric_kr(m.ior, c);
  return mk3(k, k, k);
}

// ---- GGX ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ggx_chi(float x) { return x > 0.0f ? 1.0f : 0.0f; }
__device__ __forceinline__ float ggx_sign(float v) { return v < 0 ? -1.0f : 1.0f; }
__device__ __forceinline__ float sqr32(float x) { return x * x; }
__device__ inline float ggx_g1(f3 omega, f3 omegaM, float alpha) {  // microfacetggx.go:35-52
  const float ODotN = dot3(omega, omegaM);
  const float denom = 1 + sqrtf(1 + (alpha * alpha) * ((1.0f / (omega.z * omega.z)) - 1));
  return ggx_chi(ODotN / omega.z) * 2 / denom;
}
__device__ inline float ggx_d(f3 omegaM, float alpha) {  // microfacetggx.go:54-71
  const float numer = alpha * alpha * ggx_chi(omegaM.z);
  if (omegaM.z == 1.0f) return 1.0f / (VG_PI32 * alpha * alpha);
  const float denom = VG_PI32 * sqr32(omegaM.z * omegaM.z) * sqr32(alpha * alpha + ((1.0f / (omegaM.z * omegaM.z)) - 1));
  return numer / denom;
}
struct GgxVertex {
  f3 omegaR;    // view direction in the tangent frame
  float alpha;  // sqr32(roughness*roughness): NewMicrofacetGGX stores roughness^2, every method squares it again
};
template <bool FAST>
__device__ inline f3 ggx_sample(const Frame& fr, const GgxVertex& g, double r0, double r1) {  // :91-107
  float st, ct, sp, cp;
  if (FAST) {
    const float thetaM = atan2f(g.alpha * sqrtf((float)r0), sqrtf((float)(1 - r0)));
    sincosf(thetaM, &st, &ct);
    sincospif((float)(2 * r1), &sp, &cp);
  } else {
    const double thetaM = atan2((double)g.alpha * sqrt(r0), sqrt(1 - r0));
    const double phiM = 2.0 * VG_PI64 * r1;
    st = sin32((float)thetaM); ct = cos32((float)thetaM);
    sp = sin32((float)phiM); cp = cos32((float)phiM);
  }
  const f3 omegaM = mk3(st * cp, st * sp, ct);
  const f3 o = sub3(scale3(2.0f * fabsf(dot3(omegaM, g.omegaR)), omegaM), g.omegaR);
  return basis_expand(fr.U, fr.V, fr.N, normalize3t<FAST>(o));
}
// float32(bsdf.PDF(wo)); 0 when the float64 value is NaN (:120-122). The reference's `Pdf <= 0` test on the float64 is the
// same test on this value (the float64 is an exactly converted float32).
template <bool FAST>
__device__ inline float ggx_pdf32(const Frame& fr, const GgxVertex& g, f3 wo) {  // :110-125
  const f3 o = basis_project(fr.U, fr.V, fr.N, wo);
  const f3 m = scale3(ggx_sign(g.omegaR.z), normalize3t<FAST>(add3(g.omegaR, o)));
  const float pdf = ggx_d(m, g.alpha) * m.z;
  return pdf != pdf ? 0.0f : pdf;
}
template <bool FAST>
__device__ inline Spec4 ggx_eval(const Frame& fr, const GgxVertex& g, const DevMat& mat, const Hero& hero, f3 wo) {  // :128-146
  const f3 oi = basis_project(fr.U, fr.V, fr.N, wo);
  const f3 h = scale3(ggx_sign(g.omegaR.z), normalize3t<FAST>(add3(g.omegaR, oi)));
  const f3 fres = fresnel_kr(mat, fabsf(dot3(g.omegaR, h)));
  const float numer = ggx_g1(g.omegaR, h, g.alpha) * ggx_g1(oi, h, g.alpha) * ggx_d(h, g.alpha);
  const float denom = 4 * fabsf(g.omegaR.z) * fabsf(oi.z);
  Spec4 rho = spec_from_rgb(fres, hero);
  const float k = fabsf(oi.z) * numer / denom;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    rho.c[i] *= k;
    if (rho.c[i] < 0 || rho.c[i] != rho.c[i]) rho.c[i] = 0;
  }
  return rho;
}

// ---- lights ------------------------------------------------------------------------------------------
// sphere.go:69-127 (== geom/sphere/trace.go:53-109)
__device__ inline bool ray_sphere(f3 Ro, f3 Rd, f3 P, float radius, float* tout) {
  const f3 L = sub3(Ro, P);
  const float a = dot3(Rd, Rd);
  const float b = 2 * dot3(Rd, L);
  const float c = dot3(L, L) - radius * radius;
  const float discr = b * b - 4 * a * c;
  if (discr < 0) return false;
  float x0, x1;
  if (discr == 0) {
    x1 = -0.5f * b / a;
    x0 = x1;
  } else {
    const float q = b > 0 ? -0.5f * (b + sqrtf(discr)) : -0.5f * (b - sqrtf(discr));
    x0 = q / a;
    x1 = c / q;
  }
  if (x0 > x1) { const float t = x0; x0 = x1; x1 = t; }
  if (x0 < 0) {
    x0 = x1;
    if (x0 < 0) return false;
  }
  *tout = x0;
  return true;
}

// Per (vertex, light) invariants.
struct LightVertex {
  bool by_area;  // tri: some vertex below the receiver's horizon -> area sampling (triangle.go:287-291)
  SphTri sph;    // tri: spherical triangle of the light seen from P
  f3 su, sv, sw; // sphere: cone frame (sphere.go:188-194)
  float pdf_cone;
  float cone_k;  // Sqrt(1 - sqr(Radius / l))
};
template <bool FAST>
__device__ inline LightVertex light_vertex(const DevLight& L, const ShadeCtx& c) {
  LightVertex v;
  v.by_area = false;
  v.pdf_cone = v.cone_k = 0.0f;
  v.su = v.sv = v.sw = mk3(0, 0, 1);
  if (L.type == VG_LIGHT_TRI) {
    v.by_area = dot3(c.Ng, sub3(L.p0, c.P)) < 0 || dot3(c.Ng, sub3(L.p1, c.P)) < 0 || dot3(c.Ng, sub3(L.p2, c.P)) < 0;
    if (!v.by_area) v.sph = spherical_setup<FAST>(L.p0, L.p1, L.p2, c.P);
  } else if (L.type == VG_LIGHT_SPHERE) {
    const f3 V = sub3(L.p0, c.P);
    const float l = length3(V);
    v.sw = normalize3t<FAST>(V);
    v.sv = normalize3t<FAST>(cross3(v.sw, c.Ng));
    v.su = cross3(v.sw, v.sv);
    v.cone_k = sqrtf(1 - sqr32(L.radius / l));
    v.pdf_cone = 1 / (2 * VG_PI32 * (1 - v.cone_k));  // q_2 from Shirley96 (sphere.go:176,259)
  }
  return v;
}

// Light.SampleArea, sample `i` of `n`.
template <bool FAST>
__device__ inline LightRec light_sample_any(const DevLight& L, const ShadeCtx& c, const LightVertex& lv, long long I, int n, int i, uint64_t scr0,
                                            uint64_t scr1) {
  if (L.type == VG_LIGHT_TRI) return light_sample<FAST>(L, c, lv.by_area, lv.sph, I, n, i, scr0, scr1);
  LightRec r;
  r.valid = false;
  r.Ld = mk3(0, 0, 1);
  r.Ldist = 0.0f;
  r.pdf = 0.0f;
  const uint64_t idx = (uint64_t)(I * n + i);
  const double r0 = vdc(idx, scr0);
  const double r1 = sobol(idx, scr1);
  if (L.type == VG_LIGHT_DISK) {  // disk.go:172-217; p1 = T, p2 = B
    const float sq = sqrtf((float)r0);
    const float ang = 2 * VG_PI32 * (float)r1;
    const float u = L.radius * sq * Trig<FAST>::cos(ang);
    const float v = L.radius * sq * Trig<FAST>::sin(ang);
    const f3 a = scale3(u, L.p2), b = scale3(v, L.p1);
    const f3 Pl = mk3(L.p0.x + a.x + b.x, L.p0.y + a.y + b.y, L.p0.z + a.z + b.z);
    const f3 V = sub3(Pl, c.P);
    if (dot3(V, c.Ng) > 0.0f && dot3(V, L.N) < 0.0f) {
      r.Ldist = length3t<FAST>(V);
      r.Ld = normalize3t<FAST>(V);
      r.pdf = divt<FAST>(L.inv_area * (r.Ldist * r.Ldist), fabsf(dot3(r.Ld, L.N)));
      r.valid = true;
    }
  } else {  // sphere.go:186-270
    const float r0f = (float)r0;
    const float theta = Trig<FAST>::acos(1 - r0f + r0f * lv.cone_k);
    const float phi = 2 * VG_PI32 * (float)r1;
    const float st = Trig<FAST>::sin(theta), ct = Trig<FAST>::cos(theta);
    const f3 a = mk3(Trig<FAST>::cos(phi) * st, Trig<FAST>::sin(phi) * st, ct);
    const f3 omega = basis_expand(lv.su, lv.sv, lv.sw, a);
    float t;
    if (!(dot3(omega, c.Ng) < 0) && ray_sphere(c.P, omega, L.p0, L.radius, &t)) {
      const f3 x = mad3(c.P, omega, t);
      const f3 D = sub3(x, c.P);
      r.Ldist = length3t<FAST>(D);
      r.Ld = normalize3t<FAST>(D);
      r.pdf = lv.pdf_cone;
      r.valid = true;
    }
  }
  return r;
}

// Light.ValidSample for a BSDF direction `wo` whose (float32) pdf is `pdf`.
template <bool FAST>
__device__ inline BsdfRec light_valid_sample(const DevLight& L, const ShadeCtx& c, const LightVertex& lv, f3 wo, float pdf) {
  BsdfRec r;
  r.valid = false;
  r.pdf = pdf;
  r.pdfLight = 0.0f;
  r.Ldist = 0.0f;
  r.Ld = mk3(0, 0, 1);
  if (L.type == VG_LIGHT_TRI) {  // triangle.go:136-230
    f3 Pl;
    if (!ray_triangle<FAST>(c.P, wo, L.p0, L.p1, L.p2, &Pl)) return r;
    const bool by_area = dot3(c.Ng, sub3(L.p0, Pl)) < 0 || dot3(c.Ng, sub3(L.p1, Pl)) < 0 || dot3(c.Ng, sub3(L.p2, Pl)) < 0;
    float pdfl;
    if (by_area) {
      pdfl = L.inv_area;
    } else {
      const float area = !lv.by_area ? lv.sph.area : spherical_setup<FAST>(L.p0, L.p1, L.p2, c.P).area;
      pdfl = FAST ? __fdividef(1.0f, area) : (float)(double)(1 / area);
    }
    const f3 D = sub3(Pl, c.P);
    r.Ldist = length3t<FAST>(D);
    r.Ld = normalize3t<FAST>(D);
    if (dot3(r.Ld, L.N) > 0 || dot3(r.Ld, c.Ng) < 0) return r;
    r.pdfLight = by_area ? divt<FAST>(pdfl * (r.Ldist * r.Ldist), fabsf(dot3(r.Ld, L.N))) : pdfl;
    r.valid = true;
  } else if (L.type == VG_LIGHT_DISK) {  // disk.go:38-68,123-169
    const float denom = dot3(L.N, wo);
    if (!(fabsf(denom) > 1e-6f)) return r;
    const float t = dot3(sub3(L.p0, c.P), L.N) / denom;
    if (!(t >= 0)) return r;
    const f3 p = mad3(c.P, wo, t);
    const f3 dv = sub3(p, L.p0);
    if (!(sqrtf(dot3(dv, dv)) <= L.radius)) return r;
    const f3 V = sub3(p, c.P);
    if (dot3(V, c.Ng) <= 0.0f || dot3(V, L.N) >= 0.0f) return r;
    r.Ldist = length3t<FAST>(V);
    r.Ld = normalize3t<FAST>(V);
    r.pdfLight = divt<FAST>(L.inv_area * (r.Ldist * r.Ldist), fabsf(dot3(r.Ld, L.N)));
    r.valid = true;
  } else {  // sphere.go:132-183
    float t;
    if (!ray_sphere(c.P, wo, L.p0, L.radius, &t)) return r;
    const f3 x = mad3(c.P, wo, t);
    const f3 D = sub3(x, c.P);
    r.Ldist = length3t<FAST>(D);
    r.Ld = normalize3t<FAST>(D);
    r.pdfLight = lv.pdf_cone;
    r.valid = true;
  }
  return r;
}

// ---- the kernel ----------------------------------------------------------------------------------------
// Lobe 0 = Oren-Nayar diffuse, lobe 1 = GGX glossy. Slot layout per path: [lobe][light slot_base + sample].
template <bool FAST>
__global__ void __launch_bounds__(128, 4) k_shade_generic(const RenderParams p, int level, int qin, int qout, int iter_base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = p.counts[qin];
  if ((i & ~31) >= n) return;
  bool active = i < n;
  int path = 0;
  DevHit h;
  h.prim = -1;
  f3 Ro = mk3(0, 0, 0), Rd = mk3(0, 0, 1);
  if (active) {
    path = p.pathq[qin][i];
    const float4 h0 = *reinterpret_cast<const float4*>(&p.hits[i]);
    const int4 h1 = *(reinterpret_cast<const int4*>(&p.hits[i]) + 1);
    h.t = h0.x; h.u = h0.y; h.v = h0.z; h.w = h0.w;
    h.prim = h1.x; h.geom = h1.y; h.slot = h1.z; h.xf = h1.w;
    const float4* rp = reinterpret_cast<const float4*>(p.rayq[qin] + i);
    const float4 a = rp[0], b = rp[1];
    Ro = mk3(a.x, a.y, a.z);
    Rd = mk3(a.w, b.x, b.y);
  }
  const int SL = p.S * p.nlobes;
  int matid = 255;
  if (active && h.prim >= 0) matid = p.sc.prim_material[p.sc.geoms[h.geom].prim_base + h.prim];
  if (active) {
    p.v_mat[i] = (uint8_t)matid;
    p.T[(size_t)level * p.P + path] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  active = active && matid != 255 && level <= 3;

  DevMat m;
  ShadeCtx c;
  Frame fr;
  f3 omegaI = mk3(0, 0, 1);
  OrenVertex ov;
  GgxVertex gv;
  Hero hero;
  float time = 0;
  long long I = 0;
  uint64_t scr0 = 0, scr1 = 0;
  if (active) {
    m = *((p.vmats && p.mat_tex[matid].mask) ? p.vmats + i : p.mats + matid);  // one load site: only the fields used are fetched
    if (m.bad) {
      atomicOr(p.counts + 5, m.bad);
      active = false;
    }
  }
  if (active) {
    const float lambda = p.lambda[path];
    time = p.time[path];
    int own, it;
    path_decode(p, path, own, it);
    if (h.xf > 0) {  // an instance left its transform in the context (instance.go:107-111)
      if (p.sc.geoms[h.geom].keys == 0) build_context_sphere(p, h, Ro, Rd, c);  // (its own re-normalisations are idempotent)
      else build_context<false, FAST>(p, h, time, c);
      apply_instance_transform(p, h.xf - 1, time, c);
    } else if (p.sc.geoms[h.geom].keys == 0) {
      build_context_sphere(p, h, Ro, Rd, c);
    } else {
      build_context<true, FAST>(p, h, time, c);
    }
    f3 V = cross3(c.N, c.DdPdu);
    if (len2_3(V) < 0.1f) V = cross3(c.N, c.DdPdv);
    V = normalize3t<FAST>(V);
    fr.U = normalize3t<FAST>(cross3(c.N, V));
    fr.V = V;
    fr.N = c.N;
    omegaI = basis_project(fr.U, fr.V, fr.N, neg3(Rd));
    hero = hero_setup(lambda);
    ov = oren_vertex<FAST>(omegaI, m.rough2, hero, p.qmc->white_spec);
    gv.omegaR = omegaI;
    gv.alpha = sqr32(m.spec_rough * m.spec_rough);
    const size_t srow = p.scr_by_pixel ? (size_t)p.pix[own] : (size_t)own;
    scr0 = p.scr[srow * 6 + 4];
    scr1 = p.scr[srow * 6 + 5];
    I = (long long)(iter_base + it + 1);
  }

  for (int lobe = 0; lobe < p.nlobes; lobe++) {
    // std.go:145-163 (diffuse, only lights with DiffuseShadeMult > 0: all of them) and :269-284 (glossy)
    const bool lobe_on = active && (lobe == 0 ? m.diff_weight > 0.0f : (m.spec_weight > 0.0f && m.spec_rough > 0.0f));
    for (int l = 0; l < p.nlights; l++) {
      const DevLight L = p.lights[l];
      const int NS = level > 0 ? 1 : L.nsamples;
      const bool lit = lobe_on && L.geom != h.geom;
      const int hN = NS > 1 ? NS / 2 : NS;
      Spec4 Liu;
      LightVertex lv;
      if (lit) {
        Liu = spec_from_table(p.lights[l].spec, hero);
        lv = light_vertex<FAST>(L, c);
      }
      auto bsdf_rec = [&](int s) -> BsdfRec {
        const uint64_t idx = (uint64_t)(I * hN + s);
        const double r0 = vdc(idx, scr0);
        const double r1 = sobol(idx, scr1);
        // EvaluateLightSamples normalizes the sampled direction again (core/shader.go:216)
        f3 wo;
        float pdf;
        if (lobe == 0) {
          wo = normalize3t<FAST>(basis_expand(fr.U, fr.V, fr.N, cosine_hemisphere<FAST>(r0, r1)));
          pdf = oren_pdf32<FAST>(fr, wo);
        } else {
          wo = normalize3t<FAST>(ggx_sample<FAST>(fr, gv, r0, r1));
          pdf = ggx_pdf32<FAST>(fr, gv, wo);
        }
        if (!(pdf > 0)) {
          BsdfRec r;
          r.valid = false;
          r.pdf = pdf;
          r.pdfLight = 0.0f;
          r.Ldist = 0.0f;
          r.Ld = mk3(0, 0, 1);
          return r;
        }
        return light_valid_sample<FAST>(L, c, lv, wo, pdf);
      };
      int nB = 0, nLs = 0;
      if (lit) {
        if (NS > 1)
          for (int s = 0; s < hN; s++)
            if (bsdf_rec(s).valid) nB = hN;
        for (int s = 0; s < hN; s++)
          if (light_sample_any<FAST>(L, c, lv, I, hN, s, scr0, scr1).valid) nLs = hN;
      }
      const int total = nB + nLs;
      if (active) p.v_invtot[((size_t)i * p.nlobes + lobe) * p.nlights + l] = !lit ? 0.0f : (NS > 1 ? (total > 0 ? 1.0f / (float)total : 0.0f) : 1.0f);

      for (int s = 0; s < (NS > 1 ? 2 * hN : hN); s++) {
        const bool is_bsdf = s >= hN;
        bool want = false;
        f3 Ld = mk3(0, 0, 1);
        float Ldist = 0;
        float4 rgb4 = make_float4(0, 0, 0, 0);
        if (lit && (NS == 1 || total > 0)) {
          float p_hat;
          bool valid;
          if (!is_bsdf) {
            const LightRec lr = light_sample_any<FAST>(L, c, lv, I, hN, s, scr0, scr1);
            valid = lr.valid;
            Ld = lr.Ld;
            Ldist = lr.Ldist;
            if (NS > 1) {
              const float bp = lobe == 0 ? oren_pdf32<FAST>(fr, Ld) : ggx_pdf32<FAST>(fr, gv, Ld);
              p_hat = divt<FAST>((float)nB * bp, (float)total);
              p_hat += divt<FAST>((float)nLs * lr.pdf, (float)total);
            } else {
              p_hat = lr.pdf;
            }
          } else {
            const BsdfRec br = bsdf_rec(s - hN);
            valid = br.valid;
            Ld = br.Ld;
            Ldist = br.Ldist;
            p_hat = divt<FAST>((float)nB * br.pdf, (float)total);
            p_hat += divt<FAST>((float)nLs * br.pdfLight, (float)total);
          }
          if (valid && !(dot3(Ld, c.N) <= 0)) {
            Spec4 rho = lobe == 0 ? oren_eval<FAST>(fr, ov, Ld) : ggx_eval<FAST>(fr, gv, m, hero, Ld);
            const float inv = divt<FAST>(1.0f, p_hat);
#pragma unroll
            for (int k = 0; k < 4; k++) rho.c[k] = (rho.c[k] * Liu.c[k]) * inv;
            f3 rgb = spec_to_rgb(rho, hero);
            if (NS > 1) {
              if (rgb.x < 0) rgb.x = 0;
              if (rgb.y < 0) rgb.y = 0;
              if (rgb.z < 0) rgb.z = 0;
            }
            rgb4 = make_float4(rgb.x, rgb.y, rgb.z, 0.f);
            want = true;
          }
        }
        const int slot = i * SL + lobe * p.S + L.slot_base + s;
        if (active) p.contrib[slot] = rgb4;  // zero unless `want`
        const int qi = warp_append(p.counts + 2, want);
        if (want) {
          const f3 o = offset_p(c.P, c.Poffset, dot3(Ld, c.Ng) < 0 ? -1 : 1);
          const f3 d = scale3(Ldist * (1.0f - 0.0001f), Ld);
          float4* rp = reinterpret_cast<float4*>(p.sray + qi);
          rp[0] = make_float4(o.x, o.y, o.z, d.x);
          rp[1] = make_float4(d.y, d.z, 1.0f, time);
          p.sslot[qi] = slot;
        }
      }
    }
  }

  // ---- mirror lobe (std.go:194-261, bsdf/specular.go) with either Fresnel model ----
  {
    bool want = false;
    f3 wo = mk3(0, 0, 1);
    float4 T4 = make_float4(0, 0, 0, 0);
    if (active && m.spec_weight > 0.0f && m.spec_rough == 0.0f) {
      const f3 refl = reflect_z(omegaI);
      wo = basis_expand(fr.U, fr.V, fr.N, normalize3(refl));
      const f3 o = basis_project(fr.U, fr.V, fr.N, wo);
      const double pdf = dot3(o, refl) < 0.9999f ? 0.0 : 1.0;
      if (!(dot3(wo, c.Ng) <= 0.0f)) {
        Spec4 rho;
        rho.c[0] = rho.c[1] = rho.c[2] = rho.c[3] = 0.f;
        if (!(dot3(o, refl) < 0.9999f)) {
          rho = spec_from_rgb(fresnel_kr(m, omegaI.z), hero);
          const float az = fabsf(o.z);
#pragma unroll
          for (int k = 0; k < 4; k++) rho.c[k] *= az;
        }
        const float inv = 1.0f / (float)pdf;
#pragma unroll
        for (int k = 0; k < 4; k++) rho.c[k] *= inv;
        const f3 rgb = spec_to_rgb(rho, hero);
        T4 = make_float4(rgb.x * m.spec_colour.x, rgb.y * m.spec_colour.y, rgb.z * m.spec_colour.z, m.spec_weight);
        want = (level + 1 <= 3) || p.trace_last_level;
      }
    }
    if (active) p.T[(size_t)level * p.P + path] = T4;
    const int qi = warp_append(p.counts + qout, want);
    if (want) {
      const f3 o = offset_p(c.P, c.Poffset, 1);
      float4* rp = reinterpret_cast<float4*>(p.rayq[qout] + qi);
      rp[0] = make_float4(o.x, o.y, o.z, wo.x);
      rp[1] = make_float4(wo.y, wo.z, __int_as_float(0x7f800000), time);
      p.pathq[qout][qi] = path;
    }
  }
}

}  // namespace vg
