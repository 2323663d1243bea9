// shading.cuh - per-pixel Cook-Torrance/GGX shading and its analytic adjoint, in registers.
//
// Device-side restatement (NOT a translation) of what the reference evaluates with ~100 eager
// tensor ops per render call (development/multiImage_pytorch/renderers.py:8-104).  The algebra
// is rearranged for the FP32 pipes of sm_100a:
//
//   * everything that does not depend on the maps (wi, wo, |wi+wo|, Fresnel power, light
//     falloff) is computed once per (pixel, scene record) and shared by the input and the
//     target map ("Geo");
//   * the half vector is never formed: n.h = (n.wi + n.wo) / |wi+wo| and
//     v.h = (1 + wi.wo) / |wi+wo|, |wi+wo|^2 = 2 + 2 wi.wo;
//   * the Smith term needs no division:  (1 + sqrt(1 + a2 (1-c^2)/c^2)) * c = c + sqrt(c^2 (1-a2) + a2),
//     so  F G D / (4 VN LN) = F * a2 / (pi q^2 (VN + wV)(LN + wL));
//   * the xi() Heaviside factors of renderers.py:15-16,27,38 are identically 1 on this path
//     (all their arguments are clamped to >= 1e-3 first) and are dropped;
//   * divisions / square roots / logs are single MUFU operations (rcp/rsqrt/lg2 .approx.ftz,
//     <= 2 ulp), the natural-log scale ln2 is applied once to the reduced loss.
//
// RC is the number of distinct roughness channels: the API carries three (utils.py:48-51) but
// the model and the dataset always replicate one (utils.py:78-80); RC=1 is the fast path.
#pragma once
#ifdef SVB_HOST_EMULATION
// tests/emulation/ compiles this header with g++ to check the algebra (forward and adjoint) against
// the oracle on machines without a GPU.  MUFU approximations become exact libm calls there.
#include <cmath>
#define SVB_DEV inline
#else
#include <cuda_runtime.h>
#define SVB_DEV __device__ __forceinline__
#endif

namespace svb {

constexpr float kInvPi = 0.318309886183790671538f;
constexpr float kLn2 = 0.693147180559945309417f;
constexpr float kClamp = 1e-3f;        // renderers.py:26,48-52,87
constexpr float kEpsRender = 0.1f;     // losses.py:46
constexpr float kEpsL1 = 0.01f;        // losses.py:13

#ifdef SVB_HOST_EMULATION
SVB_DEV float mufu_rcp(float x) { return 1.0f / x; }
SVB_DEV float mufu_rsqrt(float x) { return 1.0f / sqrtf(x); }
SVB_DEV float mufu_lg2(float x) { return log2f(x); }
#else
SVB_DEV float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
SVB_DEV float mufu_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
SVB_DEV float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif

// ---- per (pixel, scene record): map-independent geometry ---------------------------------------
struct Geo {
    float wix, wiy, wiz;   // unit vector to the light          (renderers.py:91-93)
    float wox, woy, woz;   // unit vector to the camera         (renderers.py:79-82)
    float ih;              // 1 / |wi + wo|                      (renderers.py:45)
    float p5, omp5;        // (1-VH)^5 and its complement        (renderers.py:32,49)
    float e0, e1, e2;      // light colour * 1/dist^2            (renderers.py:98-100)
};

// s points at one scene record (9 floats: camera xyz, light xyz, colour rgb) in the constant bank.
SVB_DEV Geo make_geo(float x, float y, const float* __restrict__ s) {
    Geo g;
    const float lx = s[3] - x, ly = s[4] - y, lz = s[5];
    const float il = mufu_rsqrt(fmaf(lx, lx, fmaf(ly, ly, lz * lz)));
    g.wix = lx * il; g.wiy = ly * il; g.wiz = lz * il;
    const float fall = il * il;
    g.e0 = s[6] * fall; g.e1 = s[7] * fall; g.e2 = s[8] * fall;
    const float vx = s[0] - x, vy = s[1] - y, vz = s[2];
    const float iv = mufu_rsqrt(fmaf(vx, vx, fmaf(vy, vy, vz * vz)));
    g.wox = vx * iv; g.woy = vy * iv; g.woz = vz * iv;
    const float c = fmaf(g.wix, g.wox, fmaf(g.wiy, g.woy, g.wiz * g.woz));
    const float t = fmaf(2.f, c, 2.f);                 // |wi+wo|^2
    g.ih = mufu_rsqrt(t);
    const float vh = fmaxf(0.5f * t * g.ih, kClamp);   // wo.h = (1+c)/|wi+wo|, clamped (renderers.py:49)
    const float m = 1.f - vh, m2 = m * m;
    g.p5 = m2 * m2 * m;
    g.omp5 = 1.f - g.p5;
    return g;
}

// ---- per pixel: quantities of one SVBRDF map that do not depend on the scene record ----------
template <int RC>
struct Pix {
    float nx, ny, nz;      // normal, used as given (not re-normalised; renderers.py:84)
    float kd[3];           // diffuse / pi                        (renderers.py:18-20)
    float s[3];            // specular albedo
    float a2[RC];          // alpha^2 = clamp(rough,1e-3)^4       (renderers.py:23-24,87)
    float oma2[RC];        // 1 - alpha^2
    float rg[RC];          // clamp(rough) where the clamp passes gradient, else 0
};

// v[12] = the pixel's 12 channels in API order.
template <int RC>
SVB_DEV Pix<RC> make_pix(const float (&v)[12]) {
    Pix<RC> p;
    p.nx = v[0]; p.ny = v[1]; p.nz = v[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) { p.kd[c] = v[3 + c] * kInvPi; p.s[c] = v[9 + c]; }
#pragma unroll
    for (int j = 0; j < RC; ++j) {
        const float r = fmaxf(v[6 + j], kClamp);
        const float a = r * r;
        p.a2[j] = a * a;
        p.oma2[j] = 1.f - p.a2[j];
        p.rg[j] = (v[6 + j] >= kClamp) ? r : 0.f;       // clamp(min) passes gradient at equality
    }
    return p;
}

// ---- forward shading of one map under one scene record -----------------------------------------
template <int RC>
struct Fwd {
    float NHr, VNr, LNr;           // unclamped dots (for the clamp masks)
    float NH, VN, LN, LN0;
    float NH2, VN2, LN2;
    float q[RC];                   // GGX denominator before the clamp
    float iq[RC], zV[RC], zL[RC];  // 1/q, 1/(wV (VN+wV)), 1/(wL (LN+wL))   (BWD only)
    float wV[RC], wL[RC];
    float iR[RC];                  // 1 / (pi q^2 (VN+wV)(LN+wL))
    float S[RC];                   // G D / (4 VN LN)
    float F[3], f[3], Smkd[3];     // Fresnel, BRDF value, S - kd
};

template <int RC, bool BWD>
SVB_DEV void shade_fwd(const Geo& g, const Pix<RC>& p, Fwd<RC>& o) {
    o.LNr = fmaf(p.nx, g.wix, fmaf(p.ny, g.wiy, p.nz * g.wiz));
    o.VNr = fmaf(p.nx, g.wox, fmaf(p.ny, g.woy, p.nz * g.woz));
    o.NHr = (o.LNr + o.VNr) * g.ih;
    o.NH = fmaxf(o.NHr, kClamp); o.VN = fmaxf(o.VNr, kClamp); o.LN = fmaxf(o.LNr, kClamp);
    o.LN0 = fmaxf(o.LNr, 0.f);                                    // renderers.py:96
    o.NH2 = o.NH * o.NH; o.VN2 = o.VN * o.VN; o.LN2 = o.LN * o.LN;
#pragma unroll
    for (int j = 0; j < RC; ++j) {
        o.q[j] = fmaf(o.NH2, -p.oma2[j], 1.f);                    // NH^2 a2 + 1 - NH^2 (renderers.py:26)
        const float qc = fmaxf(o.q[j], kClamp);
        const float tV = fmaf(o.VN2, p.oma2[j], p.a2[j]);
        const float tL = fmaf(o.LN2, p.oma2[j], p.a2[j]);
        const float rwV = mufu_rsqrt(tV), rwL = mufu_rsqrt(tL);
        o.wV[j] = tV * rwV; o.wL[j] = tL * rwL;
        const float PV = o.VN + o.wV[j], PL = o.LN + o.wL[j];
        if (BWD) {
            const float iq = mufu_rcp(qc), iPV = mufu_rcp(PV), iPL = mufu_rcp(PL);
            o.iq[j] = iq; o.zV[j] = rwV * iPV; o.zL[j] = rwL * iPL;
            o.iR[j] = (iq * iq) * (iPV * iPL) * kInvPi;
        } else {
            o.iR[j] = mufu_rcp((qc * qc) * (PV * PL)) * kInvPi;
        }
        o.S[j] = p.a2[j] * o.iR[j];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int j = (RC == 1) ? 0 : c;
        o.F[c] = fmaf(p.s[c], g.omp5, g.p5);                      // s + (1-s)(1-VH)^5 (renderers.py:32)
        o.Smkd[c] = o.S[j] - p.kd[c];
        o.f[c] = fmaf(o.F[c], o.Smkd[c], p.kd[c]);                // (1-F) kd + F S     (renderers.py:62-65)
    }
}

// ---- gradient accumulators of one pixel (summed over scene records) ----------------------------
struct Acc {
    float gn[3], gd[3], gs[3], ga2[3];
};
SVB_DEV void acc_zero(Acc& a) {
#pragma unroll
    for (int c = 0; c < 3; ++c) { a.gn[c] = 0.f; a.gd[c] = 0.f; a.gs[c] = 0.f; a.ga2[c] = 0.f; }
}

// A[c] = d loss / d radiance_c for this (pixel, record).  Adds the adjoint of shade_fwd + the
// radiance product (renderers.py:96-100) into acc.  gd is accumulated w.r.t. kd*pi ... i.e. the
// caller multiplies acc.gd by 1/pi once per pixel.
template <int RC>
SVB_DEV void shade_bwd(const Geo& g, const Pix<RC>& p, const Fwd<RC>& o,
                                          const float (&A)[3], Acc& acc) {
    const float E[3] = {g.e0, g.e1, g.e2};
    float T[RC], G[RC];
#pragma unroll
    for (int j = 0; j < RC; ++j) {
        const bool qpass = o.q[j] >= kClamp;
        // d ln S / d a2 * S  =  iR - S * rest,   rest = 2 NH^2/q [q unclamped] + (1-VN^2) zV / 2 + (1-LN^2) zL / 2
        const float tq = qpass ? 2.f * o.NH2 * o.iq[j] : 0.f;
        const float rest = fmaf(0.5f * (1.f - o.VN2), o.zV[j], fmaf(0.5f * (1.f - o.LN2), o.zL[j], tq));
        T[j] = fmaf(-o.S[j], rest, o.iR[j]);
        G[j] = 0.f;
    }
    float gLN0 = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int j = (RC == 1) ? 0 : c;
        const float AE = A[c] * E[c];
        const float gf = AE * o.LN0;
        gLN0 = fmaf(AE, o.f[c], gLN0);
        acc.gd[c] = fmaf(gf, 1.f - o.F[c], acc.gd[c]);
        acc.gs[c] = fmaf(gf * o.Smkd[c], g.omp5, acc.gs[c]);
        const float gfF = gf * o.F[c];
        acc.ga2[c] = fmaf(gfF, T[j], acc.ga2[c]);
        G[j] = fmaf(gfF, o.S[j], G[j]);                           // d loss / d ln S
    }
    float gNH = 0.f, gVN = 0.f, gLN = 0.f;
#pragma unroll
    for (int j = 0; j < RC; ++j) {
        const bool qpass = o.q[j] >= kClamp;
        const float cNH = qpass ? 4.f * o.NH * p.oma2[j] * o.iq[j] : 0.f;   // -2 dq/dNH / q
        const float cVN = fmaf(o.VN, p.oma2[j], o.wV[j]) * o.zV[j];        // -(d ln(VN+wV)/dVN), sign below
        const float cLN = fmaf(o.LN, p.oma2[j], o.wL[j]) * o.zL[j];
        gNH = fmaf(G[j], cNH, gNH);
        gVN = fmaf(-G[j], cVN, gVN);
        gLN = fmaf(-G[j], cLN, gLN);
    }
    // clamp(min=...) passes the gradient where the raw value is >= the bound (renderers.py:48-52,96)
    const float gNHr = (o.NHr >= kClamp) ? gNH : 0.f;
    const float gVNr = (o.VNr >= kClamp) ? gVN : 0.f;
    const float gLNr = ((o.LNr >= kClamp) ? gLN : 0.f) + ((o.LNr >= 0.f) ? gLN0 : 0.f);
    const float gh = gNHr * g.ih;                                  // n.h = (n.wi + n.wo) ih
    const float cw = gh + gVNr, ci = gh + gLNr;
    acc.gn[0] = fmaf(cw, g.wox, fmaf(ci, g.wix, acc.gn[0]));
    acc.gn[1] = fmaf(cw, g.woy, fmaf(ci, g.wiy, acc.gn[1]));
    acc.gn[2] = fmaf(cw, g.woz, fmaf(ci, g.wiz, acc.gn[2]));
}

// Turns the accumulators into the 12 API-order gradient channels; `scale` = upstream / element count.
template <int RC>
SVB_DEV void acc_to_grad(const Acc& a, const Pix<RC>& p, float scale, float (&out)[12]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int j = (RC == 1) ? 0 : c;
        out[c] = a.gn[c] * scale;
        out[3 + c] = a.gd[c] * (scale * kInvPi);
        const float r = p.rg[j];
        out[6 + c] = a.ga2[c] * (4.f * scale) * (r * r * r);       // d a2 / d rough = 4 r^3 [rough >= 1e-3]
        out[9 + c] = a.gs[c] * scale;
    }
}

}  // namespace svb
