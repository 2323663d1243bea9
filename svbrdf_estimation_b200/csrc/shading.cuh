// shading.cuh - per-pixel Cook-Torrance/GGX shading and its analytic adjoint, in registers.
//
// Device-side restatement (NOT a translation) of what the reference evaluates with ~100 eager
// tensor ops per render call (development/multiImage_pytorch/renderers.py:8-104).  The algebra
// is rearranged for the FP32 pipes of sm_100a:
//
//   * everything that does not depend on the maps (wi, wo, |wi+wo|, Fresnel power, light
//     falloff) is computed once per (pixel, scene record) and shared by the input and the
//     target map ("Geo");
//   * the half vector is never normalised: n.h = (n.wi + n.wo) / |wi+wo| and v.h = |wi+wo| / 2;
//   * the Smith term needs no division:  (1 + sqrt(1 + a2 (1-c^2)/c^2)) * c = c + sqrt(c^2 (1-a2) + a2),
//     so  F G D / (4 VN LN) = F * a2 / (pi q^2 (VN + wV)(LN + wL));
//   * the xi() Heaviside factors of renderers.py:15-16,27,38 are identically 1 on this path
//     (all their arguments are clamped to >= 1e-3 first) and are dropped;
//   * divisions / square roots / logs are single MUFU operations (rcp/rsqrt/lg2 .approx.ftz,
//     <= 2 ulp), the natural-log scale ln2 is applied once to the reduced loss.
//
// Two template axes:
//   T  - the lane type: `float` (one pixel per thread) or `F2` (two horizontally adjacent pixels per
//        thread, packed in a 64-bit register pair).  Blackwell's FADD2/FMUL2/FFMA2 execute one F2
//        operation per issue slot, which is what lifts this issue-bound kernel: every add/mul/fma
//        below is a single instruction for two pixels; only min/max/select/compare and the MUFU
//        operations are issued per lane.
//   NC - colour channels per pass: 3 (the three channels share one roughness value, which is what
//        the model and the dataset always produce, utils.py:78-80) or 1 (one colour channel with
//        its own roughness; the API allows three different roughness channels, utils.py:48-51, and
//        the kernels then run three single-channel passes - the loss and every gradient decompose
//        by colour channel).
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

// ---- scalar MUFU wrappers --------------------------------------------------------------------
#ifdef SVB_HOST_EMULATION
#ifdef SVB_EMU_MUFU_NOISE
// optional model of the MUFU approximation error (pseudo-random, +-SVB_EMU_MUFU_NOISE relative) for
// sensitivity studies of the algebra on the CPU
SVB_DEV float mufu_noise(float v) {
    unsigned b; __builtin_memcpy(&b, &v, 4);
    b *= 2654435761u;
    return v * (1.0f + ((float)((b >> 8) & 0xffff) / 32767.5f - 1.0f) * (float)(SVB_EMU_MUFU_NOISE));
}
#else
SVB_DEV float mufu_noise(float v) { return v; }
#endif
SVB_DEV float mufu_rcp(float x) { return mufu_noise(1.0f / x); }
SVB_DEV float mufu_rsqrt(float x) { return mufu_noise(1.0f / sqrtf(x)); }
SVB_DEV float mufu_lg2(float x) { return log2f(x); }
SVB_DEV float mufu_sqrt(float x) { return mufu_noise(sqrtf(x)); }
#else
SVB_DEV float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
SVB_DEV float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
SVB_DEV float mufu_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
SVB_DEV float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif

// ---- the packed lane type -----------------------------------------------------------------------
#ifdef SVB_HOST_EMULATION
struct F2 { float x, y; };
SVB_DEV F2 mk2(float a, float b) { return F2{a, b}; }
SVB_DEV float lo(F2 a) { return a.x; }
SVB_DEV float hi(F2 a) { return a.y; }
SVB_DEV F2 operator+(F2 a, F2 b) { return F2{a.x + b.x, a.y + b.y}; }
SVB_DEV F2 operator-(F2 a, F2 b) { return F2{a.x - b.x, a.y - b.y}; }
SVB_DEV F2 operator*(F2 a, F2 b) { return F2{a.x * b.x, a.y * b.y}; }
SVB_DEV F2 vfma(F2 a, F2 b, F2 c) { return F2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#else
struct F2 { unsigned long long v; };
SVB_DEV F2 mk2(float a, float b) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
SVB_DEV float lo(F2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); (void)y; return x; }
SVB_DEV float hi(F2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); (void)x; return y; }
SVB_DEV F2 operator+(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SVB_DEV F2 operator-(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SVB_DEV F2 operator*(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SVB_DEV F2 vfma(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
#endif
// scalar (warp-uniform or literal) operands are broadcast; sm_100a SASS takes them as UR.F32 / immediates
SVB_DEV F2 operator+(F2 a, float b) { return a + mk2(b, b); }
SVB_DEV F2 operator-(F2 a, float b) { return a - mk2(b, b); }
SVB_DEV F2 operator-(float a, F2 b) { return mk2(a, a) - b; }
SVB_DEV F2 operator*(F2 a, float b) { return a * mk2(b, b); }
SVB_DEV F2 operator*(float a, F2 b) { return mk2(a, a) * b; }
SVB_DEV F2 vfma(F2 a, F2 b, float c) { return vfma(a, b, mk2(c, c)); }
SVB_DEV F2 vfma(F2 a, float b, F2 c) { return vfma(a, mk2(b, b), c); }
SVB_DEV F2 vfma(F2 a, float b, float c) { return vfma(a, mk2(b, b), mk2(c, c)); }
SVB_DEV float vfma(float a, float b, float c) { return fmaf(a, b, c); }

// per-lane operations (no packed form in the ISA)
struct B2 { bool x, y; };
SVB_DEV float vneg(float a) { return -a; }
SVB_DEV F2 vneg(F2 a) { return mk2(-lo(a), -hi(a)); }
SVB_DEV float vmax(float a, float m) { return fmaxf(a, m); }
#ifdef SVB_ABL_NOMINMAX  // timing ablation only (wrong results)
SVB_DEV F2 vmax(F2 a, float m) { return a; }
#else
SVB_DEV F2 vmax(F2 a, float m) { return mk2(fmaxf(lo(a), m), fmaxf(hi(a), m)); }
#endif
SVB_DEV float vmin(float a, float m) { return fminf(a, m); }
SVB_DEV F2 vmin(F2 a, float m) { return mk2(fminf(lo(a), m), fminf(hi(a), m)); }
SVB_DEV float vabs(float a) { return fabsf(a); }
SVB_DEV F2 vabs(F2 a) { return mk2(fabsf(lo(a)), fabsf(hi(a))); }
#ifdef SVB_ABL_NOMUFU   // timing ablation only (wrong results): every MUFU pair becomes one packed multiply
SVB_DEV F2 operator*(F2 a, F2 b);
SVB_DEV F2 vrcp(F2 a) { return a * mk2(1.0009765625f, 1.0009765625f); }
SVB_DEV F2 vrsqrt(F2 a) { return a * mk2(0.9990234375f, 0.9990234375f); }
SVB_DEV F2 vlg2(F2 a) { return a * mk2(1.0029296875f, 1.0029296875f); }
#else
SVB_DEV F2 vrcp(F2 a) { return mk2(mufu_rcp(lo(a)), mufu_rcp(hi(a))); }
SVB_DEV F2 vrsqrt(F2 a) { return mk2(mufu_rsqrt(lo(a)), mufu_rsqrt(hi(a))); }
SVB_DEV F2 vlg2(F2 a) { return mk2(mufu_lg2(lo(a)), mufu_lg2(hi(a))); }
#endif
SVB_DEV float vsqrt(float a) { return mufu_sqrt(a); }
SVB_DEV F2 vsqrt(F2 a) { return mk2(mufu_sqrt(lo(a)), mufu_sqrt(hi(a))); }
SVB_DEV float vrcp(float a) { return mufu_rcp(a); }
SVB_DEV float vrsqrt(float a) { return mufu_rsqrt(a); }
SVB_DEV float vlg2(float a) { return mufu_lg2(a); }
SVB_DEV bool vge(float a, float b) { return a >= b; }
SVB_DEV B2 vge(F2 a, float b) { return B2{lo(a) >= b, hi(a) >= b}; }
SVB_DEV float vsel(bool m, float a, float b) { return m ? a : b; }
SVB_DEV F2 vsel(B2 m, F2 a, float b) { return mk2(m.x ? lo(a) : b, m.y ? hi(a) : b); }
// sign(d) * v with sign(0) = 0 (torch.sign / l1_loss backward)
SVB_DEV float vsigned(float d, float v) { return (d > 0.f) ? v : ((d < 0.f) ? -v : 0.f); }
SVB_DEV F2 vsigned(F2 d, F2 v) { return mk2(vsigned(lo(d), lo(v)), vsigned(hi(d), hi(v))); }
// sign(d) * v for v >= 0, sign(0) = 0: copysign + one compare-select per lane
SVB_DEV float vsignz(float d, float v) { return d != 0.f ? copysignf(v, d) : 0.f; }
SVB_DEV F2 vsignz(F2 d, F2 v) { return mk2(vsignz(lo(d), lo(v)), vsignz(hi(d), hi(v))); }
// |v| with the sign of d (one LOP3 per lane); d == 0 counts as positive - used where an exact zero of d
// is handled separately (see loss_kernel: bitwise-identical inputs are masked per pixel / per channel)
SVB_DEV float vcopysign(float v, float d) { return copysignf(v, d); }
SVB_DEV F2 vcopysign(F2 v, F2 d) { return mk2(copysignf(lo(v), lo(d)), copysignf(hi(v), hi(d))); }
// (d != 0) ? v : 0 per lane
SVB_DEV float vnonzero(float d, float v) { return d != 0.f ? v : 0.f; }
SVB_DEV F2 vnonzero(F2 d, F2 v) { return mk2(lo(d) != 0.f ? lo(v) : 0.f, hi(d) != 0.f ? hi(v) : 0.f); }
// (a >= b) ? v : 0 per lane, as a float factor
SVB_DEV float vstep(float a, float b, float v) { return a >= b ? v : 0.f; }
SVB_DEV F2 vstep(F2 a, float b, float v) { return mk2(lo(a) >= b ? v : 0.f, hi(a) >= b ? v : 0.f); }
SVB_DEV float hsum(float a) { return a; }
SVB_DEV float hsum(F2 a) { return lo(a) + hi(a); }

template <typename T> struct LaneTraits;
template <> struct LaneTraits<float> {
    static constexpr int kLanes = 1;
    typedef bool Mask;
    static SVB_DEV float splat(float c) { return c; }
};
template <> struct LaneTraits<F2> {
    static constexpr int kLanes = 2;
    typedef B2 Mask;
    static SVB_DEV F2 splat(float c) { return mk2(c, c); }
};

// ---- per (pixel, scene record): map-independent geometry ---------------------------------------
template <typename T>
struct Geo {
    T wix, wiy, wiz;   // unit vector to the light          (renderers.py:91-93)
    T wox, woy, woz;   // unit vector to the camera         (renderers.py:79-82)
    T ih;              // 1 / |wi + wo|                      (renderers.py:45)
    T p5, omp5;        // (1-VH)^5 and its complement        (renderers.py:32,49)
    T fall;            // 1 / |light - p|^2                  (renderers.py:99)
    T hx, hy, hz, ih2; // wi + wo and 1 / |wi + wo|^2 (only the accurate-highlight forward reads them)
};

// What one scene record (9 floats: camera xyz, light xyz, colour rgb) contributes that is the same for every pixel
// of an image ROW: the y/z parts of the vectors to the light and to the camera.  The kernels either form it per
// thread from the constant bank (ConstRecs) or once per warp into shared memory (RowTabRecs, kernels.cu) - both
// with exactly these operations, so the two paths are bit-identical.
struct RecScalars {
    float sx, vx;          // light x, camera x
    float ly, lz, lyz;     // light - p: y and z components, ly^2 + lz^2
    float vy, vz, vyz;     // camera - p
    float col[3];          // light colour / pi
};
SVB_DEV RecScalars rec_scalars(const float* __restrict__ s, float y) {
    RecScalars r;
    r.sx = s[3]; r.vx = s[0];
    r.ly = s[4] - y; r.lz = s[5]; r.vy = s[1] - y; r.vz = s[2];
    r.lyz = fmaf(r.ly, r.ly, r.lz * r.lz); r.vyz = fmaf(r.vy, r.vy, r.vz * r.vz);
    r.col[0] = s[6] * kInvPi; r.col[1] = s[7] * kInvPi; r.col[2] = s[8] * kInvPi;
    return r;
}
// Record source reading the records where the launcher put them (kernel parameter block / host memory).
struct ConstRecs {
    const float* rec;      // [N,9]
    float y;               // row coordinate of this thread's pixel(s)
    SVB_DEV RecScalars get(int k) const { return rec_scalars(rec + k * 9, y); }
};

// x is per lane; the row coordinate is inside r (the lanes of an F2 are horizontal neighbours).
template <typename T>
SVB_DEV Geo<T> make_geo(T x, const RecScalars& r) {
    Geo<T> g;
    const float ly = r.ly, lz = r.lz, vy = r.vy, vz = r.vz;
    const T lx = LaneTraits<T>::splat(r.sx) - x;
    const T il = vrsqrt(vfma(lx, lx, r.lyz));
    g.wix = lx * il; g.wiy = il * ly; g.wiz = il * lz;
    g.fall = il * il;
    const T vx = LaneTraits<T>::splat(r.vx) - x;
    const T iv = vrsqrt(vfma(vx, vx, r.vyz));
    g.wox = vx * iv; g.woy = iv * vy; g.woz = iv * vz;
    // |wi+wo| from the summed vector itself (not from 2 + 2 wi.wo): the rounding of the two
    // normalisations then cancels in n.h to first order exactly where the GGX lobe is sharpest
    // (mirror configuration, n.wi ~ n.wo), which keeps the fp32 result near the reference's fp64 one.
    const T hx = g.wix + g.wox, hy = g.wiy + g.woy, hz = g.wiz + g.woz;
    const T hh = vfma(hx, hx, vfma(hy, hy, hz * hz));
    g.hx = hx; g.hy = hy; g.hz = hz;
    g.ih = vrsqrt(hh);       // (a Newton step here was measured on B200: no accuracy gain, the residual error is fp32 rounding of n.h itself)
    // wo.h = |wi+wo|/2 for unit vectors, clamped at 1e-3 (renderers.py:49); 1 - max(t/2, c) = min(1 - t/2, 1 - c),
    // bit-identical (t/2 is exact) and one packed operation shorter
    const T m = vmin(vfma(hh * g.ih, -0.5f, 1.f), 1.f - kClamp), m2 = m * m;
    g.p5 = (m2 * m2) * m;
    g.omp5 = 1.f - g.p5;
    g.ih2 = g.ih * g.ih;
    return g;
}
// s points at one scene record in the constant bank / host memory
template <typename T>
SVB_DEV Geo<T> make_geo(T x, float y, const float* __restrict__ s) { return make_geo<T>(x, rec_scalars(s, y)); }

// ---- per pixel: quantities of one SVBRDF map that do not depend on the scene record ----------
// Scaling convention: everything below works with f' = pi * f (BRDF value times pi) so that the
// diffuse albedo enters unscaled and 1/pi is folded into the light term E' = colour * falloff / pi,
// which is warp-uniform (one multiply per record instead of per pixel and map).
//
// BRDF blend (renderers.py:18-20,29-32,62-65).  With F = s + (1-s) p5 = s (1-p5) + p5:
//     f' = (1-F) d + F S = (1-p5) * (d (1-s) + s S) + p5 * S
// - a sum of non-negative terms (no cancellation even where F ~ 1e-3 multiplies S ~ 1e4), and per record
// and channel only two fused operations on the per-pixel constants s and dk = d (1-s).
template <typename T, int NC>
struct Pix {
    T nx, ny, nz;      // normal, used as given (not re-normalised; renderers.py:84)
    T s[NC];           // specular albedo
    T dk[NC];          // diffuse albedo * (1 - specular albedo)
    T a2;              // alpha^2 = clamp(rough,1e-3)^4       (renderers.py:23-24,87)
    T oma2;            // 1 - alpha^2
    T omn2;            // 1 - |n|^2 to ~1e-8 absolute (only the accurate-highlight forward reads it)
};

// n[3] normals; d, s: NC diffuse / specular channels; rough: the roughness channel of this pass.
template <typename T, int NC>
SVB_DEV Pix<T, NC> make_pix(const T* n, const T* d, const T* s, T rough) {
    Pix<T, NC> p;
    p.nx = n[0]; p.ny = n[1]; p.nz = n[2];
#pragma unroll
    for (int c = 0; c < NC; ++c) { p.s[c] = s[c]; p.dk[c] = d[c] * (1.f - s[c]); }
    const T r = vmax(rough, kClamp);
    const T a = r * r;
    p.a2 = a * a;
    p.oma2 = 1.f - p.a2;
    {   // 1 - |n|^2 with the products' rounding errors recovered by FMA (exact residuals) and the dominant z term
        // subtracted first (1 - nz^2 is exact for nz^2 in [0.5, 2], the case of every upper-hemisphere unit normal)
        const T px = p.nx * p.nx, py = p.ny * p.ny, pz = p.nz * p.nz;
        const T ex = vfma(p.nx, p.nx, vneg(px)), ey = vfma(p.ny, p.ny, vneg(py)), ez = vfma(p.nz, p.nz, vneg(pz));
        p.omn2 = (((1.f - pz) - py) - px) - ((ex + ey) + ez);
    }
    return p;
}

// ---- forward shading of one map under one scene record -----------------------------------------
template <typename T>
struct Fwd {
    T NHr, VNr, LNr;       // unclamped dots (for the clamp masks)
    T NH, VN, LN, LN0;
    T VN2, LN2;
    T q;                   // GGX denominator before the clamp
    T iq, zV, zL;          // 1/q, 1/(wV (VN+wV)), 1/(wL (LN+wL))
    T wV, wL;
    T iR;                  // 1 / (q^2 (VN+wV)(LN+wL))
    T S;                   // pi * G D / (4 VN LN) = a2 * iR
};

// Everything of the specular lobe that is shared by the colour channels.
// ACC (forward only): the GGX denominator q = 1 - NH^2 (1-a2) is where fp32 loses the highlights - NH^2 -> 1 and q -> a2,
// so the ~1.5e-7 absolute rounding error of n.h = (n.wi+n.wo)/|wi+wo| becomes a relative error of up to 1e-3 in q and
// twice that in D (the reference's own fp32 evaluation has the same problem, profiles/r1_accuracy_study.txt).  With
// Lagrange's identity  1 - NH^2 = (1 - |n|^2) + |n x h|^2 / |h|^2,  h = wi + wo, the small quantity is computed from small
// quantities (cross-product components ~ sqrt(1-NH^2), relative error ~1e-6) and a per-pixel constant: +10 packed
// operations per map and record, used where the images themselves are the product (render_forward, HBM-bound anyway).
template <typename T, int NC, bool BWD, bool ACC = false>
SVB_DEV void shade_fwd(const Geo<T>& g, const Pix<T, NC>& p, Fwd<T>& o) {
    o.LNr = vfma(p.nx, g.wix, vfma(p.ny, g.wiy, p.nz * g.wiz));
    o.VNr = vfma(p.nx, g.wox, vfma(p.ny, g.woy, p.nz * g.woz));
    o.NHr = (o.LNr + o.VNr) * g.ih;
    o.NH = vmax(o.NHr, kClamp); o.VN = vmax(o.VNr, kClamp); o.LN = vmax(o.LNr, kClamp);
    o.LN0 = vmax(o.LNr, 0.f);                                     // renderers.py:96
    o.VN2 = o.VN * o.VN; o.LN2 = o.LN * o.LN;
    if (ACC) {
        const T cx = vfma(p.ny, g.hz, vneg(p.nz * g.hy)), cy = vfma(p.nz, g.hx, vneg(p.nx * g.hz)),
                cz = vfma(p.nx, g.hy, vneg(p.ny * g.hx));
        const T c2 = vfma(cx, cx, vfma(cy, cy, cz * cz));
        const T om = vfma(c2, g.ih2, p.omn2);                                        // 1 - NHr^2
        // clamp(n.h, min=1e-3) (renderers.py:48): below the bound NH^2 = 1e-6 whatever the raw value
        const typename LaneTraits<T>::Mask live = vge(o.NHr, kClamp);
        const T nh2 = vsel(live, 1.f - om, kClamp * kClamp);
        const T omc = vsel(live, om, 1.f - kClamp * kClamp);
        o.q = vfma(nh2, p.a2, omc);                                                  // NH^2 a2 + (1 - NH^2) (renderers.py:26)
    } else {
        o.q = 1.f - (o.NH * o.NH) * p.oma2;                       // NH^2 a2 + 1 - NH^2 (renderers.py:26)
    }
    const T qc = vmax(o.q, kClamp);
    const T tV = vfma(o.VN2, p.oma2, p.a2);
    const T tL = vfma(o.LN2, p.oma2, p.a2);
    if (!BWD) {
        // forward only: sqrt is one MUFU operation; the backward also needs 1/sqrt, so there rsqrt and a multiply
        o.wV = vsqrt(tV); o.wL = vsqrt(tL);
        o.iR = vrcp((qc * qc) * ((o.VN + o.wV) * (o.LN + o.wL)));
    } else {
        const T rwV = vrsqrt(tV), rwL = vrsqrt(tL);
        o.wV = tV * rwV; o.wL = tL * rwL;
        const T PV = o.VN + o.wV, PL = o.LN + o.wL;
        const T iq = vrcp(qc), iPV = vrcp(PV), iPL = vrcp(PL);
        o.iq = iq; o.zV = rwV * iPV; o.zL = rwL * iPL;
        o.iR = (iq * iq) * (iPV * iPL);
    }
    o.S = p.a2 * o.iR;
}

// pi * BRDF value of the NC channels (see Pix): f'_c = (1-p5) (dk_c + s_c S) + p5 S
template <typename T, int NC>
SVB_DEV void brdf_values(const Geo<T>& g, const Pix<T, NC>& p, const Fwd<T>& o, T (&f)[NC]) {
    const T p5S = g.p5 * o.S;
#pragma unroll
    for (int c = 0; c < NC; ++c) f[c] = vfma(g.omp5, vfma(o.S, p.s[c], p.dk[c]), p5S);
}

// EL * f'_c + eps for a light term EL shared by the channels (grey light), without forming f':
//   eps + EL p5 S + (EL (1-p5)) (dk_c + s_c S)
template <typename T, int NC>
SVB_DEV void radiance_plus_eps_grey(const Geo<T>& g, const Pix<T, NC>& p, const Fwd<T>& o, T EL, float eps, T (&x)[NC]) {
    const T ELo = EL * g.omp5;
    const T c0 = vfma(EL * o.S, g.p5, eps);
#pragma unroll
    for (int c = 0; c < NC; ++c) x[c] = vfma(ELo, vfma(o.S, p.s[c], p.dk[c]), c0);
}

// ---- gradient accumulators of one pixel (summed over scene records) ----------------------------
// U_c = sum gf_c (1-p5) and W_c = sum gf_c (1-p5) S with gf_c = d loss / d f'_c: the diffuse and specular
// albedo gradients are assembled from them once per pixel (acc_albedo_grads), because
// d f'_c / d d_c = (1-p5)(1-s_c) and d f'_c / d s_c = (1-p5)(S - d_c) only involve per-pixel constants besides.
template <typename T, int NC>
struct Acc {
    T gn[3], U[NC], W[NC], ga2[NC];
};
template <typename T, int NC>
SVB_DEV void acc_zero(Acc<T, NC>& a) {
    const T z = LaneTraits<T>::splat(0.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) a.gn[c] = z;
#pragma unroll
    for (int c = 0; c < NC; ++c) { a.U[c] = z; a.W[c] = z; a.ga2[c] = z; }
}
// d loss / d diffuse_c and d loss / d specular_c from the accumulators (d, s: the pixel's albedo channels)
template <typename T, int NC>
SVB_DEV void acc_albedo_grads(const Acc<T, NC>& a, const T* d, const T* s, T (&gd)[NC], T (&gs)[NC]) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        gd[c] = a.U[c] * (1.f - s[c]);
        gs[c] = a.W[c] - d[c] * a.U[c];
    }
}

// gf[c] = d loss / d f'_c and gLN0 = d loss / d LN0 (through the radiance product, renderers.py:96-100)
// for this (pixel, record).  Adds the adjoint of shade_fwd / brdf_values into acc.
// Clamp masks are applied as 0/1 (or 0/2) float factors: one compare-and-set per lane plus a packed
// multiply is cheaper in register-file cycles than predicated moves of register pairs.
template <typename T, int NC>
SVB_DEV void shade_bwd(const Geo<T>& g, const Pix<T, NC>& p, const Fwd<T>& o, const T (&gf)[NC], T gLN0,
                       Acc<T, NC>& acc) {
    // w = 2 NH / q where the clamp on q passes (renderers.py:26), else 0
    const T w = (o.NH * o.iq) * vstep(o.q, kClamp, 2.f);
    // S * d ln S / d a2  =  iR - S * rest,   rest = 2 NH^2/q [q unclamped] + (1-VN^2) zV / 2 + (1-LN^2) zL / 2
    const T hV = vfma(o.VN2, -0.5f, 0.5f), hL = vfma(o.LN2, -0.5f, 0.5f);
    const T rest = vfma(o.NH, w, vfma(hV, o.zV, hL * o.zL));
    const T Tk = o.iR - o.S * rest;
    const T oS = g.omp5 * o.S;
    T gFsum;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        acc.U[c] = vfma(gf[c], g.omp5, acc.U[c]);
        acc.W[c] = vfma(gf[c], oS, acc.W[c]);
        const T gfF = gf[c] * vfma(p.s[c], g.omp5, g.p5);         // gf * F, F = s + (1-s)(1-VH)^5 (renderers.py:32)
        acc.ga2[c] = vfma(gfF, Tk, acc.ga2[c]);
        gFsum = (c == 0) ? gfF : gFsum + gfF;
    }
    const T G = gFsum * o.S;                                      // d loss / d ln S
    // d ln S / d NH = 4 NH (1-a2) / q [q unclamped];  d ln S / d VN = -(wV + VN (1-a2)) zV;  same for LN.
    // clamp(min=...) passes the gradient where the raw value is >= the bound (renderers.py:48-52,96).
    const T gNHr = (G * (p.oma2 * w)) * vstep(o.NHr, kClamp, 2.f);
#ifdef SVB_MASK_SELECT          // A/B builds: compare + select per lane (ALU pipe) instead of a 0/1 factor (FMA pipe)
    const T gVNp = vsel(vge(o.VNr, kClamp), G * (vfma(o.VN, p.oma2, o.wV) * o.zV), 0.f);
    const T gLNp = vsel(vge(o.LNr, kClamp), G * (vfma(o.LN, p.oma2, o.wL) * o.zL), 0.f);
#else
    const T gVNp = (G * (vfma(o.VN, p.oma2, o.wV) * o.zV)) * vstep(o.VNr, kClamp, 1.f);   // = -gVNr
    const T gLNp = (G * (vfma(o.LN, p.oma2, o.wL) * o.zL)) * vstep(o.LNr, kClamp, 1.f);   // = -gLNr (specular part)
#endif
    const T gh = gNHr * g.ih;                                      // n.h = (n.wi + n.wo) ih
    const T cw = gh - gVNp;
    const T ci = vfma(gLN0, vstep(o.LNr, 0.f, 1.f), gh - gLNp);
    acc.gn[0] = vfma(cw, g.wox, vfma(ci, g.wix, acc.gn[0]));
    acc.gn[1] = vfma(cw, g.woy, vfma(ci, g.wiy, acc.gn[1]));
    acc.gn[2] = vfma(cw, g.woz, vfma(ci, g.wiz, acc.gn[2]));
}

// d a2 / d rough = 4 r^3 where the clamp of renderers.py:87 passes gradient (rough >= 1e-3), else 0.
template <typename T>
SVB_DEV T rough_chain(T rough) {
    const T r = vsel(vge(rough, kClamp), rough, 0.f);
    return ((r * r) * r) * 4.f;
}

}  // namespace svb
