// pixel_ops.cuh - what one thread does with its pixel(s): RenderingLoss / MixedLoss forward (+ backward)
// and LocalRenderer.render forward / backward over all scene records of a batch element.
//
// Shared by the CUDA kernels (kernels.cu) and by the host emulation used in the CPU tests
// (tests/emulation/emu.cpp, SVB_HOST_EMULATION) so that both run literally the same source.
// Memory access goes through an IO policy (device: __ldcs/__stcs; host: plain loads/stores).
//
// Reference semantics: renderers.py:67-104, losses.py:7-19,29-52 (development/multiImage_pytorch/).
#pragma once
#include "shading.cuh"

namespace svb {

constexpr int kRecFloats = 9;   // scene record: camera xyz, light xyz, light colour rgb

#ifdef SVB_HOST_EMULATION
#define SVB_WARP_ALL(pred) (pred)
#define SVB_UNROLL1
#else
#define SVB_WARP_ALL(pred) __all_sync(0xffffffffu, (pred))
#define SVB_UNROLL1 _Pragma("unroll 1")
#endif

// ---- small per-lane predicates ------------------------------------------------------------------------
SVB_DEV bool same3(const float (&v)[12]) { return v[6] == v[7] && v[7] == v[8]; }
SVB_DEV bool same3(const F2 (&v)[12]) {
    return lo(v[6]) == lo(v[7]) && lo(v[7]) == lo(v[8]) && hi(v[6]) == hi(v[7]) && hi(v[7]) == hi(v[8]);
}
SVB_DEV bool vne(float a, float b) { return a != b; }
SVB_DEV B2 vne(F2 a, F2 b) { return B2{lo(a) != lo(b), hi(a) != hi(b)}; }
SVB_DEV bool vne1(float a) { return a != 1.f; }
SVB_DEV B2 vne1(F2 a) { return B2{lo(a) != 1.f, hi(a) != 1.f}; }
SVB_DEV bool mor(bool a, bool b) { return a || b; }
SVB_DEV B2 mor(B2 a, B2 b) { return B2{a.x || b.x, a.y || b.y}; }
SVB_DEV bool mand(bool a, bool b) { return a && b; }
SVB_DEV B2 mand(B2 a, B2 b) { return B2{a.x && b.x, a.y && b.y}; }
// Does colour channel C of the two maps RENDER differently?  Decided on the parameters the shading actually consumes:
// the normal, the clamped roughness of the channel (two values below the clamp of renderers.py:87 are the same
// material), the specular albedo, and the diffuse albedo unless the specular albedo is exactly 1 on both maps
// ((1-F) d = 0 for any d then, renderers.py:18-20,32).  Where this is false the reference renders bit-identical values
// and the channel contributes exactly 0 to the loss and to every gradient (sign(0) = 0, losses.py:50).
template <typename T>
SVB_DEV typename LaneTraits<T>::Mask normals_differ(const T (&a)[12], const T (&b)[12]) {
    return mor(mor(vne(a[0], b[0]), vne(a[1], b[1])), vne(a[2], b[2]));
}
template <typename T>
SVB_DEV typename LaneTraits<T>::Mask rough_differs(T ra, T rb) { return vne(vmax(ra, kClamp), vmax(rb, kClamp)); }
template <typename T>
SVB_DEV typename LaneTraits<T>::Mask albedo_differs(T da, T sa, T db, T sb) {
    return mor(vne(sa, sb), mand(vne(da, db), vne1(sa)));      // second clause only matters when sa == sb
}
SVB_DEV bool all_or_none(bool a, bool b, bool c) { return (a == b) && (b == c); }
SVB_DEV bool all_or_none(B2 a, B2 b, B2 c) { return (a.x == b.x) && (b.x == c.x) && (a.y == b.y) && (b.y == c.y); }

// ---------------------------------------------------------------------------------------------
// RenderingLoss: forward (+ backward) of one thread's pixels over the N records of its batch element
// ---------------------------------------------------------------------------------------------
// C0 = first colour channel of the pass (0 for NC = 3; the pass's channel for NC = 1).
// GREY: every record of the launch has r == g == b light colour (always true for the scenes
// RenderingLoss samples, environment.py:27,52), so colour * falloff is formed once, not per channel.
// RS: the record source (shading.cuh ConstRecs, or the kernels' per-warp shared-memory table).
// Returns sum |log2 ratio| (ln2 and the mean are applied to the reduced loss).
// ACC: accurate-highlight GGX denominator in both forward evaluations (shading.cuh shade_fwd<..., ACC>).
template <typename T, int NC, int C0, bool BWD, bool GREY, bool ACC, typename RS>
SVB_DEV T loss_records(const Pix<T, NC>& pi, const Pix<T, NC>& pt, T x, const RS& recs, int N, Acc<T, NC>& acc) {
    T lsum = LaneTraits<T>::splat(0.f);
    SVB_UNROLL1
    for (int k = 0; k < N; ++k) {
        const RecScalars rk = recs.get(k);
        const Geo<T> g = make_geo<T>(x, rk);
        Fwd<T> fi, ft;
#ifdef SVB_ACCURATE_LOSS       // build option: accurate form in every loss kernel (~ +13 % time)
        constexpr bool kAcc = true;
#else
        constexpr bool kAcc = ACC;
#endif
#ifdef SVB_TARGET_FIRST         // A/B builds: source order of the two independent forward evaluations
        shade_fwd<T, NC, false, kAcc>(g, pt, ft);
        shade_fwd<T, NC, BWD, kAcc>(g, pi, fi);
#else
        shade_fwd<T, NC, BWD, kAcc>(g, pi, fi);
        shade_fwd<T, NC, false, kAcc>(g, pt, ft);
#endif
        // radiance + 0.1 of both maps (losses.py:46-47); E = light colour * falloff / pi
        T E[NC], fin[NC], xi[NC], xt[NC];
        if (GREY) {
            E[0] = g.fall * rk.col[C0];
            radiance_plus_eps_grey<T, NC>(g, pt, ft, E[0] * ft.LN0, kEpsRender, xt);
            if (BWD) {                                                     // the gradient of LN0 needs f' itself
                brdf_values<T, NC>(g, pi, fi, fin);
                const T ELi = E[0] * fi.LN0;
#pragma unroll
                for (int c = 0; c < NC; ++c) xi[c] = vfma(fin[c], ELi, kEpsRender);
            } else {
                radiance_plus_eps_grey<T, NC>(g, pi, fi, E[0] * fi.LN0, kEpsRender, xi);
            }
        } else {
            T ftg[NC];
            brdf_values<T, NC>(g, pi, fi, fin);
            brdf_values<T, NC>(g, pt, ft, ftg);
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                E[c] = g.fall * rk.col[C0 + c];
                xi[c] = vfma(fin[c], E[c] * fi.LN0, kEpsRender);
                xt[c] = vfma(ftg[c], E[c] * ft.LN0, kEpsRender);
            }
        }
        T A[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            // log(xt) - log(xi) as ONE lg2 of the ratio: 1/xi is needed for the gradient anyway.
            const T ix = vrcp(xi[c]);
            const T l = vlg2(xt[c] * ix);
            lsum = lsum + vabs(l);
            // d|l|/d xi = -sign(l)/xi: the accumulators carry +sign(l)/xi and the caller applies the minus
            // with the final scale.  sign(0) is taken as +1 here: an exact 0 only arises from channels that
            // render identically, and those are masked in loss_pixel (losses.py:50).
            if (BWD) A[c] = vcopysign(ix, l);
        }
        if (BWD) {
            T gf[NC], gLN0;
            if (GREY) {
                const T ELi = E[0] * fi.LN0;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    gf[c] = A[c] * ELi;
                    gLN0 = (c == 0) ? A[c] * fin[c] : vfma(A[c], fin[c], gLN0);
                }
                gLN0 = gLN0 * E[0];
            } else {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const T AE = A[c] * E[c];
                    gf[c] = AE * fi.LN0;
                    gLN0 = (c == 0) ? AE * fin[c] : vfma(AE, fin[c], gLN0);
                }
            }
            shade_bwd<T, NC>(g, pi, fi, gf, gLN0, acc);
        }
    }
    return lsum;
}

// One single-channel pass (general path): channel C of input/target with its own roughness.  Lanes whose
// channel-C parameters render identically are masked (their exact contribution is 0).
template <typename T, int C, bool BWD, bool GREY, bool ACC, typename RS>
SVB_DEV T loss_channel_pass(const T (&vi)[12], const T (&vt)[12], T x, const RS& recs, int N,
                            float nscale, typename LaneTraits<T>::Mask ndiff, T (&g)[12]) {
    const typename LaneTraits<T>::Mask live =
        mor(mor(ndiff, rough_differs<T>(vi[6 + C], vt[6 + C])), albedo_differs<T>(vi[3 + C], vi[9 + C], vt[3 + C], vt[9 + C]));
    const Pix<T, 1> pi = make_pix<T, 1>(&vi[0], &vi[3 + C], &vi[9 + C], vi[6 + C]);
    const Pix<T, 1> pt = make_pix<T, 1>(&vt[0], &vt[3 + C], &vt[9 + C], vt[6 + C]);
    Acc<T, 1> acc;
    acc_zero(acc);
    const T l = loss_records<T, 1, C, BWD, GREY, ACC, RS>(pi, pt, x, recs, N, acc);
    if (BWD) {
        const T ns = vsel(live, LaneTraits<T>::splat(nscale), 0.f);    // masked lanes: scale 0
        T gd[1], gs[1];
        acc_albedo_grads<T, 1>(acc, &vi[3 + C], &vi[9 + C], gd, gs);
#pragma unroll
        for (int j = 0; j < 3; ++j) g[j] = vfma(acc.gn[j], ns, g[j]);
        g[3 + C] = gd[0] * ns;
        g[6 + C] = (acc.ga2[0] * ns) * rough_chain(vi[6 + C]);
        g[9 + C] = gs[0] * ns;
    }
    return vsel(live, l, 0.f);
}

// Loss (log2 units, unscaled) and d loss / d input (scaled by `scale`) of one thread's pixels.
//
// Exact zeros: where a colour channel of the two maps renders identically (see albedo_differs above) the reference
// produces identical values and that channel contributes exactly 0 to the loss and to every gradient.  Here input
// and target run through differently scheduled instruction sequences, so that case is handled explicitly: the fast
// path requires that, per pixel, either all three channels differ or none does (identical pixels are zeroed at
// the end); anything else takes the channel-wise path where identical channels are masked.
template <typename T, bool BWD, bool GREY, bool ACC, typename RS>
SVB_DEV T loss_pixel_rs(const T (&vi)[12], const T (&vt)[12], T x, const RS& recs, int N, float scale, T (&g)[12]) {
    typedef typename LaneTraits<T>::Mask M;
    const M ndiff = normals_differ<T>(vi, vt);
    const float nscale = -scale;    // the accumulators carry the gradient with the opposite sign (loss_records)
    // fast path: the warp's pixels all carry one roughness value replicated on the three channels
    const M common = mor(ndiff, rough_differs<T>(vi[6], vt[6]));
    const M d0 = mor(common, albedo_differs<T>(vi[3], vi[9], vt[3], vt[9]));
    const M d1 = mor(common, albedo_differs<T>(vi[4], vi[10], vt[4], vt[10]));
    const M d2 = mor(common, albedo_differs<T>(vi[5], vi[11], vt[5], vt[11]));
    if (SVB_WARP_ALL(same3(vi) && same3(vt) && all_or_none(d0, d1, d2))) {
        const Pix<T, 3> pi = make_pix<T, 3>(&vi[0], &vi[3], &vi[9], vi[6]);
        const Pix<T, 3> pt = make_pix<T, 3>(&vt[0], &vt[3], &vt[9], vt[6]);
        Acc<T, 3> acc;
        acc_zero(acc);
        const T l = loss_records<T, 3, 0, BWD, GREY, ACC, RS>(pi, pt, x, recs, N, acc);
        if (BWD) {
            const T ns = vsel(d0, LaneTraits<T>::splat(nscale), 0.f);  // identically rendering pixels: scale 0
            const T chain = rough_chain(vi[6]) * ns;
            T gd[3], gs[3];
            acc_albedo_grads<T, 3>(acc, &vi[3], &vi[9], gd, gs);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                g[c] = acc.gn[c] * ns;
                g[3 + c] = gd[c] * ns;
                g[6 + c] = acc.ga2[c] * chain;
                g[9 + c] = gs[c] * ns;
            }
        }
        return vsel(d0, l, 0.f);
    }
    // general path: three single-channel passes (the loss and all gradients decompose by colour channel)
    if (BWD) { g[0] = g[1] = g[2] = LaneTraits<T>::splat(0.f); }
    T l = loss_channel_pass<T, 0, BWD, GREY, ACC, RS>(vi, vt, x, recs, N, nscale, ndiff, g);
    l = l + loss_channel_pass<T, 1, BWD, GREY, ACC, RS>(vi, vt, x, recs, N, nscale, ndiff, g);
    l = l + loss_channel_pass<T, 2, BWD, GREY, ACC, RS>(vi, vt, x, recs, N, nscale, ndiff, g);
    return l;
}
// records read where the launcher put them (kernel parameter block; host memory in the emulation)
template <typename T, bool BWD, bool GREY, bool ACC = false>
SVB_DEV T loss_pixel(const T (&vi)[12], const T (&vt)[12], T x, float y, const float* __restrict__ rec, int N,
                     float scale, T (&g)[12]) {
    const ConstRecs recs{rec, y};
    return loss_pixel_rs<T, BWD, GREY, ACC, ConstRecs>(vi, vt, x, recs, N, scale, g);
}

// Map-space L1 terms of SVBRDFL1Loss (losses.py:7-19) for one thread's pixels (natural-log units,
// unscaled sum); adds their gradient, scaled by `scale`, to g.
template <typename T, bool BWD>
SVB_DEV T l1_pixel(const T (&vi)[12], const T (&vt)[12], float scale, T (&g)[12]) {
    T s_lin = LaneTraits<T>::splat(0.f), s_lg2 = LaneTraits<T>::splat(0.f);
#pragma unroll
    for (int c = 0; c < 12; ++c) {
        const bool logged = (c >= 3 && c < 6) || c >= 9;      // diffuse and specular use log(x + 0.01)
        const T diff = vi[c] - vt[c];                         // sign(log a' - log b') = sign(a - b), exact at 0
        if (logged) {
            // |log a' - log b'| = |lg2(b'/a')| ln2 with b'/a' = 1 - (a - b)/a': exactly 1 (and the log exactly 0) for
            // identical values, and 1/a' is also d log(a')/da.  (This block runs once per pixel, but it is as long as a
            // record iteration, so every instruction saved here counts for MixedLoss.)
            const T ra = vrcp(vi[c] + kEpsL1);
            const T l = vlg2(vfma(vneg(diff), ra, 1.f));
            s_lg2 = s_lg2 + vabs(l);
            if (BWD) g[c] = g[c] + vsignz(diff, ra * scale);
        } else {
            s_lin = s_lin + vabs(diff);
            if (BWD) g[c] = g[c] + vsignz(diff, LaneTraits<T>::splat(scale));
        }
    }
    return vfma(s_lg2, kLn2, s_lin);
}

// ---------------------------------------------------------------------------------------------
// Model-output epilogue: the network emits 9 channels in [-1,1] (after tanh):
// normal xy, diffuse rgb, roughness, specular rgb (utils.py:52-56).  models.py:334-346 turns them into
// the 12-channel maps: n = normalize(3 e0, 3 e1, 1) (utils.py:82-86), roughness replicated x3
// (utils.py:78-80), diffuse / roughness / specular mapped to [0,1] by (x+1)/2 (utils.py:92-93).
// ---------------------------------------------------------------------------------------------
template <typename T>
SVB_DEV void decode_encoded(const T (&e)[9], T (&v)[12], T& inv_len) {
    const T vx = e[0] * 3.f, vy = e[1] * 3.f;
    const T l2 = vfma(vx, vx, vfma(vy, vy, 1.f));
    T y = vrsqrt(l2);
    y = vfma(y * 0.5f, 1.f - (l2 * y) * y, y);          // one Newton step: the reference divides by an exact sqrt
    inv_len = y;
    v[0] = vx * y; v[1] = vy * y; v[2] = y;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        v[3 + c] = vfma(e[2 + c], 0.5f, 0.5f);
        v[6 + c] = vfma(e[5], 0.5f, 0.5f);
        v[9 + c] = vfma(e[6 + c], 0.5f, 0.5f);
    }
}

// chain rule of decode_encoded: g[12] = d loss / d maps  ->  ge[9] = d loss / d encoded
template <typename T>
SVB_DEV void encode_grad(const T (&v)[12], T inv_len, const T (&g)[12], T (&ge)[9]) {
    const T ng = vfma(v[0], g[0], vfma(v[1], g[1], v[2] * g[2]));       // n . g_n
    const T k = inv_len * 3.f;
    ge[0] = (g[0] - v[0] * ng) * k;                                     // (I - n n^T) g_n / |v|, times d(3e)/de
    ge[1] = (g[1] - v[1] * ng) * k;
#pragma unroll
    for (int c = 0; c < 3; ++c) { ge[2 + c] = g[3 + c] * 0.5f; ge[6 + c] = g[9 + c] * 0.5f; }
    ge[5] = ((g[6] + g[7]) + g[8]) * 0.5f;
}

// ---------------------------------------------------------------------------------------------
// LocalRenderer.render forward / backward
// ---------------------------------------------------------------------------------------------
// out points at this thread's pixel(s) in images[b,0,0]; records advance by 3*HW floats.
// GREY: every record has r == g == b light colour (all scenes the reference samples, environment.py:27,52; dataset.py
// uses white lights too), so the radiance comes from the shared-light form without forming f' per channel.
template <typename T, int NC, int C0, bool GREY, typename IO>
SVB_DEV void render_records(const Pix<T, NC>& px, T x, float y, const float* __restrict__ rec, int N,
                            float* __restrict__ out, size_t HW, bool live) {
    SVB_UNROLL1
    for (int k = 0; k < N; ++k, rec += kRecFloats, out += 3 * HW) {
        const Geo<T> g = make_geo<T>(x, y, rec);
        Fwd<T> f;
        shade_fwd<T, NC, false, true>(g, px, f);              // accurate-highlight form: the images are the product here
        T rad[NC];                                             // renderers.py:100 (f' carries the factor pi)
        if (GREY) {
            radiance_plus_eps_grey<T, NC>(g, px, f, (g.fall * (rec[6 + C0] * kInvPi)) * f.LN0, 0.f, rad);
        } else {
            brdf_values<T, NC>(g, px, f, rad);
#pragma unroll
            for (int c = 0; c < NC; ++c) rad[c] = rad[c] * ((g.fall * (rec[6 + C0 + c] * kInvPi)) * f.LN0);
        }
        if (live) {
#pragma unroll
            for (int c = 0; c < NC; ++c) IO::st(out + (size_t)(C0 + c) * HW, rad[c]);
        }
    }
}

template <typename T, bool GREY, typename IO>
SVB_DEV void render_pixel(const T (&v)[12], T x, float y, const float* __restrict__ rec, int N, float* __restrict__ out,
                          size_t HW, bool live) {
    if (SVB_WARP_ALL(same3(v))) {
        render_records<T, 3, 0, GREY, IO>(make_pix<T, 3>(&v[0], &v[3], &v[9], v[6]), x, y, rec, N, out, HW, live);
    } else {
        render_records<T, 1, 0, GREY, IO>(make_pix<T, 1>(&v[0], &v[3], &v[9], v[6]), x, y, rec, N, out, HW, live);
        render_records<T, 1, 1, GREY, IO>(make_pix<T, 1>(&v[0], &v[4], &v[10], v[7]), x, y, rec, N, out, HW, live);
        render_records<T, 1, 2, GREY, IO>(make_pix<T, 1>(&v[0], &v[5], &v[11], v[8]), x, y, rec, N, out, HW, live);
    }
}

// The upstream gradient (N x 3 planes per batch element) is the only per-record memory traffic of the path.  It is
// streamed through a per-thread ring (IO = the kernel's RingIO: cp.async into shared memory, kDepth records in
// flight, no registers held by the loads in flight; host emulation: a plain array): `ring.fetch` enqueues the NC
// values of one record, `ring.take` waits for the oldest outstanding record and returns it.
template <typename T, int NC, int C0, typename IO>
SVB_DEV void render_bwd_records(const Pix<T, NC>& px, T x, float y, const float* __restrict__ rec, int N,
                                const float* __restrict__ gin, size_t HW, Acc<T, NC>& acc, IO& ring) {
    gin += (size_t)C0 * HW;
    ring.reset();
#pragma unroll
    for (int j = 0; j < IO::kDepth; ++j) {
        if (j < N) ring.template fetch<NC>(gin + (size_t)j * 3 * HW, HW); else ring.skip();
    }
    SVB_UNROLL1
    for (int k = 0; k < N; ++k, rec += kRecFloats) {
        if (k + IO::kDepth < N) ring.template fetch<NC>(gin + (size_t)(k + IO::kDepth) * 3 * HW, HW); else ring.skip();
        T a[NC];
        ring.template take<NC>(a);
        const Geo<T> g = make_geo<T>(x, y, rec);
        Fwd<T> f;
        shade_fwd<T, NC, true>(g, px, f);
        T fv[NC], gf[NC], gLN0;
        brdf_values<T, NC>(g, px, f, fv);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const T AE = a[c] * (g.fall * (rec[6 + C0 + c] * kInvPi));    // d loss / d radiance_c * E'_c
            gf[c] = AE * f.LN0;
            gLN0 = (c == 0) ? AE * fv[c] : vfma(AE, fv[c], gLN0);
        }
        shade_bwd<T, NC>(g, px, f, gf, gLN0, acc);
    }
}

template <typename T, int C, typename IO>
SVB_DEV void render_bwd_channel_pass(const T (&v)[12], T x, float y, const float* __restrict__ rec, int N,
                                     const float* __restrict__ gin, size_t HW, T (&g)[12], IO& ring) {
    const Pix<T, 1> px = make_pix<T, 1>(&v[0], &v[3 + C], &v[9 + C], v[6 + C]);
    Acc<T, 1> acc;
    acc_zero(acc);
    render_bwd_records<T, 1, C, IO>(px, x, y, rec, N, gin, HW, acc, ring);
    T gd[1], gs[1];
    acc_albedo_grads<T, 1>(acc, &v[3 + C], &v[9 + C], gd, gs);
#pragma unroll
    for (int j = 0; j < 3; ++j) g[j] = g[j] + acc.gn[j];
    g[3 + C] = gd[0];
    g[6 + C] = acc.ga2[0] * rough_chain(v[6 + C]);
    g[9 + C] = gs[0];
}

// grad_maps of one thread's pixels: sum over records of J^T grad_images (autograd of renderers.py:67-104)
template <typename T, typename IO>
SVB_DEV void render_bwd_pixel(const T (&v)[12], T x, float y, const float* __restrict__ rec, int N,
                              const float* __restrict__ gin, size_t HW, T (&g)[12], IO& ring) {
    if (SVB_WARP_ALL(same3(v))) {
        const Pix<T, 3> px = make_pix<T, 3>(&v[0], &v[3], &v[9], v[6]);
        Acc<T, 3> acc;
        acc_zero(acc);
        render_bwd_records<T, 3, 0, IO>(px, x, y, rec, N, gin, HW, acc, ring);
        const T chain = rough_chain(v[6]);
        T gd[3], gs[3];
        acc_albedo_grads<T, 3>(acc, &v[3], &v[9], gd, gs);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            g[c] = acc.gn[c];
            g[3 + c] = gd[c];
            g[6 + c] = acc.ga2[c] * chain;
            g[9 + c] = gs[c];
        }
    } else {
        g[0] = g[1] = g[2] = LaneTraits<T>::splat(0.f);
        render_bwd_channel_pass<T, 0, IO>(v, x, y, rec, N, gin, HW, g, ring);
        render_bwd_channel_pass<T, 1, IO>(v, x, y, rec, N, gin, HW, g, ring);
        render_bwd_channel_pass<T, 2, IO>(v, x, y, rec, N, gin, HW, g, ring);
    }
}

}  // namespace svb
