// emu.cpp - TEST INFRASTRUCTURE.  Host build of csrc/shading.cuh (SVB_HOST_EMULATION): runs the
// exact per-pixel algebra of the CUDA kernels (forward shading, log-L1, analytic adjoint) in a
// plain CPU loop so that the "-m 'not gpu'" suite can check it against the oracle and the golden
// fixtures.  Both lane types are exercised: `float` (one pixel per thread) and `F2` (pixel pairs,
// the packed FADD2/FMUL2/FFMA2 path) - selected like the CUDA launcher does (W even -> F2).
// It is NOT a fallback: nothing in svbrdf_estimation_b200 links or loads it.
#define SVB_HOST_EMULATION 1
#include "../../svbrdf_estimation_b200/csrc/shading.cuh"

#include <cstddef>

using namespace svb;

namespace {

// lane-typed access to a plane
inline void ld_lane(const float* p, float& v) { v = p[0]; }
inline void ld_lane(const float* p, F2& v) { v = mk2(p[0], p[1]); }
inline void st_lane(float* p, float v) { p[0] = v; }
inline void st_lane(float* p, F2 v) { p[0] = lo(v); p[1] = hi(v); }
inline float lane_get(float v, int) { return v; }
inline float lane_get(F2 v, int j) { return j ? hi(v) : lo(v); }
inline bool same3(const float (&v)[12]) { return v[6] == v[7] && v[7] == v[8]; }
inline bool same3(const F2 (&v)[12]) {
    return lo(v[6]) == lo(v[7]) && lo(v[7]) == lo(v[8]) && hi(v[6]) == hi(v[7]) && hi(v[7]) == hi(v[8]);
}

template <typename T, int NC, int C0>
T loss_records(const Pix<T, NC>& pi, const Pix<T, NC>& pt, T x, float y, const float* rec, int N, Acc<T, NC>& acc) {
    T lsum = LaneTraits<T>::splat(0.f);
    for (int k = 0; k < N; ++k, rec += 9) {
        const Geo<T> g = make_geo<T>(x, y, rec);
        Fwd<T, NC> fi, ft;
        shade_fwd<T, NC, true>(g, pi, fi);
        shade_fwd<T, NC, false>(g, pt, ft);
        T AE[NC];
        for (int c = 0; c < NC; ++c) {
            const T E = g.fall * rec[6 + C0 + c];
            const T xi = vfma(fi.f[c], E * fi.LN0, kEpsRender);
            const T xt = vfma(ft.f[c], E * ft.LN0, kEpsRender);
            const T ix = vrcp(xi);
            const T l = vlg2(xt * ix);
            lsum = lsum + vabs(l);
            AE[c] = vsigned(l, ix) * (-rec[6 + C0 + c] * g.fall);
        }
        shade_bwd<T, NC>(g, pi, fi, AE, acc);
    }
    return lsum;
}

template <typename T, int C>
T loss_channel_pass(const T (&vi)[12], const T (&vt)[12], T x, float y, const float* rec, int N, float scale, T (&g)[12]) {
    const Pix<T, 1> pi = make_pix<T, 1>(&vi[0], &vi[3 + C], &vi[9 + C], vi[6 + C]);
    const Pix<T, 1> pt = make_pix<T, 1>(&vt[0], &vt[3 + C], &vt[9 + C], vt[6 + C]);
    Acc<T, 1> acc;
    acc_zero(acc);
    const T l = loss_records<T, 1, C>(pi, pt, x, y, rec, N, acc);
    for (int j = 0; j < 3; ++j) g[j] = g[j] + acc.gn[j] * scale;
    g[3 + C] = acc.gd[0] * (scale * kInvPi);
    g[6 + C] = (acc.ga2[0] * scale) * rough_chain(vi[6 + C]);
    g[9 + C] = acc.gs[0] * scale;
    return l;
}

template <typename T>
T loss_pixel(const T (&vi)[12], const T (&vt)[12], T x, float y, const float* rec, int N, float scale, T (&g)[12]) {
    if (same3(vi) && same3(vt)) {
        const Pix<T, 3> pi = make_pix<T, 3>(&vi[0], &vi[3], &vi[9], vi[6]);
        const Pix<T, 3> pt = make_pix<T, 3>(&vt[0], &vt[3], &vt[9], vt[6]);
        Acc<T, 3> acc;
        acc_zero(acc);
        const T l = loss_records<T, 3, 0>(pi, pt, x, y, rec, N, acc);
        const T chain = rough_chain(vi[6]);
        for (int c = 0; c < 3; ++c) {
            g[c] = acc.gn[c] * scale;
            g[3 + c] = acc.gd[c] * (scale * kInvPi);
            g[6 + c] = (acc.ga2[c] * scale) * chain;
            g[9 + c] = acc.gs[c] * scale;
        }
        return l;
    }
    g[0] = g[1] = g[2] = LaneTraits<T>::splat(0.f);
    T l = loss_channel_pass<T, 0>(vi, vt, x, y, rec, N, scale, g);
    l = l + loss_channel_pass<T, 1>(vi, vt, x, y, rec, N, scale, g);
    l = l + loss_channel_pass<T, 2>(vi, vt, x, y, rec, N, scale, g);
    return l;
}

template <typename T>
double loss_image(const float* input, const float* target, int W, size_t HW, const float* rec, int N, const float* lin,
                  float scale, float* grad) {
    constexpr int L = LaneTraits<T>::kLanes;
    double total = 0.0;
    for (size_t p = 0; p < HW; p += L) {
        T vi[12], vt[12], g[12], x;
        for (int c = 0; c < 12; ++c) { ld_lane(input + c * HW + p, vi[c]); ld_lane(target + c * HW + p, vt[c]); }
        ld_lane(lin + p % W, x);
        const T l = loss_pixel<T>(vi, vt, x, -lin[p / W], rec, N, scale, g);
        for (int c = 0; c < 12; ++c) st_lane(grad + c * HW + p, g[c]);
        // bitwise-identical input/target pixels contribute exactly 0 (see loss_kernel)
        bool differs[2] = {false, false};
        for (int c = 0; c < 12; ++c)
            for (int j = 0; j < L; ++j) differs[j] = differs[j] || (input[c * HW + p + j] != target[c * HW + p + j]);
        for (int j = 0; j < L; ++j)
            if (!differs[j]) for (int c = 0; c < 12; ++c) grad[c * HW + p + j] = 0.f;
        for (int j = 0; j < L; ++j) if (differs[j]) total += (double)lane_get(l, j);
    }
    return total;
}

template <typename T, int NC, int C0>
void render_records(const Pix<T, NC>& px, T x, float y, const float* rec, int N, float* out, size_t HW) {
    for (int k = 0; k < N; ++k, rec += 9, out += 3 * HW) {
        const Geo<T> g = make_geo<T>(x, y, rec);
        Fwd<T, NC> f;
        shade_fwd<T, NC, false>(g, px, f);
        for (int c = 0; c < NC; ++c) st_lane(out + (C0 + c) * HW, f.f[c] * ((g.fall * rec[6 + C0 + c]) * f.LN0));
    }
}

template <typename T>
void render_image(const float* maps, int W, size_t HW, const float* rec, int N, const float* lin, float* images) {
    constexpr int L = LaneTraits<T>::kLanes;
    for (size_t p = 0; p < HW; p += L) {
        T v[12], x;
        for (int c = 0; c < 12; ++c) ld_lane(maps + c * HW + p, v[c]);
        ld_lane(lin + p % W, x);
        const float y = -lin[p / W];
        float* out = images + p;
        if (same3(v)) {
            render_records<T, 3, 0>(make_pix<T, 3>(&v[0], &v[3], &v[9], v[6]), x, y, rec, N, out, HW);
        } else {
            render_records<T, 1, 0>(make_pix<T, 1>(&v[0], &v[3], &v[9], v[6]), x, y, rec, N, out, HW);
            render_records<T, 1, 1>(make_pix<T, 1>(&v[0], &v[4], &v[10], v[7]), x, y, rec, N, out, HW);
            render_records<T, 1, 2>(make_pix<T, 1>(&v[0], &v[5], &v[11], v[8]), x, y, rec, N, out, HW);
        }
    }
}

template <typename T, int NC, int C0>
void render_bwd_records(const Pix<T, NC>& px, T x, float y, const float* rec, int N, const float* gin, size_t HW,
                        Acc<T, NC>& acc) {
    for (int k = 0; k < N; ++k, rec += 9, gin += 3 * HW) {
        const Geo<T> g = make_geo<T>(x, y, rec);
        T AE[NC];
        for (int c = 0; c < NC; ++c) {
            T a;
            ld_lane(gin + (C0 + c) * HW, a);
            AE[c] = a * (g.fall * rec[6 + C0 + c]);
        }
        Fwd<T, NC> f;
        shade_fwd<T, NC, true>(g, px, f);
        shade_bwd<T, NC>(g, px, f, AE, acc);
    }
}

template <typename T, int C>
void render_bwd_channel_pass(const T (&v)[12], T x, float y, const float* rec, int N, const float* gin, size_t HW, T (&g)[12]) {
    const Pix<T, 1> px = make_pix<T, 1>(&v[0], &v[3 + C], &v[9 + C], v[6 + C]);
    Acc<T, 1> acc;
    acc_zero(acc);
    render_bwd_records<T, 1, C>(px, x, y, rec, N, gin, HW, acc);
    for (int j = 0; j < 3; ++j) g[j] = g[j] + acc.gn[j];
    g[3 + C] = acc.gd[0] * kInvPi;
    g[6 + C] = acc.ga2[0] * rough_chain(v[6 + C]);
    g[9 + C] = acc.gs[0];
}

template <typename T>
void render_bwd_image(const float* maps, int W, size_t HW, const float* rec, int N, const float* lin,
                      const float* gimages, float* gmaps) {
    constexpr int L = LaneTraits<T>::kLanes;
    for (size_t p = 0; p < HW; p += L) {
        T v[12], g[12], x;
        for (int c = 0; c < 12; ++c) ld_lane(maps + c * HW + p, v[c]);
        ld_lane(lin + p % W, x);
        const float y = -lin[p / W];
        const float* gin = gimages + p;
        if (same3(v)) {
            const Pix<T, 3> px = make_pix<T, 3>(&v[0], &v[3], &v[9], v[6]);
            Acc<T, 3> acc;
            acc_zero(acc);
            render_bwd_records<T, 3, 0>(px, x, y, rec, N, gin, HW, acc);
            const T chain = rough_chain(v[6]);
            for (int c = 0; c < 3; ++c) {
                g[c] = acc.gn[c];
                g[3 + c] = acc.gd[c] * kInvPi;
                g[6 + c] = acc.ga2[c] * chain;
                g[9 + c] = acc.gs[c];
            }
        } else {
            g[0] = g[1] = g[2] = LaneTraits<T>::splat(0.f);
            render_bwd_channel_pass<T, 0>(v, x, y, rec, N, gin, HW, g);
            render_bwd_channel_pass<T, 1>(v, x, y, rec, N, gin, HW, g);
            render_bwd_channel_pass<T, 2>(v, x, y, rec, N, gin, HW, g);
        }
        for (int c = 0; c < 12; ++c) st_lane(gmaps + c * HW + p, g[c]);
    }
}

}  // namespace

extern "C" {

// input/target/grad [B,12,H,W]; scenes [B,N,9]; lin [W]; returns the loss (natural log, mean).
// lanes: 0 = choose like the CUDA launcher (2 when W is even), 1 = force scalar, 2 = force packed.
double emu_loss_forward_backward(const float* input, const float* target, int B, int H, int W, const float* scenes,
                                 int N, const float* lin, float* grad, int lanes) {
    const size_t HW = (size_t)H * W;
    const float scale = (float)(1.0 / ((double)B * N * 3.0 * (double)HW));
    const bool packed = lanes == 2 || (lanes == 0 && (W & 1) == 0);
    double total = 0.0;
    for (int b = 0; b < B; ++b) {
        const size_t off = (size_t)b * 12 * HW;
        const float* rec = scenes + (size_t)b * N * 9;
        total += packed ? loss_image<F2>(input + off, target + off, W, HW, rec, N, lin, scale, grad + off)
                        : loss_image<float>(input + off, target + off, W, HW, rec, N, lin, scale, grad + off);
    }
    return total * (double)kLn2 / ((double)B * N * 3.0 * (double)HW);
}

// maps [B,12,H,W]; scenes [B,N,9] (per_batch) or [N,9]; images [B,N,3,H,W].
void emu_render_forward(const float* maps, int B, int H, int W, const float* scenes, int N, int per_batch,
                        const float* lin, float* images, int lanes) {
    const size_t HW = (size_t)H * W;
    const bool packed = lanes == 2 || (lanes == 0 && (W & 1) == 0);
    for (int b = 0; b < B; ++b) {
        const float* rec = scenes + (per_batch ? (size_t)b * N * 9 : 0);
        if (packed) render_image<F2>(maps + (size_t)b * 12 * HW, W, HW, rec, N, lin, images + (size_t)b * N * 3 * HW);
        else        render_image<float>(maps + (size_t)b * 12 * HW, W, HW, rec, N, lin, images + (size_t)b * N * 3 * HW);
    }
}

void emu_render_backward(const float* maps, int B, int H, int W, const float* scenes, int N, int per_batch,
                         const float* lin, const float* grad_images, float* grad_maps, int lanes) {
    const size_t HW = (size_t)H * W;
    const bool packed = lanes == 2 || (lanes == 0 && (W & 1) == 0);
    for (int b = 0; b < B; ++b) {
        const float* rec = scenes + (per_batch ? (size_t)b * N * 9 : 0);
        const float* gin = grad_images + (size_t)b * N * 3 * HW;
        if (packed) render_bwd_image<F2>(maps + (size_t)b * 12 * HW, W, HW, rec, N, lin, gin, grad_maps + (size_t)b * 12 * HW);
        else        render_bwd_image<float>(maps + (size_t)b * 12 * HW, W, HW, rec, N, lin, gin, grad_maps + (size_t)b * 12 * HW);
    }
}

}  // extern "C"
