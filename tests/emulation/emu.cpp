// emu.cpp - TEST INFRASTRUCTURE.  Host build of csrc/shading.cuh (SVB_HOST_EMULATION): runs the
// exact per-pixel algebra of the CUDA kernels (forward shading, log-L1, analytic adjoint) in a
// plain CPU loop so that the "-m 'not gpu'" suite can check it against the oracle and the golden
// fixtures.  It is NOT a fallback: nothing in svbrdf_estimation_b200 links or loads it.
#define SVB_HOST_EMULATION 1
#include "../../svbrdf_estimation_b200/csrc/shading.cuh"

#include <cstddef>

using namespace svb;

namespace {

inline bool same3(const float (&v)[12]) { return v[6] == v[7] && v[7] == v[8]; }

template <int RC>
double loss_pixel(const float (&vi)[12], const float (&vt)[12], float x, float y, const float* rec, int N,
                  float scale, float (&gout)[12]) {
    const Pix<RC> pi = make_pix<RC>(vi), pt = make_pix<RC>(vt);
    Acc acc;
    acc_zero(acc);
    double lsum = 0.0;
    for (int k = 0; k < N; ++k, rec += 9) {
        const Geo g = make_geo(x, y, rec);
        Fwd<RC> fi, ft;
        shade_fwd<RC, true>(g, pi, fi);
        shade_fwd<RC, true>(g, pt, ft);
        const float E[3] = {g.e0, g.e1, g.e2};
        float A[3];
        for (int c = 0; c < 3; ++c) {
            const float xi = fmaf(fi.f[c], E[c] * fi.LN0, kEpsRender);
            const float xt = fmaf(ft.f[c], E[c] * ft.LN0, kEpsRender);
            const float d = mufu_lg2(xi) - mufu_lg2(xt);
            lsum += fabsf(d);
            const float ix = mufu_rcp(xi);
            A[c] = (d > 0.f) ? ix : ((d < 0.f) ? -ix : 0.f);
        }
        shade_bwd<RC>(g, pi, fi, A, acc);
    }
    acc_to_grad<RC>(acc, pi, scale, gout);
    return lsum;
}

template <int RC>
void render_pixel(const float (&v)[12], float x, float y, const float* rec, int N, float* out, size_t HW) {
    const Pix<RC> px = make_pix<RC>(v);
    for (int k = 0; k < N; ++k, rec += 9, out += 3 * HW) {
        const Geo g = make_geo(x, y, rec);
        Fwd<RC> f;
        shade_fwd<RC, false>(g, px, f);
        out[0] = f.f[0] * (g.e0 * f.LN0);
        out[HW] = f.f[1] * (g.e1 * f.LN0);
        out[2 * HW] = f.f[2] * (g.e2 * f.LN0);
    }
}

template <int RC>
void render_bwd_pixel(const float (&v)[12], float x, float y, const float* rec, int N, const float* gin, size_t HW,
                      float (&gout)[12]) {
    const Pix<RC> px = make_pix<RC>(v);
    Acc acc;
    acc_zero(acc);
    for (int k = 0; k < N; ++k, rec += 9, gin += 3 * HW) {
        const float A[3] = {gin[0], gin[HW], gin[2 * HW]};
        const Geo g = make_geo(x, y, rec);
        Fwd<RC> f;
        shade_fwd<RC, true>(g, px, f);
        shade_bwd<RC>(g, px, f, A, acc);
    }
    acc_to_grad<RC>(acc, px, 1.f, gout);
}

}  // namespace

extern "C" {

// input/target/grad [B,12,H,W]; scenes [B,N,9]; lin [W]; returns the loss (natural log, mean).
double emu_loss_forward_backward(const float* input, const float* target, int B, int H, int W, const float* scenes,
                                 int N, const float* lin, float* grad) {
    const size_t HW = (size_t)H * W;
    const float scale = (float)(1.0 / ((double)B * N * 3.0 * (double)HW));
    double total = 0.0;
    for (int b = 0; b < B; ++b)
        for (size_t p = 0; p < HW; ++p) {
            const int row = (int)(p / W), col = (int)(p % W);
            float vi[12], vt[12], g[12];
            for (int c = 0; c < 12; ++c) {
                vi[c] = input[((size_t)b * 12 + c) * HW + p];
                vt[c] = target[((size_t)b * 12 + c) * HW + p];
            }
            const float* rec = scenes + (size_t)b * N * 9;
            total += (same3(vi) && same3(vt)) ? loss_pixel<1>(vi, vt, lin[col], -lin[row], rec, N, scale, g)
                                              : loss_pixel<3>(vi, vt, lin[col], -lin[row], rec, N, scale, g);
            for (int c = 0; c < 12; ++c) grad[((size_t)b * 12 + c) * HW + p] = g[c];
        }
    return total * (double)kLn2 / ((double)B * N * 3.0 * (double)HW);
}

// maps [B,12,H,W]; scenes [B,N,9] (per_batch) or [N,9]; images [B,N,3,H,W].
void emu_render_forward(const float* maps, int B, int H, int W, const float* scenes, int N, int per_batch,
                        const float* lin, float* images) {
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (size_t p = 0; p < HW; ++p) {
            float v[12];
            for (int c = 0; c < 12; ++c) v[c] = maps[((size_t)b * 12 + c) * HW + p];
            const float* rec = scenes + (per_batch ? (size_t)b * N * 9 : 0);
            float* out = images + (size_t)b * N * 3 * HW + p;
            const float x = lin[p % W], y = -lin[p / W];
            if (same3(v)) render_pixel<1>(v, x, y, rec, N, out, HW);
            else          render_pixel<3>(v, x, y, rec, N, out, HW);
        }
}

void emu_render_backward(const float* maps, int B, int H, int W, const float* scenes, int N, int per_batch,
                         const float* lin, const float* grad_images, float* grad_maps) {
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (size_t p = 0; p < HW; ++p) {
            float v[12], g[12];
            for (int c = 0; c < 12; ++c) v[c] = maps[((size_t)b * 12 + c) * HW + p];
            const float* rec = scenes + (per_batch ? (size_t)b * N * 9 : 0);
            const float* gin = grad_images + (size_t)b * N * 3 * HW + p;
            const float x = lin[p % W], y = -lin[p / W];
            if (same3(v)) render_bwd_pixel<1>(v, x, y, rec, N, gin, HW, g);
            else          render_bwd_pixel<3>(v, x, y, rec, N, gin, HW, g);
            for (int c = 0; c < 12; ++c) grad_maps[((size_t)b * 12 + c) * HW + p] = g[c];
        }
}

}  // extern "C"
