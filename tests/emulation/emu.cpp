// emu.cpp - TEST INFRASTRUCTURE.  Host build of the kernels' per-pixel source (csrc/pixel_ops.cuh +
// csrc/shading.cuh with SVB_HOST_EMULATION): runs literally the same algebra as the CUDA kernels
// (forward shading, log-L1, analytic adjoint, exact-zero masking, channel-wise general path) in a plain
// CPU loop so that the "-m 'not gpu'" suite can check it against the oracle and the golden fixtures.
// Both lane types are exercised: `float` (one pixel per thread) and `F2` (pixel pairs, the packed
// FADD2/FMUL2/FFMA2 path) - selected like the CUDA launcher does (W even -> F2).  MUFU approximations
// are exact libm calls here.  It is NOT a fallback: nothing in svbrdf_estimation_b200 links or loads it.
#define SVB_HOST_EMULATION 1
#include "../../svbrdf_estimation_b200/csrc/pixel_ops.cuh"

#include <cstddef>

using namespace svb;

namespace {

struct HostIO {
    static void ld(const float* p, float& v) { v = p[0]; }
    static void ld(const float* p, F2& v) { v = mk2(p[0], p[1]); }
    static void st(float* p, float v) { p[0] = v; }
    static void st(float* p, F2 v) { p[0] = lo(v); p[1] = hi(v); }
};

// host stand-in of the kernels' RingIO (pixel_ops.cuh render_bwd_records): same fetch / skip / take protocol
template <typename T>
struct HostRing {
    static constexpr int kDepth = 3, kSlots = 4;
    T buf[kSlots][3];
    int head, tail;
    void reset() { head = tail = 0; }
    template <int NC> void fetch(const float* p, size_t HW) {
        for (int c = 0; c < NC; ++c) HostIO::ld(p + (size_t)c * HW, buf[head % kSlots][c]);
        ++head;
    }
    void skip() { ++head; }
    template <int NC> void take(T (&a)[NC]) {
        for (int c = 0; c < NC; ++c) a[c] = buf[tail % kSlots][c];
        ++tail;
    }
};

bool all_grey(const float* recs, int nrec) {
    for (int i = 0; i < nrec; ++i) {
        const float* c = recs + (size_t)i * kRecFloats + 6;
        if (c[0] != c[1] || c[1] != c[2]) return false;
    }
    return true;
}

// One image: rendering-loss (+ map-L1 when MIXED) forward+backward.  ENC: `input` is the 9-channel network
// output, grad has 9 channels.  Returns the unscaled sums (log2 units for the rendering part).
template <typename T, bool GREY, bool MIXED, bool ENC, bool ACC = false>
void loss_image(const float* input, const float* target, int W, size_t HW, const float* rec, int N, const float* lin,
                float scale_render, float scale_l1, float* grad, double* sum_render, double* sum_l1) {
    constexpr int L = LaneTraits<T>::kLanes;
    for (size_t p = 0; p < HW; p += L) {
        T vi[12], vt[12], g[12], x, inv_len;
        if (ENC) {
            T e[9];
            for (int c = 0; c < 9; ++c) HostIO::ld(input + c * HW + p, e[c]);
            decode_encoded<T>(e, vi, inv_len);
        } else {
            for (int c = 0; c < 12; ++c) HostIO::ld(input + c * HW + p, vi[c]);
        }
        for (int c = 0; c < 12; ++c) HostIO::ld(target + c * HW + p, vt[c]);
        HostIO::ld(lin + p % W, x);
        const T l = loss_pixel<T, true, GREY, ACC>(vi, vt, x, -lin[p / W], rec, N, scale_render, g);
        *sum_render += (double)hsum(l);
        if (MIXED) *sum_l1 += (double)hsum(l1_pixel<T, true>(vi, vt, scale_l1, g));
        if (ENC) {
            T ge[9];
            encode_grad<T>(vi, inv_len, g, ge);
            for (int c = 0; c < 9; ++c) HostIO::st(grad + c * HW + p, ge[c]);
        } else {
            for (int c = 0; c < 12; ++c) HostIO::st(grad + c * HW + p, g[c]);
        }
    }
}

template <typename T, bool MIXED, bool ENC>
void loss_image_g(bool grey, const float* input, const float* target, int W, size_t HW, const float* rec, int N,
                  const float* lin, float sr, float sl, float* grad, double* a, double* b) {
    if (grey) loss_image<T, true, MIXED, ENC>(input, target, W, HW, rec, N, lin, sr, sl, grad, a, b);
    else      loss_image<T, false, MIXED, ENC>(input, target, W, HW, rec, N, lin, sr, sl, grad, a, b);
}

template <typename T, bool GREY>
void render_image(const float* maps, int W, size_t HW, const float* rec, int N, const float* lin, float* images) {
    constexpr int L = LaneTraits<T>::kLanes;
    for (size_t p = 0; p < HW; p += L) {
        T v[12], x;
        for (int c = 0; c < 12; ++c) HostIO::ld(maps + c * HW + p, v[c]);
        HostIO::ld(lin + p % W, x);
        render_pixel<T, GREY, HostIO>(v, x, -lin[p / W], rec, N, images + p, HW, true);
    }
}

template <typename T>
void render_bwd_image(const float* maps, int W, size_t HW, const float* rec, int N, const float* lin,
                      const float* gimages, float* gmaps) {
    constexpr int L = LaneTraits<T>::kLanes;
    for (size_t p = 0; p < HW; p += L) {
        T v[12], g[12], x;
        for (int c = 0; c < 12; ++c) HostIO::ld(maps + c * HW + p, v[c]);
        HostIO::ld(lin + p % W, x);
        HostRing<T> ring;
        render_bwd_pixel<T, HostRing<T>>(v, x, -lin[p / W], rec, N, gimages + p, HW, g, ring);
        for (int c = 0; c < 12; ++c) HostIO::st(gmaps + c * HW + p, g[c]);
    }
}

}  // namespace

extern "C" {

// General entry: input [B,12,H,W] (or [B,9,H,W] when encoded), target [B,12,H,W], scenes [B,N,9], lin [W].
// mixed: add l1_weight * SVBRDFL1Loss.  out[0] = total, out[1] = rendering loss, out[2] = map-L1 loss.
// lanes: 0 = choose like the CUDA launcher (2 when W is even), 1 = force scalar, 2 = force packed.
void emu_loss(const float* input, const float* target, int B, int H, int W, const float* scenes, int N,
              const float* lin, float* grad, int lanes, int mixed, float l1_weight, int encoded, double* out) {
    const size_t HW = (size_t)H * W;
    const float sr = (float)(1.0 / ((double)B * N * 3.0 * (double)HW));
    const float sl = (float)((double)l1_weight / ((double)B * 3.0 * (double)HW));
    const bool packed = lanes == 2 || (lanes == 0 && (W & 1) == 0);
    const bool grey = all_grey(scenes, B * N);
    const int cin = encoded ? 9 : 12;
    double a = 0.0, b = 0.0;
    for (int i = 0; i < B; ++i) {
        const float* rec = scenes + (size_t)i * N * 9;
        const float *pi = input + (size_t)i * cin * HW, *pt = target + (size_t)i * 12 * HW;
        float* pg = grad + (size_t)i * cin * HW;
        if (packed) {
            if (encoded)    loss_image_g<F2, true, true>(grey, pi, pt, W, HW, rec, N, lin, sr, sl, pg, &a, &b);
            else if (mixed) loss_image_g<F2, true, false>(grey, pi, pt, W, HW, rec, N, lin, sr, sl, pg, &a, &b);
            else            loss_image_g<F2, false, false>(grey, pi, pt, W, HW, rec, N, lin, sr, sl, pg, &a, &b);
        } else {
            if (encoded)    loss_image_g<float, true, true>(grey, pi, pt, W, HW, rec, N, lin, sr, sl, pg, &a, &b);
            else if (mixed) loss_image_g<float, true, false>(grey, pi, pt, W, HW, rec, N, lin, sr, sl, pg, &a, &b);
            else            loss_image_g<float, false, false>(grey, pi, pt, W, HW, rec, N, lin, sr, sl, pg, &a, &b);
        }
    }
    out[1] = a * (double)kLn2 / ((double)B * N * 3.0 * (double)HW);
    out[2] = b / ((double)B * 3.0 * (double)HW);
    out[0] = out[1] + ((mixed || encoded) ? (double)l1_weight * out[2] : 0.0);
}

// RenderingLoss forward+backward with the accurate-highlight forward (svbrdf_b200_loss_forward_backward_accurate).
double emu_loss_forward_backward_accurate(const float* input, const float* target, int B, int H, int W, const float* scenes,
                                          int N, const float* lin, float* grad, int lanes) {
    const size_t HW = (size_t)H * W;
    const bool packed = lanes == 2 || (lanes == 0 && (W & 1) == 0);
    const bool grey = all_grey(scenes, B * N);
    const float sr = (float)(1.0 / ((double)B * N * 3.0 * (double)HW));
    double a = 0.0, b = 0.0;
    for (int i = 0; i < B; ++i) {
        const float* rec = scenes + (size_t)i * N * 9;
        const float *pi = input + (size_t)i * 12 * HW, *pt = target + (size_t)i * 12 * HW;
        float* pg = grad + (size_t)i * 12 * HW;
        if (packed) { if (grey) loss_image<F2, true, false, false, true>(pi, pt, W, HW, rec, N, lin, sr, 0.f, pg, &a, &b);
                      else      loss_image<F2, false, false, false, true>(pi, pt, W, HW, rec, N, lin, sr, 0.f, pg, &a, &b); }
        else        { if (grey) loss_image<float, true, false, false, true>(pi, pt, W, HW, rec, N, lin, sr, 0.f, pg, &a, &b);
                      else      loss_image<float, false, false, false, true>(pi, pt, W, HW, rec, N, lin, sr, 0.f, pg, &a, &b); }
    }
    return a * (double)kLn2 / ((double)B * N * 3.0 * (double)HW);
}

// RenderingLoss only (the original entry point of this file).
double emu_loss_forward_backward(const float* input, const float* target, int B, int H, int W, const float* scenes,
                                 int N, const float* lin, float* grad, int lanes) {
    double out[3];
    emu_loss(input, target, B, H, W, scenes, N, lin, grad, lanes, 0, 0.f, 0, out);
    return out[1];
}

// maps [B,12,H,W]; scenes [B,N,9] (per_batch) or [N,9]; images [B,N,3,H,W].
void emu_render_forward(const float* maps, int B, int H, int W, const float* scenes, int N, int per_batch,
                        const float* lin, float* images, int lanes) {
    const size_t HW = (size_t)H * W;
    const bool packed = lanes == 2 || (lanes == 0 && (W & 1) == 0);
    for (int b = 0; b < B; ++b) {
        const float* rec = scenes + (per_batch ? (size_t)b * N * 9 : 0);
        const float* m = maps + (size_t)b * 12 * HW;
        float* im = images + (size_t)b * N * 3 * HW;
        const bool grey = all_grey(per_batch ? scenes : rec, per_batch ? B * N : N);      // chosen per launch, like the library
        if (packed) { if (grey) render_image<F2, true>(m, W, HW, rec, N, lin, im); else render_image<F2, false>(m, W, HW, rec, N, lin, im); }
        else        { if (grey) render_image<float, true>(m, W, HW, rec, N, lin, im); else render_image<float, false>(m, W, HW, rec, N, lin, im); }
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
