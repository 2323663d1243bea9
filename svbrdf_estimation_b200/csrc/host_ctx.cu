// host_ctx.cu - host-buffer entry point of the rendering-loss path.
//
// svbrdf_b200_rendering_loss_host() is the call a caller without device memory makes: maps live
// in host memory, the loss and d loss / d input come back to host memory.  It is what bench.py
// times as "e2e".  The batch is cut into slices; slice i's uploads (two H2D streams: input, target), its kernel
// (compute stream) and the download of its gradient (D2H stream) overlap with the neighbouring
// slices', so the PCIe link - not the kernel - is the bound and both directions are busy at once.
//
// Reference call being replaced: RenderingLoss.forward + loss.backward()
// (development/multiImage_pytorch/losses.py:29-52, main.py:116-117) on CPU-resident tensors.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/svbrdf_b200.h"
#include "internal.h"

struct svbrdf_b200_ctx {
    int max_B, max_N, H, W, device;
    size_t map_floats;                       // 12*H*W
    float *d_in, *d_tg, *d_gr, *d_lin, *d_loss, *d_ws;
    float *h_in, *h_tg, *h_gr, *h_loss;      // pinned
    size_t ws_bytes;
    cudaStream_t s_h2d, s_h2d2, s_comp, s_d2h;
    cudaEvent_t* ev_up;                      // per slice: input upload done
    cudaEvent_t* ev_up2;                     // per slice: target upload done
    cudaEvent_t* ev_k;                       // per slice: kernel done
    int max_slices;
};

// torch.linspace(-1, 1, W) in fp32: step = (end-start)/(W-1); the first half counts up from the
// start, the second half counts down from the end (ATen RangeFactories), each as ONE fused
// multiply-add - which is what both the CPU and the CUDA kernels of torch compute (checked against
// torch.linspace for W = 2..4096 in tests/test_host_logic.py).  renderers.py:73.
static void fill_lin(float* lin, int W) {
    if (W == 1) { lin[0] = -1.f; return; }
    const float start = -1.f, end = 1.f;
    const float step = (end - start) / (float)(W - 1);
    const int half = W / 2;
    for (int i = 0; i < W; ++i)
        lin[i] = (i < half) ? fmaf(step, (float)i, start) : fmaf(-step, (float)(W - 1 - i), end);
}

extern "C" int svbrdf_b200_coordinate_table(float* lin_host, int W) {
    if (!lin_host || W <= 0) return svb_fail(SVBRDF_E_INVALID, "bad coordinate table arguments");
    fill_lin(lin_host, W);
    return 0;
}

// ---- scene sampler (host side) -------------------------------------------------------
// Counter-based: every draw is a pure function of (seed, batch element, record, draw index), so the
// call is stateless, thread-safe and independent of B (element b gets the same scenes whatever the
// batch size or the rank that samples it).  Distributions follow environment.py:18-55 /
// utils.py:100-111; the stream of numbers is NOT the torch generator's (use the Python samplers for
// reference-order draws).
static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline float u01(uint64_t seed, uint64_t b, uint64_t k, uint64_t j) {   // [0,1), 24 random bits
    const uint64_t h = mix64(mix64(mix64(seed) ^ (b * 0xD1342543DE82EF95ull)) ^ (k << 8 | j));
    return (float)(h >> 40) * (1.0f / 16777216.0f);
}
static inline void hemisphere_dir(float r1, float r2, float* d) {               // utils.py:104-111
    const float r = sqrtf(r1), phi = 6.283185307179586f * r2;
    d[0] = r * cosf(phi); d[1] = r * sinf(phi); d[2] = sqrtf(1.0f - r * r);
}
static inline float normal01(float u1, float u2) {                              // Box-Muller
    return sqrtf(-2.0f * logf(1.0f - u1)) * cosf(6.283185307179586f * u2);
}

extern "C" int svbrdf_b200_sample_scenes(uint64_t seed, int first_batch_element, int B, int n_random, int n_specular,
                                         float* records_host) {
    if (!records_host || B <= 0 || n_random < 0 || n_specular < 0 || n_random + n_specular <= 0 || first_batch_element < 0)
        return svb_fail(SVBRDF_E_INVALID, "bad sampler arguments");
    const int N = n_random + n_specular;
    const float lo = 0.001f, hi = 1.0f - 0.1f;                                   // environment.py:20-21,34
    for (int b = 0; b < B; ++b) {
        const uint64_t e = (uint64_t)(first_batch_element + b);
        for (int k = 0; k < N; ++k) {
            float* r = records_host + ((size_t)b * N + k) * 9;
            float view[3];
            hemisphere_dir(lo + (hi - lo) * u01(seed, e, k, 0), u01(seed, e, k, 1), view);
            if (k < n_random) {                                                  // environment.py:18-30
                hemisphere_dir(lo + (hi - lo) * u01(seed, e, k, 2), u01(seed, e, k, 3), r + 3);
                r[0] = view[0]; r[1] = view[1]; r[2] = view[2];
                r[6] = r[7] = r[8] = 20.0f;
            } else {                                                             // environment.py:32-55
                const float dv = expf(0.5f + 0.75f * normal01(u01(seed, e, k, 2), u01(seed, e, k, 3)));
                const float dl = expf(0.5f + 0.75f * normal01(u01(seed, e, k, 4), u01(seed, e, k, 5)));
                const float sx = 2.0f * u01(seed, e, k, 6) - 1.0f, sy = 2.0f * u01(seed, e, k, 7) - 1.0f, sz = 0.0001f;
                r[0] = view[0] * dv + sx;  r[1] = view[1] * dv + sy;  r[2] = view[2] * dv + sz;
                r[3] = -view[0] * dl + sx; r[4] = -view[1] * dl + sy; r[5] = view[2] * dl + sz;
                r[6] = r[7] = r[8] = 50.0f;
            }
        }
    }
    return 0;
}

#define CK(call)                                              \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) { rc = svb_cuda_status(e_, #call); goto done; } \
    } while (0)

extern "C" void svbrdf_b200_ctx_destroy(svbrdf_b200_ctx* c) {
    if (!c) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    if (c->ev_up) for (int i = 0; i < c->max_slices; ++i) if (c->ev_up[i]) cudaEventDestroy(c->ev_up[i]);
    if (c->ev_k) for (int i = 0; i < c->max_slices; ++i) if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
    if (c->ev_up2) for (int i = 0; i < c->max_slices; ++i) if (c->ev_up2[i]) cudaEventDestroy(c->ev_up2[i]);
    free(c->ev_up); free(c->ev_k); free(c->ev_up2);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_h2d2) cudaStreamDestroy(c->s_h2d2);
    if (c->s_comp) cudaStreamDestroy(c->s_comp);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    cudaFree(c->d_in); cudaFree(c->d_tg); cudaFree(c->d_gr); cudaFree(c->d_lin); cudaFree(c->d_loss); cudaFree(c->d_ws);
    cudaFreeHost(c->h_in); cudaFreeHost(c->h_tg); cudaFreeHost(c->h_gr); cudaFreeHost(c->h_loss);
    free(c);
    if (prev >= 0) cudaSetDevice(prev);
}

extern "C" int svbrdf_b200_ctx_create(svbrdf_b200_ctx** out, int max_B, int max_N, int H, int W) {
    if (!out) return svb_fail(SVBRDF_E_INVALID, "out is null");
    *out = nullptr;
    if (int e = svb_check_shape(max_B, H, W, max_N)) return e;
    int rc = 0;
    svbrdf_b200_ctx* c = (svbrdf_b200_ctx*)calloc(1, sizeof(svbrdf_b200_ctx));
    if (!c) return svb_fail(SVBRDF_E_INVALID, "out of host memory");
    float* lin = nullptr;
    c->max_B = max_B; c->max_N = max_N; c->H = H; c->W = W;
    c->map_floats = (size_t)12 * H * W;
    c->max_slices = max_B < 16 ? max_B : 16;
    {
        const size_t bytes = (size_t)max_B * c->map_floats * sizeof(float);
        CK(cudaGetDevice(&c->device));
        CK(cudaMalloc(&c->d_in, bytes));
        CK(cudaMalloc(&c->d_tg, bytes));
        CK(cudaMalloc(&c->d_gr, bytes));
        CK(cudaMalloc(&c->d_lin, (size_t)W * sizeof(float)));
        CK(cudaMalloc(&c->d_loss, 4 * sizeof(float)));
        c->ws_bytes = svbrdf_b200_workspace_bytes(max_B, max_N, H, W);
        CK(cudaMalloc(&c->d_ws, c->ws_bytes));
        CK(cudaMallocHost(&c->h_in, bytes));
        CK(cudaMallocHost(&c->h_tg, bytes));
        CK(cudaMallocHost(&c->h_gr, bytes));
        CK(cudaMallocHost(&c->h_loss, 4 * sizeof(float)));
        CK(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->s_h2d2, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
        c->ev_up = (cudaEvent_t*)calloc(c->max_slices, sizeof(cudaEvent_t));
        c->ev_k = (cudaEvent_t*)calloc(c->max_slices, sizeof(cudaEvent_t));
        c->ev_up2 = (cudaEvent_t*)calloc(c->max_slices, sizeof(cudaEvent_t));
        if (!c->ev_up || !c->ev_k || !c->ev_up2) { rc = svb_fail(SVBRDF_E_INVALID, "out of host memory"); goto done; }
        for (int i = 0; i < c->max_slices; ++i) {
            CK(cudaEventCreateWithFlags(&c->ev_up[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_up2[i], cudaEventDisableTiming));
        }
        lin = (float*)malloc((size_t)W * sizeof(float));
        if (!lin) { rc = svb_fail(SVBRDF_E_INVALID, "out of host memory"); goto done; }
        fill_lin(lin, W);
        CK(cudaMemcpy(c->d_lin, lin, (size_t)W * sizeof(float), cudaMemcpyHostToDevice));
    }
done:
    free(lin);
    if (rc) { svbrdf_b200_ctx_destroy(c); return rc; }
    *out = c;
    return 0;
}

extern "C" float* svbrdf_b200_ctx_pinned(svbrdf_b200_ctx* c, int which) {
    if (!c) return nullptr;
    return which == 0 ? c->h_in : which == 1 ? c->h_tg : which == 2 ? c->h_gr : nullptr;
}

static int loss_host_impl(svbrdf_b200_ctx* c, const float* input_host, int input_layout, const float* target_host,
                          int target_layout, int B, const float* scenes_host, int N, float l1_weight, float* out_host,
                          int n_out, float* grad_host) {
    if (!c) return svb_fail(SVBRDF_E_STATE, "context is null");
    if (!input_host || !target_host || !scenes_host || !out_host)
        return svb_fail(SVBRDF_E_INVALID, "null pointer argument");
    if (B <= 0 || B > c->max_B || N <= 0 || N > c->max_N)
        return svb_fail(SVBRDF_E_INVALID, "B or N exceeds what the context was created for");
    const int lay = svb_layout_id(input_layout, target_layout);
    const bool mixed = l1_weight >= 0.f;
    if (int e = svb_check_layout_form(lay, mixed, grad_host != nullptr)) return e;
    int rc = 0, prev = -1;
    cudaGetDevice(&prev);
    const int HW = c->H * c->W;
    const bool packed = (c->W & 1) == 0;          // cudaMalloc'd buffers: aligned
    const int cpi = svb_ctas_per_image(HW, packed);
    const size_t in_floats = (size_t)input_layout * HW, tg_floats = (size_t)target_layout * HW;   // per batch element
    float* part_render = c->d_ws;
    float* part_l1 = c->d_ws + (size_t)B * cpi;
    // slices: enough to overlap the two copy directions with compute, not so many that launch
    // overhead shows; every slice is at least one batch element.
    const int slices = B < c->max_slices ? B : c->max_slices;
    {
        CK(cudaSetDevice(c->device));
        for (int i = 0; i < slices; ++i) {
            const int b0 = (int)((long long)B * i / slices), b1 = (int)((long long)B * (i + 1) / slices);
            CK(cudaMemcpyAsync(c->d_in + b0 * in_floats, input_host + b0 * in_floats, (b1 - b0) * in_floats * sizeof(float),
                               cudaMemcpyHostToDevice, c->s_h2d));
            CK(cudaMemcpyAsync(c->d_tg + b0 * tg_floats, target_host + b0 * tg_floats, (b1 - b0) * tg_floats * sizeof(float),
                               cudaMemcpyHostToDevice, c->s_h2d2));
            CK(cudaEventRecord(c->ev_up[i], c->s_h2d));
            CK(cudaEventRecord(c->ev_up2[i], c->s_h2d2));
            CK(cudaStreamWaitEvent(c->s_comp, c->ev_up[i], 0));
            CK(cudaStreamWaitEvent(c->s_comp, c->ev_up2[i], 0));
            rc = svb_launch_loss_range(c->d_in, c->d_tg, grad_host ? c->d_gr : nullptr, B, HW, c->W, scenes_host, N,
                                       c->d_lin, part_render, part_l1, mixed, mixed ? l1_weight : 0.f, b0, b1 - b0, c->s_comp,
                                       packed, lay);
            if (rc) goto done;
            if (grad_host) {
                CK(cudaEventRecord(c->ev_k[i], c->s_comp));
                CK(cudaStreamWaitEvent(c->s_d2h, c->ev_k[i], 0));
                CK(cudaMemcpyAsync(grad_host + b0 * in_floats, c->d_gr + b0 * in_floats, (b1 - b0) * in_floats * sizeof(float),
                                   cudaMemcpyDeviceToHost, c->s_d2h));
            }
        }
        rc = svb_launch_finalize(part_render, part_l1, B, HW, packed, N, mixed, mixed ? l1_weight : 0.f, c->d_loss, 3, c->s_comp);
        if (rc) goto done;
        CK(cudaMemcpyAsync(c->h_loss, c->d_loss, 3 * sizeof(float), cudaMemcpyDeviceToHost, c->s_comp));
        CK(cudaStreamSynchronize(c->s_comp));
        CK(cudaStreamSynchronize(c->s_d2h));
        for (int i = 0; i < n_out; ++i) out_host[i] = c->h_loss[i];
    }
done:
    if (rc) { cudaStreamSynchronize(c->s_h2d); cudaStreamSynchronize(c->s_h2d2); cudaStreamSynchronize(c->s_comp); cudaStreamSynchronize(c->s_d2h); }
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

extern "C" int svbrdf_b200_rendering_loss_host(svbrdf_b200_ctx* c, const float* input_host, const float* target_host,
                                               int B, const float* scenes_host, int N, float* loss_host,
                                               float* grad_host) {
    if (!loss_host) return svb_fail(SVBRDF_E_INVALID, "null pointer argument");
    return loss_host_impl(c, input_host, SVBRDF_LAYOUT_MAPS12, target_host, SVBRDF_LAYOUT_MAPS12, B, scenes_host, N, -1.f,
                          loss_host, 1, grad_host);
}

extern "C" int svbrdf_b200_loss_host(svbrdf_b200_ctx* c, const float* input_host, int input_layout, const float* target_host,
                                     int target_layout, int B, const float* scenes_host, int N, float l1_weight,
                                     float* out_host, float* grad_host) {
    return loss_host_impl(c, input_host, input_layout, target_host, target_layout, B, scenes_host, N, l1_weight, out_host, 3,
                          grad_host);
}
