/* c_abi_demo.c - the rendering-loss path from plain C, no PyTorch: only include/svbrdf_b200.h and the shared library.
 *
 *   gcc -O2 -Iinclude examples/c_abi_demo.c -o build/c_abi_demo -Lsvbrdf_estimation_b200 -lsvbrdf_b200 \
 *       -Wl,-rpath,$PWD/svbrdf_estimation_b200 -lm
 *   build/c_abi_demo [out_dir]
 *
 * Builds B synthetic SVBRDF map pairs in the context's pinned buffers, samples the scenes with the library's
 * counter-based sampler, runs RenderingLoss forward+backward through the host entry point (what the reference
 * does with RenderingLoss.forward + loss.backward(), losses.py:29-52 / main.py:116-117) and prints the loss.
 * With an output directory it also writes input / target / records / gradient as raw float32 files so that
 * tests/test_gpu_cabi.py can replay the same problem through the Python layer. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "svbrdf_b200.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int st_ = (call);                                                                \
        if (st_ != 0) {                                                                  \
            fprintf(stderr, "%s failed (%d): %s\n", #call, st_, svbrdf_b200_last_error()); \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

enum { B = 3, N = 9, H = 40, W = 40 };

/* one map set [12,H,W]: unit upper-hemisphere normals, albedos in (0,1), one roughness value on three channels */
static void fill_maps(float* m, int b, float phase) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const float u = (float)x / W, v = (float)y / H;
            float n[3] = {0.35f * sinf(9.0f * u + phase + b), 0.35f * cosf(7.0f * v - phase), 1.0f};
            const float len = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            const float rough = 0.15f + 0.8f * (0.5f + 0.5f * sinf(5.0f * (u + v) + phase));
            for (int c = 0; c < 3; ++c) {
                m[(0 + c) * H * W + y * W + x] = n[c] / len;
                m[(3 + c) * H * W + y * W + x] = 0.5f + 0.45f * sinf(11.0f * u * (c + 1) + 3.0f * v + phase);
                m[(6 + c) * H * W + y * W + x] = rough;
                m[(9 + c) * H * W + y * W + x] = 0.5f + 0.45f * cosf(4.0f * u - 13.0f * v * (c + 1) + phase);
            }
        }
}

static int dump(const char* dir, const char* name, const float* p, size_t n) {
    char path[1024];
    snprintf(path, sizeof(path), "%s/%s", dir, name);
    FILE* f = fopen(path, "wb");
    if (!f) return 1;
    const size_t w = fwrite(p, sizeof(float), n, f);
    fclose(f);
    return w != n;
}

int main(int argc, char** argv) {
    const size_t per_map = (size_t)12 * H * W;
    svbrdf_b200_ctx* ctx = NULL;
    if (svbrdf_b200_abi_version() != SVBRDF_B200_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
    CHECK(svbrdf_b200_ctx_create(&ctx, B, N, H, W));
    float* in = svbrdf_b200_ctx_pinned(ctx, 0);
    float* tg = svbrdf_b200_ctx_pinned(ctx, 1);
    float* gr = svbrdf_b200_ctx_pinned(ctx, 2);
    for (int b = 0; b < B; ++b) {
        fill_maps(in + b * per_map, b, 0.3f);
        fill_maps(tg + b * per_map, b, 1.1f);
    }
    float records[B * N * 9];
    CHECK(svbrdf_b200_sample_scenes(/*seed*/ 2024, /*first batch element*/ 0, B, /*random*/ 3, /*specular*/ 6, records));
    float loss = 0.0f;
    CHECK(svbrdf_b200_rendering_loss_host(ctx, in, tg, B, records, N, &loss, gr));
    double gsum = 0.0;
    for (size_t i = 0; i < B * per_map; ++i) gsum += fabs((double)gr[i]);
    printf("loss %.9g  sum|grad| %.9g\n", loss, gsum);
    int rc = 0;
    if (argc > 1)
        rc = dump(argv[1], "input.f32", in, B * per_map) | dump(argv[1], "target.f32", tg, B * per_map) |
             dump(argv[1], "records.f32", records, (size_t)B * N * 9) | dump(argv[1], "grad.f32", gr, B * per_map);
    svbrdf_b200_ctx_destroy(ctx);
    return rc || !(loss > 0.0f) || !isfinite(gsum);
}
