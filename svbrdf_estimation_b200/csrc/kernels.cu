// kernels.cu - sm_100a kernels of the rendering-loss path and their C-ABI launchers.
//
// Work decomposition (DESIGN.md "Kernels"):
//   grid = (ceil(H*W / 256), batch elements of this launch), 256 threads, ONE pixel per thread.
//   A thread loads its pixel's 12 (+12 target) channels with coalesced 32-bit loads (a warp reads
//   128 contiguous bytes of each of the 24 planes), keeps them and the 12 gradient accumulators in
//   registers, loops over the N scene records of its batch element - which live in the kernel
//   parameter block (constant bank, warp-uniform addresses, no scene upload) - and never writes a
//   per-record intermediate to memory.  The log-L1 terms are reduced thread -> warp (shuffle) ->
//   CTA (shared memory) -> one partial per CTA; a 1-CTA finalize kernel adds the partials in a
//   fixed order in fp64, so the loss is run-to-run deterministic and uses no float atomics.
//
// Reference semantics: LocalRenderer.render renderers.py:67-104, RenderingLoss.forward
// losses.py:29-52, SVBRDFL1Loss/MixedLoss losses.py:7-19,54-63 (paths relative to
// development/multiImage_pytorch/ of mworchel/svbrdf-estimation).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/svbrdf_b200.h"
#include "shading.cuh"
#include "internal.h"

namespace svb {

constexpr int kThreads = 256;
constexpr int kRecFloats = 9;
// Kernel parameters may total 32,764 bytes on sm_70+ with CUDA >= 12.1.  Two capacities keep the
// parameter copy small for the common render(scene, maps) call.
constexpr int kCapSmall = 64;    // records ->  2,304 B
constexpr int kCapLarge = 796;   // records -> 28,656 B

template <int CAP>
struct SceneBlock {
    float v[CAP * kRecFloats];
};

struct LossArgs {
    const float* input;    // [Bc,12,H,W] (already offset to this launch's first batch element)
    const float* target;
    float* grad;           // may be null (forward only)
    const float* lin;      // [W]
    float* part_render;    // per-CTA partial sums of |dlog| (offset to this launch)
    float* part_l1;        // per-CTA partial sums of the map-L1 terms (MIXED only)
    int HW, W, N;
    float scale_render;    // 1 / (B N 3 H W)
    float scale_l1;        // l1_weight / (B 3 H W)
};

struct RenderArgs {
    const float* maps;     // [Bc,12,H,W]
    const float* lin;
    float* images;         // [Bc,N,3,H,W]            (forward)
    const float* gimages;  // [Bc,N,3,H,W]            (backward)
    float* gmaps;          // [Bc,12,H,W]             (backward)
    int HW, W, N, per_batch;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA; result valid in thread 0.  Fixed order => deterministic.
__device__ __forceinline__ float cta_sum(float v, float* smem /* [kThreads/32] */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += smem[w];
    }
    __syncthreads();
    return t;
}

__device__ __forceinline__ void load12(const float* __restrict__ base, int HW, float (&v)[12]) {
#pragma unroll
    for (int c = 0; c < 12; ++c) v[c] = __ldg(base + (size_t)c * HW);
}

__device__ __forceinline__ bool same3(const float (&v)[12]) { return v[6] == v[7] && v[7] == v[8]; }

// ---------------------------------------------------------------------------------------------
// RenderingLoss / MixedLoss: forward (+ backward) for one pixel over all N records
// ---------------------------------------------------------------------------------------------
template <int RC, bool BWD>
__device__ __forceinline__ float loss_pixel(const float (&vi)[12], const float (&vt)[12], float x, float y,
                                            const float* __restrict__ rec, int N, float scale, float (&gout)[12]) {
    const Pix<RC> pi = make_pix<RC>(vi);
    const Pix<RC> pt = make_pix<RC>(vt);
    Acc acc;
    acc_zero(acc);
    float lsum = 0.f;
#pragma unroll 1
    for (int k = 0; k < N; ++k, rec += kRecFloats) {
        const Geo g = make_geo(x, y, rec);
        Fwd<RC> fi, ft;
        shade_fwd<RC, BWD>(g, pi, fi);
        shade_fwd<RC, BWD>(g, pt, ft);   // same arithmetic as the input map: identical maps give exactly 0
        const float E[3] = {g.e0, g.e1, g.e2};
        float A[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float xi = fmaf(fi.f[c], E[c] * fi.LN0, kEpsRender);   // radiance + 0.1 (losses.py:46-47)
            const float xt = fmaf(ft.f[c], E[c] * ft.LN0, kEpsRender);
            const float d = mufu_lg2(xi) - mufu_lg2(xt);                 // log2 units; ln2 applied at the end
            lsum += fabsf(d);
            if (BWD) {
                const float ix = mufu_rcp(xi);
                A[c] = (d > 0.f) ? ix : ((d < 0.f) ? -ix : 0.f);         // sign(0) = 0 like torch (losses.py:50)
            }
        }
        if (BWD) shade_bwd<RC>(g, pi, fi, A, acc);
    }
    if (BWD) acc_to_grad<RC>(acc, pi, scale, gout);
    return lsum;
}

// Map-space L1 terms of SVBRDFL1Loss (losses.py:7-19) for one pixel; adds their gradient.
template <bool BWD>
__device__ __forceinline__ float l1_pixel(const float (&vi)[12], const float (&vt)[12], float scale, float (&gout)[12]) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 12; ++c) {
        const bool logged = (c >= 3 && c < 6) || c >= 9;      // diffuse and specular use log(x + 0.01)
        float d, w = 1.f;
        if (logged) {
            const float a = vi[c] + kEpsL1, b = vt[c] + kEpsL1;
            d = (mufu_lg2(a) - mufu_lg2(b)) * kLn2;
            if (BWD) w = mufu_rcp(a);
        } else {
            d = vi[c] - vt[c];
        }
        s += fabsf(d);
        if (BWD) gout[c] += (d > 0.f) ? w * scale : ((d < 0.f) ? -w * scale : 0.f);
    }
    return s;
}

template <bool BWD, bool MIXED, int CAP>
__global__ void __launch_bounds__(kThreads)
loss_kernel(const LossArgs a, const __grid_constant__ SceneBlock<CAP> sc) {
    __shared__ float red[kThreads / 32];
    const int b = blockIdx.y;
    const int pix = blockIdx.x * kThreads + threadIdx.x;
    const bool live = pix < a.HW;
    const int p = live ? pix : a.HW - 1;
    const int row = p / a.W, col = p - row * a.W;
    const float x = __ldg(a.lin + col), y = -__ldg(a.lin + row);   // renderers.py:73-76
    const size_t off = (size_t)b * 12 * a.HW + p;
    float vi[12], vt[12], g[12];
    load12(a.input + off, a.HW, vi);
    load12(a.target + off, a.HW, vt);
    const float* rec = sc.v + (size_t)b * a.N * kRecFloats;

    float lsum;
    const bool shared_rough = same3(vi) && same3(vt);
    if (__all_sync(0xffffffffu, shared_rough)) lsum = loss_pixel<1, BWD>(vi, vt, x, y, rec, a.N, a.scale_render, g);
    else                                        lsum = loss_pixel<3, BWD>(vi, vt, x, y, rec, a.N, a.scale_render, g);

    float l1 = 0.f;
    if (MIXED) l1 = l1_pixel<BWD>(vi, vt, a.scale_l1, g);

    if (BWD && live) {
        float* gp = a.grad + off;
#pragma unroll
        for (int c = 0; c < 12; ++c) __stcs(gp + (size_t)c * a.HW, g[c]);
    }
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    const float tr = cta_sum(live ? lsum : 0.f, red);
    if (threadIdx.x == 0) a.part_render[cta] = tr;
    if (MIXED) {
        const float tl = cta_sum(live ? l1 : 0.f, red);
        if (threadIdx.x == 0) a.part_l1[cta] = tl;
    }
}

// out[0] = mixed (or rendering) loss, out[1] = rendering loss, out[2] = map-L1 loss (MIXED only).
__global__ void __launch_bounds__(1024)
finalize_kernel(const float* __restrict__ part_render, const float* __restrict__ part_l1, int count,
                double mul_render, double mul_l1, float l1_weight, float* __restrict__ out, int n_out) {
    __shared__ double sm[2][32];
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        s0 += (double)part_render[i];
        if (part_l1) s1 += (double)part_l1[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = s0; sm[1][threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t0 += sm[0][w]; t1 += sm[1][w]; }
        const double render = t0 * mul_render, l1 = t1 * mul_l1;
        out[0] = (float)(render + (double)l1_weight * l1);
        if (n_out > 1) out[1] = (float)render;
        if (n_out > 2) out[2] = (float)l1;
    }
}

// ---------------------------------------------------------------------------------------------
// LocalRenderer.render forward / backward
// ---------------------------------------------------------------------------------------------
template <int RC>
__device__ __forceinline__ void render_pixel(const float (&v)[12], float x, float y, const float* __restrict__ rec,
                                             int N, float* __restrict__ out, int HW, bool live) {
    const Pix<RC> px = make_pix<RC>(v);
#pragma unroll 1
    for (int k = 0; k < N; ++k, rec += kRecFloats, out += (size_t)3 * HW) {
        const Geo g = make_geo(x, y, rec);
        Fwd<RC> f;
        shade_fwd<RC, false>(g, px, f);
        if (live) {
            __stcs(out, f.f[0] * (g.e0 * f.LN0));                // renderers.py:100
            __stcs(out + HW, f.f[1] * (g.e1 * f.LN0));
            __stcs(out + 2 * (size_t)HW, f.f[2] * (g.e2 * f.LN0));
        }
    }
}

template <int CAP>
__global__ void __launch_bounds__(kThreads)
render_fwd_kernel(const RenderArgs a, const __grid_constant__ SceneBlock<CAP> sc) {
    const int b = blockIdx.y;
    const int pix = blockIdx.x * kThreads + threadIdx.x;
    const bool live = pix < a.HW;
    const int p = live ? pix : a.HW - 1;
    const int row = p / a.W, col = p - row * a.W;
    const float x = __ldg(a.lin + col), y = -__ldg(a.lin + row);
    float v[12];
    load12(a.maps + (size_t)b * 12 * a.HW + p, a.HW, v);
    const float* rec = sc.v + (a.per_batch ? (size_t)b * a.N * kRecFloats : 0);
    float* out = a.images + (size_t)b * a.N * 3 * a.HW + p;
    if (__all_sync(0xffffffffu, same3(v))) render_pixel<1>(v, x, y, rec, a.N, out, a.HW, live);
    else                                   render_pixel<3>(v, x, y, rec, a.N, out, a.HW, live);
}

template <int RC>
__device__ __forceinline__ void render_bwd_pixel(const float (&v)[12], float x, float y, const float* __restrict__ rec,
                                                 int N, const float* __restrict__ gin, int HW, float (&gout)[12]) {
    const Pix<RC> px = make_pix<RC>(v);
    Acc acc;
    acc_zero(acc);
#pragma unroll 1
    for (int k = 0; k < N; ++k, rec += kRecFloats, gin += (size_t)3 * HW) {
        const float A[3] = {__ldcs(gin), __ldcs(gin + HW), __ldcs(gin + 2 * (size_t)HW)};
        const Geo g = make_geo(x, y, rec);
        Fwd<RC> f;
        shade_fwd<RC, true>(g, px, f);
        shade_bwd<RC>(g, px, f, A, acc);
    }
    acc_to_grad<RC>(acc, px, 1.f, gout);
}

template <int CAP>
__global__ void __launch_bounds__(kThreads)
render_bwd_kernel(const RenderArgs a, const __grid_constant__ SceneBlock<CAP> sc) {
    const int b = blockIdx.y;
    const int pix = blockIdx.x * kThreads + threadIdx.x;
    const bool live = pix < a.HW;
    const int p = live ? pix : a.HW - 1;
    const int row = p / a.W, col = p - row * a.W;
    const float x = __ldg(a.lin + col), y = -__ldg(a.lin + row);
    float v[12], g[12];
    const size_t off = (size_t)b * 12 * a.HW + p;
    load12(a.maps + off, a.HW, v);
    const float* rec = sc.v + (a.per_batch ? (size_t)b * a.N * kRecFloats : 0);
    const float* gin = a.gimages + (size_t)b * a.N * 3 * a.HW + p;
    if (__all_sync(0xffffffffu, same3(v))) render_bwd_pixel<1>(v, x, y, rec, a.N, gin, a.HW, g);
    else                                   render_bwd_pixel<3>(v, x, y, rec, a.N, gin, a.HW, g);
    if (live) {
        float* gp = a.gmaps + off;
#pragma unroll
        for (int c = 0; c < 12; ++c) __stcs(gp + (size_t)c * a.HW, g[c]);
    }
}

__global__ void __launch_bounds__(kThreads)
scale_kernel(float* __restrict__ g, size_t count, const float* __restrict__ upstream) {
    const float u = __ldg(upstream);
    if (u == 1.0f) return;                      // loss.backward(): nothing to do, decided on the device
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) g[i] *= u;
}

// ---------------------------------------------------------------------------------------------
// FP32 throughput probes (bench.py: measured denominators for the FP32 roofline)
// ---------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(kThreads)
probe_kernel(int iters, float* __restrict__ sink) {
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = 1.0f + 1e-3f * (float)((threadIdx.x + i) & 7);
    const float m = 0.9999f + 1e-7f * (float)(threadIdx.x & 3), c = 1e-4f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (KIND == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], m, c);
            } else if (KIND == 1) {
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    unsigned long long d, a2, m2, c2;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(a2) : "f"(r[i]), "f"(r[i + 1]));
                    asm("mov.b64 %0, {%1, %1};" : "=l"(m2) : "f"(m));
                    asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(c));
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a2), "l"(m2), "l"(c2));
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(r[i]), "=f"(r[i + 1]) : "l"(d));
                }
            } else if (KIND == 2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = mufu_rcp(r[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; i += 2) { r[i] = r[i] * m; r[i + 1] = r[i + 1] + c; }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    if (s == 123.456f) sink[blockIdx.x * kThreads + threadIdx.x] = s;   // keeps the chain alive
}

}  // namespace svb

// =============================================================================================
// C ABI
// =============================================================================================
using namespace svb;

static thread_local char g_err[256] = "";
#define fail svb_fail
#define cuda_status svb_cuda_status

int svb_fail(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int svb_cuda_status(cudaError_t e, const char* where) {
    if (e == cudaSuccess) return 0;
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}

int svb_check_shape(int B, int H, int W, int N) {
    if (B <= 0 || H <= 0 || W <= 0 || N <= 0) return fail(SVBRDF_E_INVALID, "B, H, W and N must be positive");
    if (H != W) return fail(SVBRDF_E_INVALID, "maps must be square (H == W), as in renderers.py:73-76");
    if ((long long)H * W > (1LL << 28)) return fail(SVBRDF_E_TOO_LARGE, "H*W exceeds 2^28 pixels");
    if (N > kCapLarge) return fail(SVBRDF_E_TOO_LARGE, "more than 796 scene records per batch element");
    return 0;
}

static inline int ctas_per_image(int HW) { return (HW + kThreads - 1) / kThreads; }

extern "C" int svbrdf_b200_abi_version(void) { return SVBRDF_B200_ABI_VERSION; }
extern "C" const char* svbrdf_b200_last_error(void) { return g_err; }

extern "C" size_t svbrdf_b200_workspace_bytes(int B, int N, int H, int W) {
    (void)N;
    if (B <= 0 || H <= 0 || W <= 0) return 256;
    const size_t ctas = (size_t)B * (size_t)ctas_per_image(H * W);
    return 2 * ctas * sizeof(float) + 256;     // render partials + map-L1 partials
}

// Batch elements per launch such that the launch's records fit the parameter block.
static inline int batch_per_launch(int N, int per_batch, int cap) {
    if (!per_batch) return 65535;
    int bc = cap / N;
    return bc > 65535 ? 65535 : bc;
}

template <int CAP, typename Args, typename Kernel>
static cudaError_t launch_with_scenes(Kernel kernel, dim3 grid, const Args& args, const float* recs, int nrec,
                                      cudaStream_t st) {
    SceneBlock<CAP> blk;
    memcpy(blk.v, recs, (size_t)nrec * kRecFloats * sizeof(float));
    kernel<<<grid, kThreads, 0, st>>>(args, blk);
    return cudaGetLastError();
}

// Enqueues the loss kernel for batch elements [b0, b0+bn) of a B-element problem (several launches if
// the records do not fit one parameter block).  Pointers are for the WHOLE problem.
int svb_launch_loss_range(const float* input, const float* target, float* grad, int B, int HW, int W,
                          const float* scenes, int N, const float* lin, float* part_render, float* part_l1,
                          bool mixed, float l1_weight, int b0, int bn, cudaStream_t st) {
    const int cpi = ctas_per_image(HW);
    LossArgs a;
    a.lin = lin; a.HW = HW; a.W = W; a.N = N;
    a.scale_render = (float)(1.0 / ((double)B * N * 3.0 * HW));
    a.scale_l1 = (float)((double)l1_weight / ((double)B * 3.0 * HW));
    const bool small = (size_t)bn * N <= (size_t)kCapSmall;
    const int bc_max = batch_per_launch(N, 1, small ? kCapSmall : kCapLarge);
    for (int s0 = b0; s0 < b0 + bn; s0 += bc_max) {
        const int bc = (b0 + bn - s0 < bc_max) ? (b0 + bn - s0) : bc_max;
        a.input = input + (size_t)s0 * 12 * HW;
        a.target = target + (size_t)s0 * 12 * HW;
        a.grad = grad ? grad + (size_t)s0 * 12 * HW : nullptr;
        a.part_render = part_render + (size_t)s0 * cpi;
        a.part_l1 = part_l1 + (size_t)s0 * cpi;
        const dim3 grid(cpi, bc);
        const float* recs = scenes + (size_t)s0 * N * kRecFloats;
        cudaError_t e;
#define SVB_LAUNCH(BWD, MIX)                                                                                 \
    (small ? launch_with_scenes<kCapSmall>(loss_kernel<BWD, MIX, kCapSmall>, grid, a, recs, bc * N, st)     \
           : launch_with_scenes<kCapLarge>(loss_kernel<BWD, MIX, kCapLarge>, grid, a, recs, bc * N, st))
        if (grad) e = mixed ? SVB_LAUNCH(true, true) : SVB_LAUNCH(true, false);
        else      e = mixed ? SVB_LAUNCH(false, true) : SVB_LAUNCH(false, false);
#undef SVB_LAUNCH
        if (e != cudaSuccess) return cuda_status(e, "loss_kernel launch");
    }
    return 0;
}

// Adds the per-CTA partials of a B-element problem in a fixed order (fp64) and writes the loss value(s).
int svb_launch_finalize(const float* part_render, const float* part_l1, int B, int HW, int N, bool mixed,
                        float l1_weight, float* out, int n_out, cudaStream_t st) {
    const size_t total_ctas = (size_t)B * ctas_per_image(HW);
    // ln2 converts the log2 differences to natural log; 1/M is the mean of losses.py:50.
    const double mul_render = (double)kLn2 / ((double)B * N * 3.0 * HW);
    const double mul_l1 = 1.0 / ((double)B * 3.0 * HW);
    finalize_kernel<<<1, 1024, 0, st>>>(part_render, mixed ? part_l1 : nullptr, (int)total_ctas, mul_render, mul_l1,
                                        mixed ? l1_weight : 0.f, out, n_out);
    return cuda_status(cudaGetLastError(), "finalize_kernel launch");
}

static int loss_impl(const float* input, const float* target, int B, int H, int W, const float* scenes, int N,
                     const float* lin, float* out, int n_out, float* grad, void* ws, size_t ws_bytes,
                     bool mixed, float l1_weight, void* stream) {
    if (int e = svb_check_shape(B, H, W, N)) return e;
    if (!input || !target || !scenes || !lin || !out || !ws) return fail(SVBRDF_E_INVALID, "null pointer argument");
    if (ws_bytes < svbrdf_b200_workspace_bytes(B, N, H, W)) return fail(SVBRDF_E_INVALID, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    float* part_render = (float*)ws;
    float* part_l1 = part_render + (size_t)B * ctas_per_image(HW);
    if (int e = svb_launch_loss_range(input, target, grad, B, HW, W, scenes, N, lin, part_render, part_l1, mixed,
                                      l1_weight, 0, B, st))
        return e;
    return svb_launch_finalize(part_render, part_l1, B, HW, N, mixed, l1_weight, out, n_out, st);
}

extern "C" int svbrdf_b200_loss_forward(const float* input_dev, const float* target_dev, int B, int H, int W,
                                        const float* scenes_host, int N, const float* lin_dev, float* loss_dev,
                                        void* workspace_dev, size_t workspace_bytes, void* stream) {
    return loss_impl(input_dev, target_dev, B, H, W, scenes_host, N, lin_dev, loss_dev, 1, nullptr, workspace_dev,
                     workspace_bytes, false, 0.f, stream);
}

extern "C" int svbrdf_b200_loss_forward_backward(const float* input_dev, const float* target_dev, int B, int H, int W,
                                                 const float* scenes_host, int N, const float* lin_dev,
                                                 float* loss_dev, float* grad_input_dev, void* workspace_dev,
                                                 size_t workspace_bytes, void* stream) {
    if (!grad_input_dev) return fail(SVBRDF_E_INVALID, "grad_input_dev is null");
    return loss_impl(input_dev, target_dev, B, H, W, scenes_host, N, lin_dev, loss_dev, 1, grad_input_dev,
                     workspace_dev, workspace_bytes, false, 0.f, stream);
}

extern "C" int svbrdf_b200_mixed_loss_forward_backward(const float* input_dev, const float* target_dev, int B, int H,
                                                       int W, const float* scenes_host, int N, float l1_weight,
                                                       const float* lin_dev, float* out_dev, float* grad_input_dev,
                                                       void* workspace_dev, size_t workspace_bytes, void* stream) {
    return loss_impl(input_dev, target_dev, B, H, W, scenes_host, N, lin_dev, out_dev, 3, grad_input_dev,
                     workspace_dev, workspace_bytes, true, l1_weight, stream);
}

static int render_impl(const float* maps, int B, int H, int W, const float* scenes, int N, int per_batch,
                       const float* lin, float* images, const float* gimages, float* gmaps, void* stream) {
    if (int e = svb_check_shape(B, H, W, N)) return e;
    if (!maps || !scenes || !lin) return fail(SVBRDF_E_INVALID, "null pointer argument");
    const bool backward = gmaps != nullptr;
    if (backward ? !gimages : !images) return fail(SVBRDF_E_INVALID, "null pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W, cpi = ctas_per_image(HW);
    RenderArgs a;
    a.lin = lin; a.HW = HW; a.W = W; a.N = N; a.per_batch = per_batch ? 1 : 0;
    const size_t nrec_total = per_batch ? (size_t)B * N : (size_t)N;
    const bool small = nrec_total <= (size_t)kCapSmall;
    const int bc_max = batch_per_launch(N, per_batch, small ? kCapSmall : kCapLarge);
    for (int b0 = 0; b0 < B; b0 += bc_max) {
        const int bc = (B - b0 < bc_max) ? (B - b0) : bc_max;
        a.maps = maps + (size_t)b0 * 12 * HW;
        a.images = images ? images + (size_t)b0 * N * 3 * HW : nullptr;
        a.gimages = gimages ? gimages + (size_t)b0 * N * 3 * HW : nullptr;
        a.gmaps = gmaps ? gmaps + (size_t)b0 * 12 * HW : nullptr;
        const dim3 grid(cpi, bc);
        const float* recs = per_batch ? scenes + (size_t)b0 * N * kRecFloats : scenes;
        const int nrec = per_batch ? bc * N : N;
        cudaError_t e;
        if (backward)
            e = small ? launch_with_scenes<kCapSmall>(render_bwd_kernel<kCapSmall>, grid, a, recs, nrec, st)
                      : launch_with_scenes<kCapLarge>(render_bwd_kernel<kCapLarge>, grid, a, recs, nrec, st);
        else
            e = small ? launch_with_scenes<kCapSmall>(render_fwd_kernel<kCapSmall>, grid, a, recs, nrec, st)
                      : launch_with_scenes<kCapLarge>(render_fwd_kernel<kCapLarge>, grid, a, recs, nrec, st);
        if (e != cudaSuccess) return cuda_status(e, backward ? "render_bwd_kernel launch" : "render_fwd_kernel launch");
    }
    return 0;
}

extern "C" int svbrdf_b200_render_forward(const float* maps_dev, int B, int H, int W, const float* scenes_host, int N,
                                          int scenes_per_batch, const float* lin_dev, float* images_dev, void* stream) {
    return render_impl(maps_dev, B, H, W, scenes_host, N, scenes_per_batch, lin_dev, images_dev, nullptr, nullptr, stream);
}

extern "C" int svbrdf_b200_render_backward(const float* maps_dev, int B, int H, int W, const float* scenes_host, int N,
                                           int scenes_per_batch, const float* lin_dev, const float* grad_images_dev,
                                           float* grad_maps_dev, void* stream) {
    if (!grad_maps_dev) return fail(SVBRDF_E_INVALID, "grad_maps_dev is null");
    return render_impl(maps_dev, B, H, W, scenes_host, N, scenes_per_batch, lin_dev, nullptr, grad_images_dev,
                       grad_maps_dev, stream);
}

extern "C" int svbrdf_b200_scale_grad(float* grad_dev, size_t count, const float* upstream_dev, void* stream) {
    if (!grad_dev || !upstream_dev) return fail(SVBRDF_E_INVALID, "null pointer argument");
    if (count == 0) return 0;
    size_t blocks = (count + kThreads * 8 - 1) / (kThreads * 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    scale_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(grad_dev, count, upstream_dev);
    return cuda_status(cudaGetLastError(), "scale_kernel launch");
}

extern "C" int svbrdf_b200_probe_launch(int kind, int blocks, int iters, float* sink_dev, int* ops_per_thread_iter,
                                        void* stream) {
    if (blocks <= 0 || iters <= 0 || !sink_dev) return fail(SVBRDF_E_INVALID, "bad probe arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int ops = 8 * 16;   // unroll 8 x 16 registers, one counted op each (f32x2 counts 2 per instruction -> still 16)
    switch (kind) {
        case 0: probe_kernel<0><<<blocks, kThreads, 0, st>>>(iters, sink_dev); break;
        case 1: probe_kernel<1><<<blocks, kThreads, 0, st>>>(iters, sink_dev); break;
        case 2: probe_kernel<2><<<blocks, kThreads, 0, st>>>(iters, sink_dev); break;
        case 3: probe_kernel<3><<<blocks, kThreads, 0, st>>>(iters, sink_dev); break;
        default: return fail(SVBRDF_E_INVALID, "unknown probe kind");
    }
    if (ops_per_thread_iter) *ops_per_thread_iter = ops;
    return cuda_status(cudaGetLastError(), "probe_kernel launch");
}
