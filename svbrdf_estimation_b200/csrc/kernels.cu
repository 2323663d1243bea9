// kernels.cu - sm_100a kernels of the rendering-loss path and their C-ABI launchers.
//
// Work decomposition (DESIGN.md "Kernels"):
//   grid = (ceil(H*W / (threads * lanes)), batch elements of this launch).  A thread owns `lanes`
//   horizontally adjacent pixels: 2 when W is even (lane type F2: packed FADD2/FMUL2/FFMA2 math,
//   64-bit coalesced loads/stores), 1 otherwise.  It loads its pixels' 12 (+12 target) channels once
//   (a warp reads 128/256 contiguous bytes of each of the 24 planes), keeps them and the 12 gradient
//   accumulators in registers, loops over the N scene records of its batch element - which live in
//   the kernel parameter block (constant bank, warp-uniform addresses, no scene upload) - and never
//   writes a per-record intermediate to memory.  The log-L1 terms are reduced thread -> warp
//   (shuffle) -> CTA (shared memory) -> one partial per CTA; a 1-CTA finalize kernel adds the
//   partials in a fixed order in fp64, so the loss is run-to-run deterministic and uses no float
//   atomics.
//
// Reference semantics: LocalRenderer.render renderers.py:67-104, RenderingLoss.forward
// losses.py:29-52, SVBRDFL1Loss/MixedLoss losses.py:7-19,54-63 (paths relative to
// development/multiImage_pytorch/ of mworchel/svbrdf-estimation).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/svbrdf_b200.h"
#include "pixel_ops.cuh"
#include "internal.h"

namespace svb {

// Kernel parameters may total 32,764 bytes on sm_70+ with CUDA >= 12.1.  Two capacities keep the
// parameter copy small for the common render(scene, maps) call.
constexpr int kCapSmall = 64;    // records ->  2,304 B
constexpr int kCapLarge = 900;   // records -> 32,400 B (+ <= 96 B of other arguments < 32,764)

// Threads per CTA and the occupancy promised to ptxas (register cap) by lane type.  The packed
// kernels hold two pixels of state per thread; 128 threads x 3 CTAs/SM (<= 170 registers) measured
// best on B200 (DESIGN.md "Tuning log"): the loop is bound by register-file operand bandwidth, not
// by occupancy, so more resident warps bought nothing once spills appeared.
#ifndef SVB_F2_THREADS
#define SVB_F2_THREADS 128
#endif
#ifndef SVB_F2_MINB
#define SVB_F2_MINB 3
#endif
template <typename T> struct Cfg;
template <> struct Cfg<float> { static constexpr int kThreads = 256; static constexpr int kMinBlocks = 2; static constexpr int kRenderBwdMinBlocks = 2; };
#ifndef SVB_RBWD_MINB
#define SVB_RBWD_MINB 4
#endif
template <> struct Cfg<F2> { static constexpr int kThreads = SVB_F2_THREADS; static constexpr int kMinBlocks = SVB_F2_MINB; static constexpr int kRenderBwdMinBlocks = SVB_RBWD_MINB; };

template <int CAP>
struct SceneBlock {
    float v[CAP * kRecFloats];
};
static_assert(sizeof(SceneBlock<kCapLarge>) + 128 <= 32764, "scene records + arguments must fit the kernel parameter space");

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

// ---- lane-typed global memory access -------------------------------------------------------------
__device__ __forceinline__ void ld_lane(const float* p, float& v) { v = __ldg(p); }
__device__ __forceinline__ void ld_lane(const float* p, F2& v) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v = mk2(t.x, t.y);
}
__device__ __forceinline__ void ld_stream(const float* p, float& v) { v = __ldcs(p); }
__device__ __forceinline__ void ld_stream(const float* p, F2& v) {
    const float2 t = __ldcs(reinterpret_cast<const float2*>(p));
    v = mk2(t.x, t.y);
}
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float* p, F2 v) { __stcs(reinterpret_cast<float2*>(p), make_float2(lo(v), hi(v))); }

template <typename T>
__device__ __forceinline__ void load12(const float* __restrict__ base, int HW, T (&v)[12]) {
#pragma unroll
    for (int c = 0; c < 12; ++c) ld_lane(base + (size_t)c * HW, v[c]);
}
template <typename T>
__device__ __forceinline__ void store12(float* __restrict__ base, int HW, const T (&v)[12]) {
#pragma unroll
    for (int c = 0; c < 12; ++c) st_stream(base + (size_t)c * HW, v[c]);
}

// 10-channel layout [n(3) d(3) r(1) s(3)]: the information content of the 12-channel contract when the three roughness
// channels are replicas of one map, which is what the model and the dataset produce (utils.py:78-80).  The kernels work
// on the 12-channel form in registers; the gradient of the single roughness channel is the sum of the three.
template <typename T>
__device__ __forceinline__ void load10(const float* __restrict__ base, int HW, T (&v)[12]) {
#pragma unroll
    for (int c = 0; c < 6; ++c) ld_lane(base + (size_t)c * HW, v[c]);
    ld_lane(base + (size_t)6 * HW, v[6]);
    v[7] = v[6]; v[8] = v[6];
#pragma unroll
    for (int c = 0; c < 3; ++c) ld_lane(base + (size_t)(7 + c) * HW, v[9 + c]);
}
template <typename T>
__device__ __forceinline__ void store10(float* __restrict__ base, int HW, const T (&v)[12]) {
#pragma unroll
    for (int c = 0; c < 6; ++c) st_stream(base + (size_t)c * HW, v[c]);
    st_stream(base + (size_t)6 * HW, (v[6] + v[7]) + v[8]);
#pragma unroll
    for (int c = 0; c < 3; ++c) st_stream(base + (size_t)(7 + c) * HW, v[9 + c]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA; result valid in thread 0.  Fixed order => deterministic.
template <int THREADS>
__device__ __forceinline__ float cta_sum(float v, float* smem /* [THREADS/32] */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) t += smem[w];
    }
    __syncthreads();
    return t;
}

// Pixel bookkeeping shared by all kernels: first pixel of this thread, its coordinates.
template <typename T>
struct Where {
    bool live;
    int p;        // first pixel (clamped into the image for dead threads)
    T x;          // lin[col] per lane      (renderers.py:73-76)
    float y;      // -lin[row]
};
template <typename T>
__device__ __forceinline__ Where<T> locate(int HW, int W, const float* __restrict__ lin) {
    constexpr int L = LaneTraits<T>::kLanes;
    Where<T> w;
    const int pix = (blockIdx.x * Cfg<T>::kThreads + threadIdx.x) * L;
    w.live = pix < HW;
    w.p = w.live ? pix : HW - L;
    const int row = w.p / W, col = w.p - row * W;
    ld_lane(lin + col, w.x);
    w.y = -__ldg(lin + row);
    return w;
}

// ---------------------------------------------------------------------------------------------
// RenderingLoss / MixedLoss kernel (per-pixel work: pixel_ops.cuh)
// ---------------------------------------------------------------------------------------------
// ENC: `input` is the network's 9-channel encoded output [B,9,H,W] (decoded on the fly, gradient written
// for the 9 encoded channels) instead of 12-channel maps.
//
// ROWTAB (packed lanes, W % 64 == 0, N <= kRowTabMaxN): the 64 pixels of a warp lie in one image row, so everything
// a scene record contributes that does not depend on the column - light/camera y and z offsets, their squared sums,
// colour / pi (shading.cuh RecScalars) - is formed ONCE per warp into a shared-memory table (lane k fills record k)
// and read back in the record loop with three broadcast LDS.128, instead of 7 indexed constant loads and 7 scalar
// FP operations per thread and record.  Same operations on the same values: bit-identical to the other path.
constexpr int kRowTabMaxN = 64;
// (input, target) channel layouts of a loss launch; the gradient has the input's layout
constexpr int kLay12 = 0;      // [B,12,H,W] maps, [B,12,H,W] target                      (the reference's tensors)
constexpr int kLayEnc12 = 1;   // [B,9,H,W] encoded network output, [B,12,H,W] target
constexpr int kLay10 = 2;      // [B,10,H,W] maps, [B,10,H,W] target                      (roughness stored once)
constexpr int kLayEnc10 = 3;   // [B,9,H,W] encoded network output, [B,10,H,W] target
struct RowTabRecs {
    const float4* tab;     // [N][3]
    __device__ __forceinline__ RecScalars get(int k) const {
        const float4 a = tab[3 * k], b = tab[3 * k + 1], c = tab[3 * k + 2];
        RecScalars r;
        r.sx = a.x; r.vx = a.y; r.ly = a.z; r.lz = a.w;
        r.lyz = b.x; r.vy = b.y; r.vz = b.z; r.vyz = b.w;
        r.col[0] = c.x; r.col[1] = c.y; r.col[2] = c.z;
        return r;
    }
};
extern __shared__ float4 svb_rowtab[];

// The per-thread work of the loss kernels (both lane types).  Register budget: the packed kernels are capped at 160
// registers per thread with __maxnreg__ (3 CTAs of 128 threads per SM would allow 168; ptxas settles on 158 and finds
// a 2 % faster schedule there - measured on B200 for caps 144 ... 168, profiles/r2_variants.txt; the MixedLoss,
// encoded-input and accurate kernels keep 168, they spill below), the one-pixel-per-thread kernels at 128 (2 CTAs of
// 256 threads).
template <typename T, bool BWD, bool MIXED, bool GREY, int LAY, int CAP, bool ACC, bool ROWTAB>
__device__ __forceinline__ void loss_body(const LossArgs& a, const SceneBlock<CAP>& sc) {
    constexpr int THREADS = Cfg<T>::kThreads;
    constexpr bool ENC = LAY == kLayEnc12 || LAY == kLayEnc10;
    constexpr int CIN = LAY == kLay12 ? 12 : (LAY == kLay10 ? 10 : 9);       // channels of `input` and of `grad`
    constexpr int CTG = (LAY == kLay10 || LAY == kLayEnc10) ? 10 : 12;         // channels of `target`
    __shared__ float red[THREADS / 32];
    // Programmatic dependent launch: the finalize kernel queued behind this grid may be scheduled as soon as
    // every CTA of this grid is resident (it then waits in griddepcontrol.wait for this grid to complete and
    // flush), so its launch latency overlaps the last wave instead of following it.
    asm volatile("griddepcontrol.launch_dependents;");
    const int b = blockIdx.y;
    const Where<T> w = locate<T>(a.HW, a.W, a.lin);
    const size_t off_tg = (size_t)b * CTG * a.HW + w.p;
    const size_t off_in = (size_t)b * CIN * a.HW + w.p;
    T vi[12], vt[12], g[12], inv_len;
    if (ENC) {
        T e[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) ld_lane(a.input + off_in + (size_t)c * a.HW, e[c]);
        decode_encoded<T>(e, vi, inv_len);
    } else if (CIN == 10) {
        load10<T>(a.input + off_in, a.HW, vi);
    } else {
        load12<T>(a.input + off_in, a.HW, vi);
    }
    if (CTG == 10) load10<T>(a.target + off_tg, a.HW, vt);
    else load12<T>(a.target + off_tg, a.HW, vt);
    const float* rec = sc.v + (size_t)b * a.N * kRecFloats;

    T lsum;
    if (ROWTAB) {
        float4* tab = svb_rowtab + (threadIdx.x >> 5) * (3 * a.N);
        for (int k = threadIdx.x & 31; k < a.N; k += 32) {
            const RecScalars r = rec_scalars(rec + k * kRecFloats, w.y);
            tab[3 * k] = make_float4(r.sx, r.vx, r.ly, r.lz);
            tab[3 * k + 1] = make_float4(r.lyz, r.vy, r.vz, r.vyz);
            tab[3 * k + 2] = make_float4(r.col[0], r.col[1], r.col[2], 0.f);
        }
        __syncwarp();
        const RowTabRecs recs{tab};
        lsum = loss_pixel_rs<T, BWD, GREY, ACC, RowTabRecs>(vi, vt, w.x, recs, a.N, a.scale_render, g);
    } else {
        lsum = loss_pixel<T, BWD, GREY, ACC>(vi, vt, w.x, w.y, rec, a.N, a.scale_render, g);
    }
    T l1 = LaneTraits<T>::splat(0.f);
    if (MIXED) l1 = l1_pixel<T, BWD>(vi, vt, a.scale_l1, g);
    if (BWD && w.live) {
        if (ENC) {
            T ge[9];
            encode_grad<T>(vi, inv_len, g, ge);
#pragma unroll
            for (int c = 0; c < 9; ++c) st_stream(a.grad + off_in + (size_t)c * a.HW, ge[c]);
        } else if (CIN == 10) {
            store10<T>(a.grad + off_in, a.HW, g);
        } else {
            store12<T>(a.grad + off_in, a.HW, g);
        }
    }

    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    const float tr = cta_sum<THREADS>(w.live ? hsum(lsum) : 0.f, red);
    if (threadIdx.x == 0) a.part_render[cta] = tr;
    if (MIXED) {
        const float tl = cta_sum<THREADS>(w.live ? hsum(l1) : 0.f, red);
        if (threadIdx.x == 0) a.part_l1[cta] = tl;
    }
}

#ifndef SVB_LOSS_MAXNREG
#define SVB_LOSS_MAXNREG 160
#endif
template <typename T, bool BWD, bool MIXED, bool GREY, int LAY, int CAP, bool ACC = false, bool ROWTAB = false>
__global__ void __maxnreg__((MIXED || LAY != kLay12 || ACC) ? 168 : SVB_LOSS_MAXNREG)
loss_kernel_packed(const LossArgs a, const __grid_constant__ SceneBlock<CAP> sc) {
    loss_body<T, BWD, MIXED, GREY, LAY, CAP, ACC, ROWTAB>(a, sc);
}
template <typename T, bool BWD, bool MIXED, bool GREY, int LAY, int CAP, bool ACC = false, bool ROWTAB = false>
__global__ void __launch_bounds__(Cfg<float>::kThreads, Cfg<float>::kMinBlocks)
loss_kernel_scalar(const LossArgs a, const __grid_constant__ SceneBlock<CAP> sc) {
    loss_body<T, BWD, MIXED, GREY, LAY, CAP, ACC, ROWTAB>(a, sc);
}

// out[0] = mixed (or rendering) loss, out[1] = rendering loss, out[2] = map-L1 loss (MIXED only).
__global__ void __launch_bounds__(1024)
finalize_kernel(const float* __restrict__ part_render, const float* __restrict__ part_l1, int count,
                double mul_render, double mul_l1, float l1_weight, float* __restrict__ out, int n_out) {
    __shared__ double sm[2][32];
    double s0 = 0.0, s1 = 0.0;
    asm volatile("griddepcontrol.wait;" ::: "memory");      // all loss-kernel grids before us are complete and visible
    // 8 independent loads in flight per thread; the summation order is fixed by (thread, i)
    for (int base = threadIdx.x; base < count; base += 8 * 1024) {
        float v[8], u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = base + j * 1024;
            v[j] = (i < count) ? __ldcg(part_render + i) : 0.f;
            u[j] = (part_l1 && i < count) ? __ldcg(part_l1 + i) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { s0 += (double)v[j]; s1 += (double)u[j]; }
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
// LocalRenderer.render forward / backward kernels (per-pixel work: pixel_ops.cuh)
// ---------------------------------------------------------------------------------------------
struct DeviceIO {
    template <typename T> static __device__ __forceinline__ void ld(const float* p, T& v) { ld_stream(p, v); }
    template <typename T> static __device__ __forceinline__ void st(float* p, T v) { st_stream(p, v); }
};

template <typename T, int CAP, bool GREY>
__global__ void __launch_bounds__(Cfg<T>::kThreads)
render_fwd_kernel(const RenderArgs a, const __grid_constant__ SceneBlock<CAP> sc) {
    const int b = blockIdx.y;
    const Where<T> w = locate<T>(a.HW, a.W, a.lin);
    T v[12];
    load12<T>(a.maps + (size_t)b * 12 * a.HW + w.p, a.HW, v);
    const float* rec = sc.v + (a.per_batch ? (size_t)b * a.N * kRecFloats : 0);
    float* out = a.images + (size_t)b * a.N * 3 * a.HW + w.p;
    render_pixel<T, GREY, DeviceIO>(v, w.x, w.y, rec, a.N, out, (size_t)a.HW, w.live);
}

// Per-thread ring of upstream-gradient records in shared memory, filled with cp.async (LDGSTS): kDepth records
// (3 planes x the thread's pixels each) are in flight per thread without holding registers, which is what this
// HBM-latency-bound kernel needs (12 N bytes per pixel stream through it).  A thread only ever reads back what it
// copied itself, so cp.async.wait_group is the only synchronisation.  Layout [slot][channel][thread]: conflict-free.
#ifndef SVB_RING_DEPTH
#define SVB_RING_DEPTH 3
#endif
template <typename T>
struct RingIO {
    static constexpr int kDepth = SVB_RING_DEPTH, kSlots = SVB_RING_DEPTH + 1;
    static constexpr int kBytes = 4 * LaneTraits<T>::kLanes;
    static constexpr int kStride = Cfg<T>::kThreads * kBytes;          // bytes per (slot, channel)
    uint32_t base;                                                      // shared-space address of this thread's cell 0
    int head, tail;
    __device__ __forceinline__ explicit RingIO(float* smem) {
        base = (uint32_t)__cvta_generic_to_shared(smem) + threadIdx.x * kBytes;
    }
    __device__ __forceinline__ void reset() { head = tail = 0; }
    template <int NC>
    __device__ __forceinline__ void fetch(const float* p, size_t HW) {
        const uint32_t dst = base + head * (3 * kStride);
#pragma unroll
        for (int c = 0; c < NC; ++c)
            asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst + c * kStride), "l"(p + (size_t)c * HW), "n"(kBytes) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        head = (head + 1 == kSlots) ? 0 : head + 1;
    }
    __device__ __forceinline__ void skip() { asm volatile("cp.async.commit_group;" ::: "memory"); }
    template <int NC>
    __device__ __forceinline__ void take(T (&a)[NC]) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kDepth) : "memory");   // all but the kDepth newest groups have landed
        const uint32_t src = base + tail * (3 * kStride);
#pragma unroll
        for (int c = 0; c < NC; ++c) ld_cell(src + c * kStride, a[c]);
        tail = (tail + 1 == kSlots) ? 0 : tail + 1;
    }
    static __device__ __forceinline__ void ld_cell(uint32_t addr, float& v) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); }
    static __device__ __forceinline__ void ld_cell(uint32_t addr, F2& v) {
        float x, y;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr));
        v = mk2(x, y);
    }
};

template <typename T, int CAP>
__global__ void __launch_bounds__(Cfg<T>::kThreads, Cfg<T>::kRenderBwdMinBlocks)
render_bwd_kernel(const RenderArgs a, const __grid_constant__ SceneBlock<CAP> sc) {
    __shared__ __align__(16) float ring_mem[RingIO<T>::kSlots * 3 * Cfg<T>::kThreads * LaneTraits<T>::kLanes];
    RingIO<T> ring(ring_mem);
    const int b = blockIdx.y;
    const Where<T> w = locate<T>(a.HW, a.W, a.lin);
    T v[12], g[12];
    const size_t off = (size_t)b * 12 * a.HW + w.p;
    load12<T>(a.maps + off, a.HW, v);
    const float* rec = sc.v + (a.per_batch ? (size_t)b * a.N * kRecFloats : 0);
    const float* gin = a.gimages + (size_t)b * a.N * 3 * a.HW + w.p;
    render_bwd_pixel<T, RingIO<T>>(v, w.x, w.y, rec, a.N, gin, (size_t)a.HW, g, ring);
    if (w.live) store12<T>(a.gmaps + off, a.HW, g);
}

__global__ void __launch_bounds__(256)
scale_kernel(float* __restrict__ g, size_t count, const float* __restrict__ upstream) {
    const float u = __ldg(upstream);
    if (u == 1.0f) return;                      // loss.backward(): nothing to do, decided on the device
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) g[i] *= u;
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
    if (N > kCapLarge) return fail(SVBRDF_E_TOO_LARGE, "more than 900 scene records per batch element");
    return 0;
}

// Packed (two pixels per thread) kernels need pixel pairs that never straddle a row and 8-byte
// aligned planes: W even (H == W, so H*W is even too) and 8-byte aligned base pointers.  Anything else - odd widths,
// or a contiguous view that starts at an odd float offset - takes the one-pixel-per-thread kernels, which only need
// the natural 4-byte alignment of a float.
static inline bool aligned8(const void* p) { return ((uintptr_t)p & 7u) == 0; }
static inline int pixels_per_cta(bool packed) { return packed ? 2 * Cfg<F2>::kThreads : Cfg<float>::kThreads; }
int svb_ctas_per_image(int HW, bool packed) { return (HW + pixels_per_cta(packed) - 1) / pixels_per_cta(packed); }
bool svb_loss_packed(int W, const void* input, const void* target, const void* grad, const void* lin) {
    return (W & 1) == 0 && aligned8(input) && aligned8(target) && aligned8(lin) && (!grad || aligned8(grad));
}

extern "C" int svbrdf_b200_abi_version(void) { return SVBRDF_B200_ABI_VERSION; }
extern "C" const char* svbrdf_b200_last_error(void) { return g_err; }
#ifndef SVB_BUILD_ID
#define SVB_BUILD_ID "unknown"
#endif
extern "C" const char* svbrdf_b200_build_id(void) { return SVB_BUILD_ID; }

extern "C" size_t svbrdf_b200_workspace_bytes(int B, int N, int H, int W) {
    (void)N;
    if (B <= 0 || H <= 0 || W <= 0) return 256;
    const int a = svb_ctas_per_image(H * W, true), b = svb_ctas_per_image(H * W, false);
    const size_t ctas = (size_t)B * (size_t)(a > b ? a : b);     // whichever kernel family the pointers will select
    return 2 * ctas * sizeof(float) + 256;                       // render partials + map-L1 partials
}

// Batch elements per launch such that the launch's records fit the parameter block.
static inline int batch_per_launch(int N, int per_batch, int cap) {
    if (!per_batch) return 65535;
    int bc = cap / N;
    return bc > 65535 ? 65535 : bc;
}

// When the records of `total` batch elements do not fit one parameter block, split into equal launches
// (e.g. 32 elements with a cap of 29 -> 16 + 16, not 29 + 3: a 3-element launch fills under two waves).
static inline int balanced_chunk(int total, int cap) {
    if (total <= cap) return total > 0 ? total : 1;
    const int launches = (total + cap - 1) / cap;
    return (total + launches - 1) / launches;
}

template <int CAP, int THREADS, typename Args, typename Kernel>
static cudaError_t launch_with_scenes(Kernel kernel, dim3 grid, const Args& args, const float* recs, int nrec,
                                      cudaStream_t st, size_t smem = 0) {
    SceneBlock<CAP> blk;
    memcpy(blk.v, recs, (size_t)nrec * kRecFloats * sizeof(float));
    kernel<<<grid, THREADS, smem, st>>>(args, blk);
    return cudaGetLastError();
}

// r == g == b light colour in every record?  (selects the GREY kernels)
static bool all_grey(const float* recs, int nrec) {
    for (int i = 0; i < nrec; ++i) {
        const float* c = recs + (size_t)i * kRecFloats + 6;
        if (c[0] != c[1] || c[1] != c[2]) return false;
    }
    return true;
}

// Runtime axes of a loss launch that are template parameters of the kernel.
struct LossSel {
    bool grey;      // all records have r == g == b light colour
    bool rowtab;    // per-warp row table in shared memory (packed lanes, W % 64 == 0, N <= kRowTabMaxN)
    bool small;     // records fit the small parameter block
};

template <typename T, bool BWD, bool MIXED, int LAY, bool ACC>
static cudaError_t launch_loss_v(const LossSel& sel, dim3 grid, const LossArgs& a, const float* recs, int nrec, cudaStream_t st) {
    constexpr int TH = Cfg<T>::kThreads;
    constexpr bool kPacked = LaneTraits<T>::kLanes == 2;
    constexpr bool ENC = LAY != kLay12;          // only the 12-channel layout has small-parameter-block kernels
    const size_t smem = (size_t)(TH / 32) * 3 * a.N * sizeof(float4);
#define SVB_GO_K(KERNEL, CAP, GREY, ROW) launch_with_scenes<CAP, TH>(KERNEL<T, BWD, MIXED, GREY, LAY, CAP, ACC, ROW>, grid, a, recs, nrec, st, (ROW) ? smem : 0)
    if constexpr (kPacked) {
#define SVB_GO(CAP, GREY, ROW) SVB_GO_K(loss_kernel_packed, CAP, GREY, ROW)
        if (sel.grey && sel.rowtab) {          // what RenderingLoss / MixedLoss launch on every power-of-two map size
            if constexpr (ENC) return SVB_GO(kCapLarge, true, true);
            else return sel.small ? SVB_GO(kCapSmall, true, true) : SVB_GO(kCapLarge, true, true);
        }
        if constexpr (ENC) return sel.grey ? SVB_GO(kCapLarge, true, false) : SVB_GO(kCapLarge, false, false);
        else {
            if (sel.grey) return sel.small ? SVB_GO(kCapSmall, true, false) : SVB_GO(kCapLarge, true, false);
            return sel.small ? SVB_GO(kCapSmall, false, false) : SVB_GO(kCapLarge, false, false);
        }
#undef SVB_GO
    } else {
#define SVB_GO(CAP, GREY, ROW) SVB_GO_K(loss_kernel_scalar, CAP, GREY, ROW)
        if constexpr (ENC) return sel.grey ? SVB_GO(kCapLarge, true, false) : SVB_GO(kCapLarge, false, false);
        else {
            if (sel.grey) return sel.small ? SVB_GO(kCapSmall, true, false) : SVB_GO(kCapLarge, true, false);
            return sel.small ? SVB_GO(kCapSmall, false, false) : SVB_GO(kCapLarge, false, false);
        }
#undef SVB_GO
    }
#undef SVB_GO_K
}

template <typename T>
static cudaError_t launch_loss_t(bool bwd, bool mixed, int lay, bool accurate, const LossSel& sel, dim3 grid,
                                 const LossArgs& a, const float* recs, int nrec, cudaStream_t st) {
    if (lay == kLayEnc12) return launch_loss_v<T, true, true, kLayEnc12, false>(sel, grid, a, recs, nrec, st);
    if (lay == kLayEnc10) return launch_loss_v<T, true, true, kLayEnc10, false>(sel, grid, a, recs, nrec, st);
    if (lay == kLay10) return bwd ? launch_loss_v<T, true, false, kLay10, false>(sel, grid, a, recs, nrec, st)
                                  : launch_loss_v<T, false, false, kLay10, false>(sel, grid, a, recs, nrec, st);
    if (accurate) return bwd ? launch_loss_v<T, true, false, kLay12, true>(sel, grid, a, recs, nrec, st)
                             : launch_loss_v<T, false, false, kLay12, true>(sel, grid, a, recs, nrec, st);
    if (bwd) return mixed ? launch_loss_v<T, true, true, kLay12, false>(sel, grid, a, recs, nrec, st)
                          : launch_loss_v<T, true, false, kLay12, false>(sel, grid, a, recs, nrec, st);
    return mixed ? launch_loss_v<T, false, true, kLay12, false>(sel, grid, a, recs, nrec, st)
                 : launch_loss_v<T, false, false, kLay12, false>(sel, grid, a, recs, nrec, st);
}

// Enqueues the loss kernel for batch elements [b0, b0+bn) of a B-element problem (several launches if
// the records do not fit one parameter block).  Pointers are for the WHOLE problem; `packed` is
// svb_loss_packed() of those pointers (the caller sizes the partial arrays with the same value).
int svb_launch_loss_range(const float* input, const float* target, float* grad, int B, int HW, int W,
                          const float* scenes, int N, const float* lin, float* part_render, float* part_l1,
                          bool mixed, float l1_weight, int b0, int bn, cudaStream_t st, bool packed, int lay,
                          bool accurate) {
    const int cin = lay == kLay12 ? 12 : (lay == kLay10 ? 10 : 9);
    const int ctg = (lay == kLay10 || lay == kLayEnc10) ? 10 : 12;
    const int cpi = svb_ctas_per_image(HW, packed);
    LossArgs a;
    a.lin = lin; a.HW = HW; a.W = W; a.N = N;
    a.scale_render = (float)(1.0 / ((double)B * N * 3.0 * HW));
    a.scale_l1 = (float)((double)l1_weight / ((double)B * 3.0 * HW));
    LossSel sel;
    sel.small = lay == kLay12 && (size_t)bn * N <= (size_t)kCapSmall;
#ifdef SVB_NO_ROWTAB            // A/B builds (scripts/variant_bench.py)
    sel.rowtab = false;
#else
    sel.rowtab = packed && (W % 64) == 0 && N <= kRowTabMaxN;
#endif
    const int bc_max = balanced_chunk(bn, batch_per_launch(N, 1, sel.small ? kCapSmall : kCapLarge));
    for (int s0 = b0; s0 < b0 + bn; s0 += bc_max) {
        const int bc = (b0 + bn - s0 < bc_max) ? (b0 + bn - s0) : bc_max;
        a.input = input + (size_t)s0 * cin * HW;
        a.target = target + (size_t)s0 * ctg * HW;
        a.grad = grad ? grad + (size_t)s0 * cin * HW : nullptr;
        a.part_render = part_render + (size_t)s0 * cpi;
        a.part_l1 = part_l1 + (size_t)s0 * cpi;
        const dim3 grid(cpi, bc);
        const float* recs = scenes + (size_t)s0 * N * kRecFloats;
        sel.grey = all_grey(recs, bc * N);
        const cudaError_t e = packed ? launch_loss_t<F2>(grad != nullptr, mixed, lay, accurate, sel, grid, a, recs, bc * N, st)
                                     : launch_loss_t<float>(grad != nullptr, mixed, lay, accurate, sel, grid, a, recs, bc * N, st);
        if (e != cudaSuccess) return cuda_status(e, "loss_kernel launch");
    }
    return 0;
}

// Adds the per-CTA partials of a B-element problem in a fixed order (fp64) and writes the loss value(s).
int svb_launch_finalize(const float* part_render, const float* part_l1, int B, int HW, bool packed, int N, bool mixed,
                        float l1_weight, float* out, int n_out, cudaStream_t st) {
    const size_t total_ctas = (size_t)B * svb_ctas_per_image(HW, packed);
    // ln2 converts the log2 differences to natural log; 1/M is the mean of losses.py:50.
    const double mul_render = (double)kLn2 / ((double)B * N * 3.0 * HW);
    const double mul_l1 = 1.0 / ((double)B * 3.0 * HW);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, finalize_kernel, part_render, mixed ? part_l1 : (const float*)nullptr,
                                             (int)total_ctas, mul_render, mul_l1, mixed ? l1_weight : 0.f, out, n_out);
    return cuda_status(e, "finalize_kernel launch");
}

static int loss_impl(const float* input, const float* target, int B, int H, int W, const float* scenes, int N,
                     const float* lin, float* out, int n_out, float* grad, void* ws, size_t ws_bytes,
                     bool mixed, float l1_weight, void* stream, int lay = kLay12, bool accurate = false) {
    if (int e = svb_check_shape(B, H, W, N)) return e;
    if (!input || !target || !scenes || !lin || !out || !ws) return fail(SVBRDF_E_INVALID, "null pointer argument");
    if (ws_bytes < svbrdf_b200_workspace_bytes(B, N, H, W)) return fail(SVBRDF_E_INVALID, "workspace too small");
    if (((uintptr_t)input | (uintptr_t)target | (uintptr_t)lin | (uintptr_t)grad | (uintptr_t)ws | (uintptr_t)out) & 3u)
        return fail(SVBRDF_E_INVALID, "device pointers must be 4-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    const bool packed = svb_loss_packed(W, input, target, grad, lin);
    float* part_render = (float*)ws;
    float* part_l1 = part_render + (size_t)B * svb_ctas_per_image(HW, packed);
    if (int e = svb_launch_loss_range(input, target, grad, B, HW, W, scenes, N, lin, part_render, part_l1, mixed,
                                      l1_weight, 0, B, st, packed, lay, accurate))
        return e;
    return svb_launch_finalize(part_render, part_l1, B, HW, packed, N, mixed, l1_weight, out, n_out, st);
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

extern "C" int svbrdf_b200_loss_forward_backward_accurate(const float* input_dev, const float* target_dev, int B, int H, int W,
                                                          const float* scenes_host, int N, const float* lin_dev,
                                                          float* loss_dev, float* grad_input_dev, void* workspace_dev,
                                                          size_t workspace_bytes, void* stream) {
    if (!grad_input_dev) return fail(SVBRDF_E_INVALID, "grad_input_dev is null");
    return loss_impl(input_dev, target_dev, B, H, W, scenes_host, N, lin_dev, loss_dev, 1, grad_input_dev,
                     workspace_dev, workspace_bytes, false, 0.f, stream, kLay12, true);
}

extern "C" int svbrdf_b200_loss_forward_accurate(const float* input_dev, const float* target_dev, int B, int H, int W,
                                                 const float* scenes_host, int N, const float* lin_dev, float* loss_dev,
                                                 void* workspace_dev, size_t workspace_bytes, void* stream) {
    return loss_impl(input_dev, target_dev, B, H, W, scenes_host, N, lin_dev, loss_dev, 1, nullptr, workspace_dev,
                     workspace_bytes, false, 0.f, stream, kLay12, true);
}

extern "C" int svbrdf_b200_mixed_loss_forward_backward(const float* input_dev, const float* target_dev, int B, int H,
                                                       int W, const float* scenes_host, int N, float l1_weight,
                                                       const float* lin_dev, float* out_dev, float* grad_input_dev,
                                                       void* workspace_dev, size_t workspace_bytes, void* stream) {
    return loss_impl(input_dev, target_dev, B, H, W, scenes_host, N, lin_dev, out_dev, 3, grad_input_dev,
                     workspace_dev, workspace_bytes, true, l1_weight, stream);
}

extern "C" int svbrdf_b200_mixed_loss_encoded_forward_backward(const float* encoded_dev, const float* target_dev, int B,
                                                               int H, int W, const float* scenes_host, int N,
                                                               float l1_weight, const float* lin_dev, float* out_dev,
                                                               float* grad_encoded_dev, void* workspace_dev,
                                                               size_t workspace_bytes, void* stream) {
    if (!grad_encoded_dev) return fail(SVBRDF_E_INVALID, "grad_encoded_dev is null");
    return loss_impl(encoded_dev, target_dev, B, H, W, scenes_host, N, lin_dev, out_dev, 3, grad_encoded_dev,
                     workspace_dev, workspace_bytes, true, l1_weight, stream, kLayEnc12);
}

// (input layout, target layout) -> kernel layout id, or -1
int svb_layout_id(int input_layout, int target_layout) {
    if (input_layout == SVBRDF_LAYOUT_MAPS12 && target_layout == SVBRDF_LAYOUT_MAPS12) return kLay12;
    if (input_layout == SVBRDF_LAYOUT_ENCODED9 && target_layout == SVBRDF_LAYOUT_MAPS12) return kLayEnc12;
    if (input_layout == SVBRDF_LAYOUT_MAPS10 && target_layout == SVBRDF_LAYOUT_MAPS10) return kLay10;
    if (input_layout == SVBRDF_LAYOUT_ENCODED9 && target_layout == SVBRDF_LAYOUT_MAPS10) return kLayEnc10;
    return -1;
}
// Which loss forms exist for a layout: the encoded-input kernels are MixedLoss forward+backward only, the 10-channel
// maps kernels RenderingLoss only.  Returns 0 or sets the error.
int svb_check_layout_form(int lay, bool mixed, bool has_grad) {
    if (lay < 0) return fail(SVBRDF_E_INVALID, "unsupported (input, target) layout combination");
    if ((lay == kLayEnc12 || lay == kLayEnc10) && !(mixed && has_grad))
        return fail(SVBRDF_E_INVALID, "encoded input: only the MixedLoss forward+backward form exists (l1_weight >= 0, gradient buffer)");
    if (lay == kLay10 && mixed) return fail(SVBRDF_E_INVALID, "10-channel maps: only the RenderingLoss forms exist (l1_weight < 0)");
    return 0;
}

extern "C" int svbrdf_b200_loss_layouts(const float* input_dev, int input_layout, const float* target_dev, int target_layout,
                                        int B, int H, int W, const float* scenes_host, int N, float l1_weight,
                                        const float* lin_dev, float* out_dev, float* grad_input_dev, void* workspace_dev,
                                        size_t workspace_bytes, void* stream) {
    const int lay = svb_layout_id(input_layout, target_layout);
    const bool mixed = l1_weight >= 0.f;
    if (int e = svb_check_layout_form(lay, mixed, grad_input_dev != nullptr)) return e;
    return loss_impl(input_dev, target_dev, B, H, W, scenes_host, N, lin_dev, out_dev, 3, grad_input_dev, workspace_dev,
                     workspace_bytes, mixed, mixed ? l1_weight : 0.f, stream, lay);
}

template <typename T>
static cudaError_t launch_render_t(bool backward, bool small, dim3 grid, const RenderArgs& a, const float* recs,
                                   int nrec, cudaStream_t st) {
    constexpr int TH = Cfg<T>::kThreads;
    if (backward)
        return small ? launch_with_scenes<kCapSmall, TH>(render_bwd_kernel<T, kCapSmall>, grid, a, recs, nrec, st)
                     : launch_with_scenes<kCapLarge, TH>(render_bwd_kernel<T, kCapLarge>, grid, a, recs, nrec, st);
    if (all_grey(recs, nrec))
        return small ? launch_with_scenes<kCapSmall, TH>(render_fwd_kernel<T, kCapSmall, true>, grid, a, recs, nrec, st)
                     : launch_with_scenes<kCapLarge, TH>(render_fwd_kernel<T, kCapLarge, true>, grid, a, recs, nrec, st);
    return small ? launch_with_scenes<kCapSmall, TH>(render_fwd_kernel<T, kCapSmall, false>, grid, a, recs, nrec, st)
                 : launch_with_scenes<kCapLarge, TH>(render_fwd_kernel<T, kCapLarge, false>, grid, a, recs, nrec, st);
}

static int render_impl(const float* maps, int B, int H, int W, const float* scenes, int N, int per_batch,
                       const float* lin, float* images, const float* gimages, float* gmaps, void* stream) {
    if (int e = svb_check_shape(B, H, W, N)) return e;
    if (!maps || !scenes || !lin) return fail(SVBRDF_E_INVALID, "null pointer argument");
    const bool backward = gmaps != nullptr;
    if (backward ? !gimages : !images) return fail(SVBRDF_E_INVALID, "null pointer argument");
    if (((uintptr_t)maps | (uintptr_t)lin | (uintptr_t)images | (uintptr_t)gimages | (uintptr_t)gmaps) & 3u)
        return fail(SVBRDF_E_INVALID, "device pointers must be 4-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const bool packed = (W & 1) == 0 && aligned8(maps) && aligned8(lin) && aligned8(images) && aligned8(gimages) && aligned8(gmaps);
    const int HW = H * W, cpi = svb_ctas_per_image(HW, packed);
    RenderArgs a;
    a.lin = lin; a.HW = HW; a.W = W; a.N = N; a.per_batch = per_batch ? 1 : 0;
    const size_t nrec_total = per_batch ? (size_t)B * N : (size_t)N;
    const bool small = nrec_total <= (size_t)kCapSmall;
    const int bc_max = balanced_chunk(B, batch_per_launch(N, per_batch, small ? kCapSmall : kCapLarge));
    for (int b0 = 0; b0 < B; b0 += bc_max) {
        const int bc = (B - b0 < bc_max) ? (B - b0) : bc_max;
        a.maps = maps + (size_t)b0 * 12 * HW;
        a.images = images ? images + (size_t)b0 * N * 3 * HW : nullptr;
        a.gimages = gimages ? gimages + (size_t)b0 * N * 3 * HW : nullptr;
        a.gmaps = gmaps ? gmaps + (size_t)b0 * 12 * HW : nullptr;
        const dim3 grid(cpi, bc);
        const float* recs = per_batch ? scenes + (size_t)b0 * N * kRecFloats : scenes;
        const int nrec = per_batch ? bc * N : N;
        const cudaError_t e = packed ? launch_render_t<F2>(backward, small, grid, a, recs, nrec, st)
                                     : launch_render_t<float>(backward, small, grid, a, recs, nrec, st);
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
    size_t blocks = (count + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    scale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad_dev, count, upstream_dev);
    return cuda_status(cudaGetLastError(), "scale_kernel launch");
}
