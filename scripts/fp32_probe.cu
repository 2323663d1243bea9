// fp32_probe.cu - BENCH-ONLY helper (not part of the product library or its ABI): register-resident FP32
// throughput probes that bench.py uses as measured denominators beside the nominal FP32 peak.
//   kind 0 = independent scalar FFMA chains, 1 = packed fma.rn.f32x2 (FFMA2), 3 = FMUL+FADD mix (1 FLOP each).
// Built by __graft_entry__.build() into scripts/libfp32_probe.so.
#include <cuda_runtime.h>

#include "../svbrdf_estimation_b200/csrc/shading.cuh"

using namespace svb;

template <int KIND>
__global__ void __launch_bounds__(256)
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
                    const F2 t = vfma(mk2(r[i], r[i + 1]), mk2(m, m), mk2(c, c));
                    r[i] = lo(t); r[i + 1] = hi(t);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; i += 2) { r[i] = r[i] * m; r[i + 1] = r[i + 1] + c; }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    if (s == 123.456f) sink[blockIdx.x * 256 + threadIdx.x] = s;   // keeps the chain alive
}

// Launches `blocks` CTAs of 256 threads running `iters` unrolled groups; *ops_per_thread_iter receives the number
// of counted operations (FMA = 1 op, an f32x2 instruction = 2) per thread per iteration.  Returns a cudaError_t.
extern "C" __attribute__((visibility("default")))
int fp32_probe_launch(int kind, int blocks, int iters, float* sink_dev, int* ops_per_thread_iter, void* stream) {
    if (blocks <= 0 || iters <= 0 || !sink_dev) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    switch (kind) {
        case 0: probe_kernel<0><<<blocks, 256, 0, st>>>(iters, sink_dev); break;
        case 1: probe_kernel<1><<<blocks, 256, 0, st>>>(iters, sink_dev); break;
        case 3: probe_kernel<3><<<blocks, 256, 0, st>>>(iters, sink_dev); break;
        default: return -1;
    }
    if (ops_per_thread_iter) *ops_per_thread_iter = 8 * 16;
    return (int)cudaGetLastError();
}
