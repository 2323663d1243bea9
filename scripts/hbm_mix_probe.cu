// hbm_mix_probe.cu - BENCH-ONLY: what does HBM deliver for the read:write plane mixes of this path's kernels?
// MEASURED_PEAKS.json's hbm_gbs is a 1:1 copy.  The kernels here do no arithmetic: each thread reads R planes and
// writes Wp planes of a [B, planes, H*W] layout with the same 64-bit coalesced accesses and cache hints as the real
// kernels (ld.global.nc / st.global.cs), so the result is the bandwidth bound of that traffic shape.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/hbm_mix_probe scripts/hbm_mix_probe.cu && build/hbm_mix_probe
#include <cuda_runtime.h>
#include <stdio.h>

template <int R, int Wp>
__global__ void __launch_bounds__(128) mix_kernel(const float2* __restrict__ in, float2* __restrict__ out, int HW2) {
    const int p = blockIdx.x * 128 + threadIdx.x;          // pixel pair
    if (p >= HW2) return;
    const size_t bi = (size_t)blockIdx.y * R * HW2, bo = (size_t)blockIdx.y * Wp * HW2;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < R; ++c) {
        const float2 v = __ldg(in + bi + (size_t)c * HW2 + p);
        acc.x += v.x; acc.y += v.y;
    }
#pragma unroll
    for (int c = 0; c < Wp; ++c) __stcs(out + bo + (size_t)c * HW2 + p, make_float2(acc.x + c, acc.y - c));
}

template <int R, int Wp>
static void run(const char* what, int B, int HW) {
    const int HW2 = HW / 2;
    float2 *in, *out;
    const size_t nin = (size_t)B * R * HW2, nout = (size_t)B * Wp * HW2;
    cudaMalloc(&in, 2 * nin * sizeof(float2));               // two rotating sets: larger than L2 together
    cudaMalloc(&out, 2 * nout * sizeof(float2));
    cudaMemset(in, 0, 2 * nin * sizeof(float2));
    dim3 grid((HW2 + 127) / 128, B);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 4; ++i) mix_kernel<R, Wp><<<grid, 128>>>(in + (i & 1) * nin, out + (i & 1) * nout, HW2);
    cudaEventRecord(e0);
    const int n = 40;
    for (int i = 0; i < n; ++i) mix_kernel<R, Wp><<<grid, 128>>>(in + (i & 1) * nin, out + (i & 1) * nout, HW2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= n;
    const double bytes = (double)(nin + nout) * sizeof(float2);
    printf("{\"mix\": \"%s\", \"read_planes\": %d, \"write_planes\": %d, \"B\": %d, \"HW\": %d, \"MB\": %.1f, \"ms\": %.4f, \"GB_s\": %.1f}\n",
           what, R, Wp, B, HW, bytes / 1e6, ms, bytes / ms / 1e6);
    cudaFree(in); cudaFree(out);
}

int main() {
    const int B = 64, HW = 256 * 256;
    run<12, 12>("copy-like 1:1 (12 read, 12 written)", B, HW);
    run<24, 12>("loss fwd+bwd (24 read, 12 written)", B, HW);
    run<24, 1>("loss forward (24 read)", B, HW);
    run<12, 27>("render_forward N=9 (12 read, 27 written)", B, HW);
    run<39, 12>("render_backward N=9 (39 read, 12 written)", B, HW);
    run<1, 27>("write-only-like (1 read, 27 written)", B, HW);
    return 0;
}
