// microbench2.cu - does an FFMA2 block the issue port for 2 cycles?  One warp per SMSP, independent chains.
#include <cstdio>
#include <cuda_runtime.h>
#define REP 128
template <int NP, int NA, int NM, int NS>   // per round: NP packed FFMA2, NA alu FMNMX, NM MUFU, NS scalar FFMA (all independent chains)
__global__ void mix(float* out, long long* cyc, float seed) {
    unsigned long long p[8]; float a[8], m[8], s[8];
    for (int i = 0; i < 8; ++i) {
        float x = seed + i + threadIdx.x * 1e-3f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(x), "f"(x + 0.5f));
        a[i] = x; m[i] = x + 2.f; s[i] = x + 3.f;
    }
    unsigned long long m2, c2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(m2) : "f"(0.999f));
    asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(1e-3f));
    long long t0 = clock64();
#pragma unroll
    for (int r = 0; r < REP; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < NP) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(m2), "l"(c2));
            if (i < NA) asm volatile("max.f32 %0, %1, %2;" : "=f"(a[i]) : "f"(a[i]), "f"(s[7 - i]));
            if (i < NM) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(m[i]) : "f"(m[i]));
            if (i < NS) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(s[i]) : "f"(s[i]), "f"(0.999f), "f"(1e-3f));
        }
    }
    long long t1 = clock64();
    float acc = 0.f;
    for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); acc += x + y + a[i] + m[i] + s[i]; }
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NP, int NA, int NM, int NS>
void run(int threads = 128) {
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 8);
    mix<NP, NA, NM, NS><<<1, threads>>>(out, cyc, 1.0f);
    mix<NP, NA, NM, NS><<<1, threads>>>(out, cyc, 1.0f);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("per round: %d FFMA2 + %d FMNMX + %d MUFU + %d FFMA, %d warp(s)/SMSP : %.2f cycles per round\n", NP, NA, NM, NS, threads / 128, (double)h / REP);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<4, 0, 0, 0>(); run<0, 4, 0, 0>(); run<0, 0, 0, 4>(); run<0, 8, 0, 0>(); run<0, 0, 0, 8>();
    run<4, 4, 0, 0>(); run<4, 8, 0, 0>(); run<4, 0, 0, 4>(); run<4, 0, 1, 0>(); run<4, 4, 1, 0>(); run<8, 8, 1, 0>();
    run<0, 4, 0, 4>(); run<0, 8, 0, 8>(); run<0, 4, 1, 4>();
    run<8, 0, 0, 0>(); run<8, 8, 0, 0>(); run<8, 4, 0, 0>();
    return 0;
}
