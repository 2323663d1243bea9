// microbench3.cu - register-file bank limits for packed FP32: FFMA2 with 3 distinct register-pair operands
// vs 2 pairs + uniform/immediate.  Independent chains (ILP 8), 1 and 4 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
#define REP 64
#define ILP 8
template <int KIND>
__global__ void k(float* out, long long* cyc, float seed) {
    unsigned long long a[ILP], b[ILP], c[ILP];
    float sa[ILP], sb[ILP], sc[ILP];
    for (int i = 0; i < ILP; ++i) {
        float x = seed + i + threadIdx.x * 1e-3f;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(x), "f"(x + 0.5f));
        asm("mov.b64 %0, {%1, %2};" : "=l"(b[i]) : "f"(0.999f + 1e-6f * x), "f"(0.998f + 1e-6f * x));
        asm("mov.b64 %0, {%1, %2};" : "=l"(c[i]) : "f"(1e-3f * x), "f"(2e-3f * x));
        sa[i] = x; sb[i] = 0.999f + 1e-6f * x; sc[i] = 1e-3f * x;
    }
    unsigned long long u1, u2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(u1) : "f"(seed * 0.999f));
    asm("mov.b64 %0, {%1, %1};" : "=l"(u2) : "f"(seed * 1e-3f));
    __syncthreads();
    long long t0 = clock64();
#pragma unroll
    for (int r = 0; r < REP; ++r) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a[i]) : "l"(a[i]), "l"(b[i]), "l"(c[i]));          // 3 distinct pairs
            if (KIND == 1) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a[i]) : "l"(a[i]), "l"(b[i]), "l"(u2));            // 2 pairs + uniform
            if (KIND == 2) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a[i]) : "l"(a[i]), "l"(u1), "l"(u2));              // 1 pair + 2 uniform
            if (KIND == 3) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(a[i]) : "l"(a[i]), "l"(b[i]));                          // FMUL2 2 pairs
            if (KIND == 4) asm volatile("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(a[i]) : "l"(a[i]), "l"(c[i]));                      // a*a+c
            if (KIND == 5) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(sa[i]) : "f"(sa[i]), "f"(sb[i]), "f"(sc[i]));        // scalar 3 regs
            if (KIND == 6) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a[i]) : "l"(a[i]), "l"(b[i]), "l"(a[(i + 1) % ILP])); // 3 pairs, operand also chain value
        }
    }
    long long t1 = clock64();
    float acc = 0.f;
    for (int i = 0; i < ILP; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); acc += x + y + sa[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == blockDim.x - 1) cyc[0] = t1 - t0;     // last warp (lowest priority in the arbiter?) - report both
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
}
template <int KIND>
void run(const char* name) {
    float* out; long long* cyc; long long h[2];
    cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 16);
    for (int threads : {128, 512}) {
        k<KIND><<<1, threads>>>(out, cyc, 1.0f);
        k<KIND><<<1, threads>>>(out, cyc, 1.0f);
        cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        double w = threads / 128.0;
        printf("%-34s %d warp(s)/SMSP: %.2f / %.2f cycles per instruction per SMSP (last/first warp)\n", name, threads / 128,
               (double)h[0] / (REP * ILP * w), (double)h[1] / (REP * ILP * w));
    }
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("FFMA2 R,R,R (3 distinct pairs)");
    run<1>("FFMA2 R,R,UR");
    run<2>("FFMA2 R,UR,UR");
    run<3>("FMUL2 R,R");
    run<4>("FFMA2 R,R(same),R");
    run<5>("FFMA  R,R,R scalar");
    run<6>("FFMA2 R,R,R (c = other chain)");
    return 0;
}
