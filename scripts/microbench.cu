// microbench.cu - development probe: dependent-chain latencies and issue rates of the instructions the
// shading loop is made of (FFMA, FFMA2/FMUL2/FADD2, MUFU, FMNMX) on sm_100a.  One warp per SM-subpartition
// for latency, many warps for throughput.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define REP 256
struct F2 { unsigned long long v; };
__device__ __forceinline__ F2 mk2(float a, float b) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(F2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return x; }
__device__ __forceinline__ float hi(F2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return y; }

template <int KIND, int ILP>
__global__ void lat(float* out, long long* cyc, float seed) {
    float a[ILP]; F2 p[ILP];
    for (int i = 0; i < ILP; ++i) { a[i] = seed + i + threadIdx.x * 1e-3f; p[i] = mk2(a[i], a[i] + 0.5f); }
    const float m = 0.999f, c = 1e-3f;
    const F2 m2 = mk2(m, m), c2 = mk2(c, c);
    long long t0 = clock64();
#pragma unroll
    for (int r = 0; r < REP; ++r) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) a[i] = fmaf(a[i], m, c);
            if (KIND == 1) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i].v) : "l"(p[i].v), "l"(m2.v), "l"(c2.v));
            if (KIND == 2) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p[i].v) : "l"(p[i].v), "l"(m2.v));
            if (KIND == 3) asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(p[i].v) : "l"(p[i].v), "l"(c2.v));
            if (KIND == 4) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(a[i]));
            if (KIND == 5) a[i] = fmaxf(a[i] * 1.0f, c);   // FMNMX (alu) after FMUL
            if (KIND == 6) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(a[i]));
            if (KIND == 7) a[i] = a[i] * m;                 // FMUL
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < ILP; ++i) s += a[i] + lo(p[i]) + hi(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int KIND, int ILP>
void run(const char* name, int threads) {
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    lat<KIND, ILP><<<1, threads>>>(out, cyc, 1.0f);
    lat<KIND, ILP><<<1, threads>>>(out, cyc, 1.0f);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-10s ILP=%d warps/SMSP=%d : %.2f cycles per instruction per warp, %.3f instr/clk/SMSP\n", name, ILP, threads / 128,
           (double)h / (REP * ILP), (double)(REP * ILP) * (threads / 128.0) / (double)h);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    const char* names[] = {"FFMA", "FFMA2", "FMUL2", "FADD2", "MUFU.RCP", "FMUL+FMNMX", "MUFU.LG2", "FMUL"};
    printf("--- dependent chains, 1 warp per SMSP (latency) ---\n");
    run<0, 1>(names[0], 128); run<1, 1>(names[1], 128); run<2, 1>(names[2], 128); run<3, 1>(names[3], 128);
    run<4, 1>(names[4], 128); run<5, 1>(names[5], 128); run<6, 1>(names[6], 128); run<7, 1>(names[7], 128);
    printf("--- ILP 2/4/8, 1 warp per SMSP ---\n");
    run<0, 2>(names[0], 128); run<0, 4>(names[0], 128); run<0, 8>(names[0], 128);
    run<1, 2>(names[1], 128); run<1, 4>(names[1], 128); run<1, 8>(names[1], 128);
    run<4, 2>(names[4], 128); run<4, 4>(names[4], 128); run<4, 8>(names[4], 128);
    printf("--- ILP 1, 2/4/8 warps per SMSP ---\n");
    run<0, 1>(names[0], 256); run<0, 1>(names[0], 512); run<0, 1>(names[0], 1024);
    run<1, 1>(names[1], 256); run<1, 1>(names[1], 512); run<1, 1>(names[1], 1024);
    run<4, 1>(names[4], 256); run<4, 1>(names[4], 512); run<4, 1>(names[4], 1024);
    printf("--- ILP 4, 4 warps per SMSP (throughput) ---\n");
    run<0, 4>(names[0], 512); run<1, 4>(names[1], 512); run<2, 4>(names[2], 512); run<3, 4>(names[3], 512);
    run<4, 4>(names[4], 512); run<6, 4>(names[6], 512); run<7, 4>(names[7], 512);
    return 0;
}
