// scene_sampler.cpp - the random draws of the reference's scene samplers, consumed from torch's global CPU
// generator IN THE REFERENCE'S ORDER, without one Python-level generator call per tensor.
//
// RenderingLoss.forward samples, per batch element, generate_random_scenes(n_random) +
// generate_specular_scenes(n_specular) (losses.py:35): 2-3 legacy `torch.Tensor(n, k).uniform_()/normal_()` calls per
// scene group (environment.py:18-55, utils.py:100-111), i.e. ~9 generator calls per batch element - 0.36 ms per
// element in the reference, 5 us in the vectorised Python sampler of this package, and still the largest host cost of
// a loss evaluation.  This file restates what those calls do to the generator: at::mt19937 (the 32-bit Mersenne
// Twister as ATen drives it), at::uniform_real_distribution<float> and at::normal_distribution<double> with its
// cached second Box-Muller sample (ATen/core/DistributionsHelper.h, the scalar path normal_() takes for tensors of
// fewer than 16 elements).  The caller hands in the serialised generator state (torch.get_rng_state(): the
// CPUGeneratorImplState struct), gets the raw draws of a whole batch back, and stores the advanced state with
// torch.set_rng_state() - after which torch continues exactly where the reference would have left it.
// Host-only code: no CUDA call.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/svbrdf_b200.h"
#include "internal.h"

namespace {

constexpr int kN = 624, kM = 397;

// byte layout of at::CPUGeneratorImplState (ATen/CPUGeneratorImpl.cpp), little endian
struct TorchCpuState {
    uint64_t seed;
    int32_t left;
    int32_t seeded;
    uint64_t next;
    uint64_t state[kN];              // 32-bit words stored as 64-bit
    double normal_x, normal_y, normal_rho;
    int32_t normal_is_valid;         // normal_y holds the cached double sample when set
    int32_t pad0;
    float next_float_normal_sample;
    uint8_t is_next_float_normal_sample_valid;
    uint8_t pad1[3];
};
static_assert(sizeof(TorchCpuState) == 5056, "must match torch.get_rng_state() of the CPU generator");

struct Engine {
    uint32_t s[kN];
    int left, next;

    void refill() {
        uint32_t* p = s;
        auto twist = [](uint32_t u, uint32_t v) {
            return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
        };
        left = kN; next = 0;
        for (int j = kN - kM + 1; --j; ++p) *p = p[kM] ^ twist(p[0], p[1]);
        for (int j = kM; --j; ++p) *p = p[kM - kN] ^ twist(p[0], p[1]);
        *p = p[kM - kN] ^ twist(p[0], s[0]);
    }
    uint32_t random() {
        if (--left == 0) refill();
        uint32_t y = s[next++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    uint64_t random64() {
        const uint32_t hi = random(), lo = random();
        return ((uint64_t)hi << 32) | lo;
    }
    // uniform_real_distribution<float>(0, 1): 24 random bits; the (from, to) affine map is the caller's
    float uniform01f() { return (float)((double)(random() & ((1u << 24) - 1)) * (1.0 / 16777216.0)); }
    double uniform01d() { return (double)(random64() & ((1ull << 53) - 1)) * (1.0 / 9007199254740992.0); }
};

}  // namespace

extern "C" int svbrdf_b200_reference_draws(void* torch_cpu_rng_state, size_t state_bytes, int B, int n_random,
                                           int n_specular, float* uniforms_host, float* normals_host) {
    if (!torch_cpu_rng_state || !uniforms_host || (n_specular > 0 && !normals_host))
        return svb_fail(SVBRDF_E_INVALID, "null pointer argument");
    if (state_bytes != sizeof(TorchCpuState))
        return svb_fail(SVBRDF_E_STATE, "unexpected size of the torch CPU generator state (expected 5056 bytes)");
    if (B <= 0 || n_random < 0 || n_specular < 0 || n_random + n_specular <= 0)
        return svb_fail(SVBRDF_E_INVALID, "bad sampler arguments");
    if (n_specular >= 16)
        return svb_fail(SVBRDF_E_INVALID, "normal_() on 16 or more elements takes ATen's vectorised path; use the torch sampler");
    TorchCpuState st;
    memcpy(&st, torch_cpu_rng_state, sizeof(st));
    if (!st.seeded || st.left < 0 || st.left > kN || st.next > (uint64_t)kN)
        return svb_fail(SVBRDF_E_STATE, "torch CPU generator state is not a seeded mt19937 state");
    Engine e;
    for (int i = 0; i < kN; ++i) e.s[i] = (uint32_t)st.state[i];
    e.left = st.left; e.next = (int)st.next;
    bool have = st.normal_is_valid != 0;
    double cached = st.normal_y;
    const double mean = 0.5, stdv = 0.75;                       // environment.py:38-39

    const int head = 4 * n_random + 2 * n_specular, tail = 2 * n_specular;
    for (int b = 0; b < B; ++b) {
        float* u = uniforms_host + (size_t)b * (head + tail);
        // generate_random_scenes: view (r1.., r2..), light (r1.., r2..); generate_specular_scenes: view (r1.., r2..)
        for (int i = 0; i < head; ++i) u[i] = e.uniform01f();
        // the two normal_(0.5, 0.75) calls (view, then light distances): scalar path, second sample cached
        float* nrm = normals_host + (size_t)b * 2 * n_specular;
        for (int i = 0; i < 2 * n_specular; ++i) {
            double v;
            if (have) {
                v = cached * stdv + mean;
                have = false;
            } else {
                const double u1 = e.uniform01d();
                const double u2 = e.uniform01d();
                const double r = sqrt(-2.0 * log1p(-u2));
                const double theta = 2.0 * 3.14159265358979323846 * u1;
                cached = r * sin(theta);
                have = true;
                v = r * cos(theta) * stdv + mean;
            }
            nrm[i] = (float)v;
        }
        // shift: Tensor(count, 2).uniform_(-1, 1), row-major
        for (int i = 0; i < tail; ++i) u[head + i] = e.uniform01f();
    }
    for (int i = 0; i < kN; ++i) st.state[i] = e.s[i];
    st.left = e.left; st.next = (uint64_t)e.next;
    st.normal_is_valid = have ? 1 : 0;
    st.normal_y = have ? cached : st.normal_y;
    memcpy(torch_cpu_rng_state, &st, sizeof(st));
    return 0;
}

// ---- the two-phase form RenderingLoss uses ------------------------------------------------------------------------
// Everything of environment.py:18-55 / utils.py:100-111 that is exactly reproducible outside torch happens here (the
// draws, the affine maps of uniform_(lo, hi), sqrt, products and sums - each rounded to float exactly where the
// reference's float32 tensor ops round); sqrt, cos, sin and exp stay with torch (its vectorised implementations are
// not libm's), applied once to the whole batch between the two calls:
//     begin  -> r1, angle phi = 2 pi r2 per direction, log-distances, shifts
//     torch  -> r = sqrt(r1), cos(phi), sin(phi), z = sqrt(1 - r*r), exp(log-distance)
//               (torch.sqrt is MKL VML on x86 builds - within 1 ulp but not the correctly rounded sqrtf)
//     finish -> records [B][N][9]
// Directions per batch element are ordered: n_random views, n_random lights, n_specular mirror views.
extern "C" int svbrdf_b200_reference_scenes_begin(void* torch_cpu_rng_state, size_t state_bytes, int B, int n_random,
                                                  int n_specular, float* r1_host, float* phi_host,
                                                  float* log_distance_host, float* shift_host) {
    if (!r1_host || !phi_host || (n_specular > 0 && (!log_distance_host || !shift_host)))
        return svb_fail(SVBRDF_E_INVALID, "null pointer argument");
    if (B <= 0 || n_random < 0 || n_specular < 0 || n_random + n_specular <= 0 || n_random > 4096 || n_specular > 4096)
        return svb_fail(SVBRDF_E_INVALID, "bad sampler arguments");
    const int nr = n_random, ns = n_specular, nd = 2 * nr + ns, head = 4 * nr + 2 * ns, tail = 2 * ns;
    float* uni = new float[(size_t)B * (head + tail)];
    const int rc = svbrdf_b200_reference_draws(torch_cpu_rng_state, state_bytes, B, nr, ns, uni, log_distance_host);
    if (rc == 0) {
        // uniform_(from, to) = x * (to - from) + from with x in double, (to - from) in float, one rounding to float
        const float lo = (float)(0.0 + 0.001), hi = (float)(1.0 - 0.1);       // environment.py:20-21,34
        const double scale = (double)(hi - lo), offset = (double)lo;
        const float two_pi = (float)(2 * 3.14159265358979323846);              // utils.py:105: python scalar times float tensor
        for (int b = 0; b < B; ++b) {
            const float* u = uni + (size_t)b * (head + tail);
            float* r = r1_host + (size_t)b * nd;
            float* ph = phi_host + (size_t)b * nd;
            // stream order: view r1[nr] r2[nr] | light r1[nr] r2[nr] | specular view r1[ns] r2[ns] | shift[ns][2]
            const int r1_at[3] = {0, 2 * nr, 4 * nr}, cnt[3] = {nr, nr, ns}, dir_at[3] = {0, nr, 2 * nr};
            for (int g = 0; g < 3; ++g)
                for (int i = 0; i < cnt[g]; ++i) {
                    const float r1 = (float)((double)u[r1_at[g] + i] * scale + offset);
                    r[dir_at[g] + i] = r1;                                      // the caller takes torch.sqrt (utils.py:104)
                    ph[dir_at[g] + i] = two_pi * u[r1_at[g] + cnt[g] + i];      // utils.py:105
                }
            float* sh = shift_host + (size_t)b * tail;
            for (int i = 0; i < tail; ++i) sh[i] = (float)((double)u[head + i] * 2.0 + -1.0);   // environment.py:44
        }
    }
    delete[] uni;
    return rc;
}

extern "C" int svbrdf_b200_reference_scenes_finish(int B, int n_random, int n_specular, const float* radius_host,
                                                   const float* cos_phi_host, const float* sin_phi_host,
                                                   const float* z_host, const float* distance_host,
                                                   const float* shift_host, float* records_host) {
    if (!radius_host || !cos_phi_host || !sin_phi_host || !z_host || !records_host || (n_specular > 0 && (!distance_host || !shift_host)))
        return svb_fail(SVBRDF_E_INVALID, "null pointer argument");
    if (B <= 0 || n_random < 0 || n_specular < 0 || n_random + n_specular <= 0)
        return svb_fail(SVBRDF_E_INVALID, "bad sampler arguments");
    const int nr = n_random, ns = n_specular, nd = 2 * nr + ns, N = nr + ns;
    for (int b = 0; b < B; ++b) {
        const float* r = radius_host + (size_t)b * nd;
        const float* c = cos_phi_host + (size_t)b * nd;
        const float* s = sin_phi_host + (size_t)b * nd;
        const float* z = z_host + (size_t)b * nd;
        float* rec = records_host + (size_t)b * N * 9;
        auto direction = [&](int i, float* d) {                                  // utils.py:107-109
            d[0] = r[i] * c[i]; d[1] = r[i] * s[i]; d[2] = z[i];
        };
        for (int k = 0; k < nr; ++k, rec += 9) {                                  // environment.py:18-30
            direction(k, rec);
            direction(nr + k, rec + 3);
            rec[6] = rec[7] = rec[8] = 20.0f;
        }
        const float* dist = distance_host + (size_t)b * 2 * ns;                   // [2][ns]: view, light
        const float* sh = shift_host + (size_t)b * 2 * ns;                        // [ns][2]
        for (int k = 0; k < ns; ++k, rec += 9) {                                  // environment.py:32-55
            float v[3];
            direction(2 * nr + k, v);
            const float off[3] = {sh[2 * k], sh[2 * k + 1], 0.0f + 0.0001f};
            const float m[3] = {v[0] * -1.0f, v[1] * -1.0f, v[2] * 1.0f};
            for (int j = 0; j < 3; ++j) {
                const float pv = v[j] * dist[k], pl = m[j] * dist[ns + k];        // product rounded, then the sum (no FMA)
                rec[j] = pv + off[j];
                rec[3 + j] = pl + off[j];
            }
            rec[6] = rec[7] = rec[8] = 50.0f;
        }
    }
    return 0;
}
