// K5 -- ancestral sampler of the IAR model (fcdiff/model.py:52-236) with a
// counter-based Philox4x32-10 generator: every random variable draws from the
// counter (element index, stream offset, variable id) under the key `seed`, so
// the output does not depend on launch geometry or on how edges are sharded.
// Parity with the reference (MT19937, consumed data-dependently) is
// distributional only.  Edges are in util order everywhere.
#include "fcd_common.cuh"

namespace fcd {

constexpr int kSmpThreads = 256;

enum { VAR_R = 1, VAR_T = 2, VAR_F = 3, VAR_FT = 4, VAR_B = 5, VAR_BT = 6 };

struct Philox {
    uint32_t k0, k1, c2, c3;
};

__host__ __device__ inline Philox philox_make(uint64_t seed, uint64_t offset, int var) {
    Philox p;
    p.k0 = (uint32_t)seed;
    p.k1 = (uint32_t)(seed >> 32);
    p.c2 = (uint32_t)offset;
    p.c3 = ((uint32_t)(offset >> 32) << 8) | (uint32_t)var;
    return p;
}

// Philox4x32-10 (Salmon et al., SC'11): counter (idx_lo, idx_hi, c2, c3).
__device__ __forceinline__ uint4 philox4x32_10(const Philox& p, uint64_t idx) {
    uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = p.c2, c3 = p.c3;
    uint32_t k0 = p.k0, k1 = p.k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// 53-bit uniform in [0, 1) from two words.
__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {
    const uint64_t v = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
    return (double)v * (1.0 / 9007199254740992.0);
}

// Standard normal by Box-Muller from one Philox block.
__device__ __forceinline__ double std_normal(uint4 w) {
    const double u1 = 1.0 - u01(w.x, w.y);           // (0, 1]
    const double u2 = u01(w.z, w.w);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__device__ __forceinline__ int onehot_state(const uint8_t* f3) {
    return f3[1] ? 1 : (f3[2] ? 2 : 0);
}

// r[n,u] ~ Bernoulli(pi)                                     (model.py:108)
__global__ void __launch_bounds__(kSmpThreads)
sample_R_kernel(Philox p, int64_t n, double pi, uint8_t* __restrict__ r) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 w = philox4x32_10(p, (uint64_t)i);
        r[i] = u01(w.x, w.y) < pi;
    }
}

// t[c,u] = r_n & r_m if r_n == r_m else Bernoulli(eta)       (model.py:133-142)
__global__ void __launch_bounds__(kSmpThreads)
sample_T_kernel(Philox p, const uint8_t* __restrict__ r, int U, double eta,
                int64_t c0, int64_t C, uint8_t* __restrict__ t) {
    const int64_t total = C * U;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cl = i / U;
        const int u = (int)(i - cl * U);
        int n, m;
        c_to_nm(c0 + cl, n, m);
        const bool rn = r[(int64_t)n * U + u] != 0, rm = r[(int64_t)m * U + u] != 0;
        uint8_t v;
        if (rn == rm) v = rn;
        else {
            const uint4 w = philox4x32_10(p, (uint64_t)((c0 + cl) * U + u));
            v = u01(w.x, w.y) < eta;
        }
        t[i] = v;
    }
}

// f[c,:] ~ one-hot Categorical(gamma)                         (model.py:160)
__global__ void __launch_bounds__(kSmpThreads)
sample_F_kernel(Philox p, int64_t c0, int64_t C, double g0, double g1, uint8_t* __restrict__ f) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < C;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 w = philox4x32_10(p, (uint64_t)(c0 + i));
        const double x = u01(w.x, w.y);
        const int k = (x >= g0) + (x >= g0 + g1);
        f[i * 3] = (k == 0);
        f[i * 3 + 1] = (k == 1);
        f[i * 3 + 2] = (k == 2);
    }
}

// f_tilde[c,u,:]: keep the template state w.p. (1-eps) if t == 0, eps if t == 1;
// the two other states share the rest equally                (model.py:181-188)
__global__ void __launch_bounds__(kSmpThreads)
sample_F_tilde_kernel(Philox p, const uint8_t* __restrict__ f, const uint8_t* __restrict__ t,
                      int64_t c0, int64_t C, int U, double epsilon, uint8_t* __restrict__ ft) {
    const int64_t total = C * U;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cl = i / U;
        const int u = (int)(i - cl * U);
        const int fk = onehot_state(f + cl * 3);
        const double p_keep = t[i] ? epsilon : 1.0 - epsilon;
        const uint4 w = philox4x32_10(p, (uint64_t)((c0 + cl) * U + u));
        int k = fk;
        if (!(u01(w.x, w.y) < p_keep)) k = (fk + 1 + (u01(w.z, w.w) < 0.5)) % 3;
        ft[i * 3] = (k == 0);
        ft[i * 3 + 1] = (k == 1);
        ft[i * 3 + 2] = (k == 2);
    }
}

struct MuSigma {
    double mu[3], sigma[3];
};

// b[c,h] ~ N(mu_f, sigma_f) clipped to [-1, 1]     (model.py:209-213; 231-236 for
// the patients, where the state is per element: PER_ELEM)
template <bool PER_ELEM>
__global__ void __launch_bounds__(kSmpThreads)
sample_B_kernel(Philox p, const uint8_t* __restrict__ f, int64_t c0, int64_t C, int S,
                const __grid_constant__ MuSigma ms, double* __restrict__ b) {
    const int64_t total = C * S;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cl = i / S;
        const int s = (int)(i - cl * S);
        const int k = onehot_state(PER_ELEM ? f + i * 3 : f + cl * 3);
        const double mu = k == 0 ? ms.mu[0] : (k == 1 ? ms.mu[1] : ms.mu[2]);
        const double sg = k == 0 ? ms.sigma[0] : (k == 1 ? ms.sigma[1] : ms.sigma[2]);
        const double z = std_normal(philox4x32_10(p, (uint64_t)((c0 + cl) * S + s)));
        b[i] = fmin(1.0, fmax(-1.0, fma(sg, z, mu)));
    }
}

static inline int smp_grid(int64_t items) {
    int64_t need = (items + kSmpThreads - 1) / kSmpThreads;
    int64_t cap = (int64_t)sm_count() * 16;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace fcd

using namespace fcd;

extern "C" {

int fcd_sample_R(uint64_t seed, uint64_t offset, int32_t N, int32_t U, double pi,
                 uint8_t* r, void* stream) {
    FCD_REQUIRE(N >= 1 && U >= 1 && r != nullptr, "fcd_sample_R: bad arguments");
    const int64_t n = (int64_t)N * U;
    sample_R_kernel<<<smp_grid(n), kSmpThreads, 0, (cudaStream_t)stream>>>(philox_make(seed, offset, VAR_R), n, pi, r);
    return check_launch("fcd_sample_R");
}

int fcd_sample_T(uint64_t seed, uint64_t offset, const uint8_t* r, int32_t N, int32_t U,
                 double eta, int64_t c0, int64_t C, uint8_t* t, void* stream) {
    FCD_REQUIRE(N >= 2 && U >= 1 && r != nullptr && t != nullptr, "fcd_sample_T: bad arguments");
    FCD_REQUIRE(c0 >= 0 && C >= 0 && c0 + C <= (int64_t)N * (N - 1) / 2, "fcd_sample_T: edge shard outside N=%d", N);
    if (C == 0) return 0;
    sample_T_kernel<<<smp_grid(C * U), kSmpThreads, 0, (cudaStream_t)stream>>>(
        philox_make(seed, offset, VAR_T), r, U, eta, c0, C, t);
    return check_launch("fcd_sample_T");
}

int fcd_sample_F(uint64_t seed, uint64_t offset, int64_t c0, int64_t C,
                 const double* gamma3_host, uint8_t* f, void* stream) {
    FCD_REQUIRE(gamma3_host != nullptr && f != nullptr && C >= 0 && c0 >= 0, "fcd_sample_F: bad arguments");
    if (C == 0) return 0;
    sample_F_kernel<<<smp_grid(C), kSmpThreads, 0, (cudaStream_t)stream>>>(
        philox_make(seed, offset, VAR_F), c0, C, gamma3_host[0], gamma3_host[1], f);
    return check_launch("fcd_sample_F");
}

int fcd_sample_F_tilde(uint64_t seed, uint64_t offset, const uint8_t* f, const uint8_t* t,
                       int64_t c0, int64_t C, int32_t U, double epsilon, uint8_t* ft, void* stream) {
    FCD_REQUIRE(f != nullptr && t != nullptr && ft != nullptr && C >= 0 && U >= 1 && c0 >= 0,
                "fcd_sample_F_tilde: bad arguments");
    if (C == 0) return 0;
    sample_F_tilde_kernel<<<smp_grid(C * U), kSmpThreads, 0, (cudaStream_t)stream>>>(
        philox_make(seed, offset, VAR_FT), f, t, c0, C, U, epsilon, ft);
    return check_launch("fcd_sample_F_tilde");
}

int fcd_sample_B(uint64_t seed, uint64_t offset, const uint8_t* f, int64_t c0, int64_t C, int32_t H,
                 const double* mu3_host, const double* sigma3_host, double* b, void* stream) {
    FCD_REQUIRE(f != nullptr && b != nullptr && mu3_host != nullptr && sigma3_host != nullptr && C >= 0 && H >= 1,
                "fcd_sample_B: bad arguments");
    if (C == 0) return 0;
    MuSigma ms;
    for (int k = 0; k < 3; ++k) { ms.mu[k] = mu3_host[k]; ms.sigma[k] = sigma3_host[k]; }
    sample_B_kernel<false><<<smp_grid(C * H), kSmpThreads, 0, (cudaStream_t)stream>>>(
        philox_make(seed, offset, VAR_B), f, c0, C, H, ms, b);
    return check_launch("fcd_sample_B");
}

int fcd_sample_B_tilde(uint64_t seed, uint64_t offset, const uint8_t* ft, int64_t c0, int64_t C, int32_t U,
                       const double* mu3_host, const double* sigma3_host, double* bt, void* stream) {
    FCD_REQUIRE(ft != nullptr && bt != nullptr && mu3_host != nullptr && sigma3_host != nullptr && C >= 0 && U >= 1,
                "fcd_sample_B_tilde: bad arguments");
    if (C == 0) return 0;
    MuSigma ms;
    for (int k = 0; k < 3; ++k) { ms.mu[k] = mu3_host[k]; ms.sigma[k] = sigma3_host[k]; }
    sample_B_kernel<true><<<smp_grid(C * U), kSmpThreads, 0, (cudaStream_t)stream>>>(
        philox_make(seed, offset, VAR_BT), ft, c0, C, U, ms, bt);
    return check_launch("fcd_sample_B_tilde");
}

}  // extern "C"
