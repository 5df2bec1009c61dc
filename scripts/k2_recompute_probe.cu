// What the E-step would cost on the 8-byte INPUT instead of the two responsibility planes (DESIGN.md "K2 ceiling"):
// every element's three Gaussians are recomputed from x (three exponentials with the library's exp_nonpos), and the
// row's running products are taken over  a_l S + b_l g_k  (S = sum_j g_j; the common factor 1/S cancels in the
// normalisation of fcdiff/fit.py:174) -- against streaming p_0, p_1 and forming  a_l + b_l p_k.  Both kernels read
// one code byte per element, keep three running products per lane and flush them rarely; no records, no logs.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Ifcdiff_b200/csrc -Iinclude -o /tmp/k2rp scripts/k2_recompute_probe.cu && /tmp/k2rp
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "fcd_math.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Th {
    double mu[3], isig[3], a[4], b[4];
};

__global__ void fill_kernel(double* x, double* p0, double* p1, uint8_t* code, int64_t n, Th th) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
        const double xv = -0.25 + 0.7 * u;
        x[i] = xv;
        double g[3], s = 0.0;
        for (int k = 0; k < 3; ++k) {
            const double z = (xv - th.mu[k]) * th.isig[k];
            g[k] = exp(-0.5 * z * z) * th.isig[k];
            s += g[k];
        }
        p0[i] = g[0] / s;
        p1[i] = g[1] / s;
        code[i] = (uint8_t)((h >> 7) % 3);
    }
}

// warp per row, 128-bit loads, three running products per lane, one flush (product -> log -> sum) per 16 elements
template <bool RECOMPUTE>
__global__ void __launch_bounds__(512)
probe_kernel(const double* __restrict__ x, const double* __restrict__ p0, const double* __restrict__ p1,
             const uint8_t* __restrict__ code, int64_t C, int U, const __grid_constant__ Th th, double* __restrict__ out) {
    __shared__ double2 s_ab[4];
    if (threadIdx.x < 4) s_ab[threadIdx.x] = make_double2(th.a[threadIdx.x], th.b[threadIdx.x]);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t W = (int64_t)gridDim.x * (blockDim.x >> 5);
    double acc[3] = {0.0, 0.0, 0.0};
    for (int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < C; c += W) {
        double pr[3] = {1.0, 1.0, 1.0};
        for (int u = 2 * lane; u < U; u += 64) {
            const int64_t i = c * U + u;
            const uint32_t c2 = *reinterpret_cast<const unsigned short*>(code + i);
            double v0[2], v1[2];
            if (RECOMPUTE) {
                const double2 xx = *reinterpret_cast<const double2*>(x + i);
                v0[0] = xx.x; v0[1] = xx.y;
            } else {
                const double2 a0 = *reinterpret_cast<const double2*>(p0 + i), a1 = *reinterpret_cast<const double2*>(p1 + i);
                v0[0] = a0.x; v0[1] = a0.y; v1[0] = a1.x; v1[1] = a1.y;
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double2 k = s_ab[(c2 >> (8 * e)) & 3];
                if (RECOMPUTE) {
                    double g[3], s = 0.0;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const double z = (v0[e] - th.mu[j]) * th.isig[j];
                        g[j] = fcd::exp_nonpos(-0.5 * z * z) * th.isig[j];
                        s += g[j];
                    }
                    const double as = k.x * s;
#pragma unroll
                    for (int j = 0; j < 3; ++j) pr[j] *= fma(k.y, g[j], as);
                } else {
                    const double p3[3] = {v0[e], v1[e], (1.0 - v0[e]) - v1[e]};
#pragma unroll
                    for (int j = 0; j < 3; ++j) pr[j] *= fma(k.y, p3[j], k.x);
                }
            }
            if (RECOMPUTE && (u & 255) == 2 * lane) {          // keep the products in range: the Gaussians are not normalised
#pragma unroll
                for (int j = 0; j < 3; ++j) { acc[j] += log(pr[j]); pr[j] = 1.0; }
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[j] += log(pr[j]);
    }
    if (acc[0] + acc[1] + acc[2] == 123.456) out[0] = acc[0];
}

int main() {
    const int64_t C = 79800;
    const int U = 500;
    const int64_t n = C * U;
    double *x, *p0, *p1, *out;
    uint8_t* code;
    CK(cudaMalloc(&x, n * 8)); CK(cudaMalloc(&p0, n * 8)); CK(cudaMalloc(&p1, n * 8)); CK(cudaMalloc(&code, n + 64)); CK(cudaMalloc(&out, 64));
    Th th;
    const double mu[3] = {-0.15, 0.0, 0.3}, sg[3] = {0.025, 0.035, 0.05}, eps[3] = {0.97, 0.03, 0.7};
    for (int k = 0; k < 3; ++k) { th.mu[k] = mu[k]; th.isig[k] = 1.0 / sg[k]; th.a[k] = (1.0 - eps[k]) / 2; th.b[k] = eps[k] - th.a[k]; }
    th.a[3] = 1.0; th.b[3] = 0.0;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    fill_kernel<<<sms * 8, 256>>>(x, p0, p1, code, n, th);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int variant = 0; variant < 2; ++variant) {
        for (int warps : {16, 32}) {
            float best = 1e9f;
            for (int r = 0; r < 8; ++r) {
                CK(cudaEventRecord(e0));
                if (variant == 0) probe_kernel<false><<<sms * (warps / 16), 512>>>(x, p0, p1, code, C, U, th, out);
                else probe_kernel<true><<<sms * (warps / 16), 512>>>(x, p0, p1, code, C, U, th, out);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (r > 1 && ms < best) best = ms;
            }
            CK(cudaGetLastError());
            printf("%-44s %2d warps/SM: %.4f ms  (%.0f GB/s of the bytes it reads)\n",
                   variant == 0 ? "two planes p_0, p_1 (16 B + 1 B per element)" : "recompute from x (8 B + 1 B, three exponentials)",
                   warps, best, (variant == 0 ? 17.0 : 9.0) * n / (best * 1e-3) / 1e9);
        }
    }
    return 0;
}
