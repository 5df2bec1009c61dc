// Read-stream probe for the E-step's memory skeleton (build + run on the GPU box):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/stream_probe scripts/stream_probe.cu && timeout 120 /tmp/stream_probe
// Streams two f64 planes P[k][c][u] (row pitch U) and a code plane (row pitch Q) the way estep_qF_coded_kernel does
// and in a few other shapes, with next to no arithmetic, to see which SHAPE of requests the memory system serves at
// what rate.  Bytes counted: 16 U + Q' per row (Q' = codes actually copied).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- V0/V1/V2: per-warp private rings (the kernel's shape).  SEG patients per stage, DEPTH stages, NW warps.
template <int SEG, int NW, bool CODES>
__global__ void __launch_bounds__(NW * 32, 1)
warp_ring_kernel(const double* __restrict__ P, int64_t planeStride, int64_t C, int U, const uint8_t* __restrict__ code,
                 int64_t Q, int depth, double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    constexpr int kStage = 2 * SEG * 8 + SEG;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* ring = s_dyn + (size_t)warp * depth * kStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_dyn + (size_t)NW * depth * kStage) + warp * depth;
    if (lane < depth) mbar_init(bars + lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int nseg = (U + SEG - 1) / SEG;
    const int64_t W = (int64_t)gridDim.x * NW, c_first = (int64_t)warp * gridDim.x + blockIdx.x;
    int64_t ic = c_first;
    int is = 0, id = 0;
    auto issue = [&]() {
        if (ic >= C) return;
        if (lane == 0) {
            const int u0 = is * SEG, np = min(SEG, U - u0), cb = (np + 15) & ~15;
            unsigned char* st = ring + id * kStage;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect(bars + id, 16 * np + (CODES ? cb : 0));
            tma_load_1d(st, P + ic * U + u0, np * 8, bars + id);
            tma_load_1d(st + SEG * 8, P + planeStride + ic * U + u0, np * 8, bars + id);
            if (CODES) tma_load_1d(st + 2 * SEG * 8, code + ic * Q + u0, cb, bars + id);
        }
        id = id + 1 == depth ? 0 : id + 1;
        if (++is == nseg) { is = 0; ic += W; }
    };
    for (int i = 0; i < depth; ++i) issue();
    double acc = 0.0;
    int d = 0;
    uint32_t phase = 0;
    for (int64_t c = c_first; c < C; c += W) {
        for (int s = 0; s < nseg; ++s) {
            mbar_wait(bars + d, phase);
            const unsigned char* st = ring + d * kStage;
            acc += *reinterpret_cast<const double*>(st + lane * 8) + *reinterpret_cast<const double*>(st + SEG * 8 + lane * 8);
            if (CODES) acc += (double)st[2 * SEG * 8 + lane];
            __syncwarp();
            issue();
            if (++d == depth) { d = 0; phase ^= 1; }
        }
    }
    if (acc == 123.456) out[0] = acc;
}

// ---- V3: CTA ring, G consecutive rows per stage (three LARGE copies per stage), one producer thread.
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1)
cta_ring_kernel(const double* __restrict__ P, int64_t planeStride, int64_t C, int U, const uint8_t* __restrict__ code,
                int64_t Q, int G, int depth, double* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rowB = U * 8;
    const int kStage = G * (2 * rowB + (int)Q);
    uint64_t* full = reinterpret_cast<uint64_t*>(s_dyn + (size_t)depth * kStage);
    uint64_t* empty = full + depth;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t ngroups = (C + G - 1) / G;
    // group g -> CTA g % grid
    int64_t ig = blockIdx.x;
    int id = 0;
    uint32_t iphase = 0;
    auto issue = [&](bool first) {
        if (ig >= ngroups) return;
        if (!first) mbar_wait(empty + id, iphase ^ 1);
        const int64_t c0 = ig * G;
        const int rows = (int)min((int64_t)G, C - c0);
        unsigned char* st = s_dyn + (size_t)id * kStage;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect(full + id, rows * (2 * rowB + (int)Q));
        tma_load_1d(st, P + c0 * U, rows * rowB, full + id);
        tma_load_1d(st + G * rowB, P + planeStride + c0 * U, rows * rowB, full + id);
        tma_load_1d(st + 2 * G * rowB, code + c0 * Q, rows * (int)Q, full + id);
        ig += gridDim.x;
        if (++id == depth) { id = 0; iphase ^= 1; }
    };
    if (threadIdx.x == 0)
        for (int i = 0; i < depth; ++i) issue(true);
    double acc = 0.0;
    int d = 0;
    uint32_t phase = 0;
    for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
        mbar_wait(full + d, phase);
        const unsigned char* st = s_dyn + (size_t)d * kStage;
        for (int r = warp; r < G; r += NW)
            acc += *reinterpret_cast<const double*>(st + r * rowB + lane * 8) +
                   *reinterpret_cast<const double*>(st + G * rowB + r * rowB + lane * 8) + (double)st[2 * G * rowB + r * Q + lane];
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + d);
        if (threadIdx.x == 0) issue(false);
        if (++d == depth) { d = 0; phase ^= 1; }
    }
    if (acc == 123.456) out[0] = acc;
}

// ---- V4: plain loads, warp per row, many warps
__global__ void __launch_bounds__(1024, 2)
ldg_kernel(const double* __restrict__ P, int64_t planeStride, int64_t C, int U, const uint8_t* __restrict__ code, int64_t Q,
           double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t W = (int64_t)gridDim.x * (blockDim.x >> 5), w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    double acc = 0.0;
    for (int64_t c = w0; c < C; c += W) {
        const double2* p0 = reinterpret_cast<const double2*>(P + c * U);
        const double2* p1 = reinterpret_cast<const double2*>(P + planeStride + c * U);
        const uint32_t* cd = reinterpret_cast<const uint32_t*>(code + c * Q);
        const int n2 = U / 2;
        double2 a[8], b[8];
        uint32_t k[4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = lane + 32 * j;
            a[j] = i < n2 ? __ldg(p0 + i) : make_double2(0, 0);
            b[j] = i < n2 ? __ldg(p1 + i) : make_double2(0, 0);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) k[j] = lane + 32 * j < U / 4 ? __ldg(cd + lane + 32 * j) : 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += a[j].x * b[j].y + a[j].y * b[j].x;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += (double)k[j];
    }
    if (acc == 123.456) out[0] = acc;
}

// fills with pseudo-random responsibilities in (0, 1) / codes in 0..5 (the planes of a real fit are not zeros)
__global__ void fill_kernel(double* P, int64_t n, uint8_t* code, int64_t nc) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t x = (uint64_t)i * 0x9E3779B97F4A7C15ull;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        P[i] = (double)(x >> 11) * (1.0 / 9007199254740992.0);
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nc; i += (int64_t)gridDim.x * blockDim.x)
        code[i] = (uint8_t)(((uint64_t)i * 0x9E3779B97F4A7C15ull >> 40) % 6);
}

template <typename F>
static float time_it(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e9f, tot = 0;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
        tot += ms;
    }
    CK(cudaGetLastError());
    printf("   best %.4f ms  mean %.4f ms", best, tot / reps);
    return best;
}

int main() {
    const int64_t C = 79800;
    const int U = 500;
    const int64_t Q = 512;
    const int64_t planeStride = C * U;
    double* P;
    uint8_t* code;
    double* out;
    CK(cudaMalloc(&P, 3 * planeStride * 8));
    CK(cudaMalloc(&code, C * Q));
    CK(cudaMalloc(&out, 64));
    CK(cudaMemset(P, 0, 3 * planeStride * 8));
    CK(cudaMemset(code, 0, C * Q));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
        printf("---- planes filled with pseudo-random values\n");
        fill_kernel<<<sms * 8, 256>>>(P, 3 * planeStride, code, C * Q);
        CK(cudaDeviceSynchronize());
    } else {
        printf("---- planes of zeros\n");
    }
    const double bytes2 = (double)C * (16.0 * U), bytesC = (double)C * 512.0;
    auto report = [&](const char* name, float ms, double bytes) { printf("  %-44s %.0f GB/s\n", name, bytes / (ms * 1e-3) / 1e9); };
    const int reps = 10;
#define RING(SEG, NW, CODES, DEPTH, NAME)                                                                       \
    do {                                                                                                        \
        const size_t smem = (size_t)NW * DEPTH * (2 * SEG * 8 + SEG + 8);                                        \
        if (smem <= 227 * 1024) {                                                                               \
            CK(cudaFuncSetAttribute(warp_ring_kernel<SEG, NW, CODES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            float ms = time_it([&] { warp_ring_kernel<SEG, NW, CODES><<<sms, NW * 32, smem>>>(P, planeStride, C, U, code, Q, DEPTH, out); }, reps); \
            report(NAME, ms, bytes2 + (CODES ? bytesC : 0));                                                    \
        }                                                                                                       \
    } while (0)
    RING(256, 16, true, 2, "warp ring seg256 nw16 depth2 +codes (K2 now)");
    RING(256, 16, true, 3, "warp ring seg256 nw16 depth3 +codes");
    RING(256, 16, false, 2, "warp ring seg256 nw16 depth2 no codes");
    RING(256, 16, false, 3, "warp ring seg256 nw16 depth3 no codes");
    RING(128, 16, true, 4, "warp ring seg128 nw16 depth4 +codes");
    RING(128, 16, true, 6, "warp ring seg128 nw16 depth6 +codes");
    RING(128, 32, true, 3, "warp ring seg128 nw32 depth3 +codes");
    RING(512, 16, true, 1, "warp ring seg512 nw16 depth1 +codes");
    RING(512, 8, true, 3, "warp ring seg512 nw8 depth3 +codes");
    RING(512, 12, true, 2, "warp ring seg512 nw12 depth2 +codes");
    RING(256, 8, true, 4, "warp ring seg256 nw8 depth4 +codes");
    RING(256, 8, true, 6, "warp ring seg256 nw8 depth6 +codes");
    for (int G : {4, 8, 16}) {
        for (int depth : {2, 3, 4}) {
            const size_t smem = (size_t)depth * G * (2 * U * 8 + Q) + 2 * depth * 8;
            if (smem > 227 * 1024) continue;
            CK(cudaFuncSetAttribute(cta_ring_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            float ms = time_it([&] { cta_ring_kernel<16><<<sms, 512, smem>>>(P, planeStride, C, U, code, Q, G, depth, out); }, reps);
            char nm[96];
            snprintf(nm, sizeof nm, "CTA ring G=%d rows depth %d (%zu KB)", G, depth, smem / 1024);
            report(nm, ms, bytes2 + bytesC);
        }
    }
    {
        float ms = time_it([&] { ldg_kernel<<<sms * 2, 1024>>>(P, planeStride, C, U, code, Q, out); }, reps);
        report("plain loads, warp per row, 64 warps/SM", ms, bytes2 + (double)C * 500);
    }
    {
        double* D;
        CK(cudaMalloc(&D, 2 * planeStride * 8));
        float ms = time_it([&] { CK(cudaMemcpyAsync(D, P, 2 * planeStride * 8, cudaMemcpyDeviceToDevice)); }, reps);
        report("cudaMemcpy D2D of the two planes (rd+wr)", ms, 2.0 * 2 * planeStride * 8);
    }
    }
    return 0;
}