// One-shot all-reduce of a few doubles over NVLink peer memory.
//
// The sharded fit exchanges, per M-step / objective evaluation / energy, <= 8
// doubles of edge-local partial sums (fcdiff/fit.py:208-220, 270-286, 142-155
// become sums over edge shards).  Such an exchange is pure latency: through NCCL
// plus a device-to-host copy it costs ~0.1-0.25 ms at 8 GPUs, more than the
// evaluation kernel it follows.  Here every rank owns a small *window* in its
// device memory, exported to the other ranks of the box by CUDA IPC; ONE kernel
// (one CTA) per exchange
//   1. stores the rank's partial sums into its slot of EVERY rank's window
//      (peer stores over NVLink / NVSwitch), fences, and raises its flag there;
//   2. waits until the flags of all ranks have arrived in its own window;
//   3. adds the slots in rank order (every rank forms bit-identical sums, so the
//      host-side optimisers of all ranks take identical steps), writes the result
//      back to the device vector and to MAPPED pinned host memory, and raises a host
//      flag the CPU spins on -- no cudaMemcpy, no stream synchronisation.
// Two slot sets (sequence parity) make back-to-back exchanges safe: a rank can be
// at most one exchange ahead of the slowest one.  With world == 1 the kernel is
// just the device -> mapped-host publication of a reduction result.
#include <chrono>
#include <cstring>

#include "fcd_comm.cuh"

namespace fcd {

// result_host: [kCommMaxVals] doubles followed by one 64-bit flag (= seq, or seq | timeout bit).
// `seq` names the PUBLICATION (what the host waits for); the slot sets of the windows follow the
// rank's own exchange counter (fcd_comm.cuh).
// keep0, nkeep: this rank's OWN values vec[keep0 .. keep0 + nkeep) (before the sum) are published behind the
// n sums -- one publication then answers "what is the total" and "what is my share" (the record counts of the
// code pass: the total decides the kernel form on all ranks alike, the share sizes this rank's lists).
__global__ void __launch_bounds__(kCommMaxWorld* kCommMaxVals)
allreduce_small_kernel(double* __restrict__ vec, int n, int keep0, int nkeep, const __grid_constant__ CommPeers peers,
                       int rank, int world, unsigned long long seq, double* __restrict__ result_host,
                       long long spin_cycles) {
    __shared__ double s_vals[kCommMaxVals];
    const int t = threadIdx.x;
    if (t < n) s_vals[t] = vec[t];
    if (t < nkeep) {
        st_relaxed_sys(result_host + n + t, vec[keep0 + t]);
        __threadfence_system();
    }
    __syncthreads();
    const bool ok = comm_exchange_cta(s_vals, n, peers, rank, world, spin_cycles);
    if (t < n) {                                              // publish (a timed-out exchange leaves vec as it was)
        if (ok && world > 1) vec[t] = s_vals[t];
        st_relaxed_sys(result_host + t, s_vals[t]);
        __threadfence_system();
    }
    __syncthreads();
    if (t == 0)
        st_release_sys(reinterpret_cast<unsigned long long*>(result_host + kCommMaxVals),
                       ok ? seq : (seq | kCommTimeoutBit));
}

// Staging of the patient all-gather: the region posteriors are (N, U, 2) arrays and a
// rank owns the patient columns [u0, u0 + Ul): not contiguous.  pack: both arrays' own columns into
// one contiguous block [2][N][ch][2] (ch = columns per rank, zero padded); unpack: the gathered
// [world][2][N][ch][2] back into the two (N, U, 2) arrays.  One launch each instead of a chain of
// strided tensor copies.
__global__ void __launch_bounds__(256)
pack_patients_kernel(const double2* __restrict__ a, const double2* __restrict__ b, int N, int U, int u0, int Ul, int ch,
                     double2* __restrict__ out) {
    const int64_t total = 2ll * N * ch;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int j = (int)(i % ch);
        const int64_t r = i / ch;
        const int n = (int)(r % N), w = (int)(r / N);
        double2 v = make_double2(0.0, 0.0);
        if (j < Ul) v = (w ? b : a)[(int64_t)n * U + u0 + j];
        out[i] = v;
    }
}

__global__ void __launch_bounds__(256)
unpack_patients_kernel(const double2* __restrict__ g, int world, int N, int U, int ch, double2* __restrict__ a,
                       double2* __restrict__ b) {
    const int64_t total = 2ll * N * U;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int u = (int)(i % U);
        const int64_t r = i / U;
        const int n = (int)(r % N), w = (int)(r / N);
        const int rk = u / ch, j = u - rk * ch;
        (w ? b : a)[(int64_t)n * U + u] = g[(((int64_t)rk * 2 + w) * N + n) * ch + j];
    }
}

}  // namespace fcd

using namespace fcd;

extern "C" {

int64_t fcd_comm_window_bytes(void) { return (int64_t)sizeof(CommWindow); }
int32_t fcd_comm_handle_bytes(void) { return (int32_t)sizeof(cudaIpcMemHandle_t); }
int32_t fcd_comm_max_world(void) { return kCommMaxWorld; }
int32_t fcd_comm_max_vals(void) { return kCommMaxVals; }

#define FCD_CUDA(call, what)                                                   \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            ::fcd::set_error("%s: %s", what, cudaGetErrorString(e_));          \
            return -2;                                                         \
        }                                                                      \
    } while (0)

int fcd_comm_window_create(void** window_out_host) {
    FCD_REQUIRE(window_out_host != nullptr, "fcd_comm_window_create: NULL argument");
    void* p = nullptr;
    FCD_CUDA(cudaMalloc(&p, sizeof(CommWindow)), "fcd_comm_window_create(cudaMalloc)");
    FCD_CUDA(cudaMemset(p, 0, sizeof(CommWindow)), "fcd_comm_window_create(cudaMemset)");
    FCD_CUDA(cudaDeviceSynchronize(), "fcd_comm_window_create(sync)");
    *window_out_host = p;
    return 0;
}

int fcd_comm_window_destroy(void* window) {
    if (window != nullptr) FCD_CUDA(cudaFree(window), "fcd_comm_window_destroy");
    return 0;
}

int fcd_comm_window_export(void* window, void* handle_host) {
    FCD_REQUIRE(window != nullptr && handle_host != nullptr, "fcd_comm_window_export: NULL argument");
    cudaIpcMemHandle_t h;
    FCD_CUDA(cudaIpcGetMemHandle(&h, window), "fcd_comm_window_export(cudaIpcGetMemHandle)");
    memcpy(handle_host, &h, sizeof(h));
    return 0;
}

int fcd_comm_window_open(const void* handle_host, void** peer_window_out_host) {
    FCD_REQUIRE(handle_host != nullptr && peer_window_out_host != nullptr, "fcd_comm_window_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    void* p = nullptr;
    FCD_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "fcd_comm_window_open(cudaIpcOpenMemHandle)");
    *peer_window_out_host = p;
    return 0;
}

int fcd_comm_window_close(void* peer_window) {
    if (peer_window != nullptr) FCD_CUDA(cudaIpcCloseMemHandle(peer_window), "fcd_comm_window_close");
    return 0;
}

/* Mapped pinned host memory for results ([fcd_comm_max_vals()] doubles + one 64-bit flag), zero-filled. */
int fcd_host_result_alloc(void** result_out_host) {
    FCD_REQUIRE(result_out_host != nullptr, "fcd_host_result_alloc: NULL argument");
    void* p = nullptr;
    const size_t bytes = (kCommMaxVals + 1) * sizeof(double);
    FCD_CUDA(cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable), "fcd_host_result_alloc");
    memset(p, 0, bytes);
    *result_out_host = p;
    return 0;
}

int fcd_host_result_free(void* result_host) {
    if (result_host != nullptr) FCD_CUDA(cudaFreeHost(result_host), "fcd_host_result_free");
    return 0;
}

int fcd_allreduce_small(double* vec, int32_t n, void* const* windows_host, int32_t rank, int32_t world,
                        uint64_t seq, double* result_host, void* stream) {
    FCD_REQUIRE(vec != nullptr && result_host != nullptr, "fcd_allreduce_small: NULL argument");
    FCD_REQUIRE(n >= 1 && n <= kCommMaxVals && world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world &&
                seq >= 1 && (seq & kCommTimeoutBit) == 0, "fcd_allreduce_small: bad shape");
    CommPeers peers;
    FCD_REQUIRE(comm_peers_from_host(windows_host, world, peers), "fcd_allreduce_small: NULL window");
    double* result_dev = nullptr;
    FCD_CUDA(cudaHostGetDevicePointer((void**)&result_dev, result_host, 0), "fcd_allreduce_small(cudaHostGetDevicePointer)");
    const long long spin_cycles = kCommSpinCycles;
    allreduce_small_kernel<<<1, kCommMaxWorld * kCommMaxVals, 0, (cudaStream_t)stream>>>(
        vec, n, 0, 0, peers, rank, world, (unsigned long long)seq, result_dev, spin_cycles);
    return check_launch("fcd_allreduce_small");
}

int fcd_allreduce_small_keep(double* vec, int32_t n, int32_t keep0, int32_t nkeep, void* const* windows_host,
                             int32_t rank, int32_t world, uint64_t seq, double* result_host, void* stream) {
    FCD_REQUIRE(vec != nullptr && result_host != nullptr, "fcd_allreduce_small_keep: NULL argument");
    FCD_REQUIRE(n >= 1 && nkeep >= 0 && keep0 >= 0 && n + nkeep <= kCommMaxVals && keep0 + nkeep <= kCommMaxVals &&
                world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world && seq >= 1 &&
                (seq & kCommTimeoutBit) == 0, "fcd_allreduce_small_keep: bad shape");
    CommPeers peers;
    FCD_REQUIRE(comm_peers_from_host(windows_host, world, peers), "fcd_allreduce_small_keep: NULL window");
    double* result_dev = nullptr;
    FCD_CUDA(cudaHostGetDevicePointer((void**)&result_dev, result_host, 0),
             "fcd_allreduce_small_keep(cudaHostGetDevicePointer)");
    allreduce_small_kernel<<<1, kCommMaxWorld * kCommMaxVals, 0, (cudaStream_t)stream>>>(
        vec, n, keep0, nkeep, peers, rank, world, (unsigned long long)seq, result_dev, kCommSpinCycles);
    return check_launch("fcd_allreduce_small_keep");
}

int fcd_pack_patients(const double* lqR, const double* qR, int32_t N, int32_t U, int32_t u0, int32_t Ul, int32_t ch,
                      double* out, void* stream) {
    FCD_REQUIRE(lqR != nullptr && qR != nullptr && out != nullptr, "fcd_pack_patients: NULL argument");
    FCD_REQUIRE(N >= 1 && U >= 1 && ch >= 1 && u0 >= 0 && Ul >= 0 && Ul <= ch && u0 + Ul <= U,
                "fcd_pack_patients: bad shape");
    int64_t grid = (2ll * N * ch + 255) / 256;
    if (grid > 4ll * sm_count()) grid = 4ll * sm_count();
    pack_patients_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double2*>(lqR), reinterpret_cast<const double2*>(qR), N, U, u0, Ul, ch,
        reinterpret_cast<double2*>(out));
    return check_launch("fcd_pack_patients");
}

int fcd_unpack_patients(const double* gathered, int32_t world, int32_t N, int32_t U, int32_t ch, double* lqR,
                        double* qR, void* stream) {
    FCD_REQUIRE(gathered != nullptr && lqR != nullptr && qR != nullptr, "fcd_unpack_patients: NULL argument");
    FCD_REQUIRE(world >= 1 && N >= 1 && U >= 1 && ch >= 1 && (int64_t)world * ch >= U, "fcd_unpack_patients: bad shape");
    int64_t grid = (2ll * N * U + 255) / 256;
    if (grid > 4ll * sm_count()) grid = 4ll * sm_count();
    unpack_patients_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double2*>(gathered), world, N, U, ch, reinterpret_cast<double2*>(lqR),
        reinterpret_cast<double2*>(qR));
    return check_launch("fcd_unpack_patients");
}

/* Spins until the kernel of exchange `seq` has published its result; copies n doubles to out_host. */
int fcd_wait_result(const double* result_host, int32_t n, uint64_t seq, double* out_host, int32_t timeout_ms) {
    FCD_REQUIRE(result_host != nullptr && out_host != nullptr && n >= 1 && n <= kCommMaxVals,
                "fcd_wait_result: bad argument");
    const volatile unsigned long long* flag =
        reinterpret_cast<const volatile unsigned long long*>(result_host + kCommMaxVals);
    const auto t0 = std::chrono::steady_clock::now();
    unsigned long long f;
    long long spins = 0;
    while (((f = *flag) & ~kCommTimeoutBit) != seq) {
        if ((++spins & 0xfff) == 0) {
            const auto dt = std::chrono::steady_clock::now() - t0;
            if (std::chrono::duration_cast<std::chrono::milliseconds>(dt).count() > timeout_ms) {
                cudaError_t e = cudaGetLastError();
                set_error("fcd_wait_result: no result for exchange %llu after %d ms (%s)", (unsigned long long)seq,
                          timeout_ms, cudaGetErrorString(e));
                return -3;
            }
        }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    FCD_REQUIRE((f & kCommTimeoutBit) == 0, "fcd_wait_result: a peer rank did not arrive (device-side timeout)");
    const volatile double* r = result_host;
    for (int i = 0; i < n; ++i) out_host[i] = r[i];
    return 0;
}

}  // extern "C"
