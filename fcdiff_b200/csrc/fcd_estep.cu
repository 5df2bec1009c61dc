// E-step kernels: healthy sufficient statistics, K2 (template posterior q_F),
// K2b (region-weight tensor W and the Gauss-Seidel sweep for q_R).
#include "fcd_common.cuh"

namespace fcd {

constexpr int kEdgeThreads = 256;       // 8 warps, one edge row per warp at a time

// ------------------------------------------------------------------ c_to_nm
__global__ void c_to_nm_kernel(int64_t c0, int64_t C, int32_t* n_out, int32_t* m_out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < C;
         i += (int64_t)gridDim.x * blockDim.x) {
        int n, m;
        c_to_nm(c0 + i, n, m);
        n_out[i] = n;
        m_out[i] = m;
    }
}

// ------------------------------------------------------------ healthy stats
// S1[c] = sum_h b[c,h], S2[c] = sum_h b[c,h]^2.  One warp per edge row,
// coalesced 64-bit loads along the subject axis.
__global__ void __launch_bounds__(kEdgeThreads)
healthy_stats_kernel(const double* __restrict__ b, int64_t C, int H, int64_t pitchH,
                     double* __restrict__ S1, double* __restrict__ S2) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t c = warp0; c < C; c += nwarps) {
        const double* row = b + c * pitchH;
        double s1 = 0.0, s2 = 0.0;
        for (int h = lane; h < H; h += 32) {
            double x = ldg_stream1(row + h);
            s1 += x;
            s2 = fma(x, x, s2);
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) {
            S1[c] = s1;
            S2[c] = s2;
        }
    }
}

// ------------------------------------------------------------ Gaussian cache
// Ea / Eb / Tm planes (see fcd_common.cuh) from the patient correlations; the
// only place where exponentials of the data are taken.  Elementwise over the
// padded [C][pitchU] rows (padding columns are written as zeros).
__global__ void __launch_bounds__(256)
gauss_cache_kernel(const double* __restrict__ bt, int64_t C, int U, int64_t pitchU,
                   const __grid_constant__ ThetaDev th,
                   double* __restrict__ Ea, double* __restrict__ Eb, double* __restrict__ Tm) {
    const int64_t total = C * pitchU;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int u = (int)(i % pitchU);
        GaussElem g;
        g.ea = g.ebc = g.tmax = 0.0;
        if (u < U) g = gauss_eval(ldg_stream1(bt + i), th);
        Ea[i] = g.ea;
        Eb[i] = g.ebc;
        if (Tm) Tm[i] = g.tmax;
    }
}

// ------------------------------------------------------------------- K2
// lqF[c,k] = log gamma_k + healthy_k(S1,S2) + sum_u sum_l w_l log M_kl(bt[c,u])
//            - logsumexp_k                                 (fcdiff/fit.py:157-174)
// The per-(c,u) term (tmax - log sqrt(2 pi)) * sum_l w_l is common to the three
// states k and cancels in the normalisation, so it is never formed.
template <bool FAST>
__device__ __forceinline__ void k2_elem(double ea, double ebc, double2 qn, double2 qm, const ThetaDev& th,
                                        const double* s_tab, double (&acc)[3]) {
    double w[3];
    pair_weights(qn, qm, w);
    const ElemM r = elem_from_cache(ea, ebc, th);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double a = acc[k];
#pragma unroll
        for (int l = 0; l < 3; ++l) a = fma(w[l], fast_log<FAST>(elem_Mp(r, th, k, l), s_tab), a);
        acc[k] = a;
    }
}

// lane 0 of the warp that owns edge c: log gamma + healthy quadratic + A, then
// scipy.special.logsumexp: a_max + log(sum exp(a - a_max))           (fit.py:165-174)
__device__ __forceinline__ void k2_finish(int64_t c, const double (&A)[3], double s1, double s2,
                                          const ThetaDev& th, double* __restrict__ lqF,
                                          double* __restrict__ qF) {
    double l[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        l[k] = th.log_gamma[k] + fma(th.hq_a[k], s2, fma(th.hq_b[k], s1, th.hq_c[k])) + A[k];
    const double mx = fmax(l[0], fmax(l[1], l[2]));
    const double lse = mx + log(exp(l[0] - mx) + exp(l[1] - mx) + exp(l[2] - mx));
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double v = l[k] - lse;
        lqF[c * 3 + k] = v;
        if (qF) qF[c * 3 + k] = exp(v);
    }
}

template <bool VEC2, bool FAST>
__global__ void __launch_bounds__(kEdgeThreads)
estep_qF_kernel(const double* __restrict__ S1, const double* __restrict__ S2,
                const double* __restrict__ Ea, const double* __restrict__ Eb,
                int64_t C, int U, int64_t pitchU,
                const double* __restrict__ qR, int N, int64_t c0,
                const __grid_constant__ ThetaDev th, const double* __restrict__ g_tab,
                double* __restrict__ lqF, double* __restrict__ qF) {
    extern __shared__ double s_tab[];               // kLogTabBytes when FAST, unused otherwise
    load_log_table<FAST>(g_tab, s_tab);
    const int lane = threadIdx.x & 31;
    double acc[3] = {0.0, 0.0, 0.0};
    auto row_end = [&](int64_t c) {
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[k] = warp_sum(acc[k]);
        if (lane == 0) k2_finish(c, acc, S1[c], S2[c], th, lqF, qF);
        acc[0] = acc[1] = acc[2] = 0.0;
    };
    if (VEC2) {
        walk_rows<false>(Ea, Eb, nullptr, C, U, pitchU, qR, c0,
                         [&](double ea, double ebc, double, double2 qn, double2 qm) {
                             k2_elem<FAST>(ea, ebc, qn, qm, th, s_tab, acc);
                         },
                         row_end);
    } else {
        const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
        const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
        const double2* qR2 = reinterpret_cast<const double2*>(qR);
        for (int64_t c = warp0; c < C; c += nwarps) {
            int n, m;
            c_to_nm(c0 + c, n, m);
            const double2* qn = qR2 + (int64_t)n * U;
            const double2* qm = qR2 + (int64_t)m * U;
            for (int u = lane; u < U; u += 32)
                k2_elem<FAST>(ldg_stream1(Ea + c * pitchU + u), ldg_stream1(Eb + c * pitchU + u),
                              __ldg(qn + u), __ldg(qm + u), th, s_tab, acc);
            row_end(c);
        }
    }
}

// K2 from per-edge sums A[c][k] = sum_u sum_l w_l log Mp_kl that a K3b pass has
// already produced with the same q_R and (eta, epsilon): no pass over the data.
__global__ void __launch_bounds__(256)
estep_qF_finish_kernel(const double* __restrict__ S1, const double* __restrict__ S2,
                       const double* __restrict__ A, int64_t C,
                       const __grid_constant__ ThetaDev th,
                       double* __restrict__ lqF, double* __restrict__ qF) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < C;
         c += (int64_t)gridDim.x * blockDim.x) {
        const double a[3] = {A[c * 3], A[c * 3 + 1], A[c * 3 + 2]};
        k2_finish(c, a, S1[c], S2[c], th, lqF, qF);
    }
}

// ------------------------------------------------------ patient-major copy
// btT[u - u0][c] = bt[c][u]; 32x32 tiles through padded shared memory so both
// sides are coalesced.
__global__ void __launch_bounds__(256)
transpose_patients_kernel(const double* __restrict__ bt, int64_t C, int U, int64_t pitchU,
                          int u0, int Ul, double* __restrict__ btT, int64_t pitchC) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    const int64_t cb = (int64_t)blockIdx.x * 32;
    const int ub = blockIdx.y * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int64_t c = cb + ty + j;
        const int u = ub + tx;
        if (c < C && u < Ul) tile[ty + j][tx] = bt[c * pitchU + u0 + u];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int u = ub + ty + j;
        const int64_t c = cb + tx;
        if (c < C && u < Ul) btT[(int64_t)u * pitchC + c] = tile[tx][ty + j];
    }
}

// ------------------------------------------------------------------- K2b/W
// WT[u][c][l] = sum_k qF[c,k] log Mp_kl(u, c) from the patient-major cache planes.
// The omitted per-(c,u) constant (tmax - log sqrt(2 pi)) sum_k qF[c,k] is the
// same for l = 0, 1, 2 and enters both states of fcdiff/fit.py:190,194 multiplied
// by (q_R[m,u,0] + q_R[m,u,1]), so it cancels at fit.py:196.
template <bool FAST>
__global__ void __launch_bounds__(256)
region_weights_kernel(const double* __restrict__ EaT, const double* __restrict__ EbT,
                      int Ul, int64_t C, int64_t pitchC,
                      const double* __restrict__ qF, const __grid_constant__ ThetaDev th,
                      const double* __restrict__ g_tab, double* __restrict__ WT) {
    extern __shared__ double s_tab[];               // kLogTabBytes when FAST, unused otherwise
    load_log_table<FAST>(g_tab, s_tab);
    // persistent CTAs over (patient, 1024-edge tile) work items: no wave tail, one
    // table load per CTA
    const int64_t tiles_per_row = (C + 1023) / 1024;
    const int64_t ntiles = tiles_per_row * Ul;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int u = (int)(t / tiles_per_row);
        const int64_t cbase = (t - (int64_t)u * tiles_per_row) * 1024;
        const double* ra = EaT + (int64_t)u * pitchC;
        const double* rb = EbT + (int64_t)u * pitchC;
        double* out = WT + (int64_t)u * C * 3;
        double ea[4], eb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {                 // issue the tile's loads first
            const int64_t c = cbase + threadIdx.x + 256 * j;
            ea[j] = eb[j] = 0.0;
            if (c < C) {
                ea[j] = ldg_stream1(ra + c);
                eb[j] = ldg_stream1(rb + c);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t c = cbase + threadIdx.x + 256 * j;
            if (c >= C) continue;
            const ElemM r = elem_from_cache(ea[j], eb[j], th);
            const double q0 = __ldg(qF + c * 3), q1 = __ldg(qF + c * 3 + 1), q2 = __ldg(qF + c * 3 + 2);
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                double w = q0 * fast_log<FAST>(elem_Mp(r, th, 0, l), s_tab);
                w = fma(q1, fast_log<FAST>(elem_Mp(r, th, 1, l), s_tab), w);
                w = fma(q2, fast_log<FAST>(elem_Mp(r, th, 2, l), s_tab), w);
                out[c * 3 + l] = w;
            }
        }
    }
}

// --------------------------------------------------------------- K2b/sweep
// Gauss-Seidel sweep of fcdiff/fit.py:184-197.  One CTA per patient; thread t
// owns the regions m = t, t+T, ... and keeps their q_R in registers.  Step n:
// every thread forms its partial sums over its m != n from the WT window that
// was prefetched during step n-1 (the loads do not depend on q_R), one warp
// shuffle reduction, ONE block barrier; then only the owner of region n
// finishes the reduction, normalises (fit.py:196) and updates its register
// copy of q_R[n] (fit.py:197) while the other threads already run step n+1.
// The two-slot s_red buffer makes the single barrier per step sufficient.
template <int T, int MPT>
__global__ void __launch_bounds__(T)
sweep_kernel(const double* __restrict__ WT, int64_t C, int N, int U, int u0,
             int lookup, double lp0, double lp1,
             double* __restrict__ qR, double* __restrict__ lqR) {
    constexpr int NW = T / 32;
    __shared__ double s_red[2][NW][2];
    const int ul = blockIdx.x;
    const int u = u0 + ul;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* Wu = WT + (int64_t)ul * C * 3;
    double q0[MPT], q1[MPT], w[MPT][3], wn[MPT][3];
#pragma unroll
    for (int j = 0; j < MPT; ++j) {
        const int m = tid + j * T;
        q0[j] = q1[j] = 0.0;
        if (m < N) {
            q0[j] = qR[((int64_t)m * U + u) * 2];
            q1[j] = qR[((int64_t)m * U + u) * 2 + 1];
        }
    }
    auto load_w = [&](int n, double (&dst)[MPT][3]) {
        const int64_t base = (int64_t)n * (n - 1) / 2;
#pragma unroll
        for (int j = 0; j < MPT; ++j) {
            const int m = tid + j * T;
            dst[j][0] = dst[j][1] = dst[j][2] = 0.0;
            if (m < N && m != n && n < N) {
                const int64_t c = (lookup == FCD_LOOKUP_REFERENCE || m < n)
                                      ? base + m
                                      : (int64_t)m * (m - 1) / 2 + n;
                dst[j][0] = __ldg(Wu + c * 3);
                dst[j][1] = __ldg(Wu + c * 3 + 1);
                dst[j][2] = __ldg(Wu + c * 3 + 2);
            }
        }
    };
    load_w(0, w);
    for (int n = 0; n < N; ++n) {
        load_w(n + 1, wn);                            // prefetch the next window
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int j = 0; j < MPT; ++j) {               // zero weights for m == n / m >= N
            s0 += fma(q0[j], w[j][0], q1[j] * w[j][2]);       // fit.py:188-190
            s1 += fma(q1[j], w[j][1], q0[j] * w[j][2]);       // fit.py:192-194
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        if (lane == 0) {
            s_red[n & 1][warp][0] = s0;
            s_red[n & 1][warp][1] = s1;
        }
        __syncthreads();
        if (tid == n % T) {                           // owner of region n
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                a += s_red[n & 1][i][0];
                b += s_red[n & 1][i][1];
            }
            // lq = l - logsumexp(l), q = exp(lq) (fit.py:196-197) with one exponential:
            // with d = -|l0 - l1| and t = exp(d):  lse = max + log1p(t),
            // q_max = 1 / (1 + t), q_min = t / (1 + t).
            const double l0r = lp0 + a, l1r = lp1 + b;
            const bool first = l0r >= l1r;
            const double d = first ? l1r - l0r : l0r - l1r;
            const double t = exp_nonpos(d);
            const double lg = log1p(t);
            const double inv = 1.0 / (1.0 + t);
            const double lmax = -lg, lmin = d - lg;
            const double qmax = inv, qmin = t * inv;
            const double l0 = first ? lmax : lmin, l1 = first ? lmin : lmax;
            const double p0 = first ? qmax : qmin, p1 = first ? qmin : qmax;
            const int slot = n / T;
#pragma unroll
            for (int j = 0; j < MPT; ++j)
                if (j == slot) {
                    q0[j] = p0;
                    q1[j] = p1;
                }
            const int64_t o = ((int64_t)n * U + u) * 2;
            lqR[o] = l0;
            lqR[o + 1] = l1;
            qR[o] = p0;
            qR[o + 1] = p1;
        }
#pragma unroll
        for (int j = 0; j < MPT; ++j) {
            w[j][0] = wn[j][0];
            w[j][1] = wn[j][1];
            w[j][2] = wn[j][2];
        }
    }
}

static inline int grid_for_rows(int64_t rows, int rows_per_block, int waves) {
    int64_t need = (rows + rows_per_block - 1) / rows_per_block;
    int64_t cap = (int64_t)sm_count() * waves;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace fcd

using namespace fcd;

extern "C" {

int fcd_c_to_nm(int64_t c0, int64_t C, int32_t* n_out, int32_t* m_out, void* stream) {
    FCD_REQUIRE(C >= 0 && c0 >= 0, "fcd_c_to_nm: bad range");
    if (C == 0) return 0;
    c_to_nm_kernel<<<grid_for_rows(C, 256, 8), 256, 0, (cudaStream_t)stream>>>(c0, C, n_out, m_out);
    return check_launch("fcd_c_to_nm");
}

int fcd_healthy_stats(const double* b, int64_t C, int32_t H, int64_t pitchH,
                      double* S1, double* S2, void* stream) {
    FCD_REQUIRE(C >= 0 && H >= 1 && pitchH >= H, "fcd_healthy_stats: bad shape C=%lld H=%d pitch=%lld",
                (long long)C, H, (long long)pitchH);
    if (C == 0) return 0;
    healthy_stats_kernel<<<grid_for_rows(C, kEdgeThreads / 32, 8), kEdgeThreads, 0, (cudaStream_t)stream>>>(
        b, C, H, pitchH, S1, S2);
    return check_launch("fcd_healthy_stats");
}

int fcd_gauss_cache(const double* bt, int64_t C, int32_t U, int64_t pitchU,
                    const fcd_theta* theta_host, double* Ea, double* Eb, double* Tm, void* stream) {
    FCD_REQUIRE(theta_host != nullptr && bt != nullptr && Ea != nullptr && Eb != nullptr,
                "fcd_gauss_cache: NULL argument");
    FCD_REQUIRE(C >= 0 && U >= 1 && pitchU >= U, "fcd_gauss_cache: bad shape");
    if (C == 0) return 0;
    const ThetaDev th = make_theta_dev(*theta_host, 0);
    gauss_cache_kernel<<<grid_for_rows(C * pitchU, 256, 16), 256, 0, (cudaStream_t)stream>>>(
        bt, C, U, pitchU, th, Ea, Eb, Tm);
    return check_launch("fcd_gauss_cache");
}

int fcd_estep_qF(const double* S1, const double* S2, int32_t H,
                 const double* Ea, const double* Eb, int64_t C, int32_t U, int64_t pitchU,
                 const double* qR, int32_t N, int64_t c0,
                 const fcd_theta* theta_host, double* lqF, double* qF, void* stream) {
    FCD_REQUIRE(theta_host != nullptr, "fcd_estep_qF: theta is NULL");
    FCD_REQUIRE(C >= 0 && U >= 1 && pitchU >= U && N >= 2, "fcd_estep_qF: bad shape");
    FCD_REQUIRE(c0 >= 0 && c0 + C <= (int64_t)N * (N - 1) / 2, "fcd_estep_qF: edge shard [%lld, %lld) outside N=%d",
                (long long)c0, (long long)(c0 + C), N);
    if (C == 0) return 0;
    const ThetaDev th = make_theta_dev(*theta_host, H);
    const double* tab = log_table((cudaStream_t)stream);
    FCD_REQUIRE(tab != nullptr, "fcd_estep_qF: log table initialisation failed");
    const int grid = grid_for_rows(C, kEdgeThreads / 32, 2);
    const bool vec2 = (pitchU % 2 == 0) && (((reinterpret_cast<uintptr_t>(Ea) | reinterpret_cast<uintptr_t>(Eb)) & 15) == 0);
    const bool fast = log_table_covers(th.epsl, th.al);
#define FCD_K2(V, F)                                                                              \
    do {                                                                                          \
        cudaFuncSetAttribute(estep_qF_kernel<V, F>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             (int)kLogTabBytes);                                                  \
        estep_qF_kernel<V, F><<<grid, kEdgeThreads, (F) ? kLogTabBytes : 0, (cudaStream_t)stream>>>( \
            S1, S2, Ea, Eb, C, U, pitchU, qR, N, c0, th, tab, lqF, qF);                            \
    } while (0)
    if (vec2) { if (fast) FCD_K2(true, true); else FCD_K2(true, false); }
    else      { if (fast) FCD_K2(false, true); else FCD_K2(false, false); }
#undef FCD_K2
    return check_launch("fcd_estep_qF");
}

int fcd_estep_qF_finish(const double* S1, const double* S2, int32_t H, const double* A, int64_t C,
                        const fcd_theta* theta_host, double* lqF, double* qF, void* stream) {
    FCD_REQUIRE(theta_host != nullptr && A != nullptr && lqF != nullptr, "fcd_estep_qF_finish: NULL argument");
    FCD_REQUIRE(C >= 0 && H >= 1, "fcd_estep_qF_finish: bad shape");
    if (C == 0) return 0;
    const ThetaDev th = make_theta_dev(*theta_host, H);
    estep_qF_finish_kernel<<<grid_for_rows(C, 256, 8), 256, 0, (cudaStream_t)stream>>>(S1, S2, A, C, th, lqF, qF);
    return check_launch("fcd_estep_qF_finish");
}

int fcd_transpose_patients(const double* bt, int64_t C, int32_t U, int64_t pitchU,
                           int32_t u0, int32_t Ul, double* btT, int64_t pitchC, void* stream) {
    FCD_REQUIRE(C >= 0 && U >= 1 && pitchU >= U && u0 >= 0 && Ul >= 0 && u0 + Ul <= U && pitchC >= C,
                "fcd_transpose_patients: bad shape");
    if (C == 0 || Ul == 0) return 0;
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)((Ul + 31) / 32));
    transpose_patients_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(bt, C, U, pitchU, u0, Ul, btT, pitchC);
    return check_launch("fcd_transpose_patients");
}

int fcd_region_weights(const double* EaT, const double* EbT, int32_t Ul, int64_t C, int64_t pitchC,
                       const double* qF, const fcd_theta* theta_host, double* WT, void* stream) {
    FCD_REQUIRE(theta_host != nullptr, "fcd_region_weights: theta is NULL");
    FCD_REQUIRE(C >= 0 && Ul >= 0 && pitchC >= C, "fcd_region_weights: bad shape");
    if (C == 0 || Ul == 0) return 0;
    const ThetaDev th = make_theta_dev(*theta_host, 0);
    const double* tab = log_table((cudaStream_t)stream);
    FCD_REQUIRE(tab != nullptr, "fcd_region_weights: log table initialisation failed");
    int64_t ntiles = ((C + 1023) / 1024) * (int64_t)Ul;
    int64_t grid = (int64_t)sm_count() * 2;            // 2 CTAs / SM resident (86 KB table each)
    if (grid > ntiles) grid = ntiles;
    if (log_table_covers(th.epsl, th.al)) {
        cudaFuncSetAttribute(region_weights_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kLogTabBytes);
        region_weights_kernel<true><<<(unsigned)grid, 256, kLogTabBytes, (cudaStream_t)stream>>>(
            EaT, EbT, Ul, C, pitchC, qF, th, tab, WT);
    } else {
        region_weights_kernel<false><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
            EaT, EbT, Ul, C, pitchC, qF, th, tab, WT);
    }
    return check_launch("fcd_region_weights");
}

int fcd_estep_qR(const double* WT, int64_t C, int32_t N, int32_t U, int32_t u0, int32_t Ul,
                 const double* log_pi2_host, int32_t edge_lookup,
                 double* qR, double* lqR, void* stream) {
    FCD_REQUIRE(log_pi2_host != nullptr, "fcd_estep_qR: log_pi2 is NULL");
    FCD_REQUIRE(N >= 2 && C == (int64_t)N * (N - 1) / 2, "fcd_estep_qR: C=%lld is not N(N-1)/2 for N=%d",
                (long long)C, N);
    FCD_REQUIRE(edge_lookup == FCD_LOOKUP_REFERENCE || edge_lookup == FCD_LOOKUP_SYMMETRIC,
                "fcd_estep_qR: unknown edge_lookup %d", edge_lookup);
    FCD_REQUIRE(!(edge_lookup == FCD_LOOKUP_REFERENCE && N < 3),
                "fcd_estep_qR: edge_lookup='reference' indexes out of bounds for N < 3 (fit.py:186)");
    FCD_REQUIRE(u0 >= 0 && Ul >= 0 && u0 + Ul <= U, "fcd_estep_qR: bad patient range");
    if (Ul == 0) return 0;
    FCD_REQUIRE(N <= 8192, "fcd_estep_qR: N=%d exceeds the 8192 regions the sweep kernel supports", N);
    cudaStream_t st = (cudaStream_t)stream;
    const double lp0 = log_pi2_host[0], lp1 = log_pi2_host[1];
#define FCD_SWEEP(T, M)                                                                            \
    sweep_kernel<T, M><<<Ul, T, 0, st>>>(WT, C, N, U, u0, edge_lookup, lp0, lp1, qR, lqR)
    if (N <= 32) FCD_SWEEP(32, 1);
    else if (N <= 64) FCD_SWEEP(32, 2);
    else if (N <= 128) FCD_SWEEP(64, 2);
    else if (N <= 256) FCD_SWEEP(128, 2);
    else if (N <= 512) FCD_SWEEP(128, 4);
    else if (N <= 1024) FCD_SWEEP(256, 4);
    else if (N <= 2048) FCD_SWEEP(256, 8);
    else if (N <= 4096) FCD_SWEEP(512, 8);
    else FCD_SWEEP(1024, 8);
#undef FCD_SWEEP
    return check_launch("fcd_estep_qR");
}

}  // extern "C"
