// K1 tensor-core path: per-subject Gram matrices R = Z Z^T on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by
// TMA with the 128-byte swizzle), error-compensated split-TF32:
//     Z = Zh + Zl (both exactly representable in TF32),
//     R ~= Zh Zh^T + Zh Zl^T + Zl Zh^T          (3 MMAs per product; Zl Zl^T ~ 2^-22 dropped)
//
// One CTA computes one 128x128 tile (I >= J) of FOUR consecutive subjects: the
// four fp32 accumulators fill the CTA's 512 TMEM columns, so that the epilogue
// holds, for every edge (n, m) of the tile, four subject-adjacent values and
// writes one full 32-byte sector of the edge-major (C, S) output (the (C, S)
// layout is what the EM kernels stream; a per-subject epilogue would write
// 8 bytes per sector).
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM
// allocator + MMA issuer (one lane), warps 2..5 = epilogue (TMEM -> registers ->
// clip / atanh in fp64 -> global).  3-stage smem ring of {A_hi, A_lo, B_hi, B_lo}
// 128 x 32-float tiles (64 KB per stage), full/empty mbarriers, tcgen05.commit
// releases a stage when the MMAs that read it have completed.
#include <cuda.h>

#include "fcd_corr.cuh"

namespace fcd {

constexpr int kTile = 128;                       // tile rows of A and of B
constexpr int kStages = 3;
constexpr int kSubj = 4;                         // subjects per CTA (4 x 128 TMEM columns)
constexpr uint32_t kTileBytes = kTile * kKChunk * 4;      // 16 KB
constexpr int kTcThreads = 192;

struct TcSmem {
    float a_hi[kStages][kTile * kKChunk];
    float a_lo[kStages][kTile * kKChunk];
    float b_hi[kStages][kTile * kKChunk];
    float b_lo[kStages][kTile * kKChunk];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full;
    uint32_t tmem_base;
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, tf32 inputs, fp32 accumulate; one thread issues.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start >> 4 | LBO(16 B, unused for swizzled K-major) | SBO = 8 rows x 128 B | version 1 | layout 2.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(kTcThreads, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
               int S, int N, int Tp, double* __restrict__ out, int64_t pitch, int s0, int fisher) {
    extern __shared__ uint8_t smem_raw[];
    TcSmem& sm = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int ti, tj;
    c_to_nm((int64_t)blockIdx.x, ti, tj);          // (ti - 1, tj) enumerates the lower-triangular tiles
    ti -= 1;
    const int sg0 = blockIdx.y * kSubj;
    const int nsub = min(kSubj, S - sg0);
    const int num_k = Tp / kKChunk;
    const bool diag = (ti == tj);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&sm.full[i], 1);
            mbar_init(&sm.empty[i], 1);
        }
        mbar_init(&sm.tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == 0 && lane == 0) {
        // ------------------------------------------------------------ TMA producer
        const uint32_t bytes = diag ? 2 * kTileBytes : 4 * kTileBytes;
        int it = 0;
        for (int sg = 0; sg < nsub; ++sg) {
            const int rowA = (sg0 + sg) * N + ti * kTile;
            const int rowB = (sg0 + sg) * N + tj * kTile;
            for (int kc = 0; kc < num_k; ++kc, ++it) {
                const int st = it % kStages;
                mbar_wait(&sm.empty[st], ((it / kStages) & 1) ^ 1);
                mbar_expect_tx(&sm.full[st], bytes);
                tma_load_2d(sm.a_hi[st], &map_hi, kc * kKChunk, rowA, &sm.full[st]);
                tma_load_2d(sm.a_lo[st], &map_lo, kc * kKChunk, rowA, &sm.full[st]);
                if (!diag) {
                    tma_load_2d(sm.b_hi[st], &map_hi, kc * kKChunk, rowB, &sm.full[st]);
                    tma_load_2d(sm.b_lo[st], &map_lo, kc * kKChunk, rowB, &sm.full[st]);
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = umma_idesc_tf32(kTile, kTile);
        int it = 0;
        for (int sg = 0; sg < nsub; ++sg) {
            const uint32_t d = tmem + (uint32_t)(sg * kTile);
            for (int kc = 0; kc < num_k; ++kc, ++it) {
                const int st = it % kStages;
                mbar_wait(&sm.full[st], (it / kStages) & 1);
                tc_fence_after();
                const uint64_t ah = umma_desc_sw128(smem_u32(sm.a_hi[st]));
                const uint64_t al = umma_desc_sw128(smem_u32(sm.a_lo[st]));
                const uint64_t bh = diag ? ah : umma_desc_sw128(smem_u32(sm.b_hi[st]));
                const uint64_t bl = diag ? al : umma_desc_sw128(smem_u32(sm.b_lo[st]));
#pragma unroll
                for (int k = 0; k < kKChunk / 8; ++k) {          // UMMA_K = 8 tf32 = 32 bytes = +2 in the address field
                    const uint64_t o = (uint64_t)(2 * k);
                    umma_tf32(d, ah + o, bh + o, idesc, (kc | k) != 0);
                    umma_tf32(d, ah + o, bl + o, idesc, 1);
                    umma_tf32(d, al + o, bh + o, idesc, 1);
                }
                umma_commit(&sm.empty[st]);                       // stage reusable once these MMAs have read it
            }
        }
        umma_commit(&sm.tmem_full);
    } else if (warp >= 2) {
        // ------------------------------------------------------------ epilogue
        mbar_wait(&sm.tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;                                   // TMEM sub-partition of this warp
        const int n = ti * kTile + q * 32 + lane;
        const int64_t rowbase = (int64_t)n * (n - 1) / 2;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        const bool vec_ok = (nsub == kSubj) && ((pitch & 3) == 0) && (((s0 + sg0) & 3) == 0) &&
                            ((reinterpret_cast<uintptr_t>(out) & 31) == 0);
        for (int col0 = 0; col0 < kTile; col0 += 8) {
            uint32_t v[kSubj][8];
#pragma unroll
            for (int sg = 0; sg < kSubj; ++sg)
                if (sg < nsub) tmem_ld8(tlane + (uint32_t)(sg * kTile + col0), v[sg]);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int m = tj * kTile + col0 + j;
                if (n < N && m < n) {
                    double* dst = out + (rowbase + m) * pitch + s0 + sg0;
                    if (vec_ok) {
                        double2 lo2, hi2;
                        lo2.x = corr_epilogue_f32(__uint_as_float(v[0][j]), fisher);
                        lo2.y = corr_epilogue_f32(__uint_as_float(v[1][j]), fisher);
                        hi2.x = corr_epilogue_f32(__uint_as_float(v[2][j]), fisher);
                        hi2.y = corr_epilogue_f32(__uint_as_float(v[3][j]), fisher);
                        reinterpret_cast<double2*>(dst)[0] = lo2;
                        reinterpret_cast<double2*>(dst)[1] = hi2;
                    } else {
#pragma unroll
                        for (int sg = 0; sg < kSubj; ++sg)
                            if (sg < nsub) dst[sg] = corr_epilogue_f32(__uint_as_float(v[sg][j]), fisher);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static bool make_map(CUtensorMap* map, const float* base, int64_t rows, int Tp) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)Tp, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)Tp * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kKChunk, (cuuint32_t)kTile};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool corr_tc_supported(int N) { return N >= 64; }

int corr_gram_tc(const float* Zh, const float* Zl, int S, int N, int Tp, double* out, int64_t pitch, int s0,
                 int fisher, cudaStream_t st) {
    alignas(64) CUtensorMap map_hi, map_lo;
    if (!make_map(&map_hi, Zh, (int64_t)S * N, Tp) || !make_map(&map_lo, Zl, (int64_t)S * N, Tp)) {
        set_error("fcd_corr_fisherz: cuTensorMapEncodeTiled failed");
        return -2;
    }
    const size_t smem = sizeof(TcSmem) + 1024;
    cudaError_t e = cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("fcd_corr_fisherz: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return -2;
    }
    const int nt = (N + kTile - 1) / kTile;
    dim3 grid((unsigned)(nt * (nt + 1) / 2), (unsigned)((S + kSubj - 1) / kSubj));
    gram_tc_kernel<<<grid, kTcThreads, smem, st>>>(map_hi, map_lo, S, N, Tp, out, pitch, s0, fisher);
    return check_launch("fcd_corr_fisherz(gram_tc)");
}

}  // namespace fcd
