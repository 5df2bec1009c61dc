// K1 tensor-core path: per-subject Gram matrices R = Z Z^T on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by
// TMA with the 128-byte swizzle), error-compensated split-TF32:
//     Z = Zh + Zl (both exactly representable in TF32),
//     R ~= Zh Zh^T + Zh Zl^T + Zl Zh^T          (3 MMAs per product; Zl Zl^T ~ 2^-22 dropped)
//
// Persistent CTAs (one per SM) walk a list of work items = (128 x NB tile of the lower
// triangle, PAIR of consecutive subjects).  The 512 TMEM columns hold two buffers of
// two 128-column fp32 accumulators: while the epilogue warps drain the pair of item i
// (TMEM -> registers -> clip / atanh -> one 16-byte store per edge of the edge-major
// (C, S) output, the layout the EM kernels stream), the MMA warp already accumulates
// item i + 1 into the other buffer -- the epilogue (~70 instructions per value, as
// long as the item's MMAs) is off the tensor pipe's critical path.
//
// Tail tiles.  With N = 400 the fourth block of rows has 16 rows; as the M side of an
// MMA it would cost a full 128-row tile (4 of the 10 tiles of the lower triangle, 87 %
// padding).  Such tiles are computed TRANSPOSED: A = the 128 rows of the column block,
// B = the tail rows, padded to a multiple of 16 (UMMA N = 16 .. 128) -- an eighth of
// the MMAs and of the operand bytes; the epilogue swaps the roles of lane and column.
//
// Warp roles (320 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator
// + MMA issuer (one lane), warps 2..9 = epilogue (two warps per TMEM sub-partition,
// each takes half of the tile's columns).  3-stage smem ring of {A_hi, A_lo, B_hi,
// B_lo} 128 x 32-float tiles (64 KB per stage), full/empty mbarriers; tcgen05.commit
// releases a stage when the MMAs that read it have completed, and hands a finished
// TMEM buffer to the epilogue (tmem_full); the epilogue warps give it back (tmem_empty).
#include <cuda.h>

#include <cstdlib>

#include "fcd_corr.cuh"

namespace fcd {

constexpr int kTile = 128;                       // tile rows of A (UMMA M)
constexpr int kStages = 3;
constexpr int kSubj = 2;                         // subjects per work item (2 x 128 TMEM columns per buffer)
constexpr int kEpiWarps = 8;
constexpr uint32_t kTileBytes = kTile * kKChunk * 4;      // 16 KB
constexpr int kTcThreads = 64 + kEpiWarps * 32;

struct TcSmem {
    float a_hi[kStages][kTile * kKChunk];
    float a_lo[kStages][kTile * kKChunk];
    float b_hi[kStages][kTile * kKChunk];
    float b_lo[kStages][kTile * kKChunk];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
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

// One work item: tile (ti >= tj) of the lower triangle x subjects [sg0, sg0 + nsub).
struct TcItem {
    int ti, tj, sg0, nsub, nb;      // nb = UMMA N: 128, or the padded tail for a transposed tile
    bool swap, diag;
};

__device__ __forceinline__ TcItem tc_item(int item, int ntl, int nt, int tailpad, int S) {
    TcItem w;
    const int g = item / ntl;
    c_to_nm((int64_t)(item - g * ntl), w.ti, w.tj);          // (ti - 1, tj) enumerates the lower-triangular tiles
    w.ti -= 1;
    w.sg0 = g * kSubj;
    w.nsub = min(kSubj, S - w.sg0);
    w.swap = (w.ti == nt - 1) && tailpad < kTile;
    w.nb = w.swap ? tailpad : kTile;
    w.diag = (w.ti == w.tj);
    return w;
}

__global__ void __launch_bounds__(kTcThreads, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
               const __grid_constant__ CUtensorMap map_hi_t, const __grid_constant__ CUtensorMap map_lo_t,
               int S, int N, int Tp, int tailpad, double* __restrict__ out, int64_t pitch, int s0, int fisher) {
    extern __shared__ uint8_t smem_raw[];
    TcSmem& sm = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = (N + kTile - 1) / kTile, ntl = nt * (nt + 1) / 2;
    const int nitems = ntl * ((S + kSubj - 1) / kSubj);
    const int num_k = Tp / kKChunk;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&sm.full[i], 1);
            mbar_init(&sm.empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.tmem_full[i], 1);
            mbar_init(&sm.tmem_empty[i], kEpiWarps);
        }
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
        int it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const TcItem w = tc_item(item, ntl, nt, tailpad, S);
            const uint32_t bytes = 2 * kTileBytes + (w.diag ? 0u : 2u * (uint32_t)w.nb * kKChunk * 4u);
            for (int sg = 0; sg < w.nsub; ++sg) {
                const int rowA = (w.sg0 + sg) * N + (w.swap ? w.tj : w.ti) * kTile;
                const int rowB = (w.sg0 + sg) * N + (w.swap ? w.ti : w.tj) * kTile;
                for (int kc = 0; kc < num_k; ++kc, ++it) {
                    const int st = it % kStages;
                    mbar_wait(&sm.empty[st], ((it / kStages) & 1) ^ 1);
                    mbar_expect_tx(&sm.full[st], bytes);
                    tma_load_2d(sm.a_hi[st], &map_hi, kc * kKChunk, rowA, &sm.full[st]);
                    tma_load_2d(sm.a_lo[st], &map_lo, kc * kKChunk, rowA, &sm.full[st]);
                    if (!w.diag) {                           // (a diagonal tile's B rows are A's, or a prefix of them)
                        tma_load_2d(sm.b_hi[st], w.swap ? &map_hi_t : &map_hi, kc * kKChunk, rowB, &sm.full[st]);
                        tma_load_2d(sm.b_lo[st], w.swap ? &map_lo_t : &map_lo, kc * kKChunk, rowB, &sm.full[st]);
                    }
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ------------------------------------------------------------ MMA issuer
        int it = 0, li = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++li) {
            const TcItem w = tc_item(item, ntl, nt, tailpad, S);
            const uint32_t idesc = umma_idesc_tf32(kTile, w.nb);
            const int buf = li & 1;
            mbar_wait(&sm.tmem_empty[buf], ((li >> 1) & 1) ^ 1);          // the epilogue has drained this buffer
            tc_fence_after();
            for (int sg = 0; sg < w.nsub; ++sg) {
                const uint32_t d = tmem + (uint32_t)(buf * kSubj * kTile + sg * kTile);
                for (int kc = 0; kc < num_k; ++kc, ++it) {
                    const int st = it % kStages;
                    mbar_wait(&sm.full[st], (it / kStages) & 1);
                    tc_fence_after();
                    const uint64_t ah = umma_desc_sw128(smem_u32(sm.a_hi[st]));
                    const uint64_t al = umma_desc_sw128(smem_u32(sm.a_lo[st]));
                    const uint64_t bh = w.diag ? ah : umma_desc_sw128(smem_u32(sm.b_hi[st]));
                    const uint64_t bl = w.diag ? al : umma_desc_sw128(smem_u32(sm.b_lo[st]));
#pragma unroll
                    for (int k = 0; k < kKChunk / 8; ++k) {      // UMMA_K = 8 tf32 = 32 bytes = +2 in the address field
                        const uint64_t o = (uint64_t)(2 * k);
                        umma_tf32(d, ah + o, bh + o, idesc, (kc | k) != 0);
                        umma_tf32(d, ah + o, bl + o, idesc, 1);
                        umma_tf32(d, al + o, bh + o, idesc, 1);
                    }
                    umma_commit(&sm.empty[st]);                   // stage reusable once these MMAs have read it
                }
            }
            umma_commit(&sm.tmem_full[buf]);                      // the item's accumulators are complete
        }
    } else if (warp >= 2) {
        // ------------------------------------------------------------ epilogue
        const int q = warp & 3;                                   // TMEM sub-partition of this warp
        const int half = (warp - 2) >> 2;                         // which half of the tile's columns
        const int row = q * 32 + lane;                            // accumulator row (TMEM lane)
        int li = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++li) {
            const TcItem w = tc_item(item, ntl, nt, tailpad, S);
            const int buf = li & 1;
            mbar_wait(&sm.tmem_full[buf], (li >> 1) & 1);
            tc_fence_after();
            const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kSubj * kTile);
            const bool vec_ok = (w.nsub == kSubj) && ((pitch & 1) == 0) && (((s0 + w.sg0) & 1) == 0) &&
                                ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
            const int cols = w.nb / 2;
            for (int col0 = half * cols; col0 < (half + 1) * cols; col0 += 8) {
                uint32_t v[kSubj][8];
#pragma unroll
                for (int sg = 0; sg < kSubj; ++sg)
                    if (sg < w.nsub) tmem_ld8(tlane + (uint32_t)(sg * kTile + col0), v[sg]);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // plain tile: row <-> n, column <-> m; transposed (tail) tile: the other way round
                    const int n = w.ti * kTile + (w.swap ? col0 + j : row);
                    const int m = w.tj * kTile + (w.swap ? row : col0 + j);
                    if (n < N && m < n) {
                        double* dst = out + ((int64_t)n * (n - 1) / 2 + m) * pitch + s0 + w.sg0;
                        if (vec_ok) {
                            double2 r2;
                            r2.x = corr_epilogue_f32(__uint_as_float(v[0][j]), fisher);
                            r2.y = corr_epilogue_f32(__uint_as_float(v[1][j]), fisher);
                            *reinterpret_cast<double2*>(dst) = r2;
                        } else {
#pragma unroll
                            for (int sg = 0; sg < kSubj; ++sg)
                                if (sg < w.nsub) dst[sg] = corr_epilogue_f32(__uint_as_float(v[sg][j]), fisher);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tmem_empty[buf]);
        }
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

static bool make_map(CUtensorMap* map, const float* base, int64_t rows, int Tp, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)Tp, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)Tp * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kKChunk, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool corr_tc_supported(int N) { return N >= 64; }

int corr_gram_tc(const float* Zh, const float* Zl, int S, int N, int Tp, double* out, int64_t pitch, int s0,
                 int fisher, cudaStream_t st) {
    alignas(64) CUtensorMap map_hi, map_lo, map_hi_t, map_lo_t;
    const int nt = (N + kTile - 1) / kTile;
    const int tail = N - (nt - 1) * kTile;                       // rows of the last block
    int tailpad = (tail + 15) / 16 * 16;                         // UMMA N of the transposed tail tiles
    if (getenv("FCD_CORR_NO_TAIL") != nullptr) tailpad = kTile;  // (experiments: every tile a full 128 x 128)
    const int64_t rows = (int64_t)S * N;
    if (!make_map(&map_hi, Zh, rows, Tp, kTile) || !make_map(&map_lo, Zl, rows, Tp, kTile) ||
        !make_map(&map_hi_t, Zh, rows, Tp, tailpad) || !make_map(&map_lo_t, Zl, rows, Tp, tailpad)) {
        set_error("fcd_corr_fisherz: cuTensorMapEncodeTiled failed");
        return -2;
    }
    const size_t smem = sizeof(TcSmem) + 1024;
    cudaError_t e = cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("fcd_corr_fisherz: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return -2;
    }
    const int64_t nitems = (int64_t)(nt * (nt + 1) / 2) * ((S + kSubj - 1) / kSubj);
    const int grid = (int)(nitems < sm_count() ? nitems : sm_count());      // persistent: one CTA per SM
    gram_tc_kernel<<<grid, kTcThreads, smem, st>>>(map_hi, map_lo, map_hi_t, map_lo_t, S, N, Tp, tailpad, out, pitch,
                                                   s0, fisher);
    return check_launch("fcd_corr_fisherz(gram_tc)");
}

}  // namespace fcd
