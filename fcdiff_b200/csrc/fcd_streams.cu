// K3b as dense streams.  Within one EM iteration the (eta, epsilon) optimiser
// evaluates E_lM(eta, epsilon) 6-16 times with q_F, q_R and the responsibility
// planes fixed (fcdiff/fit.py:228-241, 270-286).  Everything of an evaluation
// that does not depend on (eta, epsilon) is therefore hoisted into ONE pass per
// iteration, the *bucket pass*:
//   * an element (c,u) whose edge and regions are peaked (fcd_common.cuh, "Tiers")
//     contributes log(a_l + b_l p) for ONE (k*, l*) pair with weight 1: its p is
//     appended to the stream G_l of its l* (three streams, 8 bytes per element);
//   * every other element is expanded into records {p_k, w_0, w_1, w_2} with
//     w_l = q_F[c,k] * pair weight_l (one record if the edge is peaked, three if not);
//   * the theta-free part sum w L of E_lM is summed on the way.
// An evaluation is then a flat reduction over the streams with warp-uniform
// constants (a_l, b_l): no row structure, no weights, no peak-state decoding --
// 8 bytes and ~25 instructions per edge-patient, which is what lets it run at
// the HBM roofline.  Stream order is fixed by per-row counts and an exclusive
// scan, so results are deterministic.
#include "fcd_common.cuh"

namespace fcd {

constexpr int kBucketThreads = 256;

// code of element (c,u): 0..2 = l* (both regions peaked), 3 = a region is mixed, 4 = padding
__device__ __forceinline__ int pair_code(int sn, int sm) {
    const int o = sn | sm;
    return o < 2 ? (((sn ^ sm) << 1) + (sn & sm)) : ((o & 6) == 2 ? 3 : 4);
}

// Rows are handled in blocks of kRowBlock consecutive rows per CTA (one warp per
// row, kRowBlock / 8 rows per warp), so that the position of a row in the streams
// is (offset of its block) + (prefix inside the block): the block totals are
// scanned by one small kernel, the in-block prefix is recomputed in shared memory
// by the fill kernel -- no pass over per-row arrays by a single CTA.
constexpr int kRowBlock = 16;

// counts[c] = {n_0, n_1, n_2, n_records} of row c; blocktot[b] their sums over the rows of block b.
__global__ void __launch_bounds__(kBucketThreads)
bucket_count_kernel(const uint8_t* __restrict__ fstate, const uint8_t* __restrict__ rstate, int64_t pitchS,
                    const int32_t* __restrict__ nm, int64_t C, int U, int4* __restrict__ counts,
                    longlong4* __restrict__ blocktot) {
    __shared__ int4 s_cnt[kRowBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t cb = (int64_t)blockIdx.x * kRowBlock;
    for (int r = warp; r < kRowBlock; r += kBucketThreads / 32) {
        const int64_t c = cb + r;
        int4 cnt = make_int4(0, 0, 0, 0);
        if (c < C) {
            const int k = fstate[c];
            if (k == kStateMixedF) {
                cnt.w = 3 * U;
            } else {
                const int v = __ldg(nm + c);
                const uint8_t* rn = rstate + (int64_t)(v & 0xffff) * pitchS;
                const uint8_t* rm = rstate + (int64_t)((v >> 16) & 0xffff) * pitchS;
                // 4 patients per lane: one 32-bit load per state row (pitchS is a multiple of 256, padding = 4)
                for (int u0 = 0; u0 < U; u0 += 128) {
                    const int u = u0 + 4 * lane;
                    uint32_t sn4 = 0x04040404u, sm4 = 0x04040404u;
                    if (u < U) {
                        sn4 = __ldg(reinterpret_cast<const uint32_t*>(rn + u));
                        sm4 = __ldg(reinterpret_cast<const uint32_t*>(rm + u));
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int code = pair_code((sn4 >> (8 * e)) & 0xff, (sm4 >> (8 * e)) & 0xff);
                        cnt.x += code == 0;
                        cnt.y += code == 1;
                        cnt.z += code == 2;
                        cnt.w += code == 3;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    cnt.x += __shfl_xor_sync(0xffffffffu, cnt.x, o);
                    cnt.y += __shfl_xor_sync(0xffffffffu, cnt.y, o);
                    cnt.z += __shfl_xor_sync(0xffffffffu, cnt.z, o);
                    cnt.w += __shfl_xor_sync(0xffffffffu, cnt.w, o);
                }
            }
            if (lane == 0) counts[c] = cnt;
        }
        if (lane == 0) s_cnt[r] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        longlong4 t = make_longlong4(0, 0, 0, 0);
        for (int r = 0; r < kRowBlock; ++r) {
            t.x += s_cnt[r].x;
            t.y += s_cnt[r].y;
            t.z += s_cnt[r].z;
            t.w += s_cnt[r].w;
        }
        blocktot[blockIdx.x] = t;
    }
}

// blockoff[b] = exclusive prefix sums of blocktot, totals[0..3] the sums.  One CTA, coalesced.
__global__ void __launch_bounds__(1024)
bucket_scan_kernel(const longlong4* __restrict__ blocktot, int64_t nblocks, longlong4* __restrict__ blockoff,
                   long long* __restrict__ totals) {
    __shared__ long long s_part[4][1024];
    const int t = threadIdx.x;
    long long carry[4] = {0, 0, 0, 0};
    for (int64_t b0 = 0; b0 < nblocks; b0 += 1024) {
        const int64_t b = b0 + t;
        longlong4 v = make_longlong4(0, 0, 0, 0);
        if (b < nblocks) v = blocktot[b];
        const long long mine[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) s_part[i][t] = mine[i];
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {                 // Hillis-Steele inclusive scan
            long long w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = t >= d ? s_part[i][t - d] : 0;
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) s_part[i][t] += w[i];
            __syncthreads();
        }
        if (b < nblocks)
            blockoff[b] = make_longlong4(carry[0] + s_part[0][t] - mine[0], carry[1] + s_part[1][t] - mine[1],
                                         carry[2] + s_part[2][t] - mine[2], carry[3] + s_part[3][t] - mine[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) carry[i] += s_part[i][1023];
        __syncthreads();
    }
    if (t == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) totals[i] = carry[i];
    }
}

struct Record {
    double p, w0, w1, w2;
};

// out[0] = sum of the [C][U] plane (pitch pitchU): the L plane's total, once per cache.
__global__ void __launch_bounds__(kBucketThreads)
plane_sum_kernel(const double* __restrict__ X, int64_t C, int U, int64_t pitchU, double* __restrict__ out,
                 double* __restrict__ ws) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * (kBucketThreads / 32);
    double s0 = 0.0, s1 = 0.0;
    for (int64_t c = (int64_t)blockIdx.x * (kBucketThreads / 32) + (threadIdx.x >> 5); c < C; c += nw) {
        const double* row = X + c * pitchU;
        for (int u = 2 * lane; u < U; u += 64) {             // pitchU even: u + 1 < pitchU
            const double2 v = ldg_stream2(row + u);
            s0 += v.x;
            s1 += u + 1 < U ? v.y : 0.0;
        }
    }
    double vv[1] = {s0 + s1};
    grid_reduce_store<1, kBucketThreads>(vv, ws, out);
}

// Fills the streams.  G_l starts at G + base[l]; records at D.  The L plane is not read here: the
// theta-free part of E_lM over the stream elements is (sum of L over all local elements, once per
// cache) - (L of the record elements), finished by bucket_records_kernel.
// Same row blocks as bucket_count_kernel.  Per 64-patient chunk a lane owns the
// patients 2*lane, 2*lane+1 (128-bit loads); within a chunk the stream order is
// "all first elements, then all second elements" -- any fixed order will do.
__global__ void __launch_bounds__(kBucketThreads, 2)
bucket_fill_kernel(const double* __restrict__ P, int64_t planeStride,
                   int64_t C, int U, int64_t pitchU,
                   const double* __restrict__ qF, const uint8_t* __restrict__ fstate,
                   const double* __restrict__ qR, const uint8_t* __restrict__ rstate, int64_t pitchS,
                   const int32_t* __restrict__ nm, const int4* __restrict__ counts,
                   const longlong4* __restrict__ blockoff,
                   long long base0, long long base1, long long base2,
                   double* __restrict__ G, Record* __restrict__ D) {
    __shared__ longlong4 s_off[kRowBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int64_t blk = blockIdx.x; blk * kRowBlock < C; blk += gridDim.x) {
        const int64_t cb = blk * kRowBlock;
        __syncthreads();
        if (threadIdx.x == 0) {                              // in-block exclusive prefix of the row counts
            longlong4 run = blockoff[blk];
            for (int r = 0; r < kRowBlock; ++r) {
                s_off[r] = run;
                if (cb + r < C) {
                    const int4 v = counts[cb + r];
                    run.x += v.x;
                    run.y += v.y;
                    run.z += v.z;
                    run.w += v.w;
                }
            }
        }
        __syncthreads();
        for (int r = warp; r < kRowBlock; r += kBucketThreads / 32) {
            const int64_t c = cb + r;
            if (c >= C) break;
            const int k = fstate[c];
            const int v = __ldg(nm + c);
            const int n = v & 0xffff, m = (v >> 16) & 0xffff;
            const longlong4 o = s_off[r];
            long long pos[4] = {base0 + o.x, base1 + o.y, base2 + o.z, o.w};
            // 32-bit running offsets from the row's four stream positions
            double* const g0 = G + pos[0];
            double* const g1 = G + pos[1];
            double* const g2 = G + pos[2];
            Record* const d3 = D + pos[3];
            uint32_t off0 = 0, off1 = 0, off2 = 0, off3 = 0;
            if (k == kStateMixedF) {                         // three records per element, tagged with their state
                for (int u = lane; u < U; u += 32) {
                    const int64_t i = c * pitchU + u;
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk) {
                        Record rec;
                        rec.p = ldg_stream1(P + kk * planeStride + i);
                        rec.w0 = 0.0;
                        rec.w1 = __hiloint2double((int)c, u);
                        rec.w2 = 1.0 + kk;
                        D[pos[3] + 3 * (int64_t)u + kk] = rec;
                    }
                }
                continue;
            }
            const double* row = P + (int64_t)k * planeStride + c * pitchU;
            const uint8_t* rn = rstate + (int64_t)n * pitchS;
            const uint8_t* rm = rstate + (int64_t)m * pitchS;
            struct Chunk {
                uint32_t sn2, sm2;
                double2 p2;
            };
            auto load = [&](int u0) {
                Chunk ch;
                ch.sn2 = ch.sm2 = 0x0404u;
                ch.p2 = make_double2(0.0, 0.0);
                const int u = u0 + 2 * lane;
                if (u < U) {                                 // pitchU is even: u + 1 < pitchU
                    ch.sn2 = __ldg(reinterpret_cast<const unsigned short*>(rn + u));
                    ch.sm2 = __ldg(reinterpret_cast<const unsigned short*>(rm + u));
                    ch.p2 = ldg_stream2(row + u);
                }
                return ch;
            };
            auto process = [&](const Chunk& cur, int u0) {
                if (u0 >= U) return;                         // warp-uniform
                const int u = u0 + 2 * lane;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int code = pair_code((cur.sn2 >> (8 * e)) & 0xff, (cur.sm2 >> (8 * e)) & 0xff);
                    const double p = e ? cur.p2.y : cur.p2.x;
                    // every element has exactly one destination: one store per lane
                    const unsigned b0 = __ballot_sync(0xffffffffu, code == 0);
                    const unsigned b1 = __ballot_sync(0xffffffffu, code == 1);
                    const unsigned b2 = __ballot_sync(0xffffffffu, code == 2);
                    const unsigned b3 = __ballot_sync(0xffffffffu, code == 3);
                    const unsigned mine = code == 0 ? b0 : (code == 1 ? b1 : (code == 2 ? b2 : b3));
                    const uint32_t at = (code == 0 ? off0 : (code == 1 ? off1 : (code == 2 ? off2 : off3))) +
                                        __popc(mine & lt);
                    if (code < 3) (code == 0 ? g0 : (code == 1 ? g1 : g2))[at] = p;
                    if (code == 3) {                         // weights are filled in by bucket_records_kernel
                        Record rec;
                        rec.p = p;
                        rec.w0 = 0.0;
                        rec.w1 = __hiloint2double((int)c, u + e);
                        rec.w2 = 0.0;
                        d3[at] = rec;
                    }
                    off0 += __popc(b0);
                    off1 += __popc(b1);
                    off2 += __popc(b2);
                    off3 += __popc(b3);
                }
            };
            // two register sets of two chunks each, loaded in turn (no register rotation:
            // a move of a loaded value would wait for the load and defeat the prefetch)
            Chunk a0 = load(0), a1 = load(64);
            for (int u0 = 0; u0 < U; u0 += 256) {
                const Chunk b0 = load(u0 + 128), b1 = load(u0 + 192);
                process(a0, u0);
                process(a1, u0 + 64);
                a0 = load(u0 + 256);
                a1 = load(u0 + 320);
                process(b0, u0 + 128);
                process(b1, u0 + 192);
            }
        }
    }
}

// Second half of the bucket pass: the records of unpeaked elements were left as {p, -, (c, u), tag};
// one thread per record gathers the two regions' posteriors and the element's L (massively parallel,
// nothing on a row walker's critical path), writes {p, w_0, w_1, w_2} and accumulates
//   out[0] = Lsum - sum_{record elements} L + sum_records (sum_l w_l) L
// (Lsum = sum of L over all local elements, fcd_plane_sum: the stream elements have weight exactly 1).
__global__ void __launch_bounds__(kBucketThreads)
bucket_records_kernel(Record* __restrict__ D, long long nd, const double* __restrict__ L, int64_t pitchU,
                      const double* __restrict__ Lsum, const double* __restrict__ qF,
                      const double* __restrict__ qR, int U, const int32_t* __restrict__ nm,
                      double* __restrict__ out, double* __restrict__ ws) {
    const double2* qR2 = reinterpret_cast<const double2*>(qR);
    double cs = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nd;
         i += (long long)gridDim.x * blockDim.x) {
        Record rec = D[i];                                   // {p, -, (c, u), tag}: tag 0 = peaked edge, 1 + k otherwise
        const int c = __double2hiint(rec.w1), u = __double2loint(rec.w1);
        const int tag = (int)rec.w2;
        const double lv = __ldg(L + (int64_t)c * pitchU + u);
        const int v = __ldg(nm + c);
        double w[3];
        pair_weights(__ldg(qR2 + (int64_t)(v & 0xffff) * U + u), __ldg(qR2 + (int64_t)((v >> 16) & 0xffff) * U + u), w);
        double scale = 1.0, cscale = 1.0;
        if (tag > 0) {
            const double qf[3] = {__ldg(qF + (int64_t)c * 3), __ldg(qF + (int64_t)c * 3 + 1), __ldg(qF + (int64_t)c * 3 + 2)};
            scale = qf[tag - 1];
            cscale = tag == 1 ? qf[0] + qf[1] + qf[2] : 0.0;  // the element's theta-free term is counted once
        }
        // the element's theta-free term is counted once (tag <= 1), and it leaves the weight-1 sum
        cs = fma(cscale * (w[0] + w[1] + w[2]) - (tag <= 1 ? 1.0 : 0.0), lv, cs);
        rec.w0 = scale * w[0];
        rec.w1 = scale * w[1];
        rec.w2 = scale * w[2];
        D[i] = rec;
    }
    double vv[1] = {cs + ((blockIdx.x == 0 && threadIdx.x == 0) ? Lsum[0] : 0.0)};
    grid_reduce_store<1, kBucketThreads>(vv, ws, out);
}

// One evaluation: out[0] = sum_l sum_{p in G_l} log(a_l + b_l p) + sum_records sum_l w_l log(a_l + b_l p),
// out[1] = dE/d eta, out[2] = dE/d eps as in elm_kernel (fcd_mstep.cu).
// FAST: the stream elements have weight exactly 1, so the objective is the log of a running product
// (fcd_math.cuh "Sum of logs as the log of a product"; flushed by the caller every kProdMax factors)
// and the gradient needs 1 / M only: MUFU.RCP64H + one Newton step.  7 fp64 instructions per element.
struct StreamAcc {
    double obj, g, prod;
};

template <bool GRAD, bool FAST>
__device__ __forceinline__ void stream_elem(double p, double a, double b, const double* s_tab, StreamAcc& acc) {
    const double M = fma(b, p, a);
    if (FAST) {
        acc.prod *= M;
        if (GRAD) acc.g = fma(mix_num(p), rcp_newton(M), acc.g);
    } else if (GRAD) {
        double rcp;
        acc.obj += fast_log_rcp<FAST>(M, s_tab, rcp);
        acc.g = fma(mix_num(p), rcp, acc.g);
    } else {
        acc.obj += fast_log<FAST>(M, s_tab);
    }
}

template <bool FAST>
__device__ __forceinline__ void stream_flush(StreamAcc& a0, StreamAcc& a1, const double* s_tab) {
    if (FAST) {
        a0.obj += log_pos<FAST>(a0.prod * a1.prod, s_tab);
        a0.prod = a1.prod = 1.0;
    }
}

// The streams are staged through shared memory by the TMA unit: the bytes in flight do not
// depend on registers (a register-prefetching version stalled on its loads at ~55 % of the HBM peak).  One persistent CTA of 16 warps per
// SM; the three streams are cut into chunks of kEvChunk doubles, numbered across the
// streams, and warp g takes the chunks g, g + G, ...; every warp runs a private ring
// of kEvDepth stages (one 1-D bulk copy and one mbarrier per stage) and refills a
// stage as soon as it has consumed it.
constexpr int kEvChunk = 256;
constexpr int kEvDepth = 4;
constexpr int kEvWarps = 16;
constexpr int kEvThreads = kEvWarps * 32;

template <bool GRAD, bool FAST>
__global__ void __launch_bounds__(kEvThreads, 1)
elm_streams_tma_kernel(const double* __restrict__ G, long long base0, long long base1, long long base2,
                       long long n0, long long n1, long long n2,
                       const Record* __restrict__ D, long long nd,
                       const __grid_constant__ ThetaDev th, const __grid_constant__ LogTabWindow tab,
                       double* __restrict__ out, double* __restrict__ ws) {
    extern __shared__ __align__(128) double s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* ring = s_dyn + (FAST ? ((tab.n + 15) & ~15) : 0) + (size_t)warp * kEvDepth * kEvChunk;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_dyn + (FAST ? ((tab.n + 15) & ~15) : 0) +
                                                 (size_t)kEvWarps * kEvDepth * kEvChunk) + warp * kEvDepth;
    if (lane < kEvDepth) mbar_init(bars + lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    // chunks are numbered across the FOUR streams (G_0, G_1, G_2, then the records as a stream of
    // 4-double elements); a warp's position is (stream l, chunk i of that stream)
    const long long n3 = nd * 4;
    const long long nq0 = (n0 + kEvChunk - 1) / kEvChunk, nq1 = (n1 + kEvChunk - 1) / kEvChunk,
                    nq2 = (n2 + kEvChunk - 1) / kEvChunk, nq3 = (n3 + kEvChunk - 1) / kEvChunk;
    const double* const Dd = reinterpret_cast<const double*>(D);
    const long long W = (long long)gridDim.x * kEvWarps;
    struct Pos {
        int l;
        long long i;
    };
    auto chunks_of = [&](int l) { return l == 0 ? nq0 : (l == 1 ? nq1 : (l == 2 ? nq2 : nq3)); };
    auto normalise = [&](Pos& p) {                           // carry into the next stream(s)
        while (p.l < 4 && p.i >= chunks_of(p.l)) {
            p.i -= chunks_of(p.l);
            ++p.l;
        }
    };
    Pos ip = {0, (long long)blockIdx.x * kEvWarps + warp};   // next chunk to issue
    normalise(ip);
    int pd = 0;
    auto issue = [&]() {
        if (ip.l >= 4) return;
        if (lane == 0) {
            const long long n = ip.l == 0 ? n0 : (ip.l == 1 ? n1 : (ip.l == 2 ? n2 : n3));
            const double* src = ip.l == 0 ? G + base0 : (ip.l == 1 ? G + base1 : (ip.l == 2 ? G + base2 : Dd));
            const long long left = n - ip.i * kEvChunk;
            const uint32_t cnt = (uint32_t)(left < kEvChunk ? left : kEvChunk);
            const uint32_t bytes = ((cnt + 1) & ~1u) * 8;    // even element count: 16-byte granules
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bars + pd, bytes);
            tma_load_1d(ring + pd * kEvChunk, src + ip.i * kEvChunk, bytes, bars + pd);
        }
        ip.i += W;
        normalise(ip);
        if (++pd == kEvDepth) pd = 0;
    };
#pragma unroll 1
    for (int i = 0; i < kEvDepth; ++i) issue();
    // the table is staged while the first chunks are in flight (the stream bodies need it only for
    // the rare product flush)
    const double* s_tab = load_log_table<FAST>(tab, s_dyn);

    double obj = 0.0, gl[3] = {0.0, 0.0, 0.0};
    const double al[3] = {th.al[0], th.al[1], th.al[2]}, bl[3] = {th.bl[0], th.bl[1], th.bl[2]};
    const long long nn[3] = {n0, n1, n2};
    int d = 0;
    uint32_t phase = 0;
    Pos cp = {0, (long long)blockIdx.x * kEvWarps + warp};
    normalise(cp);
    auto advance = [&]() {
        __syncwarp();
        issue();
        if (++d == kEvDepth) {
            d = 0;
            phase ^= 1;
        }
        cp.i += W;
        normalise(cp);
    };
#pragma unroll
    for (int l = 0; l < 3; ++l) {                            // the warp's chunks of stream l: constants in registers
        const double a = al[l], b = bl[l];
        StreamAcc a0 = {0.0, 0.0, 1.0}, a1 = {0.0, 0.0, 1.0};
        int nf = 0;                                          // factors in each running product
        while (cp.l == l) {
            const long long left = nn[l] - cp.i * kEvChunk;
            mbar_wait(bars + d, phase);
            const double* st = ring + d * kEvChunk;
            if (left >= kEvChunk) {
#pragma unroll
                for (int j = 0; j < kEvChunk / 64; ++j) {
                    const double2 v = *reinterpret_cast<const double2*>(st + 64 * j + 2 * lane);
                    stream_elem<GRAD, FAST>(v.x, a, b, s_tab, a0);
                    stream_elem<GRAD, FAST>(v.y, a, b, s_tab, a1);
                }
            } else {
                for (int e = lane; e < (int)left; e += 32) stream_elem<GRAD, FAST>(st[e], a, b, s_tab, a0);
            }
            advance();
            nf += left >= kEvChunk ? kEvChunk / 64 : kEvChunk / 32;      // a tail chunk goes to a0 only
            if (nf + kEvChunk / 32 > kProdMax) {
                stream_flush<FAST>(a0, a1, s_tab);
                nf = 0;
            }
        }
        stream_flush<FAST>(a0, a1, s_tab);
        obj += a0.obj + a1.obj;
        gl[l] = a0.g + a1.g;
    }
    while (cp.l == 3) {                                      // records {p, w_0, w_1, w_2}: real weights, three logs each
        const long long left = (n3 - cp.i * kEvChunk) >> 2;
        const int cnt = (int)(left < kEvChunk / 4 ? left : kEvChunk / 4);
        mbar_wait(bars + d, phase);
        const double4* st = reinterpret_cast<const double4*>(ring + d * kEvChunk);
#pragma unroll
        for (int j = 0; j < kEvChunk / 128; ++j) {
            const int e = 32 * j + lane;
            if (e < cnt) {
                const double4 r = st[e];
                const double w[3] = {r.y, r.z, r.w};
                const double num = mix_num(r.x);
#pragma unroll
                for (int l = 0; l < 3; ++l) {
                    const double M = fma(bl[l], r.x, al[l]);
                    if (GRAD) {
                        double rcp;
                        obj = fma(w[l], fast_log_rcp<FAST>(M, s_tab, rcp), obj);
                        gl[l] = fma(w[l], num * rcp, gl[l]);
                    } else {
                        obj = fma(w[l], fast_log<FAST>(M, s_tab), obj);
                    }
                }
            }
        }
        advance();
    }
    // dE/d eta = -(2 eps - 1) g_2;  dE/d eps = -(-g_0 + g_1 + (2 eta - 1) g_2)      (fit.py:600-697)
    double v[3] = {obj, -(2.0 * th.epsilon - 1.0) * gl[2], -(-gl[0] + gl[1] + (2.0 * th.eta - 1.0) * gl[2])};
    grid_reduce_store<3, kEvThreads>(v, ws, out);
}

static inline int rows_grid(int64_t rows, int rows_per_block, int waves) {
    int64_t need = (rows + rows_per_block - 1) / rows_per_block;
    int64_t cap = (int64_t)sm_count() * waves;
    if (cap > kMaxReduceBlocks) cap = kMaxReduceBlocks;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace fcd

using namespace fcd;

extern "C" {

int fcd_bucket_count(const uint8_t* fstate, const uint8_t* rstate, int64_t pitchS, const int32_t* nm,
                     int64_t C, int32_t U, int32_t* counts, int64_t* blockoff, int64_t* totals, void* stream) {
    FCD_REQUIRE(fstate != nullptr && rstate != nullptr && nm != nullptr && counts != nullptr && blockoff != nullptr &&
                totals != nullptr, "fcd_bucket_count: NULL argument");
    FCD_REQUIRE(C >= 1 && U >= 1 && pitchS >= U && pitchS % 4 == 0, "fcd_bucket_count: bad shape");
    FCD_REQUIRE((reinterpret_cast<uintptr_t>(counts) & 15) == 0 && (reinterpret_cast<uintptr_t>(blockoff) & 31) == 0 &&
                (reinterpret_cast<uintptr_t>(rstate) & 3) == 0,
                "fcd_bucket_count: counts / blockoff must be 16 / 32-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nblocks = (C + kRowBlock - 1) / kRowBlock;
    // blockoff doubles as the block totals' storage: [nblocks] totals followed by [nblocks] offsets
    longlong4* bt = reinterpret_cast<longlong4*>(blockoff);
    bucket_count_kernel<<<(unsigned)nblocks, kBucketThreads, 0, st>>>(fstate, rstate, pitchS, nm, C, U,
                                                                      reinterpret_cast<int4*>(counts), bt);
    int rc = check_launch("fcd_bucket_count");
    if (rc) return rc;
    bucket_scan_kernel<<<1, 1024, 0, st>>>(bt, nblocks, bt + nblocks, reinterpret_cast<long long*>(totals));
    return check_launch("fcd_bucket_count(scan)");
}

int64_t fcd_bucket_blocks(int64_t C) { return (C + kRowBlock - 1) / kRowBlock; }

int fcd_plane_sum(const double* X, int64_t C, int32_t U, int64_t pitchU, double* out1, double* ws, void* stream) {
    FCD_REQUIRE(X != nullptr && out1 != nullptr && ws != nullptr, "fcd_plane_sum: NULL argument");
    FCD_REQUIRE(C >= 1 && U >= 1 && pitchU >= U && pitchU % 2 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0,
                "fcd_plane_sum: bad shape (even pitch, 16-byte aligned)");
    plane_sum_kernel<<<rows_grid(C, kBucketThreads / 32, 8), kBucketThreads, 0, (cudaStream_t)stream>>>(
        X, C, U, pitchU, out1, ws);
    return check_launch("fcd_plane_sum");
}

int fcd_bucket_fill(const double* P, int64_t planeStride, const double* L, const double* Lsum,
                    int64_t C, int32_t U, int64_t pitchU, const double* qF, const uint8_t* fstate, const double* qR, const uint8_t* rstate,
                    int64_t pitchS, int32_t N, const int32_t* nm, const int32_t* counts, const int64_t* blockoff,
                    const int64_t* base3_host, double* G, double* D, int64_t nd, double* out1, double* ws,
                    void* stream) {
    FCD_REQUIRE(P != nullptr && L != nullptr && Lsum != nullptr && qF != nullptr && fstate != nullptr && qR != nullptr &&
                rstate != nullptr && nm != nullptr && counts != nullptr && blockoff != nullptr &&
                base3_host != nullptr && G != nullptr && D != nullptr && out1 != nullptr && ws != nullptr,
                "fcd_bucket_fill: NULL argument");
    FCD_REQUIRE(C >= 1 && U >= 1 && pitchU >= U && pitchS >= U && N >= 2 && N < 65536, "fcd_bucket_fill: bad shape");
    FCD_REQUIRE((reinterpret_cast<uintptr_t>(D) & 31) == 0 && ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(L)) & 15) == 0 &&
                pitchU % 2 == 0 && planeStride % 2 == 0 && pitchS % 2 == 0,
                "fcd_bucket_fill: planes must be 16-byte aligned with even pitches, records 32-byte aligned");
    const int64_t nblocks = (C + kRowBlock - 1) / kRowBlock;
    int64_t grid = nblocks;                                  // persistent: the resident CTAs share the row blocks
    if (grid > (int64_t)sm_count() * 2) grid = (int64_t)sm_count() * 2;
    bucket_fill_kernel<<<(unsigned)grid, kBucketThreads, 0, (cudaStream_t)stream>>>(
        P, planeStride, C, U, pitchU, qF, fstate, qR, rstate, pitchS, nm, reinterpret_cast<const int4*>(counts),
        reinterpret_cast<const longlong4*>(blockoff) + nblocks, base3_host[0], base3_host[1], base3_host[2], G,
        reinterpret_cast<Record*>(D));
    int rc = check_launch("fcd_bucket_fill");
    if (rc) return rc;
    int64_t rgrid = (nd + kBucketThreads - 1) / kBucketThreads;       // nd == 0: one CTA writes out1[0] = Lsum
    if (rgrid > (int64_t)sm_count() * 8) rgrid = (int64_t)sm_count() * 8;
    if (rgrid < 1) rgrid = 1;
    bucket_records_kernel<<<(unsigned)rgrid, kBucketThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<Record*>(D), nd, L, pitchU, Lsum, qF, qR, U, nm, out1, ws);
    return check_launch("fcd_bucket_fill(records)");
}

int fcd_elm_streams(const double* G, const int64_t* base3_host, const int64_t* count3_host,
                    const double* D, int64_t nd, const fcd_theta* theta_host, int32_t want_grad,
                    double* out3, double* ws, void* stream) {
    FCD_REQUIRE(G != nullptr && base3_host != nullptr && count3_host != nullptr && theta_host != nullptr &&
                out3 != nullptr && ws != nullptr && (nd == 0 || D != nullptr), "fcd_elm_streams: NULL argument");
    FCD_REQUIRE((reinterpret_cast<uintptr_t>(G) & 15) == 0 && base3_host[0] % 2 == 0 && base3_host[1] % 2 == 0 &&
                base3_host[2] % 2 == 0 && (reinterpret_cast<uintptr_t>(D) & 31) == 0,
                "fcd_elm_streams: streams must start 16-byte aligned (even bases), records 32-byte aligned");
    const ThetaDev th = make_theta_dev(*theta_host, 0);
    cudaStream_t st = (cudaStream_t)stream;
    LogTabWindow tab;
    FCD_REQUIRE(log_table_window(th.epsl, th.al, st, tab, true), "fcd_elm_streams: log table initialisation failed");
    const bool fast = log_table_covers(th.epsl, th.al);
    const size_t tbytes = fast ? (size_t)((tab.n + 15) & ~15) * sizeof(double) : 0;
    const size_t smem = tbytes + (size_t)kEvWarps * kEvDepth * (kEvChunk * 8 + 8);
    const long long chunks = (count3_host[0] + kEvChunk - 1) / kEvChunk + (count3_host[1] + kEvChunk - 1) / kEvChunk +
                             (count3_host[2] + kEvChunk - 1) / kEvChunk + (nd * 4 + kEvChunk - 1) / kEvChunk;
    long long grid = (chunks + kEvWarps - 1) / kEvWarps;
    if (grid > sm_count()) grid = sm_count();                // one persistent CTA per SM
    if (grid < 1) grid = 1;
#define FCD_ES(G_, F_)                                                                                   \
    do {                                                                                                 \
        cudaFuncSetAttribute(elm_streams_tma_kernel<G_, F_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             (int)(kLogTabBytes + 128 + (size_t)kEvWarps * kEvDepth * (kEvChunk * 8 + 8))); \
        elm_streams_tma_kernel<G_, F_><<<(unsigned)grid, kEvThreads, smem, st>>>(                        \
            G, base3_host[0], base3_host[1], base3_host[2], count3_host[0], count3_host[1], count3_host[2], \
            reinterpret_cast<const Record*>(D), nd, th, tab, out3, ws);                                   \
    } while (0)
    if (want_grad) { if (fast) FCD_ES(true, true); else FCD_ES(true, false); }
    else           { if (fast) FCD_ES(false, true); else FCD_ES(false, false); }
#undef FCD_ES
    return check_launch("fcd_elm_streams");
}

}  // extern "C"
