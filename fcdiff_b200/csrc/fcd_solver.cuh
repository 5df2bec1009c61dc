// Device-side bounded Newton solve of the (eta, epsilon) sub-problem.
//
// The reference minimises -E_lM(eta, epsilon) on [1e-5, 1 - 1e-5]^2 with scipy.optimize.minimize
// (fcdiff/fit.py:228-241): a host optimiser that needs the objective back on the CPU after every
// evaluation.  Each evaluation is one reduction kernel over the patient data (K3b); between two
// of them the GPU waits for a launch -> mapped-memory spin -> Python -> launch round trip, and
// with edge shards for one more exchange kernel.  Here the optimiser's step runs in the LAST CTA
// of the evaluation kernel itself: the partial sums are reduced (and exchanged over the peer
// windows when the edges are sharded), one thread takes a safeguarded projected Newton step on
// the two parameters and stores the next iterate in device memory, where the next evaluation --
// already enqueued behind this one -- reads it.  The host enqueues a batch of evaluations, waits
// once, and reads the state; launches after convergence exit at once.
//
// With y_l = A + eps_l D, A = (1 - p)/2, D = (3 p - 1)/2 the mixture weight relative to the
// total density (fit.py:427-430) and eps_0 = 1 - eps, eps_1 = eps, eps_2 = eta eps + (1 - eta)(1 - eps)
// (fit.py:433-444), every term of E_lM is w_l log y_l and, with g_l = D / y_l,
//     G_l = sum w_l g_l,   Q_l = sum w_l g_l^2
// give the gradient (the reference's fit.py:600-697) AND the Hessian of f = -E_lM:
//     df/d eta = -(2 eps - 1) G_2                     df/d eps = G_0 - G_1 - (2 eta - 1) G_2
//     d2f/d eta2 = (2 eps - 1)^2 Q_2                  d2f/d eps2 = Q_0 + Q_1 + (2 eta - 1)^2 Q_2
//     d2f/d eta d eps = (2 eps - 1)(2 eta - 1) Q_2 - 2 G_2
// An evaluation delivers {obj, ge, gh, qa, qb} = {sum w_l log y_l, sum s_l w_l g_l with
// s = (-1, 1, 2 eta - 1), G_2, Q_0 + Q_1, Q_2}: two accumulators more than the gradient alone.
// f is convex along each parameter (a sum of -log of affine functions), not jointly (eps_2 is
// bilinear): the step falls back to the diagonal when the 2 x 2 Hessian is not safely positive
// definite, never moves a parameter more than 3/4 of the way to 0 or 1 (the scale on which
// -log(A + eps D) changes), projects onto the box, and is halved back towards the last accepted
// iterate whenever the objective went up.
#pragma once

#include "fcd_comm.cuh"

namespace fcd {

struct SolverState {
    double x[2];          // (eta, epsilon) the NEXT evaluation is made at; the solution when done
    double xprev[2];      // last accepted iterate
    double fprev;         // objective there
    double f;             // objective at the last accepted iterate (when done: f(x), by the quadratic model
                          //   for a final step below `tol`)
    double g[2];          // gradient at the last evaluated point
    double lo[2], hi[2];  // box
    double tol;           // a Newton step with |dx_i| <= tol * min(x_i, 1 - x_i) for both parameters is taken WITHOUT
                          //   another evaluation
    double step;          // max |dx| of the last step
    int32_t have_prev;    // an accepted iterate exists
    int32_t nfev;         // evaluations consumed
    int32_t nback;        // consecutive halvings
    int32_t done;         // 1: converged; 2: stopped (backtracking exhausted / evaluation budget); 3: exchange timeout
    int32_t max_evals;
    int32_t pad_;
};

// Mapped host copy of the state, written after every evaluation: the state followed by the flag
// the host spins on (= the launch's sequence number).
struct SolverPublished {
    SolverState st;
    unsigned long long seq;
};

constexpr int kSolverVals = 6;        // obj, ge, gh, qa, qb, theta-free part of E_lM

#ifdef __CUDACC__
#define FCD_HD __host__ __device__
#else
#define FCD_HD
#endif

// One optimiser transition from the GLOBAL sums of the evaluation made at st.x.  One thread.
// (Plain C++: also compiled for the host, where the CPU tests drive it with NumPy-made sums.)
FCD_HD inline void solver_step(SolverState& st, const double* sums) {
    const double eta = st.x[0], eps = st.x[1];
    const double obj = sums[0], ge = sums[1], G2 = sums[2], QA = sums[3], Q2 = sums[4], konst = sums[5];
    const double f = -(obj + konst);
    const double te = 2.0 * eps - 1.0, th = 2.0 * eta - 1.0;
    const double g0 = -te * G2, g1 = -ge;
    st.nfev += 1;
    st.g[0] = g0;
    st.g[1] = g1;
    if (!(fabs(f) <= 1.7976931348623157e308) || !(fabs(g0) + fabs(g1) + fabs(QA) + fabs(Q2) <= 1.7976931348623157e308)) {
        st.f = f;                                             // NaN / inf sums (non-finite inputs): nothing to iterate on
        st.done = 2;
        return;
    }
    if (st.have_prev && f > st.fprev + 1e-12 * fabs(st.fprev)) {
        // the objective went up: halve the step back towards the last accepted iterate
        if (++st.nback > 12 || st.nfev >= st.max_evals) {
            st.x[0] = st.xprev[0];
            st.x[1] = st.xprev[1];
            st.f = st.fprev;
            st.done = 2;
            return;
        }
        st.x[0] = st.xprev[0] + 0.5 * (st.x[0] - st.xprev[0]);
        st.x[1] = st.xprev[1] + 0.5 * (st.x[1] - st.xprev[1]);
        return;
    }
    st.nback = 0;
    const double H00 = te * te * Q2, H11 = QA + th * th * Q2, H01 = te * th * Q2 - 2.0 * G2;
    const double x[2] = {eta, eps}, g[2] = {g0, g1}, Hd[2] = {H00, H11};
    bool fr[2];
    const double flat = 1e-14 * (H00 + H11);                  // a parameter the objective does not depend on stays put
    for (int i = 0; i < 2; ++i)                               // active set: at a bound with the gradient pointing out
        fr[i] = !((x[i] <= st.lo[i] && g[i] > 0.0) || (x[i] >= st.hi[i] && g[i] < 0.0)) && Hd[i] > flat;
    double d[2] = {0.0, 0.0};
    const double det = H00 * H11 - H01 * H01;
    if (fr[0] && fr[1] && det > 1e-8 * H00 * H11) {
        d[0] = -(H11 * g0 - H01 * g1) / det;
        d[1] = -(H00 * g1 - H01 * g0) / det;
    } else {
        for (int i = 0; i < 2; ++i)
            if (fr[i]) d[i] = -g[i] / Hd[i];
    }
    // never more than 3/4 of the way to 0 or to 1 (one common factor: the direction is kept)
    double a = 1.0;
    for (int i = 0; i < 2; ++i) {
        const double room = d[i] < 0.0 ? 0.75 * x[i] : 0.75 * (1.0 - x[i]);
        if (fabs(d[i]) > room) a = fmin(a, room / fabs(d[i]));
    }
    double xn[2], step = 0.0;
    bool small = true;
    for (int i = 0; i < 2; ++i) {
        xn[i] = fmin(st.hi[i], fmax(st.lo[i], x[i] + a * d[i]));
        step = fmax(step, fabs(xn[i] - x[i]));
        // RELATIVE to the distance from 0 / 1: the mixture weights are linear in eps_l and the terms with
        // y ~ eps_l dominate the curvature, so f''' / f'' ~ 1 / min(x, 1 - x) and the error a Newton step of
        // relative size t leaves behind is ~ t^2 of the parameter (t = 2e-4: 4e-8, against the 1e-6 of north_star)
        small = small && fabs(xn[i] - x[i]) <= st.tol * fmin(x[i], 1.0 - x[i]);
    }
    st.step = step;
    st.xprev[0] = x[0];
    st.xprev[1] = x[1];
    st.fprev = f;
    st.have_prev = 1;
    st.x[0] = xn[0];
    st.x[1] = xn[1];
    if (small) {
        // final step: second-order model of f at the new point (the cubic term, ~ n t^3 / 3 for n elements with
        // y ~ eps_l, is below 1e-11 of |f| ~ n for t = 2e-4)
        const double e0 = xn[0] - x[0], e1 = xn[1] - x[1];
        st.f = f + (g0 * e0 + g1 * e1) + 0.5 * (H00 * e0 * e0 + 2.0 * H01 * e0 * e1 + H11 * e1 * e1);
        st.done = 1;
        return;
    }
    st.f = f;
    if (st.nfev >= st.max_evals) {                            // budget: stay at the evaluated point
        st.x[0] = x[0];
        st.x[1] = x[1];
        st.done = 2;
    }
}

#ifdef __CUDACC__
// theta_sub-dependent constants of an evaluation, formed by every thread from the iterate
struct SubTheta {
    double eta, epsilon;
    double al[3], bl[3];
};
__device__ __forceinline__ SubTheta sub_theta(double eta, double epsilon) {
    SubTheta t;
    t.eta = eta;
    t.epsilon = epsilon;
    double e2 = eta * epsilon;
    e2 += (1.0 - eta) * (1.0 - epsilon);                      // fit.py:442-443
    const double epsl[3] = {1.0 - epsilon, epsilon, e2};
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        t.al[l] = (1.0 - epsl[l]) * 0.5;
        t.bl[l] = epsl[l] - t.al[l];
    }
    return t;
}

__device__ __forceinline__ void solver_publish(const SolverState& st, SolverPublished* pub, unsigned long long seq) {
    if (pub == nullptr) return;
    const double* src = reinterpret_cast<const double*>(&st);
    double* dst = reinterpret_cast<double*>(&pub->st);
    static_assert(sizeof(SolverState) % 8 == 0, "SolverState is copied as doubles");
#pragma unroll
    for (int i = 0; i < (int)(sizeof(SolverState) / 8); ++i) st_relaxed_sys(dst + i, src[i]);
    __threadfence_system();
    st_release_sys(&pub->seq, seq);
}

// Epilogue of an evaluation kernel, run by ALL threads of the CTA that arrived last: s_sums holds
// this rank's {obj, ge, gh, qa, qb} (shared memory, kSolverVals slots); the theta-free part of
// E_lM is read from konst_dev (this rank's share); exchange, step, publication.
__device__ __forceinline__ void solver_epilogue(double* s_sums, SolverState* state, const double* konst_dev,
                                                const CommPeers& peers, int rank, int world,
                                                SolverPublished* pub, unsigned long long seq) {
    if (threadIdx.x == 0) s_sums[5] = konst_dev ? *konst_dev : 0.0;
    __syncthreads();
    const bool ok = comm_exchange_cta(s_sums, kSolverVals, peers, rank, world, kCommSpinCycles);
    if (threadIdx.x == 0) {
        SolverState st = *state;
        if (ok) solver_step(st, s_sums);
        else st.done = 3;
        *state = st;
        __threadfence();
        solver_publish(st, pub, seq);
    }
}

// A launch made after the solve has finished: publishes the (unchanged) state and exits.
__device__ __forceinline__ bool solver_finished(const SolverState* state, SolverPublished* pub, unsigned long long seq) {
    const int done = *reinterpret_cast<const volatile int32_t*>(&state->done);
    if (done && blockIdx.x == 0 && threadIdx.x == 0) solver_publish(*state, pub, seq);
    return done != 0;
}
#endif  // __CUDACC__

}  // namespace fcd
