"""
CPU tests of the (eta, epsilon) solver's state machine (csrc/fcd_solver.cuh): the
SAME ``solver_step`` the evaluation kernels' last CTA runs is compiled for the host
(``fcd_solver_step_host``) and driven here with sums formed by NumPy from the
oracle's arrays -- no GPU, no compute kernel.  Checks the closed forms of the
gradient / Hessian sums and the safeguards against the oracle's polished optimum.
"""
import ctypes

import numpy as np
import numpy.testing as nptest
import pytest

from oracle import iar_oracle as O
from fcdiff_b200 import _lib


def _sums(p3, L, q_F, q_R, eta, eps):
    """{obj, ge, G2, QA, Q2, konst} of csrc/fcd_solver.cuh from dense arrays: p3 (C,U,3)
    responsibilities, L (C,U) log total density."""
    C = q_F.shape[0]
    (n, m) = O.edge_pairs(q_R.shape[0])
    w = O.eval_q_R_w(q_R, n[:C], m[:C])                                  # (C,U,3)
    epsl = [1 - eps, eps, eta * eps + (1 - eta) * (1 - eps)]
    sl = [-1.0, 1.0, 2 * eta - 1]
    out = np.zeros(6)
    for k in range(3):
        D = 1.5 * p3[:, :, k] - 0.5
        for l in range(3):
            W = q_F[:, 0, k][:, None] * w[:, :, l]
            y = (1 - epsl[l]) / 2 + (epsl[l] - (1 - epsl[l]) / 2) * p3[:, :, k]
            g = D / y
            out[0] += np.sum(W * np.log(y))
            out[1] += sl[l] * np.sum(W * g)
            if l == 2:
                out[2] += np.sum(W * g)
                out[4] += np.sum(W * g * g)
            else:
                out[3] += np.sum(W * g * g)
            out[5] += np.sum(W * L) if l >= 0 else 0.0
    return out


def _solve(p, q_F, q_R, x0, tol=1e-7, max_evals=60, lo=1e-5):
    lib = _lib.load()
    S = p.sum(axis=2)
    p3 = p / S[:, :, None]
    L = np.log(S)
    st = _lib.SolverState()
    lo2 = _lib.d3([lo, lo])
    hi2 = _lib.d3([1 - lo, 1 - lo])
    assert lib.fcd_solver_init_host(ctypes.byref(st), x0[0], x0[1], lo2, hi2, tol, max_evals) == 0
    trace = []
    while not st.done:
        s = _sums(p3, L, q_F, q_R, st.x[0], st.x[1])
        trace.append((st.x[0], st.x[1], -(s[0] + s[5])))
        assert lib.fcd_solver_step_host(ctypes.byref(st), _lib.d3(list(s))) == 0
    return st, trace


def _problem(N, H, U, seed, th=None):
    th = th or O.Theta()
    (_, _, _, _, b, bt) = O.sample(th, N, H, U, np.random.RandomState(seed))
    thf = O.Theta()
    (lq_F, lq_R) = O.init_lps(N, U)
    (lpB, p, lM) = O.update_lps(b, bt, thf)
    lq_F = O.update_lq_F(thf.gamma, lpB, lM, lq_R)
    lq_R = O.update_lq_R(np.array([1 - thf.pi, thf.pi]), lq_F, lM, lq_R)
    return p, np.exp(lq_F), np.exp(lq_R)


def test_struct_layout_matches_the_library():
    lib = _lib.load()
    assert ctypes.sizeof(_lib.SolverState) == lib.fcd_solver_state_bytes()
    assert lib.fcd_solver_published_bytes() == lib.fcd_solver_state_bytes() + 8


def test_sums_reproduce_the_reference_objective_and_gradient():
    (p, q_F, q_R) = _problem(9, 6, 11, 1)
    S = p.sum(axis=2)
    for (eta, eps) in ((0.3, 0.03), (0.7, 0.4), (0.05, 0.9)):
        s = _sums(p / S[:, :, None], np.log(S), q_F, q_R, eta, eps)
        (f, g) = O.elm_objective_and_grad(p, q_F, q_R, [eta, eps])
        nptest.assert_allclose(-(s[0] + s[5]), f, rtol=1e-12)
        nptest.assert_allclose([-(2 * eps - 1) * s[2], -s[1]], g, rtol=1e-9, atol=1e-9 * abs(f))
        # Hessian of csrc/fcd_solver.cuh against central differences of the reference's analytic gradient
        (te, th) = (2 * eps - 1, 2 * eta - 1)
        Hm = np.array([[te * te * s[4], te * th * s[4] - 2 * s[2]], [te * th * s[4] - 2 * s[2], s[3] + th * th * s[4]]])
        h = 1e-6
        fd = np.zeros((2, 2))
        for i in range(2):
            (xp, xm) = (np.array([eta, eps]), np.array([eta, eps]))
            xp[i] += h
            xm[i] -= h
            fd[:, i] = (O.elm_objective_and_grad(p, q_F, q_R, xp)[1] - O.elm_objective_and_grad(p, q_F, q_R, xm)[1]) / (2 * h)
        nptest.assert_allclose(Hm, fd, rtol=2e-6, atol=1e-6 * np.abs(fd).max())


@pytest.mark.parametrize("N,H,U,seed", [(10, 20, 20, 0), (16, 8, 30, 3), (30, 20, 25, 5)])
@pytest.mark.parametrize("x0", [(0.4, 0.03), (0.6, 0.3), (0.02, 0.0004), (0.7, 0.002)])
def test_newton_state_machine_finds_the_polished_optimum(N, H, U, seed, x0):
    (p, q_F, q_R) = _problem(N, H, U, seed)
    (st, trace) = _solve(p, q_F, q_R, x0)
    ref = O.minimize_eta_epsilon(lambda x: O.elm_objective_and_grad(p, q_F, q_R, x), x0, polish=True)
    assert st.done == 1 and st.nfev <= 14, (st.done, st.nfev, trace)
    nptest.assert_allclose([st.x[0], st.x[1]], ref.x, rtol=1e-7)
    nptest.assert_allclose(st.f, ref.fun, rtol=1e-12)
    assert trace[-1][2] <= trace[0][2]


def test_newton_state_machine_stops_on_active_bounds():
    """pi, epsilon ~ 0 in the data: the minimiser sits on the box (SciPy's L-BFGS-B agrees)."""
    (p, q_F, q_R) = _problem(12, 10, 30, 8, O.Theta(epsilon=1e-9, pi=0.0001))
    (st, trace) = _solve(p, q_F, q_R, (0.3, 0.03))
    ref = O.minimize_eta_epsilon(lambda x: O.elm_objective_and_grad(p, q_F, q_R, x), (0.3, 0.03), polish=True)
    assert st.done == 1
    nptest.assert_allclose([st.x[0], st.x[1]], ref.x, rtol=1e-6, atol=1e-12)
    nptest.assert_allclose(st.f, ref.fun, rtol=1e-10)


def test_budget_and_non_finite_sums_stop_the_solve():
    lib = _lib.load()
    st = _lib.SolverState()
    lo2 = _lib.d3([1e-5, 1e-5])
    hi2 = _lib.d3([1 - 1e-5, 1 - 1e-5])
    lib.fcd_solver_init_host(ctypes.byref(st), 0.3, 0.03, lo2, hi2, 1e-7, 60)
    lib.fcd_solver_step_host(ctypes.byref(st), _lib.d3([float("nan"), 0, 0, 1, 1, 0]))
    assert st.done == 2 and st.nfev == 1
    (p, q_F, q_R) = _problem(10, 20, 20, 0)
    (st, _) = _solve(p, q_F, q_R, (0.02, 0.0004), max_evals=2)
    assert st.done == 2 and st.nfev == 2
