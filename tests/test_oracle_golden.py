"""CPU: the NumPy restatement (oracle/iar_oracle.py) against the golden vectors
that oracle/make_golden.py produced by executing the REFERENCE's own step
functions (fcdiff/fit.py) -- the pin SURVEY 8c asks for."""
import numpy as np
import numpy.testing as nptest
import pytest

from oracle import iar_oracle as O
from oracle.make_golden import golden_inputs, checksum

RTOL = 1e-12


def ideal():
    return O.Theta.ideal()


def test_update_lps_matches_reference(unit_vectors):
    g = unit_vectors
    (lpB, pBt, lM) = O.update_lps(g["lps_b"], g["lps_bt"], ideal())
    # the reference's own test asserts bit-equality (test_fit.py:164-166)
    nptest.assert_array_equal(lpB, g["lps_lp_B_g_F"])
    nptest.assert_array_equal(pBt, g["lps_p_Bt_g_Ft"])
    nptest.assert_array_equal(lM, g["lps_lM"])


def test_eval_M_all_kl(unit_vectors):
    g = unit_vectors
    (eta, eps) = g["M_eta_eps"]
    for k in range(3):
        for l in range(3):
            nptest.assert_array_equal(O.eval_M(g["M_p"], eta, eps, k, l), g["M_out"][:, :, k, l])


def test_update_lq_F(unit_vectors):
    g = unit_vectors
    out = O.update_lq_F(g["lqF_gamma"], g["lqF_lp_B_g_F"], g["lqF_lM"], np.log(g["lqF_q_R"]))
    nptest.assert_allclose(out, g["lqF_out"], rtol=RTOL)


def test_update_lq_R_gauss_seidel_and_quirk(unit_vectors):
    g = unit_vectors
    out = O.update_lq_R(g["lqR_pi"], np.log(g["lqR_q_F"]), g["lqR_lM"], np.log(g["lqR_q_R"]), "reference")
    nptest.assert_allclose(out, g["lqR_out"], rtol=RTOL)
    sym = O.update_lq_R(g["lqR_pi"], np.log(g["lqR_q_F"]), g["lqR_lM"], np.log(g["lqR_q_R"]), "symmetric")
    assert np.max(np.abs(sym - g["lqR_out"])) > 1e-3      # the quirk is observable


def test_pi_gamma(unit_vectors):
    g = unit_vectors
    nptest.assert_allclose(O.update_pi(np.log(g["lqR_q_R"])), g["pi_out"], rtol=RTOL)
    nptest.assert_allclose(O.update_gamma(np.log(g["lqR_q_F"])), g["gamma_out"], rtol=RTOL)


def test_energy_terms(unit_vectors):
    g = unit_vectors
    (q_F, q_R, lM) = (g["lqR_q_F"], g["lqR_q_R"], g["lqR_lM"])
    nptest.assert_allclose(O.eval_E_lp_F(q_F, g["E_gamma"]), g["E_lp_F"], rtol=RTOL)
    nptest.assert_allclose(O.eval_E_lp_B_g_F(q_F, g["E_lpB"]), g["E_lp_B_g_F"], rtol=RTOL)
    nptest.assert_allclose(O.eval_E_lp_R(q_R, g["lqR_pi"]), g["E_lp_R"], rtol=RTOL)
    nptest.assert_allclose(O.eval_E_lM(q_F, q_R, lM), g["E_lM"], rtol=RTOL)
    nptest.assert_allclose(O.eval_E_lq_F(q_F, np.log(q_F)), g["E_lq_F"], rtol=RTOL)
    nptest.assert_allclose(O.eval_E_lq_R(q_R, np.log(q_R)), g["E_lq_R"], rtol=RTOL)


def test_derivatives(unit_vectors):
    g = unit_vectors
    (eta, eps) = g["d_eta_eps"]
    (q_F, q_R) = (g["lqR_q_F"], g["lqR_q_R"])
    nptest.assert_allclose(O.eval_dE_dh(q_R, q_F, g["d_norm"], g["d_mix"], eps), g["dE_dh"], rtol=1e-11)
    nptest.assert_allclose(O.eval_dE_de(q_R, q_F, g["d_norm"], g["d_mix"], eta), g["dE_de"], rtol=1e-11)
    for k in range(3):
        nptest.assert_allclose(O.eval_dlM_dh(g["d_norm"], g["d_mix"][:, :, k, 2], eps, k),
                               g["dlM_dh"][:, :, k], rtol=RTOL)
        for l in range(3):
            nptest.assert_allclose(O.eval_dlM_de(g["d_norm"], g["d_mix"][:, :, k, l], eta, k, l),
                                   g["dlM_de"][:, :, k, l], rtol=RTOL)


def test_q_R_w(unit_vectors):
    g = unit_vectors
    w = O.eval_q_R_w(g["lqR_q_R"], np.array([3]), np.array([1]))[0]
    nptest.assert_array_equal(w, g["qRw_out"])


def _theta0(g):
    th = O.Theta(mu=g["mu"], sigma=g["sigma"], gamma=g["gamma0"])
    (th.pi, th.eta, th.epsilon) = [float(v) for v in g["theta0"]]
    return th


def test_cfg1_full_run_trajectory(cfg1):
    g = cfg1
    th = _theta0(g)
    rec = []
    out = O.run(g["b"], g["bt"], th, record=lambda i, t, lqF, lqR, e, nfev: rec.append(
        (lqF.copy(), lqR.copy(), t.pi, t.eta, t.epsilon, t.gamma.copy())))
    nptest.assert_allclose(out["energy"], g["energy"], rtol=1e-10)
    assert len(rec) == len(g["pi"])
    for i, (lqF, lqR, pi, eta, eps, gamma) in enumerate(rec):
        nptest.assert_allclose(lqF, g["lq_F"][i], rtol=1e-7, atol=1e-9)
        nptest.assert_allclose(lqR, g["lq_R"][i], rtol=1e-7, atol=1e-9)
        nptest.assert_allclose([pi, eta, eps], [g["pi"][i], g["eta"][i], g["epsilon"][i]], rtol=1e-7)
        nptest.assert_allclose(gamma, g["gamma"][i], rtol=1e-9)


def test_cfg1_no_optimiser_iterations(cfg1):
    g = cfg1
    th = O.Theta()
    (lqF, lqR) = O.init_lps(10, 20)
    (lpB, pBt, lM) = O.update_lps(g["b"], g["bt"], th)
    for it in range(2):
        (lqF, lqR, lM, e, _) = O.em_iteration(g["b"], g["bt"], th, lqF, lqR, lpB, pBt, lM, optimise=False)
        nptest.assert_allclose(lqF, g["noopt_it%d_lq_F" % it], rtol=1e-9, atol=1e-11)
        nptest.assert_allclose(lqR, g["noopt_it%d_lq_R" % it], rtol=1e-9, atol=1e-11)
        nptest.assert_allclose(e, g["noopt_it%d_energy" % it], rtol=1e-12)
        nptest.assert_allclose(th.pi, g["noopt_it%d_pi" % it], rtol=1e-12)
    (q_F, q_R) = (np.exp(lqF), np.exp(lqR))
    for x, fg in zip(g["noopt_obj_pts"], g["noopt_obj_fg"]):
        (f, grad) = O.elm_objective_and_grad(pBt, q_F, q_R, x)
        nptest.assert_allclose([f, grad[0], grad[1]], fg, rtol=1e-10)


def test_cfg2_inputs_and_run(cfg2):
    g = cfg2
    (b, bt) = golden_inputs(90, 50, 50)
    nptest.assert_allclose(checksum(b), g["b_checksum"], rtol=1e-13)
    nptest.assert_allclose(checksum(bt), g["bt_checksum"], rtol=1e-13)
    th = _theta0(g)
    out = O.run(b, bt, th)
    nptest.assert_allclose(out["energy"], g["energy"], rtol=1e-9)
    nptest.assert_allclose(out["lq_F"], g["lq_F_final"], rtol=1e-6, atol=1e-8)
    nptest.assert_allclose(out["lq_R"], g["lq_R_final"], rtol=1e-6, atol=1e-8)
    nptest.assert_allclose([th.pi, th.eta, th.epsilon], [g["pi"][-1], g["eta"][-1], g["epsilon"][-1]], rtol=1e-6)
    # MAP labels bit-exact
    nptest.assert_array_equal(np.argmax(out["lq_F"], axis=2), np.argmax(g["lq_F_final"], axis=2))
    nptest.assert_array_equal(out["lq_R"][:, :, 1] > out["lq_R"][:, :, 0],
                              g["lq_R_final"][:, :, 1] > g["lq_R_final"][:, :, 0])


def test_is_converged_semantics():
    # test_fcdiff/test_fit.py:86-129
    assert O.is_converged([1, 1.25], 1, 0.5)
    assert O.is_converged([1, 1], 1, 0.5)
    assert O.is_converged([1, 0.501], 1, 0.5)
    assert not O.is_converged([1, 0.5], 1, 0.5)
    assert not O.is_converged([1, 0.499], 1, 0.5)


def test_util_maps():
    # test_fcdiff/test_util.py:5-36
    c = 0
    for n in range(10):
        for m in range(n):
            assert O.nm_to_c(n, m) == c
            assert O.c_to_nm(c) == (n, m)
            c += 1
    for N in range(2, 10):
        assert O.C_to_N(O.N_to_C(N)) == N
    (n, m) = O.edge_pairs(10)
    assert [O.c_to_nm(c) for c in range(45)] == list(zip(n.tolist(), m.tolist()))


def test_sampler_moments():
    # distributional checks of test_fcdiff/test_model.py:37-245 on the vectorised sampler
    th = O.Theta()
    rng = np.random.RandomState(0)
    (r, t, f, ft, b, bt) = O.sample(th, 30, 40, 60, rng)
    assert r.shape == (30, 60) and t.shape == (435, 60) and f.shape == (435, 3)
    assert ft.shape == (435, 60, 3) and b.shape == (435, 40) and bt.shape == (435, 60)
    assert np.all(f.sum(1) == 1) and np.all(ft.sum(2) == 1)
    nptest.assert_allclose(r.mean(), th.pi, atol=0.02)
    nptest.assert_allclose(f.mean(0), th.gamma, atol=0.05)
    assert b.min() >= -1 and b.max() <= 1
    k = np.argmax(f, 1)
    for s in range(3):
        nptest.assert_allclose(b[k == s].mean(), th.mu[s], atol=0.01)
        nptest.assert_allclose(b[k == s].std(), th.sigma[s], atol=0.01)


def test_corr_fisherz_oracle():
    rng = np.random.RandomState(1)
    ts = rng.standard_normal((3, 6, 50)).astype(np.float32)
    z = O.corr_fisherz(ts)
    assert z.shape == (15, 3)
    r = np.corrcoef(ts[1].astype(np.float64))
    nptest.assert_allclose(z[O.nm_to_c(4, 2), 1], np.arctanh(r[4, 2]), rtol=1e-13)
