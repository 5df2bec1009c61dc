"""CPU, build container only: the oracle against the real reference loaded
in memory (oracle/ref_compat.py).  Skipped where /root/reference is absent (the
GPU box); the committed golden vectors carry the same pin there."""
import numpy as np
import numpy.testing as nptest
import pytest

from oracle import iar_oracle as O
from oracle import ref_compat

pytestmark = pytest.mark.skipif(not ref_compat.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_compat.load()


def _inputs(N, H, U, seed):
    th = O.Theta()
    (_, _, _, _, b, bt) = O.sample(th, N, H, U, np.random.RandomState(seed))
    return b, bt


@pytest.mark.parametrize("N,H,U,seed", [(5, 3, 4, 1), (8, 7, 5, 2), (12, 9, 11, 3)])
def test_step_functions(ref, N, H, U, seed):
    (b, bt) = _inputs(N, H, U, seed)
    rng = np.random.RandomState(seed + 100)
    fit = ref.fit.UnsharedRegionFit()
    fit.b, fit.bt = b, bt
    fit.model = ref.UnsharedRegionModel()
    fit._init_lps(N, H, U)
    fit._update_lps()
    th = O.Theta()
    (lpB, pBt, lM) = O.update_lps(b, bt, th)
    nptest.assert_array_equal(lpB, fit._lp_B_g_F)
    nptest.assert_array_equal(lM, fit._lM)
    # random normalised posteriors
    C = O.N_to_C(N)
    q_R = rng.dirichlet([1, 1], size=(N, U))
    q_F = rng.dirichlet([1, 1, 1], size=(C, 1))
    fit._lq_R = np.log(q_R)
    fit._lq_F = np.log(q_F)
    fit._update_lq_F()
    nptest.assert_allclose(O.update_lq_F(th.gamma, lpB, lM, np.log(q_R)), fit._lq_F, rtol=1e-11, atol=1e-12)
    lqF = fit._lq_F.copy()
    fit.model.pi = np.array([0.95, 0.05])
    fit._update_lq_R()
    nptest.assert_allclose(O.update_lq_R(np.array([0.95, 0.05]), lqF, lM, np.log(q_R)), fit._lq_R,
                           rtol=1e-11, atol=1e-12)
    e_ref = fit._eval_energy()
    th.pi = 0.05
    terms = O.eval_energy_terms(th, fit._lq_F, fit._lq_R, lpB, lM)
    nptest.assert_allclose(O.energy_from_terms(terms), e_ref, rtol=1e-12)
    fit.model.pi = 0.05
    (f_ref, g_ref) = ref_compat.elm_objective_and_grad(ref, fit, np.exp(fit._lq_F), np.exp(fit._lq_R), [0.4, 0.1])
    (f, g) = O.elm_objective_and_grad(pBt, np.exp(fit._lq_F), np.exp(fit._lq_R), [0.4, 0.1])
    nptest.assert_allclose(f, f_ref, rtol=1e-12)
    nptest.assert_allclose(g, g_ref, rtol=1e-10)


def test_full_run_small(ref):
    (b, bt) = _inputs(7, 6, 9, 11)
    fit = ref.fit.UnsharedRegionFit()
    fit.b, fit.bt = b, bt
    fit.model = ref.UnsharedRegionModel()
    fit.model.eta += 0.1
    fit.max_iters = 4
    ref_compat.run_reference(ref, fit)
    th = O.Theta()
    th.eta += 0.1
    out = O.run(b, bt, th, max_iters=4)
    nptest.assert_allclose(out["energy"], fit.energy, rtol=1e-9)
    nptest.assert_allclose(out["lq_F"], fit._lq_F, rtol=1e-6, atol=1e-8)
    nptest.assert_allclose(out["lq_R"], fit._lq_R, rtol=1e-6, atol=1e-8)


def test_reference_errors(ref):
    fit = ref.fit.UnsharedRegionFit()
    fit.b = np.zeros((4, 2))
    fit.bt = np.zeros((4, 2))
    with pytest.raises(ValueError):
        ref_compat.run_reference(ref, fit)
    with pytest.raises(ValueError):
        O.run(np.zeros((4, 2)), np.zeros((4, 2)), O.Theta())
