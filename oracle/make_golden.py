"""
TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by executing the
REFERENCE's own step functions (loaded unmodified-in-arithmetic through
``oracle/ref_compat.py``) on seeded inputs.  Run in the build container, where
/root/reference exists:

    python -m oracle.make_golden

The committed vectors travel to the GPU box; /root/reference does not.

Files written
-------------
ref_unit_vectors.npz  inputs + reference outputs of the step functions at the
                      sizes/seeds the reference's own unit tests use
                      (test_fcdiff/test_fit.py:131-166, 233-557, 790-1087).
cfg1_run.npz          config 1 (N=10, H=U=20): b, bt and the full trajectory of
                      the repaired run() driver.
cfg2_run.npz          config 2 (N=90, H=U=50): trajectory only; the inputs are
                      regenerated from the seed (``golden_inputs``) and checked
                      against a stored checksum.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import iar_oracle as O      # noqa: E402
from oracle import ref_compat           # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


# -- helpers identical in behaviour to test_fcdiff/test_fit.py:25-64 -----------
def rand(lower, upper, shape, seed=0):
    return np.random.RandomState(seed).uniform(lower, upper, size=shape)


def rand_prob(shape, seed=0):
    return rand(1e-7, 1, shape, seed=seed)


def rand_prob_vector(shape, seed=0):
    prob = rand_prob(shape, seed=seed)
    prob /= np.sum(prob, axis=-1, keepdims=True)
    return prob


def golden_inputs(N, H, U, seed=0, theta=None):
    """Synthetic correlations for configs 1-2: model defaults
    (fcdiff/model.py:33-38), vectorised sampler in util edge order, seed 0."""
    th = theta or O.Theta()
    rng = np.random.RandomState(seed)
    (_, _, _, _, b, bt) = O.sample(th, N, H, U, rng)
    return np.ascontiguousarray(b), np.ascontiguousarray(bt)


def checksum(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return np.array([a.sum(), np.abs(a).sum(), (a * np.arange(1, a.size + 1).reshape(a.shape)).sum()])


def ideal_model(ref):
    m = ref.UnsharedRegionModel()
    m.pi = 0.1
    m.epsilon = 0.01
    m.eta = 0.3
    m.gamma = np.ones((3,)) / 3
    m.mu = np.array([-0.5, 0, 0.5])
    m.sigma = np.ones((3,)) * 0.05
    return m


def unit_vectors(ref):
    out = {}
    F = ref.fit
    # --- _update_lps (test_fit.py:131-166)
    (N, C, H, U) = (4, 6, 7, 5)
    fit = F.UnsharedRegionFit()
    fit.b = 1 - 2 * rand_prob((C, H), seed=0)
    fit.bt = 1 - 2 * rand_prob((C, U), seed=1)
    fit.model = ideal_model(ref)
    fit._init_lps(N, H, U)
    fit._update_lps()
    out.update(lps_b=fit.b, lps_bt=fit.bt, lps_lp_B_g_F=fit._lp_B_g_F,
               lps_p_Bt_g_Ft=fit._p_Bt_g_Ft, lps_lM=fit._lM)
    # --- _eval_M for all (k, l) (test_fit.py:233-386)
    p = rand_prob((3, 4, 3), seed=3)
    out["M_p"] = p
    out["M_eta_eps"] = np.array([0.3, 0.01])
    out["M_out"] = np.stack([np.stack([F._eval_M(p, 0.3, 0.01, k, l) for l in range(3)], -1)
                             for k in range(3)], -2)               # (3,4,3k,3l)
    # --- _update_lq_F (test_fit.py:428-467)
    (N, H, U) = (6, 5, 4)
    C = ref.N_to_C(N)
    q_R = rand_prob_vector((N, U, 2))
    lp_B_g_F = np.log(rand_prob((C, H, 3)))
    lM = np.log(rand_prob((C, U, 3, 3)))
    fit = F.UnsharedRegionFit()
    fit._lq_R = np.log(q_R)
    fit._lp_B_g_F = lp_B_g_F
    fit._lM = lM
    fit.model = ref.UnsharedRegionModel()
    fit.model.gamma = rand_prob_vector((3,))
    fit._update_lq_F()
    out.update(lqF_q_R=q_R, lqF_lp_B_g_F=lp_B_g_F, lqF_lM=lM, lqF_gamma=fit.model.gamma,
               lqF_out=fit._lq_F)
    # --- _update_lq_R (test_fit.py:470-510)
    (N, U) = (6, 4)
    pi = rand_prob_vector((2,))
    q_R = rand_prob_vector((N, U, 2))
    q_F = rand_prob_vector((C, 1, 3))
    fit = F.UnsharedRegionFit()
    fit._lq_R = np.log(q_R)
    fit._lq_F = np.log(q_F)
    fit._lM = lM
    fit.model = ref.UnsharedRegionModel()
    fit.model.pi = pi
    fit._update_lq_R()
    out.update(lqR_pi=pi, lqR_q_R=q_R, lqR_q_F=q_F, lqR_lM=lM, lqR_out=fit._lq_R)
    # --- _update_pi / _update_gamma (test_fit.py:513-557)
    fit = F.UnsharedRegionFit()
    fit.model = ref.UnsharedRegionModel()
    fit._lq_R = np.log(q_R)
    fit._lq_F = np.log(q_F)
    fit._update_pi()
    fit._update_gamma()
    out.update(pi_out=np.float64(fit.model.pi), gamma_out=fit.model.gamma)
    # --- free-energy terms (test_fit.py:170-230, 389-425)
    lq_F = np.log(q_F)
    lq_R = np.log(q_R)
    gamma = rand_prob_vector((3,), seed=5)
    lpB = np.log(rand_prob((C, 5, 3), seed=6))
    out.update(E_gamma=gamma, E_lpB=lpB,
               E_lp_F=np.float64(F._eval_E_lp_F(q_F, gamma)),
               E_lp_B_g_F=np.float64(F._eval_E_lp_B_g_F(q_F, lpB)),
               E_lp_R=np.float64(F._eval_E_lp_R(q_R, pi)),
               E_lM=np.float64(F._eval_E_lM(q_F, q_R, lM)),
               E_lq_F=np.float64(F._eval_E_lq_F(q_F, lq_F)),
               E_lq_R=np.float64(F._eval_E_lq_R(q_R, lq_R)))
    # --- analytic derivatives (test_fit.py:790-1087)
    norm = rand_prob((C, U, 3), seed=7)
    (eta, epsilon) = (0.3, 0.01)
    mix = np.stack([np.stack([F._eval_M(norm, eta, epsilon, k, l) for l in range(3)], -1)
                    for k in range(3)], -2)                         # (C,U,3,3)
    out.update(d_norm=norm, d_mix=mix, d_eta_eps=np.array([eta, epsilon]),
               dE_dh=np.float64(F._eval_dE_dh(q_R, q_F, norm, mix, epsilon)),
               dE_de=np.float64(F._eval_dE_de(q_R, q_F, norm, mix, eta)),
               dlM_dh=np.stack([F._eval_dlM_dh(norm, mix[:, :, k, 2], epsilon, k) for k in range(3)], -1),
               dlM_de=np.stack([np.stack([F._eval_dlM_de(norm, mix[:, :, k, l], eta, k, l)
                                          for l in range(3)], -1) for k in range(3)], -2))
    # --- _eval_q_R_w (fit.py:382-406)
    out["qRw_out"] = F._eval_q_R_w(q_R, 3, 1)
    return out


def run_trajectory(ref, b, bt, max_iters=10, eta_shift=0.1, model=None):
    fit = ref.fit.UnsharedRegionFit()
    fit.b = b
    fit.bt = bt
    fit.model = model or ref.UnsharedRegionModel()
    fit.model.eta += eta_shift          # as the reference's disabled test perturbs (test_fit.py:1373-1377)
    fit.max_iters = max_iters
    theta0 = np.array([fit.model.pi, fit.model.eta, fit.model.epsilon])
    traj = dict(lq_F=[], lq_R=[], pi=[], eta=[], epsilon=[], gamma=[], nfev=[])

    def record(i, f, res):
        traj["lq_F"].append(f._lq_F.copy())
        traj["lq_R"].append(f._lq_R.copy())
        traj["pi"].append(f.model.pi)
        traj["eta"].append(f.model.eta)
        traj["epsilon"].append(f.model.epsilon)
        traj["gamma"].append(np.array(f.model.gamma))
        traj["nfev"].append(res.nfev)

    ref_compat.run_reference(ref, fit, record=record)
    out = {k: np.array(v) for (k, v) in traj.items()}
    out["energy"] = np.array(fit.energy, dtype=np.float64)
    out["theta0"] = theta0
    out["gamma0"] = np.array([0.1, 0.8, 0.1])
    out["mu"] = np.array(fit.model.mu, dtype=np.float64)
    out["sigma"] = np.array(fit.model.sigma, dtype=np.float64)
    return out


def one_iteration_no_opt(ref, b, bt):
    """One loop body with (eta, epsilon) held fixed: pins the E-step, pi, gamma
    and energy kernels without any optimiser in the way."""
    fit = ref.fit.UnsharedRegionFit()
    fit.b = b
    fit.bt = bt
    fit.model = ref.UnsharedRegionModel()
    (C, H) = b.shape
    U = bt.shape[1]
    N = int(ref.util.C_to_N(C))
    fit._init_lps(N, H, U)
    fit._update_lps()
    out = {}
    for it in range(2):
        fit._update_lq_F()
        pi = fit.model.pi
        fit.model.pi = np.array([1 - pi, pi])
        fit._update_lq_R()
        fit.model.pi = pi
        fit._update_pi()
        fit._update_gamma()
        pi = fit.model.pi
        fit.model.pi = np.array([1 - pi, pi])
        e = fit._eval_energy()
        q_F = np.exp(fit._lq_F)
        q_R = np.exp(fit._lq_R)
        terms = np.array([
            ref.fit._eval_E_lp_F(q_F, fit.model.gamma),
            ref.fit._eval_E_lp_B_g_F(q_F, fit._lp_B_g_F),
            ref.fit._eval_E_lp_R(q_R, fit.model.pi),
            ref.fit._eval_E_lM(q_F, q_R, fit._lM),
            ref.fit._eval_E_lq_F(q_F, fit._lq_F),
            ref.fit._eval_E_lq_R(q_R, fit._lq_R)])
        fit.model.pi = pi
        out["it%d_lq_F" % it] = fit._lq_F.copy()
        out["it%d_lq_R" % it] = fit._lq_R.copy()
        out["it%d_pi" % it] = np.float64(fit.model.pi)
        out["it%d_gamma" % it] = np.array(fit.model.gamma)
        out["it%d_energy" % it] = np.float64(e)
        out["it%d_terms" % it] = terms
    # objective + gradient of the (eta, epsilon) sub-problem at two points
    pts = np.array([[0.3, 0.03], [0.55, 0.2]])
    fg = []
    for x in pts:
        f, g = ref_compat.elm_objective_and_grad(ref, fit, q_F, q_R, x)
        fg.append([f, g[0], g[1]])
    out["obj_pts"] = pts
    out["obj_fg"] = np.array(fg)
    return out


def main():
    warnings.simplefilter("ignore")
    ref = ref_compat.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)

    np.savez_compressed(os.path.join(GOLDEN_DIR, "ref_unit_vectors.npz"), **unit_vectors(ref))
    print("wrote ref_unit_vectors.npz")

    b, bt = golden_inputs(10, 20, 20)
    out = run_trajectory(ref, b, bt)
    out.update({"noopt_" + k: v for (k, v) in one_iteration_no_opt(ref, b, bt).items()})
    out["b"] = b
    out["bt"] = bt
    np.savez_compressed(os.path.join(GOLDEN_DIR, "cfg1_run.npz"), **out)
    print("wrote cfg1_run.npz: energy", out["energy"])

    b, bt = golden_inputs(90, 50, 50)
    out = run_trajectory(ref, b, bt)
    keep = {k: out[k] for k in ("pi", "eta", "epsilon", "gamma", "nfev", "energy", "theta0", "gamma0", "mu", "sigma")}
    keep["lq_F_final"] = out["lq_F"][-1]
    keep["lq_R_final"] = out["lq_R"][-1]
    keep["lq_F_first"] = out["lq_F"][0]
    keep["lq_R_first"] = out["lq_R"][0]
    keep.update({"noopt_" + k: v for (k, v) in one_iteration_no_opt(ref, b, bt).items()
                 if not k.startswith("it0_lq")})
    keep["b_checksum"] = checksum(b)
    keep["bt_checksum"] = checksum(bt)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "cfg2_run.npz"), **keep)
    print("wrote cfg2_run.npz: energy", out["energy"])


if __name__ == "__main__":
    main()
