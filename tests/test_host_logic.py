"""CPU: host-side logic of the drop-in package that needs no device -- index
maps, convergence rule, parameter packing, argument validation, lazy caches,
shard arithmetic (cases from test_fcdiff/test_util.py and test_fit.py:72-129)."""
import numpy as np
import numpy.testing as nptest
import pytest

import fcdiff_b200 as fcdiff
from fcdiff_b200.dist import EdgeShards


def test_package_surface():
    # fcdiff/__init__.py:1-5
    for name in ("UnsharedRegionModel", "fit", "N_to_C", "nm_to_c", "c_to_nm"):
        assert hasattr(fcdiff, name)
    import importlib
    alias = importlib.import_module("fcdiff")
    assert alias.fit.UnsharedRegionFit is fcdiff.fit.UnsharedRegionFit
    assert importlib.import_module("fcdiff.util").c_to_nm(5) == (3, 2)


def test_N_to_C_to_N():
    for N in range(2, 10):
        C = fcdiff.util.N_to_C(N)
        assert fcdiff.util.C_to_N(C) == N
    assert fcdiff.util.C_to_N(4) % 1 != 0


def test_nm_to_c_and_back():
    c = 0
    for n in range(60):
        for m in range(n):
            assert fcdiff.util.nm_to_c(n, m) == c
            assert fcdiff.util.c_to_nm(c) == (n, m)
            c += 1
    big = 499500 - 1
    (n, m) = fcdiff.util.c_to_nm(big)
    assert (n, m) == (999, 998)
    assert isinstance(n, int) and isinstance(m, int)


def test_init_lps_shape():
    fit = fcdiff.fit.UnsharedRegionFit()
    (N, C, H, U) = (4, 6, 7, 5)
    fit._init_lps(N, H, U)
    nptest.assert_equal(fit._lq_R.shape, (N, U, 2))
    nptest.assert_allclose(np.sum(np.exp(fit._lq_R), axis=2), 1)
    nptest.assert_equal(fit._lq_F.shape, (C, 1, 3))
    nptest.assert_allclose(np.sum(np.exp(fit._lq_F), axis=2), 1)
    nptest.assert_equal(fit._lp_B_g_F.shape, (C, H, 3))
    nptest.assert_equal(fit._p_Bt_g_Ft.shape, (C, U, 3))
    nptest.assert_equal(fit._lM.shape, (C, U, 3, 3))


@pytest.mark.parametrize("energy,expect", [([1, 1.25], True), ([1, 1], True), ([1, 0.501], True),
                                           ([1, 0.5], False), ([1, 0.499], False)])
def test_is_converged(energy, expect):
    fit = fcdiff.fit.UnsharedRegionFit()
    fit.rel_tol = 0.5
    fit.energy = energy
    assert bool(fit._is_converged(1)) is expect


def test_convergence_rule_with_negative_energy():
    fit = fcdiff.fit.UnsharedRegionFit()
    fit.energy = [-100.0, -110.0]                 # a 10 % decrease of a negative energy
    assert bool(fit._is_converged(1)) is True     # the reference's rule stops here (ratio < 0)
    fit.convergence_rule = "magnitude"
    assert bool(fit._is_converged(1)) is False
    fit.energy = [-100.0, -100.0000001]
    assert bool(fit._is_converged(1)) is True


def test_pack_unpack_theta_sub():
    fit = fcdiff.fit.UnsharedRegionFit()
    fit.model = fcdiff.UnsharedRegionModel()
    nptest.assert_array_equal(fit._pack_theta_sub(), [0.3, 0.03])
    fit._unpack_theta_sub(np.array([0.4, 0.2]))
    assert (fit.model.eta, fit.model.epsilon) == (0.4, 0.2)


def test_run_validation_errors():
    fit = fcdiff.fit.UnsharedRegionFit()
    fit.b = np.zeros((4, 3))
    fit.bt = np.zeros((4, 2))
    fit.model = fcdiff.UnsharedRegionModel()
    with pytest.raises(ValueError, match="triangular"):
        fit.run()
    fit.b = np.zeros((6, 3))
    fit.bt = np.zeros((6, 2))
    fit.model = None
    with pytest.raises(ValueError, match="initialized"):
        fit.run()


def test_defaults_match_reference():
    fit = fcdiff.fit.UnsharedRegionFit()
    assert (fit.max_iters, fit.rel_tol, fit.energy, fit.edge_lookup) == (10, 1e-5, [], "reference")
    m = fcdiff.UnsharedRegionModel()
    assert (m.pi, m.eta, m.epsilon) == (0.05, 0.3, 0.03)
    nptest.assert_array_equal(m.gamma, [0.1, 0.8, 0.1])
    nptest.assert_array_equal(m.mu, [-0.15, 0, 0.3])
    nptest.assert_array_equal(m.sigma, [0.025, 0.035, 0.05])
    assert isinstance(str(m), str)


def test_eval_M_eps():
    f = fcdiff.fit._eval_M_eps
    assert f(0.3, 0.01, 0) == 0.99 and f(0.3, 0.01, 1) == 0.01
    nptest.assert_allclose(f(0.3, 0.01, 2), 0.3 * 0.01 + 0.7 * 0.99)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fcdiff_b200 import _lib
    fit = fcdiff.fit.UnsharedRegionFit()
    fit.b = np.zeros((6, 3))
    fit.bt = np.zeros((6, 2))
    fit.model = fcdiff.UnsharedRegionModel()
    with pytest.raises(_lib.FcdError, match="no CPU fallback"):
        fit.run()
    with pytest.raises(_lib.FcdError):
        fcdiff.UnsharedRegionModel().sample(4, 2, 2)


def test_shard_spans_cover_everything():
    for world in (1, 2, 3, 4, 8):
        for total in (1, 5, 45, 4005, 79800):
            covered = []
            for r in range(world):
                (s, l) = EdgeShards(rank=r, world=world).span(total)
                covered += list(range(s, s + l))
            assert covered == list(range(total))


def test_product_does_not_import_oracle():
    import os
    root = os.path.dirname(os.path.abspath(fcdiff.__file__))
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            src = open(os.path.join(root, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_direct_lbfgsb_driver_is_bit_identical_to_scipy_minimize():
    """fcdiff_b200/_opt.py drives scipy's compiled L-BFGS-B without the Python
    front end; the oracle (and the reference, fit.py:239) call
    scipy.optimize.minimize -- both must produce the same iterates."""
    import scipy.optimize
    from fcdiff_b200 import _opt
    rng = np.random.RandomState(0)
    for _ in range(25):
        A = rng.rand(2, 2)
        A = A @ A.T + 0.1 * np.eye(2)
        c = rng.rand(2) * 2 - 0.5
        sc = 10 ** rng.uniform(0, 8)

        def fun(x):
            d = x - c
            return sc * (0.5 * d @ A @ d - 0.1 * np.sum(np.log(x))), sc * (A @ d - 0.1 / x)

        x0 = rng.rand(2) * 0.8 + 0.1
        r = _opt.minimize_lbfgsb(fun, x0, [1e-5, 1e-5], [1 - 1e-5, 1 - 1e-5])
        n = [0]

        def counted(x):
            n[0] += 1
            return fun(x)

        r2 = scipy.optimize.minimize(counted, x0, jac=True, method="L-BFGS-B", bounds=[(1e-5, 1 - 1e-5)] * 2)
        np.testing.assert_array_equal(r.x, r2.x)
        assert r.nfev == n[0]


def test_replica_split_and_label_permutations():
    """Host logic of the replica sweeps (fcdiff_b200/sweep.py): every replica is
    handled by exactly one rank; permutations preserve the group sizes and are
    reproducible from the seed."""
    import ast
    import os
    src = open(os.path.join(os.path.dirname(__file__), "..", "fcdiff_b200", "sweep.py")).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("replica_indices", "permuted_labels"):
            exec(compile(ast.Module([node], []), "sweep.py", "exec"), ns)
    for (n, world) in ((1001, 8), (5, 8), (16, 4)):
        seen = sorted(i for r in range(world) for i in ns["replica_indices"](n, r, world))
        assert seen == list(range(n))
        sizes = [len(ns["replica_indices"](n, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    labels = np.r_[np.zeros(7, bool), np.ones(5, bool)]
    a = ns["permuted_labels"](labels, 6, seed=3)
    assert a.shape == (7, 12) and np.array_equal(a[0], labels) and np.all(a.sum(axis=1) == 5)
    assert np.array_equal(a, ns["permuted_labels"](labels, 6, seed=3))
    assert not np.array_equal(a[1], a[2])


def test_time_series_file_loaders(tmp_path):
    """Host half of the input adapters (fcdiff_b200/io.py): .npy / .csv / .tsv
    parsing, header detection, time-major transposition, shape checks.  The
    module is loaded without its package (the package needs CUDA at call time
    only, but importing torch is enough of a dependency for this test)."""
    import ast
    import os
    src = open(os.path.join(os.path.dirname(__file__), "..", "fcdiff_b200", "io.py")).read()
    tree = ast.parse(src)
    ns = {"np": np, "os": os}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("load_subject", "load_timeseries"):
            exec(compile(ast.Module([node], []), "io.py", "exec"), ns)
    rng = np.random.RandomState(0)
    a = rng.randn(5, 12).astype(np.float32)
    np.save(tmp_path / "s0.npy", a)
    np.savetxt(tmp_path / "s1.csv", a, delimiter=",", header="t0,t1", comments="")     # header line
    np.savetxt(tmp_path / "s2.tsv", a.T, delimiter="\t")                                # time-major
    got0 = ns["load_subject"](str(tmp_path / "s0.npy"))
    got1 = ns["load_subject"](str(tmp_path / "s1.csv"))
    got2 = ns["load_subject"](str(tmp_path / "s2.tsv"), time_major=True)
    for g in (got0, got1, got2):
        assert g.dtype == np.float32 and g.shape == (5, 12)
        np.testing.assert_allclose(g, a, rtol=1e-6)
    ts = ns["load_timeseries"]([str(tmp_path / "s0.npy"), str(tmp_path / "s1.csv")])
    assert ts.shape == (2, 5, 12)
    np.save(tmp_path / "bad.npy", a[:, :7])
    with pytest.raises(ValueError):
        ns["load_timeseries"]([str(tmp_path / "s0.npy"), str(tmp_path / "bad.npy")])


def test_uniform_start_is_lazy_on_the_host():
    """_init_lps (fit.py:84-102) records a fill; the host array appears only when somebody reads it,
    the values are the reference's, assignments replace it and bump the version."""
    fit = fcdiff.fit.UnsharedRegionFit()
    fit._init_lps(5, 3, 4)
    (mR, mF) = (fit._mR, fit._mF)
    assert mR.host is None and mR.fill == ((5, 4, 2), -np.log(2)) and mF.fill == ((10, 1, 3), -np.log(3))
    v = mR.version
    nptest.assert_array_equal(fit._lq_R, np.full((5, 4, 2), -np.log(2)))
    nptest.assert_array_equal(fit._lq_F, np.full((10, 1, 3), -np.log(3)))
    fit._lq_R = np.zeros((5, 4, 2))
    assert mR.fill is None and mR.version == v + 1 and not fit._lq_R.any()


def test_padded_edge_chunks_for_the_in_place_gather():
    for (C, world) in [(45, 2), (79800, 8), (7, 3), (10, 1)]:
        sh = EdgeShards(rank=0, world=world)
        n = sh.edge_buffer_len(C)
        ch = sh.chunk(C, world)
        assert n == world * ch * 3 and n >= 3 * C
        for r in range(world):
            (start, length) = EdgeShards(rank=r, world=world).span(C)
            assert start == min(r * ch, C) and 0 <= length <= ch and (start + length) * 3 <= n


def test_small_vector_exchange_abi_limits():
    """The exchange window covers 16 ranks x 8 doubles (include/fcdiff_b200.h 'small exchanges')."""
    from fcdiff_b200 import _lib
    lib = _lib.load()
    assert lib.fcd_comm_max_world() >= 8 and lib.fcd_comm_max_vals() >= 6
    assert lib.fcd_comm_window_bytes() >= 2 * lib.fcd_comm_max_world() * (lib.fcd_comm_max_vals() + 1) * 8
    assert lib.fcd_comm_handle_bytes() == 64
    assert lib.fcd_code_pitch(500) == 512 and lib.fcd_code_pitch(16) == 16 and lib.fcd_code_pitch(17) == 32


def test_expect_stop_predicts_the_convergence_rules():
    """`_expect_stop` decides whether run() launches the next iteration's E-step ahead (fit.py:124-140 is the
    rule it predicts): never with the test disabled, always once the energy is negative under the reference's
    signed rule, under the magnitude rule when the last relative decrease is within 4 x of the tolerance."""
    fit = fcdiff.fit.UnsharedRegionFit()
    fit.rel_tol = -1.0
    fit.energy = [6.0e8, -1.4e8]
    assert not fit._expect_stop()
    fit.rel_tol = 1e-5
    assert fit._expect_stop()                      # reference rule: any decrease of a negative energy stops
    fit.energy = [6.0e8]
    assert not fit._expect_stop()                  # positive energy, no history
    fit.energy = [6.0e8, 5.9e8]
    assert not fit._expect_stop()                  # positive energy still falling fast
    fit.convergence_rule = "magnitude"
    fit.energy = [6.0e8, -1.4e8]
    assert not fit._expect_stop()
    fit.energy = [-1.4452e8, -1.44521e8]           # relative decrease 7e-6 < 4e-5
    assert fit._expect_stop()
    fit.energy = [-1.44e8, -1.45e8]                # 7e-3
    assert not fit._expect_stop()
    fit.energy = [float("nan")]
    assert not fit._expect_stop()


def test_speculative_estep_handle_is_dropped_on_any_change():
    """`_spec_matches`: the E-step launched behind the device solve is adopted only for the same inputs / planes /
    q_R and exactly the published (eta, epsilon); no GPU involved -- the handle is a dict."""
    from fcdiff_b200 import _lib
    fit = fcdiff.fit.UnsharedRegionFit()
    fit._dims = (4, 3, 5)
    inp = {"cache_key": ("k",), "code_verR": fit._mR.version}
    th = _lib.make_theta(0.05, 0.3, 0.03, [0.1, 0.8, 0.1], [-0.15, 0.0, 0.3], [0.025, 0.035, 0.05])
    rest = (tuple(th.gamma), tuple(th.mu), tuple(th.sigma))
    spec = dict(bufs=None, inp=inp, cache_key=("k",), verR=fit._mR.version, dims=(4, 3, 5), rest=rest, x=(0.3, 0.03))
    assert fit._spec_matches(spec, inp, th)
    assert not fit._spec_matches(dict(spec, x=None), inp, th)                   # the solve had not finished
    assert not fit._spec_matches(dict(spec, x=(0.3, 0.030000001)), inp, th)      # another solution
    assert not fit._spec_matches(spec, dict(inp), th)                            # other input set
    assert not fit._spec_matches(dict(spec, cache_key=("other",)), inp, th)      # planes rebuilt
    th2 = _lib.make_theta(0.05, 0.3, 0.03, [0.2, 0.7, 0.1], [-0.15, 0.0, 0.3], [0.025, 0.035, 0.05])
    assert not fit._spec_matches(spec, inp, th2)                                 # gamma changed
    fit._mR.set_host(np.zeros((4, 5, 2)))                                        # q_R assigned: version moves on
    assert not fit._spec_matches(spec, inp, th)
    fit.coded_estep = False
    assert not fit._spec_matches(dict(spec, verR=fit._mR.version), dict(inp, code_verR=fit._mR.version), th)


def test_mirror_completes_a_pending_gather_exactly_once():
    """Edge shards inside run(): lq_F's other-rank rows are gathered lazily (`_Mirror.finish`).  The pending
    collective runs once, before the array is read whole, and any new assignment drops it."""
    from fcdiff_b200.fit import _Mirror
    calls = []
    m = _Mirror()
    m.set_dev("lq", "q", (3, 1, 3), complete=lambda: calls.append(1))
    m.finish()
    m.finish()
    assert calls == [1]
    m.set_dev("lq", "q", (3, 1, 3), complete=lambda: calls.append(2))
    m.set_dev("lq2", "q2", (3, 1, 3))                 # a newer E-step result: the old array is never completed
    m.finish()
    assert calls == [1]
    m.set_dev("lq", "q", (3, 1, 3), complete=lambda: calls.append(3))
    m.set_host(np.zeros((3, 1, 3)))                   # user assignment
    m.finish()
    assert calls == [1] and m.complete is None
