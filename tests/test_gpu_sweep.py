"""Replica sweeps (BASELINE.json configs[4]) and posterior summaries on the GPU."""
import numpy as np
import numpy.testing as nptest
import pytest
import torch

from oracle import iar_oracle as O

pytestmark = pytest.mark.gpu

import fcdiff_b200 as fcdiff          # noqa: E402
from fcdiff_b200 import sweep         # noqa: E402


def _problem(N=12, H=14, U=10, seed=3):
    (_, _, _, _, b, bt) = O.sample(O.Theta(), N, H, U, np.random.RandomState(seed))
    corr = np.concatenate([b, bt], axis=1)
    labels = np.r_[np.zeros(H, bool), np.ones(U, bool)]
    return b, bt, corr, labels


def test_permutation_sweep_matches_individual_fits_and_the_oracle():
    (b, bt, corr, labels) = _problem()
    opts = dict(max_iters=3, rel_tol=-1.0)
    res = sweep.permutation_sweep(corr, labels, 4, seed=7, fit_options=opts)
    assert sorted(res) == [0, 1, 2, 3, 4]
    lab = sweep.permuted_labels(labels, 4, seed=7)
    assert np.all(lab.sum(axis=1) == labels.sum())
    for i in (0, 3):
        th = O.Theta()
        out = O.run(np.ascontiguousarray(corr[:, ~lab[i]]), np.ascontiguousarray(corr[:, lab[i]]), th,
                    max_iters=3, rel_tol=-1.0, polish=True)       # the package's default solver converges fully
        nptest.assert_allclose(res[i]["energy"], out["energy"], rtol=1e-6)
        nptest.assert_allclose([res[i]["pi"], res[i]["eta"], res[i]["epsilon"]], [th.pi, th.eta, th.epsilon], rtol=1e-6)
    # replicas can be split over ranks without changing any of them
    parts = {}
    for r in range(2):
        parts.update(sweep.permutation_sweep(corr, labels, 4, seed=7, fit_options=opts, rank=r, world=2, gather=False))
    assert sorted(parts) == sorted(res)
    for i in res:
        assert parts[i]["energy"] == res[i]["energy"]
    assert 0.0 < sweep.permutation_p_value(res) <= 1.0


@pytest.mark.parametrize("N,H,U", [(12, 14, 10), (70, 9, 40)])
def test_replicas_in_flight_do_not_change_results(N, H, U):
    """Two replicas in flight on the GPU (two host threads, two CUDA streams, per-stream reduction workspace /
    publication window / solver block) give bit for bit what one after the other gives."""
    (b, bt, corr, labels) = _problem(N, H, U, seed=N + 1)
    opts = dict(max_iters=4, rel_tol=-1.0)
    one = sweep.permutation_sweep(corr, labels, 6, seed=5, fit_options=opts, streams=1)
    for streams in (2, 3):
        two = sweep.permutation_sweep(corr, labels, 6, seed=5, fit_options=opts, streams=streams)
        assert sorted(one) == sorted(two)
        for i in one:
            assert one[i]["energy"] == two[i]["energy"], (streams, i)
            assert (one[i]["pi"], one[i]["eta"], one[i]["epsilon"]) == (two[i]["pi"], two[i]["eta"], two[i]["epsilon"])
            assert one[i]["expected_anomalous_regions"] == two[i]["expected_anomalous_regions"]


@pytest.mark.parametrize("N,H,U", [(12, 14, 10), (9, 13, 10), (15, 8, 9)])
def test_shared_planes_equal_per_replica_planes(N, H, U):
    """configs[4]: planes built once for all subjects + per-replica column selection give the fit of
    the plain path (own upload, own planes) -- replica 0 is the ordinary fit of (b, bt)."""
    (b, bt, corr, labels) = _problem(N, H, U, seed=N)
    opts = dict(max_iters=3, rel_tol=-1.0)
    shared = sweep.permutation_sweep(corr, labels, 3, seed=2, fit_options=opts, shared_planes=True)
    plain = sweep.permutation_sweep(corr, labels, 3, seed=2, fit_options=opts, shared_planes=False)
    fit = fcdiff.fit.UnsharedRegionFit()
    fit.model = fcdiff.UnsharedRegionModel()
    (fit.b, fit.bt) = (b, bt)
    (fit.max_iters, fit.rel_tol) = (3, -1.0)
    fit.run()
    nptest.assert_allclose(shared[0]["energy"], fit.energy, rtol=1e-12)
    for i in plain:
        nptest.assert_allclose(shared[i]["energy"], plain[i]["energy"], rtol=1e-12)
        for k in ("pi", "eta", "epsilon", "expected_anomalous_regions"):
            nptest.assert_allclose(shared[i][k], plain[i][k], rtol=1e-9)
        nptest.assert_allclose(shared[i]["gamma"], plain[i]["gamma"], rtol=1e-9)
    # the private caches of a shared-input fit are the reference's arrays of its columns
    sp = sweep.SharedPlanes(corr, fcdiff.UnsharedRegionModel())
    f2 = fcdiff.fit.UnsharedRegionFit()
    f2.model = fcdiff.UnsharedRegionModel()
    f2.set_shared_inputs(sp, np.flatnonzero(~labels), np.flatnonzero(labels))
    f2._init_lps(N, H, U)
    f2._update_lps()
    f3 = fcdiff.fit.UnsharedRegionFit()
    f3.model = fcdiff.UnsharedRegionModel()
    (f3.b, f3.bt) = (b, bt)
    f3._init_lps(N, H, U)
    f3._update_lps()
    nptest.assert_array_equal(f2._lM, f3._lM)
    nptest.assert_array_equal(f2._lp_B_g_F, f3._lp_B_g_F)
    with pytest.raises(ValueError):
        f2.update_mu_sigma = True
        f2._update_mu_sigma()


def test_restart_sweep_returns_the_lowest_energy():
    (b, bt, _, _) = _problem()
    (res, best) = sweep.restart_sweep(b, bt, 3, seed=1, fit_options=dict(max_iters=4))
    assert sorted(res) == [0, 1, 2, 3]
    assert res[best]["energy"][-1] == min(r["energy"][-1] for r in res.values())


def test_map_labels_and_ranking():
    (b, bt, _, _) = _problem(N=9, H=8, U=7, seed=5)
    fit = fcdiff.fit.UnsharedRegionFit()
    fit.model = fcdiff.UnsharedRegionModel()
    (fit.b, fit.bt) = (b, bt)
    fit.max_iters = 2
    fit.run()
    nptest.assert_array_equal(fit.map_template(), np.argmax(fit._lq_F[:, 0, :], axis=1))
    nptest.assert_array_equal(fit.map_anomalous_regions(), np.argmax(fit._lq_R, axis=2).astype(bool))
    rank = fit.anomalous_region_ranking()
    assert rank.shape == (7, 9)
    q1 = np.exp(fit._lq_R[:, :, 1])
    for u in range(7):
        assert np.all(np.diff(q1[rank[u], u]) <= 0)
    # ties resolve to the first maximum, like numpy.argmax
    fit._lq_F = np.log(np.full((36, 1, 3), 1.0 / 3))
    nptest.assert_array_equal(fit.map_template(), np.zeros(36, dtype=np.uint8))


def test_time_series_files_to_fit(tmp_path):
    """Input adapter end to end: per-subject CSV time series -> K1 (tensor-core or
    SIMT Gram + Fisher z) -> (b, bt) equal to numpy.corrcoef + arctanh, usable by the fit."""
    from fcdiff_b200 import io as fio
    rng = np.random.RandomState(1)
    (N, T, H, U) = (9, 50, 6, 5)
    paths = []
    for s in range(H + U):
        p = tmp_path / ("sub%02d.csv" % s)
        np.savetxt(p, rng.randn(N, T), delimiter=",")
        paths.append(str(p))
    (b, bt) = fio.correlations_from_files(paths[:H], paths[H:])
    assert b.shape == (36, H) and bt.shape == (36, U)
    ts = fio.load_timeseries(paths)
    nptest.assert_allclose(np.concatenate([b, bt], axis=1), O.corr_fisherz(ts.astype(np.float64)), rtol=1e-5, atol=2e-6)
    fit = fcdiff.fit.UnsharedRegionFit()
    fit.model = fcdiff.UnsharedRegionModel()
    (fit.b, fit.bt) = (b, bt)
    fit.max_iters = 2
    fit.run()
    assert len(fit.energy) >= 2 and np.all(np.isfinite(fit.energy))
