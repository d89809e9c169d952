"""Multi-GPU ensemble path (chemtools_b200/sharded.py over ct_ensemble_*).

CPU tier: the partition rule (host arithmetic, no device).  GPU tier: a sharded, chunked run on min(2, device_count)
devices is bit-identical to one batch on one device — for both partitions, with the recovery pass, for the table
reactor — and its gathered host arrays go through the HDF5 writer."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def test_partition_rule_matches_survey(ctb):
    """block: reactor i -> shard floor(i * G / n) in contiguous slices; cyclic: i -> i % G.
    Both are exact disjoint covers; the library's rule equals the Python one used by bench.py under torchrun."""
    from chemtools_b200 import sharded
    from chemtools_b200.workloads import shard_range
    for n in (0, 1, 5, 7, 4096, 65537, 16777216):
        for G in (1, 2, 3, 4, 8):
            seen_b, seen_c = 0, 0
            for g in range(G):
                f, s, c = sharded.partition_range(n, G, g, "block")
                lo, hi = shard_range(n, g, G)
                assert s == 1 and c == hi - lo and (c == 0 or f == lo)
                if c and n:
                    assert (f * G) // n == g and ((f + c - 1) * G) // n == g
                seen_b += c
                f, s, c = sharded.partition_range(n, G, g, "cyclic")
                assert s == G and c == len(range(g, n, G)) and (c == 0 or f == g)
                seen_c += c
            assert seen_b == n and seen_c == n
    with pytest.raises(ctb.reactors.CtError):
        sharded.partition_range(10, 2, 2, "block")


def _h2_air(chem, T0, phi):
    X = np.zeros((len(T0), chem.n_species))
    X[:, chem.speciesIndex("H2")] = 2.0 * np.asarray(phi)
    X[:, chem.speciesIndex("O2")] = 1.0
    X[:, chem.speciesIndex("N2")] = 3.76
    return chem.mass_fractions(X)


def _single(ctb, ch, rt, T0, p0, Y0, rtol, atol, t_end, mx, retry=False, table=None):
    b = ctb.ReactorBatch(rt, ch, T0, p0, Y0, rtol, atol, t_end, max_num_steps=mx)
    if table is not None:
        b.SetTable(*table)
    code = b.IntegrateWithRetry()["code"] if retry else b.Integrate(t_end)
    return b.Solution(), b.IgnitionDelay(), b.Status(), b.Statistics(), code


@pytest.mark.gpu
@pytest.mark.parametrize("partition", ["block", "cyclic"])
def test_sharded_equals_single_batch_bit_for_bit(ctb, chem, partition):
    from chemtools_b200 import sharded
    ch = chem("gri30")
    n = 3001  # odd: ragged shards and a ragged last chunk
    rng = np.random.default_rng(7)
    T0 = rng.uniform(950.0, 1500.0, n); phi = rng.uniform(0.5, 2.0, n); p0 = np.full(n, 101325.0)
    Y0 = _h2_air(ch, T0, phi)
    devs = list(range(min(2, ctb.device_count())))
    res = sharded.integrate(ch, ctb.Type.Const_Vol, T0, p0, Y0, 1e-8, 1e-14, 1e-3, devices=devs, max_num_steps=20000,
                            partition=partition, chunk=700, retry=False)
    st, tau, status, stats, code = _single(ctb, ch, ctb.Type.Const_Vol, T0, p0, Y0, 1e-8, 1e-14, 1e-3, 20000)
    assert np.array_equal(res["state"], st)
    assert np.array_equal(res["tau"], tau, equal_nan=True)
    assert np.array_equal(res["status"], status)
    assert np.array_equal(res["stats"], stats)
    assert res["code"] == code and code >= 0 and all(s["failed_final"] == 0 for s in res["shards"])
    assert sum(s["reactors"] for s in res["shards"]) == n and len(res["shards"]) == len(devs)
    per = -(-n // len(devs))
    assert all(s["chunks"] == -(-s["reactors"] // 700) and s["reactors"] <= per for s in res["shards"])
    assert all(s["launches"] >= 2 * s["chunks"] and s["kernel_ms"] > 0 for s in res["shards"])
    # scheduling options of the ensemble change nothing either: one launch in flight per device, no cost ordering
    e = sharded.Ensemble(ch, ctb.Type.Const_Vol, n, devs, partition, 1200)
    e.SetTolerances(1e-8, 1e-14); e.SetStopTime(1e-3); e.SetMaxNumSteps(20000); e.SetRetry(False); e.SetStreams(1); e.SetHotFirst(False)
    res1 = e.Integrate(T0, p0, Y0)
    assert np.array_equal(res1["state"], st) and np.array_equal(res1["stats"], stats) and np.array_equal(res1["tau"], tau, equal_nan=True)
    # a small step budget for the chunk launches: the reactors cut off there are integrated again with the full budget in
    # the shard's recovery batch; same answer, and the counters show that the second pass really ran
    e.SetRetry(True); e.SetFirstPassSteps(int(np.median(stats[:, 0])))
    res2 = e.Integrate(T0, p0, Y0)
    assert np.array_equal(res2["state"], st) and np.array_equal(res2["stats"], stats) and np.array_equal(res2["status"], status)
    cut = sum(s["failed_first"] for s in res2["shards"])
    assert n // 4 < cut < 3 * n // 4 and all(s["failed_final"] == 0 for s in res2["shards"])


@pytest.mark.gpu
def test_sharded_recovery_pass_and_options(ctb, chem):
    """The per-chunk recovery pass (ct_batch_integrate_with_retry) gives what the single batch's recovery pass gives, and
    an option applied through the configure hook (difference-quotient Jacobian + LU, the reference's configuration)
    reaches every internal batch."""
    from chemtools_b200 import sharded
    L = ctb.reactors.lib()
    ch = chem("gri12")
    n = 600
    rng = np.random.default_rng(3)
    T0 = rng.uniform(1100.0, 1800.0, n); p0 = np.exp(rng.uniform(np.log(0.5), np.log(20.0), n)) * 101325.0
    X = np.zeros((n, ch.n_species)); X[:, ch.speciesIndex("CH4")] = rng.uniform(0.5, 2.0, n); X[:, ch.speciesIndex("O2")] = 2.0; X[:, ch.speciesIndex("N2")] = 7.52
    Y0 = ch.mass_fractions(X)
    devs = list(range(min(2, ctb.device_count())))
    res = sharded.integrate(ch, ctb.Type.Const_Pres, T0, p0, Y0, 1e-8, 1e-14, 1e-3, devices=devs, max_num_steps=20000, chunk=250, retry=True)
    assert (res["status"] >= 0).all() and sum(s["failed_final"] for s in res["shards"]) == 0
    st, tau, status, stats, code = _single(ctb, ch, ctb.Type.Const_Pres, T0, p0, Y0, 1e-8, 1e-14, 1e-3, 20000, retry=True)
    assert res["code"] == code
    # retried reactors run in differently composed small batches, under the same ladder: same arithmetic per reactor
    assert np.array_equal(res["state"], st) and np.array_equal(res["tau"], tau, equal_nan=True) and np.array_equal(res["stats"], stats)

    def dq_lu(handle):
        L.ct_batch_set_jacobian(handle, 0)
        L.ct_batch_set_linear_solver(handle, 0)
    sub = slice(0, 96)
    res2 = sharded.integrate(ch, ctb.Type.Const_Pres, T0[sub], p0[sub], Y0[sub], 1e-8, 1e-14, 1e-3, devices=devs, max_num_steps=20000, chunk=40,
                             retry=False, configure=dq_lu)
    b = ctb.ReactorBatch(ctb.Type.Const_Pres, ch, T0[sub], p0[sub], Y0[sub], 1e-8, 1e-14, 1e-3, max_num_steps=20000)
    b.SetJacobian(0); b.SetLinearSolver(0)
    b.Integrate(1e-3)
    assert np.array_equal(res2["state"], b.Solution()) and np.array_equal(res2["stats"], b.Statistics())
    assert res2["stats"][:, 7].min() > 0  # nfeLS > 0: difference quotients really ran


@pytest.mark.gpu
def test_sharded_table_reactor_and_hdf5(ctb, chem, tmp_path):
    from chemtools_b200 import sharded, workloads as W
    from h5_min_reader import H5Min
    ch = chem("gri30")
    n = 500
    mech, rt, Ta, pa, X, t_end, extra = W.config4(ch, n, 21, 30)
    Y0 = ch.mass_fractions(X)
    devs = list(range(min(2, ctb.device_count())))
    res = sharded.integrate(ch, rt, Ta, pa, Y0, 1e-8, 1e-14, t_end, devices=devs, max_num_steps=20000, partition="cyclic", chunk=128,
                            retry=False, table=extra["table"])
    st, tau, status, stats, _ = _single(ctb, ch, rt, Ta, pa, Y0, 1e-8, 1e-14, t_end, 20000, table=extra["table"])
    assert np.array_equal(res["state"], st) and np.array_equal(res["status"], status) and np.array_equal(res["stats"], stats)
    f = str(tmp_path / "ensemble.h5")
    sharded.write_hdf5(f, ch, rt, res, T0=Ta, p0=pa, extra={"/Initial conditions/Equivalence ratio": (extra["phi"], "phi", "-")})
    h = H5Min(f).root
    assert np.array_equal(h["Mass fractions"]["CH4"]["data"], st[:, ch.speciesIndex("CH4")])
    assert np.array_equal(h["Solver"]["Status"]["data"], status.astype(float))
    assert np.array_equal(h["Initial conditions"]["Equivalence ratio"]["data"], extra["phi"])
    assert h["Initial conditions"]["Temperature"]["attrs"]["Unit"] == "K"


@pytest.mark.gpu
def test_ensemble_errors_are_reported_per_shard(ctb, chem):
    from chemtools_b200 import sharded
    ch = chem("gri30")
    with pytest.raises(ctb.reactors.CtError):
        sharded.Ensemble(ch, 0, 10, devices=[ctb.device_count()])
    with pytest.raises(ctb.reactors.CtError):
        sharded.Ensemble(ch, 0, 10, devices=[0, 0])
    e = sharded.Ensemble(ch, 4, 4, devices=[0])
    with pytest.raises(ctb.reactors.CtError) as err:  # table reactor without a table
        e.Integrate(np.full(4, 1000.0), np.full(4, 1e5), np.full((4, ch.n_species), 1.0 / ch.n_species))
    assert err.value.code == -22
