"""bench.py host logic on CPU: both arms describe the same workload (same `config` object, same initial conditions although
the reference arm builds them from oracle/ + numpy only), and at N > 1 the ranks' slices are a disjoint cover of ONE global
index set (the library's cyclic partition rule)."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("_bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name,batch", [("config2", 0), ("config5", 4096), ("config3_gri12", 512), ("config4", 256)])
def test_arms_agree_and_ranks_partition_one_index_set(bench, ctb, chem, oracle, name, batch):
    W = bench.load_workloads()
    assert "chemtools_b200.workloads" not in (getattr(W, "__name__", ""),)  # loaded as a plain file
    ocache = {}

    def ochem(m):
        if m not in ocache:
            ocache[m] = bench.OracleChem(m)
        return ocache[m]
    for world in (1, 2, 4):
        ours = [bench.workload(W, name, chem, batch, r, world) for r in range(world)]
        ref0 = bench.workload(W, name, ochem, batch, 0, world)
        # same config object and same rank-0 initial conditions in both arms
        assert bench.config_of(ours[0], world, bench.RTOL, bench.ATOL) == bench.config_of(ref0, world, bench.RTOL, bench.ATOL)
        assert np.array_equal(ours[0]["T0"], ref0["T0"]) and np.array_equal(ours[0]["p0"], ref0["p0"])
        assert np.abs(ours[0]["Y"] - ref0["Y"]).max() < 1e-15
        # the ranks' slices: equal sizes, disjoint, together the global set
        n = len(ours[0]["T0"])
        assert all(len(w["T0"]) == n for w in ours) and ours[0]["n_global"] == n * world
        if name == "config5":
            allidx = np.concatenate([w["extra"]["map_index"] for w in ours])
            assert len(np.unique(allidx)) == len(allidx) == n * world
            design = W.config5_design(n * world)
            for r, w in enumerate(ours):
                assert np.array_equal(w["extra"]["map_index"], design[r::world])  # reactor i of the global set -> rank i % world
        if name == "config2":
            rows = np.concatenate([np.stack([w["T0"], w["Y"][:, 0]], 1) for w in ours])
            assert len(np.unique(rows, axis=0)) == n * world  # every (T0, phi) point exactly once
    # world 1 = the workload BASELINE.json names
    one = bench.workload(W, name, chem, batch, 0, 1)
    if name == "config2":
        assert len(one["T0"]) == 4096 and one["T0"].min() == 900.0 and one["T0"].max() == 1500.0 and one["rt"] == 1


def test_cyclic_indices_match_the_library_rule(bench, ctb):
    from chemtools_b200 import sharded
    for n, world in ((4096, 1), (8192, 2), (65536 * 8, 8), (10, 4)):
        for r in range(world):
            first, stride, count = sharded.partition_range(n, world, r, "cyclic")
            idx = bench.cyclic_indices(n, world, r)
            assert len(idx) == count and (count == 0 or (idx[0] == first and (len(idx) < 2 or idx[1] - idx[0] == stride)))
