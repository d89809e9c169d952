"""Golden fixture for BASELINE config 5 (the north-star workload): the ORACLE (oracle/, the CPU restatement of the
reference's Cantera + CVODE path, dense difference-quotient Jacobian like Direct::BindUserFunctions) integrated on a
128-reactor design over the 256^3 constant-pressure GRI-3.0 CH4/air map, at the reference test's tolerances
(rtol 1e-9 / atol 1e-15, /root/reference/tests/src/reactor.cpp:62-63), at the headline setting (1e-8 / 1e-14) and at a
tight setting (1e-11 / 1e-17) that serves as "truth" for the two.

Run here (CPU, ~ minutes): python tools/make_config5_golden.py -> tests/golden/config5_oracle.npz
The design = the 8 corners of the map + 120 points of a 3-D Kronecker (additive-recurrence) sequence, so all of
T0 in [1000, 2000] K, p in [1, 40] atm and phi in [0.5, 2] is covered, extremes included.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

ONE_ATM = 101325.0
N_SIDE = 256


def design(n=128):
    """Global reactor indices (iT * 65536 + ip * 256 + iphi) of the parity sample."""
    pts = [(a, b, c) for a in (0, N_SIDE - 1) for b in (0, N_SIDE - 1) for c in (0, N_SIDE - 1)]
    g = 1.22074408460575947536  # real root of x^4 = x + 1: the 3-D analogue of the golden ratio
    al = np.array([1.0 / g, 1.0 / g ** 2, 1.0 / g ** 3])
    k = np.arange(1, n - len(pts) + 1)[:, None]
    u = (0.5 + k * al[None, :]) % 1.0
    for row in np.floor(u * N_SIDE).astype(int):
        pts.append(tuple(int(v) for v in row))
    idx = np.array([a * N_SIDE * N_SIDE + b * N_SIDE + c for a, b, c in pts], dtype=np.int64)
    assert len(np.unique(idx)) == n
    return idx


def conditions(idx):
    iT, ip, iphi = idx // (N_SIDE * N_SIDE), (idx // N_SIDE) % N_SIDE, idx % N_SIDE
    T0 = np.linspace(1000.0, 2000.0, N_SIDE)[iT]
    p0 = np.geomspace(1.0, 40.0, N_SIDE)[ip] * ONE_ATM
    phi = np.linspace(0.5, 2.0, N_SIDE)[iphi]
    return T0, p0, phi


def main():
    from oracle import oracle as O
    om = O.Mechanism("gri30")
    idx = design(128)
    T0, p0, phi = conditions(idx)
    X = np.zeros((len(idx), om.K))
    X[:, om.index["CH4"]] = phi; X[:, om.index["O2"]] = 2.0; X[:, om.index["N2"]] = 7.52
    Y0 = om.X_to_Y(X)
    out = dict(idx=idx, T0=T0, p0=p0, phi=phi, Y0=Y0)
    for tag, rtol, atol in (("ref", 1e-9, 1e-15), ("head", 1e-8, 1e-14), ("tight", 1e-11, 1e-17)):
        t0 = time.time()
        r = om.integrate_batch("Const_Pres", T0, p0, Y0, 1e-3, rtol=rtol, atol=atol, max_steps=200000, nthreads=os.cpu_count() or 1)
        print("%s rtol %.0e: %.1f s, failed %d, mean nst %.0f max nst %d" % (tag, rtol, time.time() - t0, int((r["status"] < 0).sum()),
                                                                               r["stats"][:, 0].mean(), r["stats"][:, 0].max()), flush=True)
        out["state_" + tag] = r["state"]; out["tau_" + tag] = r["tau"]; out["status_" + tag] = r["status"]; out["stats_" + tag] = r["stats"]
    for tag in ("ref", "head"):
        ign = np.isfinite(out["tau_tight"]) & np.isfinite(out["tau_" + tag])
        print("oracle %s vs tight: max tau dev %.2e, max T dev %.2e" % (tag, np.abs(out["tau_" + tag][ign] / out["tau_tight"][ign] - 1).max(),
                                                                        np.abs(out["state_" + tag][:, 0] / out["state_tight"][:, 0] - 1).max()))
    dst = os.path.join(ROOT, "tests", "golden", "config5_oracle.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst)


if __name__ == "__main__":
    main()
