"""Pins the oracle's double-precision kinetics arithmetic against an independent 40-digit evaluation
(tests/mp_reference.py, written from oracle/mech/*.json and the formulas of SURVEY.md section 8c): the reference-held
golden vectors resolve ~1e-5, this resolves 1e-13.  Bars: q_f, q_r <= 1e-13 relative; net production rates <= 1e-12 of
each species' gross rate (net rates cancel near partial equilibrium); ydot likewise."""
import numpy as np
import pytest

from mp_reference import MpMechanism


def _states(om, oracle_mod, name, n_rand, seed):
    """n_rand random hot / cold states + near-equilibrium ones (burnt gas from the oracle's own integration)."""
    rng = np.random.default_rng(seed)
    K = om.K
    Y = rng.random((n_rand, K)) ** 4 * (rng.random((n_rand, K)) < 0.7)
    Y[:, om.index["N2"]] += 1.0
    Y /= Y.sum(1, keepdims=True)
    T = rng.uniform(700.0, 3000.0, n_rand)
    T[:3] = rng.uniform(280.0, 420.0, 3)
    p = np.exp(rng.uniform(np.log(0.3), np.log(40.0), n_rand)) * 101325.0
    # near-equilibrium: stoichiometric and rich fuel/air burnt to 5 ms at constant pressure
    fuel = "H2" if name == "gri30" else "CH4"
    X = np.zeros((4, K))
    for i, phi in enumerate((0.7, 1.0, 1.0, 1.6)):
        X[i, om.index[fuel]] = (2.0 if fuel == "H2" else 1.0) * phi
        X[i, om.index["O2"]] = 1.0 if fuel == "H2" else 2.0
        X[i, om.index["N2"]] = 3.76 if fuel == "H2" else 7.52
    Ye = om.X_to_Y(X)
    Te0 = np.array([1400.0, 1500.0, 1800.0, 1600.0]); pe = np.array([1.0, 1.0, 20.0, 5.0]) * 101325.0
    r = om.integrate_batch("Const_Pres", Te0, pe, Ye, 5e-3, rtol=1e-10, atol=1e-16, max_steps=200000, nthreads=4)
    assert (r["status"] >= 0).all() and (r["state"][:, 0] > Te0 + 500).all()
    T = np.concatenate([T, r["state"][:, 0]]); p = np.concatenate([p, pe]); Y = np.concatenate([Y, r["state"][:, 1:]])
    return T, p, Y


@pytest.mark.parametrize("name,rtype", [("gri30", 0), ("gri30", 1), ("gri30", 2), ("gri211", 0), ("gri12", 1), ("gri12", 3)])
def test_oracle_rhs_against_mpmath(oracle, omech, name, rtype):
    om = omech(name)
    mpm = MpMechanism(name, oracle.CONSTANT_SETS[om.constants])
    n_rand = 16
    T, p, Y = _states(om, oracle, name, n_rand, 11 + rtype)
    n = len(T)
    rng = np.random.default_rng(5)
    Ys = Y * rng.uniform(0.9, 1.1, Y.shape)
    Ys[:n_rand][rng.random((n_rand, om.K)) < 0.03] *= -1e-6  # slightly negative entries: setMassFractions clips them
    Ys[n_rand:] = Y[n_rand:]                                  # the near-equilibrium states stay as they are
    Ts = T * np.concatenate([rng.uniform(0.95, 1.05, n_rand), np.ones(n - n_rand)])
    st = np.concatenate([Ts[:, None], Ys], 1) if rtype <= 1 else Ys
    ydot, wdot, qf, qr = om.rhs_batch(rtype, T, p, Y, st)
    worst = dict(qf=0.0, qr=0.0, wdot=0.0, ydot=0.0, near_eq_cancellation=1.0)
    for i in range(n):
        m_ydot, m_wdot, m_qf, m_qr = mpm.rhs(rtype, T[i], p[i], Y[i], st[i])
        for key, got, ref in (("qf", qf[i], m_qf), ("qr", qr[i], m_qr)):
            for a, b in zip(got, ref):
                if b == 0:
                    assert a == 0.0
                else:
                    worst[key] = max(worst[key], abs(float((a - b) / b)))
        gross = [0.0] * om.K  # per species: sum |nu| (q_f + q_r)
        for r, f, b in zip(mpm.rx, m_qf, m_qr):
            for k, nu in list(r["nu_r"].items()) + list(r["nu_p"].items()):
                gross[k] += nu * (f + b)
        off = 1 if rtype <= 1 else 0
        rho_scale = [float(abs(yd / w)) if w != 0 else 0.0 for yd, w in zip(m_ydot[off:], m_wdot)]
        for k in range(om.K):
            if gross[k] > 0:
                worst["wdot"] = max(worst["wdot"], abs(float((wdot[i, k] - m_wdot[k]) / gross[k])))
                if rho_scale[k] > 0:
                    worst["ydot"] = max(worst["ydot"], abs(float(ydot[i, off + k] - m_ydot[off + k])) / (rho_scale[k] * float(gross[k])))
                if i >= n_rand and abs(m_wdot[k]) > 0:
                    worst["near_eq_cancellation"] = min(worst["near_eq_cancellation"], abs(float(m_wdot[k] / gross[k])))
        if rtype <= 1:  # dT/dt = -sum_k e_k wdot_k / (rho c): relative to the gross size of its terms
            scale = sum(abs(float(e)) * float(gk) for e, gk in zip(mpm.last["e_molar"], gross)) / float(mpm.last["rho_c"])
            worst["Tdot"] = max(worst.get("Tdot", 0.0), abs(float(ydot[i, 0] - m_ydot[0])) / scale)
    print("oracle vs mpmath %s rtype %d: %s" % (name, rtype, worst))
    assert worst["qf"] <= 1e-13 and worst["qr"] <= 1e-13, worst
    assert worst["wdot"] <= 1e-12 and worst["ydot"] <= 1e-12 and worst.get("Tdot", 0.0) <= 1e-12, worst
    assert worst["near_eq_cancellation"] < 1e-3  # the near-equilibrium states really do cancel
