"""CPU tier: the product's kernel SOURCE (chemtools_b200/csrc/kernels.cuh) compiled with g++ under the lock-step
SIMT emulation of tests/simt_emu and checked against the oracle and the reference's golden vectors.  This is how the
warp-level logic (shuffles, shared-memory carving, device BDF stepper, work queue) is exercised where no GPU exists.
The emulation library is test infrastructure; the product never loads it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_longlong)


def P(a):
    return None if a is None else a.ctypes.data_as(dp)


@pytest.fixture(scope="session")
def emu():
    d = os.path.join(ROOT, "tests", "simt_emu")
    so = os.path.join(d, "libct_emu.so")
    srcs = [os.path.join(d, "emu_driver.cpp"), os.path.join(d, "simt_emu.h")] + [
        os.path.join(ROOT, "chemtools_b200", "csrc", f) for f in ("kernels.cuh", "cti_parser.cpp", "mech_compile.cpp", "mechanism.hpp", "mech_shapes.inc")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-D__align__(x)=alignas(x)", "-o", so,
                               srcs[0], srcs[3], srcs[4]])
    L = C.CDLL(so)
    L.emu_mech_load.restype = C.c_void_p
    L.emu_mech_load.argtypes = [C.c_char_p, C.c_char_p]
    L.emu_last_error.restype = C.c_char_p
    L.emu_rhs.argtypes = [C.c_void_p, C.c_int, C.c_longlong] + [dp] * 8
    L.emu_jacobian.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong] + [dp] * 5
    L.emu_lu_solve.argtypes = [C.c_int, C.c_longlong, dp, dp, dp, ip, C.c_int]
    L.emu_integrate.argtypes = [C.c_void_p, C.c_int, C.c_longlong, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_longlong,
                                C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, lp, ip, C.c_int,
                                C.c_int, C.c_double, C.c_int, C.c_int, dp]
    L.emu_roberts.argtypes = [C.c_int, dp, C.c_double, dp, dp, lp]
    L.emu_set_stats_ex_buffer.argtypes = [dp]
    L.emu_static_shape.argtypes = [C.c_void_p]
    L.emu_use_static_shapes.argtypes = [C.c_int]
    return L


@pytest.fixture(scope="session")
def emech(emu, ctb):
    cache = {}

    def get(name):
        if name not in cache:
            h = emu.emu_mech_load(ctb.mechanism_path(name).encode(), None)
            assert h, emu.emu_last_error()
            cache[name] = h
        return cache[name]
    return get


OPTION_KEYS = ("init_step", "min_step", "max_step", "nonlin_conv_coef", "max_err_test_fails", "max_nonlin_iters",
               "max_conv_fails", "max_order")


def _integrate(emu, h, om, rt, T0, p0, Y, rtol, atol, t_end, jac=1, linsol=0, n_out=0, n_calls=1, warps=1, mxstep=20000,
               ign=(1, 400.0), pool_n=0, slice_steps=0, options=None, table=None):
    n = len(T0)
    N = om.K + 1 if rt in (0, 1, 5) else om.K
    T0 = np.ascontiguousarray(T0, dtype=float); p0 = np.ascontiguousarray(p0, dtype=float); Y = np.ascontiguousarray(Y, dtype=float)
    nt = max(n_out, n_calls if n_calls > 1 else 0)
    traj = np.zeros((n, nt, N)) if nt else None
    state = np.zeros((n, N)); tau = np.zeros(n); stats = np.zeros((n, 8), dtype=np.int64); status = np.zeros(n, dtype=np.int32)
    opts = None
    if options:
        opts = np.array([options.get(k, 0.0) for k in OPTION_KEYS], dtype=float)
    tab = (0, None, None, None)
    if table is not None:
        tab_arrays = [np.ascontiguousarray(a, dtype=float) for a in table]
        tab = (len(tab_arrays[0]), P(tab_arrays[0]), P(tab_arrays[1]), P(tab_arrays[2]))
    emu.emu_integrate(h, rt, n, P(T0), P(p0), P(Y), rtol, atol, t_end, mxstep, jac, ign[0], ign[1], n_out, n_calls, tab[0], tab[1], tab[2], tab[3],
                      P(traj), P(state), P(tau), stats.ctypes.data_as(lp), status.ctypes.data_as(ip), warps, linsol, 1.0, pool_n, slice_steps, P(opts))
    return dict(state=state, tau=tau, stats=stats, status=status, traj=traj)


@pytest.mark.parametrize("name", ["gri30", "gri12"])
def test_rhs_kernel_source_vs_oracle(emu, emech, omech, name):
    om, h = omech(name), emech(name)
    rng = np.random.default_rng(1)
    n, K = 4, om.K
    for rt in (0, 1, 2, 3):
        T0 = rng.uniform(900, 2500, n); p0 = rng.uniform(0.5, 20, n) * 101325
        T0[0] = 320.0  # below 400 K: Cantera's own 1/Kc arithmetic (recip_kc slow path)
        Y = rng.random((n, K)) ** 4 * (rng.random((n, K)) < 0.7)
        Y[:, om.index["N2"]] += 1
        Y /= Y.sum(1, keepdims=True)
        st = np.ascontiguousarray(np.concatenate([T0[:, None], Y], 1) if rt <= 1 else Y)
        yd, wd, qf, qr = om.rhs_batch(rt, T0, p0, Y, st)
        yd2, wd2, qf2, qr2 = np.zeros_like(yd), np.zeros_like(wd), np.zeros_like(qf), np.zeros_like(qr)
        emu.emu_rhs(h, rt, n, P(T0), P(p0), P(np.ascontiguousarray(Y)), P(st), P(yd2), P(wd2), P(qf2), P(qr2))
        assert (np.abs(qf2 - qf) / (np.abs(qf) + 1e-300)).max() < 1e-12
        assert (np.abs(qr2 - qr) / (np.abs(qr) + 1e-300)).max() < 1e-12
        assert (np.abs(yd2 - yd) / np.abs(yd).max(1, keepdims=True)).max() < 1e-12


@pytest.mark.parametrize("name,shape", [("gri30", 1), ("gri211", 2), ("gri12", 3)])
def test_static_shapes_match_the_mechanism_compiler(emu, emech, omech, name, shape):
    """chemtools_b200/csrc/mech_shapes.inc (sizes and blob offsets as compile-time constants, generated at build time) is
    what the mechanism compiler produces for the shipped mechanisms, and the specialised instantiations give the same
    bits as the run-time shape: right-hand side with rates, and a fused integration (pool layout, two warps)."""
    om, h = omech(name), emech(name)
    assert emu.emu_static_shape(h) == shape
    rng = np.random.default_rng(5)
    n, K = 3, om.K
    T0 = rng.uniform(900, 2500, n); p0 = rng.uniform(0.5, 20, n) * 101325
    T0[0] = 350.0
    Y = rng.random((n, K)) ** 4
    Y /= Y.sum(1, keepdims=True)
    st = np.ascontiguousarray(np.concatenate([T0[:, None], Y], 1))
    out = []
    try:
        for static in (0, 1):
            emu.emu_use_static_shapes(static)
            yd, wd, qf, qr = np.zeros((n, K + 1)), np.zeros((n, K)), np.zeros((n, om.nR)), np.zeros((n, om.nR))
            emu.emu_rhs(h, 0, n, P(T0), P(p0), P(np.ascontiguousarray(Y)), P(st), P(yd), P(wd), P(qf), P(qr))
            Yf = om.X_to_Y(om.composition({"H2": 2.0, "O2": 1.0, "N2": 3.76}))
            r = _integrate(emu, h, om, 1, [1250.0, 1100.0], [101325.0, 2e5], np.stack([Yf, Yf]), 1e-8, 1e-14, 4e-5, linsol=1, warps=2, pool_n=1)
            out.append((yd, wd, qf, qr, r["state"], r["stats"], r["tau"]))
    finally:
        emu.emu_use_static_shapes(0)
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.all(out[0][5][:, 0] > 20)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("N", [3, 33, 53, 54, 64])
def test_linear_algebra_kernels(emu, N, mode):
    rng = np.random.default_rng(N)
    n = 2
    A = rng.standard_normal((n, N, N)); b = rng.standard_normal((n, N))
    x = np.zeros((n, N)); info = np.zeros(n, dtype=np.int32)
    emu.emu_lu_solve(N, n, P(np.ascontiguousarray(A.transpose(0, 2, 1))), P(b), P(x), info.ctypes.data_as(ip), mode)
    xr = np.linalg.solve(A, b[..., None])[..., 0]
    assert np.all(info == 0)
    assert np.abs(x - xr).max() / np.abs(xr).max() < 1e-11


def test_device_stepper_source_on_roberts_golden(emu, goldens):
    """The device stepper's source, run with the host's libm, reproduces the reference's CVODE golden file through the
    first nine outputs at the reference's own tolerance (beyond that the values sit below the absolute tolerances and
    a one-ulp change moves them; see tests/test_gpu_parity.py::test_roberts_device_stepper)."""
    g = goldens["roberts"]
    t = np.array(g["t_out"]); a = np.array(g["atol"]); y = np.zeros((12, 3)); st = np.zeros(8, dtype=np.int64)
    assert emu.emu_roberts(12, P(t), g["rtol"], P(a), P(y), st.ctypes.data_as(lp)) == 0
    gold = np.array(g["y"])
    assert np.abs(y[:9] / gold[:9] - 1).max() < g["tolerance"]
    assert np.abs(y[9:, 2] / gold[9:, 2] - 1).max() < g["rtol"]


def test_analytic_jacobian_source_vs_oracle_fd(emu, emech, omech):
    om, h = omech("gri12"), emech("gri12")
    T0 = np.array([1500.0]); p0 = np.array([3.0 * 101325.0])
    Y0 = om.X_to_Y(om.composition({"CH4": 1, "O2": 2, "N2": 7.52}))[None, :]
    r = om.integrate_batch(0, T0, p0, Y0, 1e-4, rtol=1e-8, atol=1e-14)
    T = r["state"][:, 0].copy()
    Y = np.maximum(r["state"][:, 1:], 0)
    Y /= Y.sum(1, keepdims=True)
    for rt in (0, 1):
        s = np.ascontiguousarray(np.concatenate([T[:, None], Y], 1))
        N = s.shape[1]
        J = np.zeros((1, N, N))
        emu.emu_jacobian(h, rt, 1, 1, P(T), P(p0), P(np.ascontiguousarray(Y)), P(s), P(J))
        J = J.transpose(0, 2, 1)
        Jf = np.zeros_like(J)
        for j in range(N):
            d = np.maximum(np.abs(s[:, j]) * 1e-6, 1e-12)
            sp, sm = s.copy(), s.copy()
            sp[:, j] += d; sm[:, j] -= d
            neg = sm[:, j] < 0
            sm[neg, j] = s[neg, j]
            Jf[:, :, j] = (om.rhs_batch(rt, T, p0, Y, sp)[0] - om.rhs_batch(rt, T, p0, Y, sm)[0]) / (sp[:, j] - sm[:, j])[:, None]
        scale = np.abs(Jf).max(axis=2, keepdims=True) + 1e-300
        assert (np.abs(J - Jf) / scale).max() < 2e-5


def test_fused_integrator_source_vs_oracle(emu, emech, omech, goldens):
    """The reference's reactor test scenario through the emulated fused kernel: analytic Jacobian + LU, trajectory on
    a 20-point grid against the oracle (1e-4) and ignition delay."""
    g = goldens["slide_trajectory"]
    om, h = omech("gri30"), emech("gri30")
    Y = om.X_to_Y(om.composition(g["X"]))[None, :]
    ref = om.integrate_batch("Const_Vol", [g["T0"]], [g["p0"]], Y, g["t_end"], rtol=g["rtol"], atol=g["atol"], max_steps=20000,
                             ign_mode=0, ign_value=g["ignition_threshold"], n_out=20)
    r = _integrate(emu, h, om, 1, [g["T0"]], [g["p0"]], Y, g["rtol"], g["atol"], g["t_end"], n_out=20, ign=(0, g["ignition_threshold"]))
    assert r["status"][0] >= 0
    assert np.abs(r["traj"][0, :, 0] / ref["traj"][0, :, 0] - 1).max() < 1e-4
    assert abs(r["tau"][0] / ref["tau"][0] - 1) < 1e-4
    assert abs(r["stats"][0, 0] / ref["stats"][0, 0] - 1) < 0.25  # step counts stay statistically close to CVODE's


def test_extended_statistics_vs_oracle(emu, emech, omech, oracle, goldens):
    """MainSolverStatistics beyond the counters (include/sundials/cvode/types.hpp:175-215 of the reference: last / current
    order, first / last / next step size, current time).  On the CVODE Roberts problem the device stepper and the oracle
    take the same steps over the first three outputs (later, last-bit differences move the step sequence), so every entry must agree there; on a reactor (where last-bit differences move the step sequence)
    the first step, the current time and the ranges are checked."""
    g = goldens["roberts"]
    t = np.array(g["t_out"][:3]); atol = np.array(g["atol"])
    y = np.zeros((3, 3)); st = np.zeros(8, dtype=np.int64); ex = np.zeros(6)
    emu.emu_set_stats_ex_buffer(P(ex))
    assert emu.emu_roberts(3, P(t), g["rtol"], P(atol), P(y), st.ctypes.data_as(lp)) == 0
    f = lambda tt, yy: np.array([-0.04 * yy[0] + 1e4 * yy[1] * yy[2], 0.04 * yy[0] - 1e4 * yy[1] * yy[2] - 3e7 * yy[1] ** 2, 3e7 * yy[1] ** 2])
    jac = lambda tt, yy: np.array([[-0.04, 1e4 * yy[2], 1e4 * yy[1]], [0.04, -1e4 * yy[2] - 6e7 * yy[1], -1e4 * yy[1]], [0.0, 6e7 * yy[1], 0.0]])
    o = oracle.BDF(3, f, jac)
    o.init(0.0, g["y0"], g["rtol"], atol)
    for tk in t:
        assert o.integrate(tk)[0] == 0
    so = o.stats()
    assert [int(ex[0]), int(ex[1])] == [so["qlast"], so["qcur"]]
    assert np.abs(ex[2:6] / np.array([so["hinused"], so["hlast"], so["hcur"], so["tcur"]]) - 1).max() < 1e-4, (ex, so)
    assert int(st[0]) == so["nst"]
    om, h = omech("gri12"), emech("gri12")
    n = 2
    T0 = np.array([1150.0, 1400.0]); p0 = np.full(n, 2 * 101325.0)
    Y = np.ascontiguousarray(om.X_to_Y(np.tile(om.composition({"H2": 2, "O2": 1, "N2": 3.76}), (n, 1))))
    ref = om.integrate_batch(1, T0, p0, Y, 2e-4, rtol=1e-8, atol=1e-14, max_steps=20000)
    ex = np.zeros((n, 6))
    emu.emu_set_stats_ex_buffer(P(ex))
    r = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, jac=0, linsol=0)
    assert np.array_equal(r["status"], ref["status"])
    ex_ref = ref["stats_ex"]
    assert np.abs(ex[:, 2] / ex_ref[:, 2] - 1).max() < 1e-9                  # first step size: same state, same heuristic
    assert np.array_equal(ex[:, 5], ex_ref[:, 5]) and np.all(ex[:, 5] == 2e-4)  # current time = stop time
    assert np.all((ex[:, :2] >= 1) & (ex[:, :2] <= 5)) and np.all(np.abs(ex[:, :2] - ex_ref[:, :2]) <= 1)
    assert np.all((ex[:, 3:5] > 0) & (ex[:, 3:5] < 2e-4))
    assert np.abs(r["stats"][:, 0] / ref["stats"][:, 0] - 1).max() < 0.10  # statistical closeness only (analytic Jacobian + inverse vs difference quotients + LU)


def test_work_queue_resume_and_failure_codes(emu, emech, omech):
    """Several reactors over two warps (atomic work queue), repeated Integrate(t) calls through the saved stepper
    state, and the per-reactor CV_TOO_MUCH_WORK code."""
    om, h = omech("gri12"), emech("gri12")
    n = 3
    T0 = np.array([1100.0, 1500.0, 1300.0]); p0 = np.full(n, 2 * 101325.0)
    Y = om.X_to_Y(np.tile(om.composition({"H2": 2, "O2": 1, "N2": 3.76}), (n, 1)))
    ref = om.integrate_batch(1, T0, p0, Y, 2e-4, rtol=1e-8, atol=1e-14, max_steps=20000)
    one = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, warps=2)
    assert np.all(one["status"] >= 0)
    assert np.abs(one["state"][:, 0] / ref["state"][:, 0] - 1).max() < 1e-4
    many = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, n_calls=4, warps=2)
    assert np.abs(many["state"][:, 0] / ref["state"][:, 0] - 1).max() < 1e-4
    few = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, mxstep=20)
    assert np.all(few["status"] == -1)
    # pool mode: Newton matrix in a shared pool of work slabs (fewer slabs than warps), inverse in global scratch
    pooled = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, warps=3, linsol=1, pool_n=2)
    assert np.all(pooled["status"] >= 0)
    assert np.abs(pooled["state"][:, 0] / ref["state"][:, 0] - 1).max() < 1e-4
    inv = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, warps=2, linsol=1)
    assert np.abs(inv["state"][:, 0] / pooled["state"][:, 0] - 1).max() < 1e-6


def test_time_slicing_is_exact(emu, emech, omech):
    """Sliced scheduling (park the stepper every few steps, queue the reactor again, possibly resume on another warp)
    must reproduce the uninterrupted run bit for bit: states, ignition delays, every counter, trajectories, and the
    CV_TOO_MUCH_WORK accounting across slices."""
    om, h = omech("gri12"), emech("gri12")
    n = 5
    T0 = np.array([1100.0, 1500.0, 1300.0, 1250.0, 1400.0]); p0 = np.full(n, 2 * 101325.0)
    Y = om.X_to_Y(np.tile(om.composition({"H2": 2, "O2": 1, "N2": 3.76}), (n, 1)))
    base = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, warps=2, linsol=1, pool_n=1)
    for steps in (7, 40):
        sl = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, warps=2, linsol=1, pool_n=1, slice_steps=steps)
        assert np.array_equal(sl["status"], base["status"])
        assert np.array_equal(sl["stats"], base["stats"])
        assert np.array_equal(sl["state"], base["state"])
        assert np.array_equal(sl["tau"], base["tau"], equal_nan=True)
    tb = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, warps=2, linsol=1, pool_n=1, n_out=5)
    ts = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, warps=2, linsol=1, pool_n=1, n_out=5, slice_steps=9)
    assert np.array_equal(ts["traj"], tb["traj"]) and np.array_equal(ts["stats"], tb["stats"])
    few = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, warps=2, linsol=1, pool_n=1, mxstep=20, slice_steps=7)
    fewb = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, warps=2, linsol=1, pool_n=1, mxstep=20)
    assert np.all(few["status"] == -1) and np.array_equal(few["stats"], fewb["stats"])
    # repeated Integrate(t) calls (user-visible resumable state) on top of slicing
    many = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, n_calls=3, warps=2, linsol=1, pool_n=1, slice_steps=11)
    manyb = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, n_calls=3, warps=2, linsol=1, pool_n=1)
    assert np.array_equal(many["state"], manyb["state"])


def test_solver_option_surface_vs_oracle(emu, emech, omech):
    """IStrategy::Set* options (interface.cpp:61-219 of the reference): initial / minimum / maximum step, Newton
    iterations and coefficient, failure limits, maximum order — device stepper source against the oracle's CVODE
    restatement with the same options."""
    om, h = omech("gri12"), emech("gri12")
    T0 = np.array([1250.0, 1500.0]); p0 = np.full(2, 2 * 101325.0)
    Y = om.X_to_Y(np.tile(om.composition({"H2": 2, "O2": 1, "N2": 3.76}), (2, 1)))
    opts = dict(init_step=1e-10, max_step=4e-6, nonlin_conv_coef=0.05, max_nonlin_iters=4, max_err_test_fails=5,
                max_conv_fails=8, max_order=3)
    ref = om.integrate_batch(1, T0, p0, Y, 2e-4, rtol=1e-8, atol=1e-14, max_steps=20000, options=opts)
    dflt = om.integrate_batch(1, T0, p0, Y, 2e-4, rtol=1e-8, atol=1e-14, max_steps=20000)
    got = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, jac=0, options=opts)  # DQ Jacobian + LU like the oracle
    assert np.all(got["status"] == ref["status"]) and np.all(ref["status"] >= 0)
    assert np.abs(got["state"][:, 0] / ref["state"][:, 0] - 1).max() < 1e-6
    assert np.abs(got["tau"] / ref["tau"] - 1).max() < 1e-5
    assert np.abs(got["stats"][:, 0] / ref["stats"][:, 0] - 1).max() < 0.05
    assert np.all(ref["stats"][:, 0] > 1.2 * dflt["stats"][:, 0])  # the options really changed the integration
    # a minimum step the ignition front cannot live with: same failure class as CVODE's
    bad = dict(min_step=2e-6)
    refb = om.integrate_batch(1, T0, p0, Y, 2e-4, rtol=1e-8, atol=1e-14, max_steps=20000, options=bad)
    gotb = _integrate(emu, h, om, 1, T0, p0, Y, 1e-8, 1e-14, 2e-4, jac=0, options=bad)
    assert np.all(refb["status"] < 0) and np.array_equal(gotb["status"], refb["status"])


def _rcm_tables(n, nt, t_end, ratio, t_c):
    """Rapid-compression-machine style volume history: smooth compression V0 -> V0 / ratio over t_c, then held."""
    t = np.linspace(0.0, t_end, nt)
    x = np.clip(t / t_c, 0.0, 1.0)
    s_ = 0.5 * (1.0 - np.cos(np.pi * x))                       # 0 -> 1, zero slope at both ends
    ds = 0.5 * np.pi * np.sin(np.pi * x) / t_c * (t < t_c)
    V = 1.0 - (1.0 - 1.0 / np.asarray(ratio)[:, None]) * s_[None, :]
    dV = -(1.0 - 1.0 / np.asarray(ratio)[:, None]) * ds[None, :]
    return t, V * 3.0e-4, dV * 3.0e-4                          # an arbitrary volume unit: only V / V(0) enters


@pytest.mark.parametrize("jac,linsol,n", [(1, 1, 3), (0, 0, 1)])
def test_prescribed_volume_reactor_vs_oracle(emu, emech, omech, jac, linsol, n):
    """Type 5 (RT_TABLE_V): adiabatic reactor in a prescribed volume V(t), dV/dt (planned by the reference, README.md:18-20).
    The kernel source against the oracle on compression-ignition cases: H2/air compressed 8-12 x in 0.3 ms, then held;
    analytic Jacobian + inverse (its energy row leaves out the composition dependence of the work term) and the
    reference's configuration (difference quotients + LU)."""
    om, h = omech("gri30"), emech("gri30")
    t_end = 8e-4
    T0 = np.array([500.0, 450.0, 420.0])[:n]; p0 = (np.array([1.0, 1.5, 0.8]) * 101325.0)[:n]
    Y = np.repeat(om.X_to_Y(om.composition({"H2": 2.0, "O2": 1.0, "N2": 3.76}))[None, :], n, 0)
    table = _rcm_tables(n, 41, t_end, [8.0, 10.0, 12.0][:n], 3e-4)
    ref = om.integrate_batch("Table_V", T0, p0, Y, t_end, rtol=1e-9, atol=1e-15, max_steps=100000, table=table)
    assert np.all(ref["status"] >= 0) and np.isfinite(ref["tau"]).sum() >= min(n, 2)  # compression ignites them
    r = _integrate(emu, h, om, 5, T0, p0, Y, 1e-9, 1e-15, t_end, jac=jac, linsol=linsol, mxstep=100000, table=table)
    assert np.all(r["status"] >= 0)
    assert np.array_equal(np.isfinite(r["tau"]), np.isfinite(ref["tau"]))
    ign = np.isfinite(ref["tau"])
    assert np.abs(r["tau"][ign] / ref["tau"][ign] - 1).max() < 1e-4
    assert np.abs(r["state"][:, 0] / ref["state"][:, 0] - 1).max() < 1e-4
    assert np.abs(r["state"][:, 1:] - ref["state"][:, 1:]).max() < 1e-6
