"""ORACLE — test infrastructure, NOT product code.

ctypes binding of oracle/libct_oracle.so (plain-C restatement of the
reference's Cantera + CVODE reactor path, see oracle/ct_oracle.h) plus the
unit conversion from CTI file units to Cantera's kmol-m-s-K-J system.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# Cantera's constant set is release dependent and the reference does not pin a
# 2.x release.  Named, selectable tables.  Default = the
# Cantera 2.5/2.6 set (CODATA-2018 gas constant, abridged IUPAC weights): with
# it the oracle reproduces the reference's published slide trajectory
# (doc/presentation/figures/output.png) to 1.3e-5 at the ignition front, with
# the 2.4 set only to 1.2e-4 (tests/test_oracle_golden.py).
CONSTANT_SETS = {
    "cantera-2.6": dict(gas_constant=8314.46261815324, one_atm=101325.0, cal=4.184,
                        weights={"H": 1.008, "C": 12.011, "N": 14.007, "O": 15.999, "Ar": 39.948}),
    "cantera-2.4": dict(gas_constant=8314.4621, one_atm=101325.0, cal=4.184,
                        weights={"H": 1.00794, "C": 12.0107, "N": 14.0067, "O": 15.9994, "Ar": 39.948}),
}
DEFAULT_CONSTANTS = "cantera-2.6"

REACTOR_TYPES = {"Const_Pres": 0, "Const_Vol": 1, "Const_Temp_Pres": 2, "Const_Temp_Vol": 3, "Table_TP": 4, "Table_V": 5}


def build(force=False):
    so = os.path.join(_HERE, "libct_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ct_kinetics.c", "ct_bdf.c", "ct_batch.c", "ct_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_long)
        L.om_create.restype = C.c_void_p
        L.om_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double] + [dp] * 4 + [ip, ip] + [dp] * 6 + [ip, dp] + [ip] * 6 + [dp]
        L.om_destroy.argtypes = [C.c_void_p]
        L.or_create.restype = C.c_void_p
        L.or_create.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, dp]
        L.or_set_table.argtypes = [C.c_void_p, C.c_int, dp, dp, dp]
        L.or_destroy.argtypes = [C.c_void_p]
        L.or_nstate.argtypes = [C.c_void_p]
        L.or_initial_state.argtypes = [C.c_void_p, dp]
        L.or_density.restype = C.c_double
        L.or_density.argtypes = [C.c_void_p]
        L.or_pressure.restype = C.c_double
        L.or_pressure.argtypes = [C.c_void_p]
        L.or_rhs.argtypes = [C.c_void_p, C.c_double, dp, dp]
        L.or_last_rates.argtypes = [C.c_void_p, dp, dp, dp]
        L.om_mixture_props.argtypes = [C.c_void_p, C.c_double, C.c_double, dp] + [dp] * 5
        L.om_species_thermo.argtypes = [C.c_void_p, C.c_double, dp, dp, dp]
        L.cvo_create.restype = C.c_void_p
        L.cvo_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cvo_destroy.argtypes = [C.c_void_p]
        L.cvo_init.argtypes = [C.c_void_p, C.c_double, dp]
        L.cvo_set_tolerances.argtypes = [C.c_void_p, C.c_double, C.c_double, dp]
        L.cvo_set_stop_time.argtypes = [C.c_void_p, C.c_double]
        L.cvo_set_max_steps.argtypes = [C.c_void_p, C.c_long]
        L.cvo_set_max_order.argtypes = [C.c_void_p, C.c_int]
        L.cvo_integrate.argtypes = [C.c_void_p, C.c_double, dp, dp]
        L.cvo_stats.argtypes = [C.c_void_p, lp, dp]
        L.cvo_watch.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.cvo_watch_time.restype = C.c_double
        L.cvo_watch_time.argtypes = [C.c_void_p]
        L.or_integrate_batch.argtypes = [C.c_void_p, C.c_int, C.c_long, dp, dp, dp, C.c_int, dp, dp, dp,
                                         C.c_double, C.c_double, C.c_double, C.c_long, C.c_int, C.c_double,
                                         C.c_int, C.c_int, dp, dp, dp, lp, ip]
        L.or_integrate_batch_opts.argtypes = L.or_integrate_batch.argtypes + [dp]
        L.or_rhs_batch.argtypes = [C.c_void_p, C.c_int, C.c_long, dp, dp, dp, dp, dp, dp, dp, dp]
        L.or_set_stats_ex_buffer.argtypes = [dp]
        _LIB = L
    return _LIB


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def mech_path(name):
    """Committed JSON dump (file units) written by oracle/cti_frontend.py."""
    return os.path.join(_HERE, "mech", name + ".json")


class Mechanism:
    """Flat SI-unit mechanism for the oracle + handle on the C copy."""

    def __init__(self, src, constants=DEFAULT_CONSTANTS):
        if isinstance(src, str):
            if src.endswith(".cti"):
                from . import cti_frontend
                src = cti_frontend.load_cti(src)
            else:
                if not os.path.exists(src):
                    src = mech_path(src)
                with open(src) as f:
                    src = json.load(f)
        self.raw = src
        cs = CONSTANT_SETS[constants]
        self.constants = constants
        self.R = cs["gas_constant"]
        self.p_ref = cs["one_atm"]
        u = src["units"]
        if (u.get("length"), u.get("quantity"), u.get("act_energy"), u.get("time")) != ("cm", "mol", "cal/mol", "s"):
            raise ValueError("unsupported unit system %r" % (u,))
        self.species = [s["name"] for s in src["species"]]
        self.index = {n: i for i, n in enumerate(self.species)}
        K = self.K = len(self.species)
        self.W = np.array([sum(n * cs["weights"][el] for el, n in s["atoms"]) for s in src["species"]])
        self.elements = src["elements"]
        self.natoms = np.zeros((K, len(self.elements)))
        for k, s in enumerate(src["species"]):
            for el, n in s["atoms"]:
                self.natoms[k, [e.lower() for e in self.elements].index(el.lower())] = n
        self.nasa_lo = np.array([s["nasa_lo"]["a"] for s in src["species"]])
        self.nasa_hi = np.array([s["nasa_hi"]["a"] for s in src["species"]])
        self.Tmid = np.array([s["nasa_lo"]["T"][1] for s in src["species"]])
        rx = src["reactions"]
        nR = self.nR = len(rx)
        self.file_equations = [r["equation"] for r in rx]
        # Kinetics::reactionString: Cantera rebuilds the equation from its species->coefficient std::map,
        # i.e. species in ASCII order on each side (tests/src/chemistry.cpp:287 expects exactly that form)
        self.equations = [self._reaction_string(r) for r in rx]
        kind = {"elementary": 0, "three_body": 1, "falloff": 2}
        self.rtype = np.array([kind[r["kind"]] for r in rx], dtype=np.int32)
        self.rev = np.array([1 if r["reversible"] else 0 for r in rx], dtype=np.int32)
        to_EaR = cs["cal"] * 1000.0 / self.R  # cal/mol -> J/kmol -> K
        A = np.zeros(nR); b = np.zeros(nR); E = np.zeros(nR)
        A0 = np.zeros(nR); b0 = np.zeros(nR); E0 = np.zeros(nR)
        ftype = np.zeros(nR, dtype=np.int32); troe = np.zeros((nR, 4))
        rp = [0]; rs = []; pp = [0]; ps = []; ep = [0]; es = []; ev = []
        for i, r in enumerate(rx):
            order = sum(c for _, c in r["reactants"])
            for name, c in r["reactants"]:
                if c != int(c):
                    raise ValueError("non-integer stoichiometry")
                rs += [self.index[name]] * int(c)
            for name, c in r["products"]:
                if c != int(c):
                    raise ValueError("non-integer stoichiometry")
                ps += [self.index[name]] * int(c)
            rp.append(len(rs)); pp.append(len(ps))
            m = order + (1 if r["kind"] == "three_body" else 0)
            # 1 mol/cm^3 = 1e3 kmol/m^3
            A[i] = r["kf"][0] * (1e-3) ** (m - 1); b[i] = r["kf"][1]; E[i] = r["kf"][2] * to_EaR
            if r["kind"] == "falloff":
                A0[i] = r["kf0"][0] * (1e-3) ** m; b0[i] = r["kf0"][1]; E0[i] = r["kf0"][2] * to_EaR
                if r["falloff"] is not None:
                    f = r["falloff"]
                    ftype[i] = 1
                    troe[i] = [f["A"], f["T3"], f["T1"], 0.0 if f["T2"] is None else f["T2"]]
            for name, e in r["efficiencies"]:
                es.append(self.index[name]); ev.append(e)
            ep.append(len(es))
        self.A, self.b, self.EaR, self.A0, self.b0, self.EaR0 = A, b, E, A0, b0, E0
        self.ftype, self.troe = ftype, troe
        self.rp = np.array(rp, dtype=np.int32); self.rs = np.array(rs, dtype=np.int32)
        self.pp = np.array(pp, dtype=np.int32); self.ps = np.array(ps, dtype=np.int32)
        self.ep = np.array(ep, dtype=np.int32); self.es = np.array(es, dtype=np.int32)
        self.ev = np.array(ev, dtype=np.float64)
        L = lib()
        self._h = L.om_create(K, nR, self.R, self.p_ref, _dp(self.W), _dp(_f64(self.nasa_lo)), _dp(_f64(self.nasa_hi)),
                              _dp(self.Tmid), _ip(self.rtype), _ip(self.rev), _dp(A), _dp(b), _dp(E), _dp(A0), _dp(b0),
                              _dp(E0), _ip(ftype), _dp(_f64(troe)), _ip(self.rp), _ip(self.rs), _ip(self.pp), _ip(self.ps),
                              _ip(self.ep), _ip(self.es), _dp(self.ev))

    @staticmethod
    def _reaction_string(r):
        def side(terms):
            acc = {}
            for name, c in terms:
                acc[name] = acc.get(name, 0) + c
            parts = [(name if c == 1 else "%g %s" % (c, name)) for name, c in sorted(acc.items())]
            s = " + ".join(parts)
            if r["kind"] == "three_body":
                s += " + M"
            elif r["kind"] == "falloff":
                s += " (+M)"
            return s
        return side(r["reactants"]) + (" <=> " if r["reversible"] else " => ") + side(r["products"])

    def __del__(self):
        try:
            if self._h:
                lib().om_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # --- composition helpers (the reference test does this through Cantera's
    #     setState_TPX + getMassFractions, tests/src/reactor.cpp:53-59) ---
    def X_to_Y(self, X):
        X = np.asarray(X, dtype=np.float64)
        X = X / X.sum(axis=-1, keepdims=True)
        Y = X * self.W
        return Y / Y.sum(axis=-1, keepdims=True)

    def Y_to_X(self, Y):
        Y = np.asarray(Y, dtype=np.float64)
        n = Y / self.W
        return n / n.sum(axis=-1, keepdims=True)

    def composition(self, spec):
        x = np.zeros(self.K)
        for k, v in spec.items():
            x[self.index[k]] = v
        return x

    def species_thermo(self, T):
        cp, h, s = np.zeros(self.K), np.zeros(self.K), np.zeros(self.K)
        lib().om_species_thermo(self._h, T, _dp(cp), _dp(h), _dp(s))
        return cp, h, s

    def mixture_props(self, T, p, Y):
        Y = _f64(Y)
        out = [C.c_double() for _ in range(5)]
        lib().om_mixture_props(self._h, T, p, _dp(Y), *[C.byref(o) for o in out])
        return dict(zip(("rho", "h_mole", "s_mole", "cp_mole", "mmw"), (o.value for o in out)))

    def rhs_batch(self, rtype, T0, p0, Y0, states):
        """RightHandSide for n reactors. Returns ydot[n,N], wdot[n,K], qf[n,R], qr[n,R]."""
        rtype = REACTOR_TYPES.get(rtype, rtype)
        T0, p0, Y0, states = _f64(np.atleast_1d(T0)), _f64(np.atleast_1d(p0)), _f64(np.atleast_2d(Y0)), _f64(np.atleast_2d(states))
        n = T0.shape[0]
        ydot = np.zeros_like(states); wdot = np.zeros((n, self.K)); qf = np.zeros((n, self.nR)); qr = np.zeros((n, self.nR))
        lib().or_rhs_batch(self._h, rtype, n, _dp(T0), _dp(p0), _dp(Y0), _dp(states), _dp(ydot), _dp(wdot), _dp(qf), _dp(qr))
        return ydot, wdot, qf, qr

    def integrate_batch(self, rtype, T0, p0, Y0, t_end, rtol=1e-8, atol=1e-14, max_steps=20000, ign_mode=1,
                        ign_value=400.0, nthreads=1, n_out=0, table=None, options=None):
        """One Apps::Reactor + CVODE(BDF, dense, DQ Jacobian) per case, 0 -> t_end."""
        rtype = REACTOR_TYPES.get(rtype, rtype)
        T0, p0, Y0 = _f64(np.atleast_1d(T0)), _f64(np.atleast_1d(p0)), _f64(np.atleast_2d(Y0))
        n = T0.shape[0]
        N = self.K + 1 if rtype in (0, 1, 5) else self.K
        state = np.zeros((n, N)); tau = np.zeros(n); stats = np.zeros((n, 8), dtype=np.int64); status = np.zeros(n, dtype=np.int32)
        traj = np.zeros((n, n_out, N)) if n_out else None
        nt, tt, tT, tp = 0, None, None, None
        if table is not None:
            tt, tT, tp = (_f64(a) for a in table)
            nt = tt.shape[0]
        opts = None
        if options:
            # solver option surface of the reference (types.hpp:68-112); 0 = CVODE default
            opts = _f64([options.get(k, 0.0) for k in ("init_step", "min_step", "max_step", "nonlin_conv_coef",
                                                        "max_err_test_fails", "max_nonlin_iters", "max_conv_fails", "max_order")])
        stats_ex = np.zeros((n, 6))
        lib().or_set_stats_ex_buffer(_dp(stats_ex))
        lib().or_integrate_batch_opts(self._h, rtype, n, _dp(T0), _dp(p0), _dp(Y0), nt, _dp(tt), _dp(tT), _dp(tp), rtol, atol,
                                      t_end, max_steps, ign_mode, ign_value, nthreads, n_out, _dp(traj), _dp(state), _dp(tau),
                                      stats.ctypes.data_as(C.POINTER(C.c_long)), _ip(status), _dp(opts))
        # stats_ex columns: qlast, qcur, hinused, hlast, hcur, tcur (include/sundials/cvode/types.hpp:175-215 of the reference)
        return dict(state=state, tau=tau, stats=stats, status=status, traj=traj, stats_ex=stats_ex)


RHS_FN = C.CFUNCTYPE(C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)
JAC_FN = C.CFUNCTYPE(C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)


class BDF:
    """The oracle's CVODE restatement on an arbitrary Python RHS (for the Roberts KAT)."""

    def __init__(self, N, f, jac=None):
        self.N = N

        def _f(t, y, ydot, _u):
            ya = np.ctypeslib.as_array(y, (N,))
            out = np.ctypeslib.as_array(ydot, (N,))
            out[:] = f(t, ya)
            return 0

        self._f = RHS_FN(_f)
        self._j = None
        if jac is not None:
            def _j(t, y, fy, J, _u):
                ya = np.ctypeslib.as_array(y, (N,))
                Ja = np.ctypeslib.as_array(J, (N * N,))
                Ja[:] = np.asarray(jac(t, ya)).T.reshape(-1)  # column-major
                return 0
            self._j = JAC_FN(_j)
        self._h = lib().cvo_create(N, C.cast(self._f, C.c_void_p), C.cast(self._j, C.c_void_p) if self._j else None, None)

    def init(self, t0, y0, rtol, atol, tstop=None, max_steps=500):
        y0 = _f64(y0)
        lib().cvo_init(self._h, t0, _dp(y0))
        if np.ndim(atol) == 0:
            lib().cvo_set_tolerances(self._h, rtol, float(atol), None)
        else:
            lib().cvo_set_tolerances(self._h, rtol, 0.0, _dp(_f64(atol)))
        if tstop is not None:
            lib().cvo_set_stop_time(self._h, tstop)
        lib().cvo_set_max_steps(self._h, max_steps)

    def integrate(self, tout):
        y = np.zeros(self.N)
        tret = C.c_double()
        code = lib().cvo_integrate(self._h, tout, _dp(y), C.cast(C.byref(tret), C.POINTER(C.c_double)))
        return code, tret.value, y

    def stats(self):
        io = np.zeros(10, dtype=np.int64); ro = np.zeros(4)
        lib().cvo_stats(self._h, io.ctypes.data_as(C.POINTER(C.c_long)), _dp(ro))
        d = dict(zip(("nst", "nfe", "nsetups", "netf", "nni", "ncfn", "nje", "nfeDQ", "qlast", "qcur"), io.tolist()))
        d.update(zip(("hlast", "hcur", "hinused", "tcur"), ro.tolist()))
        return d

    def __del__(self):
        try:
            lib().cvo_destroy(self._h)
        except Exception:
            pass
