"""Host-side mirror of the reference's reactor application interface, batched.

Names and argument meaning follow the reference:
  Chemistry::ThermoKinetics            /root/reference/include/chemistry.hpp:207-226
  Apps::Reactor::Type / Create         /root/reference/include/applications/reactors/base.hpp:21-27,
                                       /root/reference/include/applications/reactors/factory.hpp:15-72
  SUNDIALS::CVODE::Solver::Integrate   /root/reference/include/sundials/cvode/interface.hpp:297
  Solver::Solution / Client::State     /root/reference/include/sundials/cvode/client.hpp:110-120
  solver statistics                    /root/reference/include/sundials/cvode/types.hpp:175-228
but every object holds n reactors instead of one; all arithmetic happens in the
CUDA library behind the C ABI (include/chemtools_b200.h).
"""
import ctypes as C
import enum

import numpy as np

from ._lib import CtError, check, dp, ip, lib, lp

# CVODE return codes passed through unchanged (interface.cpp:448-475 of the reference)
CV_SUCCESS, CV_TSTOP_RETURN = 0, 1
CV_TOO_MUCH_WORK, CV_TOO_MUCH_ACC, CV_ERR_FAILURE, CV_CONV_FAILURE = -1, -2, -3, -4

STAT_NAMES = ("nst", "nfe", "nsetups", "netf", "nni", "ncfn", "nje", "nfeLS")
STAT_EX_NAMES = ("qlast", "qcur", "hinused", "hlast", "hcur", "tcur")


class Type(enum.IntEnum):
    Const_Pres = 0
    Const_Vol = 1
    Const_Temp_Pres = 2
    Const_Temp_Vol = 3
    Table_Temp_Pres = 4  # prescribed T(t), p(t) (planned by the reference, README.md:18-20)
    Table_Vol = 5        # adiabatic, prescribed volume V(t) and dV/dt (planned on the same lines); state [T, Y]


def has_energy(reactor_type):
    """State [T, Y...] (EnergyEnabled reactors) or [Y...]."""
    return int(reactor_type) in (Type.Const_Pres, Type.Const_Vol, Type.Table_Vol)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(dp)


class ThermoKinetics:
    """Chemistry::ThermoKinetics(mechanism): first ideal_gas phase of a .cti (or a compiled .ctm)."""

    def __init__(self, mechanism, phase_id=None, constants=None):
        st = C.c_int(0)
        self._h = lib().ct_mech_load_ex(str(mechanism).encode(), None if phase_id is None else phase_id.encode(),
                                        None if constants is None else constants.encode(), C.byref(st))
        if not self._h:
            raise CtError(st.value, lib().ct_last_error().decode())
        L = lib()
        self.n_species = L.ct_mech_n_species(self._h)
        self.n_reactions = L.ct_mech_n_reactions(self._h)
        self.species_names = [L.ct_mech_species_name(self._h, k).decode() for k in range(self.n_species)]
        self.molecular_weights = np.zeros(self.n_species)
        check(L.ct_mech_molecular_weights(self._h, _p(self.molecular_weights)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ct_mech_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # Cantera-style accessors used by the reference's tests (tests/src/chemistry.cpp:250-322)
    def nSpecies(self):
        return self.n_species

    def nReactions(self):
        return self.n_reactions

    def speciesIndex(self, name):
        return lib().ct_mech_species_index(self._h, name.encode())

    def speciesName(self, k):
        return self.species_names[k]

    def reactionString(self, i):
        s = lib().ct_mech_reaction_equation(self._h, i)
        return None if s is None else s.decode()

    def getMolecularWeights(self):
        return self.molecular_weights.copy()

    def elements(self):
        L = lib()
        return [L.ct_mech_element_name(self._h, e).decode() for e in range(L.ct_mech_n_elements(self._h))]

    def nAtoms(self, k, e):
        return lib().ct_mech_n_atoms(self._h, k, e)

    def gas_constant(self):
        return lib().ct_mech_gas_constant(self._h)

    def static_shape(self):
        """1 / 2 / 3 = GRI-3.0 / 2.11 / 1.2 (kernels with compile-time sizes and offsets), 0 = run-time shape."""
        return lib().ct_mech_static_shape(self._h)

    def save_ctm(self, path):
        check(lib().ct_mech_save_ctm(self._h, str(path).encode()))

    def compiled_arrays(self):
        R = self.n_reactions
        perm = np.zeros(R, dtype=np.int32); A = np.zeros(R); b = np.zeros(R); E = np.zeros(R)
        check(lib().ct_mech_compiled_arrays(self._h, perm.ctypes.data_as(ip), _p(A), _p(b), _p(E)))
        return perm, A, b, E

    def flop_model(self, reactor_type):
        out = np.zeros(4)
        check(lib().ct_mech_flop_model(self._h, int(reactor_type), _p(out)))
        return dict(zip(("rhs", "jac", "lu", "solve"), out.tolist()))

    def composition(self, spec):
        """{'H2': 2, 'O2': 1} -> dense vector in species order (not normalised)."""
        x = np.zeros(self.n_species)
        for name, v in spec.items():
            k = self.speciesIndex(name)
            if k < 0:
                raise KeyError("unknown species " + name)
            x[k] = v
        return x

    def mass_fractions(self, X):
        """setState_TPX + getMassFractions of the reference's caller (tests/src/reactor.cpp:57-59)."""
        X = _f64(np.atleast_2d(X))
        Y = np.zeros_like(X)
        check(lib().ct_mech_mole_to_mass(self._h, X.shape[0], _p(X), _p(Y)))
        return Y


class ReactorBatch:
    """n reactors of one Apps::Reactor::Type sharing one mechanism, integrated on the GPU."""

    def __init__(self, reactor_type, chemistry, temperature, pressure, mass_fractions,
                 relative_solver_tolerance=1e-8, absolute_solver_tolerance=1e-14, total_simulation_time=1e-3,
                 max_num_steps=500, device_arrays=False):
        self.chem = chemistry
        self.type = Type(int(reactor_type))
        self._keep = None
        if device_arrays:
            # torch CUDA tensors already resident in HBM (no H2D inside)
            T, p, Y = temperature, pressure, mass_fractions
            n = int(T.shape[0])
            self._keep = (T, p, Y)
        else:
            T = _f64(np.atleast_1d(temperature)); p = _f64(np.atleast_1d(pressure)); Y = _f64(np.atleast_2d(mass_fractions))
            n = T.shape[0]
            if p.shape != (n,) or Y.shape != (n, chemistry.n_species):
                raise ValueError("temperature/pressure/mass_fractions shapes disagree")
        self.n = n
        st = C.c_int(0)
        self._h = lib().ct_batch_create(chemistry._h, int(self.type), n, C.byref(st))
        if not self._h:
            raise CtError(st.value, lib().ct_last_error().decode())
        self.n_state = lib().ct_batch_n_state(self._h)
        check(lib().ct_batch_set_tolerances(self._h, relative_solver_tolerance, absolute_solver_tolerance))
        check(lib().ct_batch_set_stop_time(self._h, total_simulation_time))
        check(lib().ct_batch_set_max_steps(self._h, max_num_steps))
        self.t_end = total_simulation_time
        self._ctor = (relative_solver_tolerance, absolute_solver_tolerance, max_num_steps)
        self._host_ic = None if device_arrays else (T, p, Y)
        self._table = None
        if device_arrays:
            check(lib().ct_batch_set_ic_device(self._h, T.data_ptr(), p.data_ptr(), Y.data_ptr()))
        else:
            check(lib().ct_batch_set_ic(self._h, _p(T), _p(p), _p(Y)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ct_batch_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- options (Solver::Set* of interface.hpp:77-89)
    def SetTolerances(self, relative, absolute):
        """Client::SetTolerances (client.hpp:122-134): scalar or per-component (length n_state) absolute tolerance."""
        if np.ndim(absolute) == 0:
            check(lib().ct_batch_set_tolerances(self._h, float(relative), float(absolute)))
        else:
            a = _f64(absolute)
            if a.shape != (self.n_state,):
                raise ValueError("absolute tolerance vector must have n_state entries")
            check(lib().ct_batch_set_tolerances_vec(self._h, float(relative), _p(a)))

    def SetMaxNumSteps(self, n):
        check(lib().ct_batch_set_max_steps(self._h, int(n)))

    def SetMaxOrder(self, q):
        check(lib().ct_batch_set_max_order(self._h, int(q)))

    def SetInitStepSize(self, h):
        check(lib().ct_batch_set_init_step(self._h, float(h)))

    def SetMinStepSize(self, h):
        check(lib().ct_batch_set_min_step(self._h, float(h)))

    def SetMaxStepSize(self, h):
        check(lib().ct_batch_set_max_step(self._h, float(h)))

    def SetMaxErrorTestFailures(self, n):
        check(lib().ct_batch_set_max_err_test_fails(self._h, int(n)))

    def SetMaxNonlinearIterations(self, n):
        check(lib().ct_batch_set_max_nonlin_iters(self._h, int(n)))

    def SetMaxConvergenceFailures(self, n):
        check(lib().ct_batch_set_max_conv_fails(self._h, int(n)))

    def SetNonlinConvergenceCoefficient(self, c):
        check(lib().ct_batch_set_nonlin_conv_coef(self._h, float(c)))

    def SetMaxWarnMessages(self, n):
        check(lib().ct_batch_set_max_warn_messages(self._h, int(n)))

    def SetStabilityLimitDetection(self, on):
        check(lib().ct_batch_set_stability_limit_detection(self._h, 1 if on else 0))

    def ReInitialise(self, time_init, state):
        """IStrategy::ReInitialise (interface.cpp:429-446): restart from state[n, N] at time_init."""
        st = _f64(np.atleast_2d(state))
        if st.shape != (self.n, self.n_state):
            raise ValueError("state must have shape (n, n_state)")
        check(lib().ct_batch_reinit(self._h, float(time_init), _p(st)))

    def SetJacobian(self, mode):
        check(lib().ct_batch_set_jacobian(self._h, {"dq": 0, "analytic": 1}.get(mode, mode)))

    def SetLinearSolver(self, mode):
        check(lib().ct_batch_set_linear_solver(self._h, {"lu": 0, "inverse": 1}.get(mode, mode)))

    def SetJacobianClip(self, thr):
        check(lib().ct_batch_set_jacobian_clip(self._h, float(thr)))

    def SetStaticShape(self, on):
        check(lib().ct_batch_set_static_shape(self._h, 1 if on else 0))

    def SetPhaseAlignment(self, on):
        check(lib().ct_batch_set_phase_alignment(self._h, int(on)))

    def SetAlignmentPollNs(self, ns):
        check(lib().ct_batch_set_alignment_poll_ns(self._h, int(ns)))

    def SetAlignmentPeriod(self, k):
        check(lib().ct_batch_set_alignment_period(self._h, int(k)))

    def SetAlignmentMaxPolls(self, n):
        check(lib().ct_batch_set_alignment_max_polls(self._h, int(n)))

    def SetMatrixPool(self, on, warps_per_sm=0):
        check(lib().ct_batch_set_matrix_pool(self._h, 1 if on else 0, int(warps_per_sm)))

    def SetTimeSlicing(self, steps):
        check(lib().ct_batch_set_time_slicing(self._h, int(steps)))

    def SetTrace(self, reactor, capacity=4096):
        self._trace_cap = int(capacity)
        check(lib().ct_batch_set_trace(self._h, int(reactor), int(capacity)))

    def Trace(self):
        """[count, 12] step-attempt records of the traced reactor (see ct_batch_set_trace)."""
        out = np.zeros((self._trace_cap, 12)); cnt = C.c_int(0)
        check(lib().ct_batch_get_trace(self._h, out.ctypes.data_as(C.c_void_p), C.byref(cnt)))
        return out[:min(cnt.value, self._trace_cap)]

    def SetTimeline(self, on=True):
        check(lib().ct_batch_set_timeline(self._h, 1 if on else 0))

    def Timeline(self):
        """[n, 3] int64: start ns, end ns (GPU global timer), sm * 64 + warp of every reactor of the last Integrate."""
        out = np.zeros((self.n, 3), dtype=np.int64)
        check(lib().ct_batch_get_timeline(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def SetIgnition(self, mode, value):
        check(lib().ct_batch_set_ignition(self._h, {"absolute": 0, "rise": 1}.get(mode, mode), float(value)))

    def SetTable(self, t, T, p):
        t = _f64(t); T = _f64(np.atleast_2d(T)); p = _f64(np.atleast_2d(p))
        if T.shape != (self.n, t.shape[0]) or p.shape != T.shape:
            raise ValueError("table shapes must be [n, nt]")
        check(lib().ct_batch_set_tp_table(self._h, t.shape[0], _p(t), _p(T), _p(p)))
        self._table = (t, T, p)

    def SetVolumeTable(self, t, V, dVdt):
        """Prescribed volume and its time derivative (Type.Table_Vol): t[nt], V[n, nt] > 0 in any unit, dVdt[n, nt]."""
        t = _f64(t); V = _f64(np.atleast_2d(V)); dVdt = _f64(np.atleast_2d(dVdt))
        if V.shape != (self.n, t.shape[0]) or dVdt.shape != V.shape:
            raise ValueError("table shapes must be [n, nt]")
        check(lib().ct_batch_set_v_table(self._h, t.shape[0], _p(t), _p(V), _p(dVdt)))
        self._table = (t, V, dVdt)

    def SetTableFromFunction(self, function, n_points=101, module_path=None):
        """Prescribed T(t), p(t) — or, for Type.Table_Vol, V(t), dV/dt(t) — from a FunctionP1R2-style Python callable
        (include/embedded_python/functions.hpp:10-41 of the reference: one parameter in, two values out; README.md:18-20
        names exactly these two uses).  `function` is a callable f(t) -> (T, p), or a function
        name to load from the module file `module_path` like FunctionP1R2::Load(module_path, function_name).  The
        callable is sampled ONCE on the host at n_points equidistant times in [0, t_end] into the device table (linear
        interpolation in the kernel); T and p may be scalars (all reactors) or arrays of length n."""
        if module_path is not None:
            import importlib.util
            spec = importlib.util.spec_from_file_location("ct_user_profile", module_path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            function = getattr(mod, function)
        t = np.linspace(0.0, self.t_end, int(n_points))
        T = np.zeros((self.n, t.size)); p = np.zeros((self.n, t.size))
        for k, tk in enumerate(t):
            val = function(float(tk))
            if len(val) != 2:
                raise ValueError("the profile function must return two values (T, p) or (V, dV/dt)")
            T[:, k] = val[0]; p[:, k] = val[1]
        if self.type == Type.Table_Vol:
            self.SetVolumeTable(t, T, p)
        else:
            self.SetTable(t, T, p)

    def SetOrder(self, order):
        o = np.ascontiguousarray(order, dtype=np.int32)
        if o.shape != (self.n,):
            raise ValueError("order must be a permutation of range(n)")
        check(lib().ct_batch_set_order(self._h, o.ctypes.data_as(ip)))

    def SetStream(self, cuda_stream_ptr):
        check(lib().ct_batch_set_stream(self._h, cuda_stream_ptr))

    # ---- solve
    def Integrate(self, time):
        """CVode(..., CV_NORMAL) for every reactor; returns 0 / 1 / most negative per-reactor code."""
        code = lib().ct_batch_integrate(self._h, float(time))
        if code <= -100:
            check(code)
        return code

    def IntegrateWithRetry(self):
        """ct_batch_integrate_with_retry: Integrate to the stop time, then the library re-runs the reactors that ended with a
        negative CVODE status (about 0.05 % of stiff CH4 ensembles crawl into CV_TOO_MUCH_WORK: a property of the clipped
        right-hand side shared with the reference's algorithm, chaotic in which reactor it hits) as a small
        batch under other Jacobian / Newton options and merges what succeeded.  Returns dict(code, state, ignition_delay,
        statistics, status, failed_first, failed_final)."""
        n0, n1 = C.c_int64(0), C.c_int64(0)
        code = lib().ct_batch_integrate_with_retry(self._h, C.byref(n0), C.byref(n1))
        if code <= -100:
            check(code)
        return dict(code=code, state=self.Solution(), ignition_delay=self.IgnitionDelay(), statistics=self.Statistics(), status=self.Status(),
                    failed_first=n0.value, failed_final=n1.value)

    def IntegrateTrajectory(self, n_out):
        traj = np.zeros((self.n, n_out, self.n_state))
        code = lib().ct_batch_integrate_trajectory(self._h, int(n_out), _p(traj))
        if code <= -100:
            check(code)
        return code, traj

    def Solution(self):
        out = np.zeros((self.n, self.n_state))
        check(lib().ct_batch_get_state(self._h, _p(out)))
        return out

    State = Solution

    def Temperature(self):
        if not has_energy(self.type):
            raise ValueError("temperature is not a state of this reactor type")
        return self.Solution()[:, 0]

    def MassFractions(self):
        s = self.Solution()
        return s[:, 1:] if has_energy(self.type) else s

    def Time(self):
        out = np.zeros(self.n)
        check(lib().ct_batch_get_time(self._h, _p(out)))
        return out

    def IgnitionDelay(self):
        out = np.zeros(self.n)
        check(lib().ct_batch_get_ignition(self._h, _p(out)))
        return out

    def Status(self):
        out = np.zeros(self.n, dtype=np.int32)
        check(lib().ct_batch_get_status(self._h, out.ctypes.data_as(ip)))
        return out

    def Statistics(self):
        out = np.zeros((self.n, 8), dtype=np.int64)
        check(lib().ct_batch_get_stats(self._h, out.ctypes.data_as(lp)))
        return out

    def StatisticsEx(self):
        """[n, 6]: last_order_used, current_order, first / last / next internal step size, current time
        (MainSolverStatistics, include/sundials/cvode/types.hpp:175-215 of the reference)."""
        out = np.zeros((self.n, 6))
        check(lib().ct_batch_get_stats_ex(self._h, _p(out)))
        return out

    def LastKernelMs(self):
        return lib().ct_batch_last_kernel_ms(self._h)

    def LaunchCount(self):
        return lib().ct_batch_launch_count(self._h)

    def WarpsPerSM(self):
        return lib().ct_batch_warps_per_sm(self._h)

    # ---- stand-alone kernels
    def RightHandSide(self, states, times=None, rates=False):
        """Base/EnergyEnabled::RightHandSide for all reactors: returns (ydot, wdot[, qf, qr])."""
        states = _f64(states)
        if states.shape != (self.n, self.n_state):
            raise ValueError("states must be [n, n_state]")
        K, R = self.chem.n_species, self.chem.n_reactions
        ydot = np.zeros_like(states); wdot = np.zeros((self.n, K))
        qf = np.zeros((self.n, R)) if rates else None
        qr = np.zeros((self.n, R)) if rates else None
        tt = None if times is None else _f64(times)
        check(lib().ct_rhs_eval(self._h, _p(states), _p(tt), _p(ydot), _p(wdot), _p(qf), _p(qr)))
        return (ydot, wdot, qf, qr) if rates else (ydot, wdot)

    def Jacobian(self, states, mode="analytic", times=None):
        states = _f64(states)
        J = np.zeros((self.n, self.n_state, self.n_state))
        tt = None if times is None else _f64(times)
        check(lib().ct_jacobian_eval(self._h, {"dq": 0, "analytic": 1}[mode], _p(states), _p(tt), _p(J)))
        return J.transpose(0, 2, 1)  # column-major on the device -> J[n, i, j] = d f_i / d y_j

    def time_rhs(self, states, reps=20):
        ms = C.c_double()
        check(lib().ct_rhs_time(self._h, _p(_f64(states)), reps, C.cast(C.byref(ms), dp)))
        return ms.value

    def time_rhs_cfg(self, states, warps_per_block, blocks_per_sm, reps=20):
        """K1 in the fused kernel's slab layout at a chosen occupancy (ms per launch)."""
        ms = C.c_double()
        check(lib().ct_rhs_time_cfg(self._h, _p(_f64(states)), reps, int(warps_per_block), int(blocks_per_sm), C.cast(C.byref(ms), dp)))
        return ms.value

    def time_jacobian(self, states, mode="analytic", reps=5):
        ms = C.c_double()
        check(lib().ct_jacobian_time(self._h, {"dq": 0, "analytic": 1}[mode], _p(_f64(states)), reps, C.cast(C.byref(ms), dp)))
        return ms.value


def Create(reactor_type, chemistry, temperature, pressure, mass_fractions, relative_solver_tolerance,
           absolute_solver_tolerance, total_simulation_time, write_results=False, **kw):
    """Apps::Reactor::Create (factory.hpp:15-72); scalars give a batch of one reactor."""
    try:
        rt = Type(int(reactor_type))
    except ValueError:
        raise ValueError("Invalid reactor")  # factory.hpp:68-70
    return ReactorBatch(rt, chemistry, temperature, pressure, mass_fractions, relative_solver_tolerance,
                        absolute_solver_tolerance, total_simulation_time, **kw)


def lu_solve(A, b, mode="lu"):
    """Batched dense solve on the device (A[n, N, N] row-major as numpy, b[n, N]); mode "lu" or "inverse"."""
    A = _f64(A); b = _f64(b)
    n, N = b.shape
    Acm = np.ascontiguousarray(A.transpose(0, 2, 1))
    x = np.zeros_like(b); info = np.zeros(n, dtype=np.int32)
    check(lib().ct_lu_solve(N, n, _p(Acm), _p(b), _p(x), info.ctypes.data_as(ip), {"lu": 0, "inverse": 1}[mode]))
    return x, info


def lu_time(N, n, reps=10, solves_per_factor=1, mode="lu"):
    ms = C.c_double()
    check(lib().ct_lu_time(N, n, reps, solves_per_factor, {"lu": 0, "inverse": 1}[mode], C.cast(C.byref(ms), dp)))
    return ms.value


def measure_fp64_peak():
    t = C.c_double()
    check(lib().ct_measure_fp64_peak(C.cast(C.byref(t), dp)))
    return t.value


def roberts_selftest(t_out, rtol, atol3):
    t_out = _f64(t_out); atol3 = _f64(atol3)
    y = np.zeros((t_out.shape[0], 3)); st = np.zeros(8, dtype=np.int64)
    check(lib().ct_roberts_selftest(t_out.shape[0], _p(t_out), rtol, _p(atol3), _p(y), st.ctypes.data_as(lp)))
    return y, dict(zip(STAT_NAMES, st.tolist()))


def device_count():
    return lib().ct_device_count()


def set_device(i):
    check(lib().ct_set_device(int(i)))
