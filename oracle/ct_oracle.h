/* ORACLE — test infrastructure, NOT product code.
 *
 * Scalar fp64 CPU restatement of the reference's hot path
 * (Apps::Reactor::* over Cantera IdealGasPhase/GasKinetics, integrated by
 * SUNDIALS CVODE BDF + dense LU + difference-quotient Jacobian).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may link or call this.  The product (chemtools_b200/csrc)
 * never does.
 *
 * Parity status: Cantera and SUNDIALS are un-vendored third-party
 * dependencies absent from /root/reference (pins: Cantera 2.x series, the
 * code uses the 2.x raw-pointer factories, /root/reference/include/chemistry.hpp:75;
 * SUNDIALS v5.8.0, /root/reference/scripts/dependencies_installer/InstallDependencies.ps1:44).
 * Their published algorithms are restated here.  Pinned against the
 * reference's own golden vectors: Roberts .dat (stepper), equilibrium thermo
 * KAT (thermo), mechanism-structure KAT, slide trajectory (whole path, 7 s.f.).
 */
#ifndef CT_ORACLE_H
#define CT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- mechanism (Cantera kmol-m-s-K-J units) ---- */
typedef struct omech omech;

omech* om_create(int K, int R, double gas_constant, double p_ref,
                 const double* W, const double* nasa_lo /*K*7*/, const double* nasa_hi /*K*7*/,
                 const double* Tmid,
                 const int* rtype /*0 elementary,1 three-body,2 falloff*/, const int* reversible,
                 const double* A, const double* b, const double* EaR,
                 const double* A0, const double* b0, const double* EaR0,
                 const int* ftype /*0 Lindemann,1 Troe*/, const double* troe /*R*4: A,T3,T1,T2*/,
                 const int* reac_ptr, const int* reac_sp, const int* prod_ptr, const int* prod_sp,
                 const int* eff_ptr, const int* eff_sp, const double* eff_val);
void om_destroy(omech*);
int om_nspecies(const omech*);
int om_nreactions(const omech*);

/* reactor types: values 0..3 follow Apps::Reactor::Type
 * (/root/reference/include/applications/reactors/base.hpp:21-27); 4 is the
 * prescribed T(t),p(t) table reactor of BASELINE config 4; 5 is the adiabatic
 * reactor with prescribed volume V(t) and dV/dt that the reference plans next to
 * it (README.md:18-20; no reference implementation exists: the equations are the
 * closed-system energy balance  rho = rho0 V(t0)/V(t),  dT/dt = (-sum u_k wdot_k
 * - p Vdot/V) / (rho cv), i.e. ConstVol::EnergyDerivative, reactors.cpp:96-109,
 * plus the boundary work; for OR_TABLE_V the two table columns hold V and dV/dt). */
enum { OR_CONST_PRES = 0, OR_CONST_VOL = 1, OR_CONST_TEMP_PRES = 2, OR_CONST_TEMP_VOL = 3, OR_TABLE_TP = 4, OR_TABLE_V = 5 };

typedef struct oreactor oreactor;
/* T,p,Y as in Apps::Reactor::Create (factory.hpp:15-24). */
oreactor* or_create(const omech*, int type, double T, double p, const double* Y);
void or_set_table(oreactor*, int nt, const double* t, const double* T, const double* p);
void or_destroy(oreactor*);
int or_nstate(const oreactor*);
void or_initial_state(const oreactor*, double* y);
double or_density(const oreactor*);
double or_pressure(const oreactor*);
/* RightHandSide (base.cpp:62-69, :192-200). wdot/qf/qr may be NULL. */
int or_rhs(oreactor*, double t, const double* y, double* ydot);
void or_last_rates(const oreactor*, double* wdot, double* qf, double* qr);
/* thermo helpers for KATs: mass-specific/molar mixture properties at (T,p,Y) */
void om_mixture_props(const omech*, double T, double p, const double* Y,
                      double* rho, double* h_mole, double* s_mole, double* cp_mole, double* mmw);
void om_species_thermo(const omech*, double T, double* cp_R, double* h_RT, double* s_R);

/* ---- CVODE-like BDF integrator ---- */
typedef int (*cvo_rhs_fn)(double t, const double* y, double* ydot, void* user);
/* dense column-major J (ld = N): J[j*N+i] = df_i/dy_j */
typedef int (*cvo_jac_fn)(double t, const double* y, const double* fy, double* J, void* user);

typedef struct cvo cvo;
cvo* cvo_create(int N, cvo_rhs_fn f, cvo_jac_fn jac_or_null, void* user);
void cvo_destroy(cvo*);
void cvo_init(cvo*, double t0, const double* y0);
void cvo_set_tolerances(cvo*, double rtol, double atol_scalar, const double* atol_vec_or_null);
void cvo_set_stop_time(cvo*, double tstop);
void cvo_set_max_steps(cvo*, long mxstep);
void cvo_set_max_order(cvo*, int q);
void cvo_set_step_options(cvo*, double h_init, double h_min, double h_max); /* 0 = CVODE default */
void cvo_set_solver_options(cvo*, int maxnef, int maxcor, int maxncf, double nlscoef); /* <= 0 = default */
/* CVode(..., CV_NORMAL). returns 0 CV_SUCCESS, 1 CV_TSTOP_RETURN, <0 failure */
int cvo_integrate(cvo*, double tout, double* yout, double* tret);
/* nst nfe nsetups netf nni ncfn nje nfeDQ qlast qcur ; hlast hcur h0u tcur */
void cvo_stats(const cvo*, long* iout /*10*/, double* rout /*4*/);
/* ignition monitor on component `comp`: first upward crossing of `threshold` */
void cvo_watch(cvo*, int comp, double threshold);
double cvo_watch_time(const cvo*); /* NaN if not crossed */

enum { CVO_SUCCESS = 0, CVO_TSTOP_RETURN = 1, CVO_TOO_MUCH_WORK = -1, CVO_TOO_MUCH_ACC = -2,
       CVO_ERR_FAILURE = -3, CVO_CONV_FAILURE = -4, CVO_LSETUP_FAIL = -6, CVO_ILL_INPUT = -22,
       CVO_TOO_CLOSE = -27 };

/* ---- batch driver (CPU baseline: one reactor per host thread) ----
 * ign_mode 0: threshold = ign_value (absolute K); 1: threshold = T0 + ign_value.
 * out_state n*N, out_tau n, out_stats n*8 (nst,nfe,nsetups,netf,nni,ncfn,nje,nfeDQ), out_status n.
 * If n_out>0, trajectories at t_k=(k+1)*t_end/n_out go to out_traj[n][n_out][N]. */
/* optional: [n x 6] qlast, qcur, hinused, hlast, hcur, tcur of every reactor, filled by the NEXT or_integrate_batch* call of this thread */
void or_set_stats_ex_buffer(double* buf);
int or_integrate_batch(const omech*, int type, long n, const double* T0, const double* p0,
                       const double* Y0, int nt, const double* tab_t, const double* tab_T,
                       const double* tab_p, double rtol, double atol, double t_end, long max_steps,
                       int ign_mode, double ign_value, int nthreads, int n_out, double* out_traj,
                       double* out_state, double* out_tau, long* out_stats, int* out_status);
/* same with the solver option surface of the reference (types.hpp:68-112): opts[8] = h_init, h_min, h_max, nls_coef,
 * max_err_test_fails, max_nonlin_iters, max_conv_fails, max_order (0 = default); NULL = all defaults */
int or_integrate_batch_opts(const omech*, int type, long n, const double* T0, const double* p0,
                            const double* Y0, int nt, const double* tab_t, const double* tab_T,
                            const double* tab_p, double rtol, double atol, double t_end, long max_steps,
                            int ign_mode, double ign_value, int nthreads, int n_out, double* out_traj,
                            double* out_state, double* out_tau, long* out_stats, int* out_status,
                            const double* opts);
int or_rhs_batch(const omech*, int type, long n, const double* T0, const double* p0, const double* Y0,
                 const double* states, double* ydot, double* wdot, double* qf, double* qr);

#ifdef __cplusplus
}
#endif
#endif
