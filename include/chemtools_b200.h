/* chemtools_b200 — C ABI of the B200-native batched homogeneous-reactor integrator.
 *
 * Drop-in boundary for the reactor path of ahmades/ChemTools.  Each entry
 * point names the reference interface it replaces (paths relative to the
 * reference repository).  Plain pointers and sizes only; no exceptions cross
 * this boundary: every call returns an int (0 = ok, 1 = reached stop time,
 * < 0 = failure class, CVODE-style where one exists) and ct_last_error()
 * gives the message.  Handles are not thread-safe individually; distinct
 * batches / GPUs may be driven from distinct host threads.
 *
 * There is NO CPU fallback: every compute entry point fails with
 * CT_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef CHEMTOOLS_B200_H
#define CHEMTOOLS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ct_mech ct_mech;   /* Chemistry::ThermoKinetics (include/chemistry.hpp:207-226) */
typedef struct ct_batch ct_batch; /* N x { Apps::Reactor::Base + SUNDIALS::CVODE::Solver } */

/* Apps::Reactor::Type (include/applications/reactors/base.hpp:21-27); 4 = prescribed T(t),p(t) table reactor
 * (BASELINE config 4; the reference only plans it, README.md:18-20). */
enum { CT_CONST_PRES = 0, CT_CONST_VOL = 1, CT_CONST_TEMP_PRES = 2, CT_CONST_TEMP_VOL = 3, CT_TABLE_TP = 4,
       /* adiabatic closed reactor in a prescribed volume V(t), dV/dt (planned next to it, README.md:18-20); state [T, Y] */
       CT_TABLE_V = 5 };

/* return / status codes. 0, 1 and the negative CV_* values are CVODE's own, which the reference returns
 * unchanged from IStrategy::Integrate (src/sundials/cvode/interface.cpp:454-475). */
enum {
  CT_SUCCESS = 0, CT_TSTOP_RETURN = 1, CT_TOO_MUCH_WORK = -1, CT_TOO_MUCH_ACC = -2, CT_ERR_FAILURE = -3,
  CT_CONV_FAILURE = -4, CT_LSETUP_FAIL = -6, CT_ILL_INPUT = -22, CT_TOO_CLOSE = -27,
  CT_ERR_INVALID = -100, CT_ERR_MECH = -101, CT_ERR_NO_DEVICE = -102, CT_ERR_CUDA = -103, CT_ERR_UNSUPPORTED = -104,
  /* Input::Parser failures, one code per exception class the reference throws (tests/src/input_parser.cpp:27-94) */
  CT_ERR_YAML_BAD_FILE = -110 /* YAML::BadFile */, CT_ERR_YAML_INVALID_NODE = -111 /* YAML::InvalidNode */,
  CT_ERR_YAML_BAD_CONVERSION = -112 /* YAML::BadConversion */, CT_ERR_FILESYSTEM = -113 /* boost::filesystem::filesystem_error */,
  CT_ERR_INPUT = -114 /* std::runtime_error */, CT_ERR_INPUT_UNIT = -115 /* std::invalid_argument */
};

enum { CT_LINSOL_LU = 0, CT_LINSOL_INVERSE = 1 };
enum { CT_JAC_DQ = 0 /* CVODE dense difference quotient, what the reference uses (interface.cpp:698-703) */,
       CT_JAC_ANALYTIC = 1 /* default */ };

const char* ct_last_error(void);
const char* ct_version(void);

/* ---- mechanism: Chemistry::Thermo::Create + Chemistry::Kinetics::Create (include/chemistry.hpp:71-78, :166-174).
 * path: Cantera .cti (first ideal_gas phase when phase_id is NULL, as newPhase(file) does) or a compiled ".ctm".
 * constants: NULL = "cantera-2.6"; "cantera-2.4" selects the older gas constant / atomic weights. */
ct_mech* ct_mech_load(const char* path, const char* phase_id_or_null, int* status);
ct_mech* ct_mech_load_ex(const char* path, const char* phase_id_or_null, const char* constants_or_null, int* status);
void ct_mech_destroy(ct_mech*);
int ct_mech_save_ctm(const ct_mech*, const char* path);
int ct_mech_n_species(const ct_mech*);   /* ThermoPhase::nSpecies */
int ct_mech_n_reactions(const ct_mech*); /* Kinetics::nReactions */
int ct_mech_n_elements(const ct_mech*);
int ct_mech_species_index(const ct_mech*, const char* name); /* ThermoPhase::speciesIndex; -1 if absent */
const char* ct_mech_species_name(const ct_mech*, int k);     /* ThermoPhase::speciesName */
const char* ct_mech_reaction_equation(const ct_mech*, int i);/* Kinetics::reactionString */
const char* ct_mech_element_name(const ct_mech*, int e);
double ct_mech_n_atoms(const ct_mech*, int k, int e);
int ct_mech_molecular_weights(const ct_mech*, double* W /*K*/); /* ThermoPhase::getMolecularWeights (base.cpp:102) */
double ct_mech_gas_constant(const ct_mech*);
/* compiled-array access for tests (device reaction order + permutation to file order) */
int ct_mech_compiled_arrays(const ct_mech*, int* perm /*R*/, double* A /*R*/, double* b /*R*/, double* EaR /*R*/);
/* mole <-> mass fractions: what the reference's caller does with setState_TPX + getMassFractions
 * (tests/src/reactor.cpp:57-59); n rows of K, normalised. */
int ct_mech_mole_to_mass(const ct_mech*, int64_t n, const double* X, double* Y);
/* algorithmic flops per call for the roofline: rhs, jac, lu, solve */
int ct_mech_flop_model(const ct_mech*, int reactor_type, double out[4]);

/* ---- YAML case-set input: Input::Parser(config) = Parse + Validate (include/input/parser.hpp:9-28,
 * src/input_parser.cpp:147-170, :345-351).  reactor_type 0..3 selects Parameters.reactor.{const_P, const_V, const_TP,
 * const_TV} (include/input/types.hpp:310-346).  Values come back converted: Pa, K, s, mol/m^3. */
typedef struct ct_input ct_input;
ct_input* ct_input_parse(const char* config_file, int* status);
void ct_input_destroy(ct_input*);
int ct_input_n_case_sets(ct_input*, int reactor_type);
int ct_input_n_cases(ct_input*, int reactor_type, int set);
const char* ct_input_mechanism_path(ct_input*, int reactor_type, int set);
const char* ct_input_mechanism_description(ct_input*, int reactor_type, int set); /* NULL if not defined */
const char* ct_input_set_description(ct_input*, int reactor_type, int set);       /* NULL if not defined */
const char* ct_input_case_description(ct_input*, int reactor_type, int set, int cas);
/* composition_type: 1 concentration, 2 mass fraction, 3 mole fraction (Input::CompositionInputType) */
int ct_input_case(ct_input*, int reactor_type, int set, int cas, double* pressure, double* temperature, double* time,
                  int* composition_type, int* n_pairs);
const char* ct_input_case_species(ct_input*, int reactor_type, int set, int cas, int i, double* value);
/* what the reference leaves to its (missing) driver: a case's composition as mass fractions in species order */
int ct_input_case_mass_fractions(ct_input*, int reactor_type, int set, int cas, const ct_mech*, double* Y /*K*/);
int ct_input_n_warnings(ct_input*);            /* temperature outside the mechanism's range warns, src/input_parser.cpp:199-211 */
const char* ct_input_warning(ct_input*, int i);

/* ---- results: Results::HDF5Writer (src/results/hdf5_writer.cpp:136-217) — /<group>/<set> f64 datasets with the string
 * attributes Name / Notation / Unit registered through Results::Meta::Register (include/results/observer.hpp:59-72).
 * Groups are created from the path. rank 1 = one reactor's series, rank 2 = [reactor][sample]. */
typedef struct ct_h5 ct_h5;
ct_h5* ct_h5_create(void);
void ct_h5_destroy(ct_h5*);
int ct_h5_add_dataset(ct_h5*, const char* path, int rank, const int64_t* dims, const double* data, const char* name,
                      const char* notation, const char* unit);
/* storage of the datasets added afterwards: chunk_elems > 0 = rank-1 datasets unlimited, chunked and deflated as the
 * reference creates them (chunk 1000, level 6: hdf5_writer.cpp:156-181), 0 = contiguous (default).  vlen_strings (file-wide,
 * default 1): attributes as variable-length strings like the reference's H5::StrType(0, H5T_VARIABLE) (:194-208), so that
 * h5py hands scripts/results/results.py `str` values; 0 = fixed-length strings */
int ct_h5_set_layout(ct_h5*, int64_t chunk_elems, int deflate_level, int vlen_strings);
int ct_h5_write(ct_h5*, const char* filename);

/* ---- device selection / plumbing (no reference counterpart; the reference is single-threaded CPU) */
int ct_device_count(void);
int ct_set_device(int ordinal);
/* cudaStream_t to launch on (e.g. torch's current stream); NULL = default stream */
int ct_batch_set_stream(ct_batch*, void* cuda_stream);

/* ---- batch of reactors: Apps::Reactor::Create(Type, thermo, kinetics, T, p, Y, rtol, atol, t_end, write)
 * (include/applications/reactors/factory.hpp:15-72) for n cases at once. The batch borrows the mechanism. */
ct_batch* ct_batch_create(ct_mech*, int reactor_type, int64_t n, int* status);
void ct_batch_destroy(ct_batch*);
int ct_batch_n_state(const ct_batch*); /* K+1 ([T, Y...], base.cpp:202-208) or K (base.cpp:84-87) */
/* initial conditions, host arrays: T [K], p [Pa], Y mass fractions row-major n x K */
int ct_batch_set_ic(ct_batch*, const double* T, const double* p, const double* Y);
/* same, arrays already resident in device memory (no H2D inside) */
int ct_batch_set_ic_device(ct_batch*, const double* dT, const double* dp, const double* dY);
/* Client::SetTolerances (client.hpp:122-134) via Base::SetSolverParameters (base.cpp:112-121) */
int ct_batch_set_tolerances(ct_batch*, double rtol, double atol);
/* Client::SetTolerances(rtol, std::vector<realtype> atol) (client.hpp:122-134, types.hpp:114-138) -> CVodeSVtolerances:
 * one absolute tolerance per state component (N entries), the same for every reactor of the batch */
int ct_batch_set_tolerances_vec(ct_batch*, double rtol, const double* atol /*N*/);
/* Client::SetIntegrationTime(0, t_end) (base.cpp:119-120) -> CVodeSetStopTime (interface.cpp:334-343) */
int ct_batch_set_stop_time(ct_batch*, double t_end);
/* Solver::SetMaxNumSteps (interface.hpp:80; 500 default, the reactor test uses 20000) */
int ct_batch_set_max_steps(ct_batch*, int64_t mxstep);
/* Solver::SetMaxOrder (interface.hpp:83), 1..5 */
int ct_batch_set_max_order(ct_batch*, int q);
/* Solver option surface, same names, validation and defaults as IStrategy::Set* (interface.hpp:77-89,
 * src/sundials/cvode/interface.cpp:61-219, defaults include/sundials/cvode/types.hpp:68-112).  Step sizes: negative ->
 * CT_ILL_INPUT, 0 = CVODE default (estimated h0, hmin 0, hmax inf).  Counts / coefficient: must be positive; defaults
 * 7 error-test failures, 3 Newton iterations, 10 convergence failures, coefficient 0.1 */
int ct_batch_set_init_step(ct_batch*, double h);
int ct_batch_set_min_step(ct_batch*, double h);
int ct_batch_set_max_step(ct_batch*, double h);
int ct_batch_set_max_err_test_fails(ct_batch*, int n);
int ct_batch_set_max_nonlin_iters(ct_batch*, int n);
int ct_batch_set_max_conv_fails(ct_batch*, int n);
int ct_batch_set_nonlin_conv_coef(ct_batch*, double coef);
/* SetMaxWarnMessages: validated (>= 0), no effect (the device stepper prints nothing).  SetStabilityLimitDetection:
 * only `off` (the reference's default, types.hpp:99) is supported, `on` returns CT_ERR_UNSUPPORTED */
int ct_batch_set_max_warn_messages(ct_batch*, int n);
int ct_batch_set_stability_limit_detection(ct_batch*, int on);
/* IStrategy::ReInitialise(time_init, state) (interface.cpp:429-446): restart every reactor from state[n x N] at t0,
 * keeping options, tolerances and the frozen reactor parameters of the original initial conditions */
int ct_batch_reinit(ct_batch*, double t0, const double* state);
int ct_batch_set_jacobian(ct_batch*, int mode);
/* Newton linear algebra: CT_LINSOL_LU = LU + triangular solves (what SUNLinSol_Dense does, interface.cpp:714-721);
 * CT_LINSOL_INVERSE (default) = Gauss-Jordan inverse of the error-weight-equilibrated I - gamma*J once per setup, each
 * Newton solve one mat-vec (no 108-step serial dependency): 15-20 % faster on config 2, same step counts */
int ct_batch_set_linear_solver(ct_batch*, int mode);
/* analytic Jacobian: the column of species j is zeroed iff y_j * ewt_j < -thr.  Cantera clips negative mass
 * fractions, so the right-hand side does not depend on them (the true column is zero, as CVODE's difference quotients
 * see it); but a species hovering a few atol below zero with positive production is about to re-enter the unclipped
 * branch, and a zero column then makes modified Newton overshoot across the kink every step (period-2 sawtooth, step
 * size pinned at ~atol/production: 6e4..2e5 steps per ms for burnt lean CH4, in the reference algorithm too).
 * Default 10: measured optimum on the config-5 slice (max steps 61285 -> 4842, 4x throughput); 0 = CVODE-like */
int ct_batch_set_jacobian_clip(ct_batch*, double thr);
/* scheduling only (results are unaffected): align the warps of a CTA in front of every Newton residual so that they
 * share instruction fetches (the fused kernel is instruction-cache bound); default on */
/* kernel selection only (results are unaffected): the three shipped GRI mechanisms run kernels whose sizes and blob offsets
 * are compile-time constants (ct_mech_static_shape: 1 GRI-3.0, 2 GRI-2.11, 3 GRI-1.2, 0 = any other mechanism: run-time
 * shape); off = run-time shape for every mechanism; default on */
int ct_batch_set_static_shape(ct_batch*, int on);
int ct_mech_static_shape(const ct_mech*);
int ct_batch_set_phase_alignment(ct_batch*, int on); /* 0 off, 1 one group per CTA (default), 2 two groups (even/odd warps) */
int ct_batch_set_alignment_poll_ns(ct_batch*, int ns); /* 0 (default): waiting warps block in hardware on an mbarrier; > 0: they poll with __nanosleep(ns) */
int ct_batch_set_alignment_period(ct_batch*, int period); /* meet at every period-th Newton residual; default 1 */
int ct_batch_set_alignment_max_polls(ct_batch*, int polls); /* bounded wait: give up after that many polls; 0 = unbounded */
/* scheduling only: keep the N x N Newton matrix out of the per-warp shared-memory slab (shared pool of work slabs for
 * Jacobian assembly + inversion, inverse parked in L2-resident global scratch) so that warps_per_sm (default 12)
 * instead of 6 warps fit one SM for GRI-3.0; applies with CT_LINSOL_INVERSE; default on */
int ct_batch_set_matrix_pool(ct_batch*, int on, int warps_per_sm);
/* scheduling only (results are unaffected, the step sequence of every reactor is the uninterrupted one): a warp
 * integrates a reactor for `steps` steps, parks its stepper state in global memory and queues it again, so that all
 * reactors of the batch advance together and no warp idles while late starters finish.  Applies with the matrix pool
 * when the batch exceeds one wave of warps and the reactor-indexed scratch fits 4 GiB.  Default 64; 0 = off */
int ct_batch_set_time_slicing(ct_batch*, int steps);
/* solver diagnostics: one record of 12 doubles per step ATTEMPT of one reactor (nst, t, h, q, Newton flag, error-test
 * flag, dsm, acnrm, component with the largest weighted correction, that correction, its predicted value,
 * 100 * convergence failures + error-test failures of the step); turns time slicing off for the batch */
int ct_batch_set_trace(ct_batch*, int64_t reactor, int capacity);
int ct_batch_get_trace(ct_batch*, double* out /* capacity x 12 */, int* count);
/* scheduling diagnostics: record per reactor [start ns, end ns (GPU global timer), sm * 64 + warp] during Integrate */
int ct_batch_set_timeline(ct_batch*, int on);
int ct_batch_get_timeline(ct_batch*, int64_t* out /* n x 3 */);
/* ignition monitor: mode 0 = first crossing of T >= value (reference test: 1756 K, tests/src/reactor.cpp:130),
 * mode 1 = T >= T0 + value (default, value = 400 K); located on the integrator's dense output */
int ct_batch_set_ignition(ct_batch*, int mode, double value);
/* prescribed T(t), p(t): common time grid t[nt], per-reactor values T[n x nt], p[n x nt] (CT_TABLE_TP only) */
int ct_batch_set_tp_table(ct_batch*, int nt, const double* t, const double* T, const double* p);
/* prescribed volume and its time derivative (CT_TABLE_V only): common grid t[nt], V[n x nt] > 0 in any unit (only
 * V(t) / V(t[0]) enters: rho = rho0 V(t0) / V(t)), dVdt[n x nt] in that unit per second; the energy equation is
 * ConstVol::EnergyDerivative (reactors.cpp:96-109) plus the boundary work -p dV/dt / (V rho cv) */
int ct_batch_set_v_table(ct_batch*, int nt, const double* t, const double* V, const double* dVdt);
/* optional processing order of the work queue (a permutation of 0..n-1), e.g. most expensive first */
int ct_batch_set_order(ct_batch*, const int32_t* order);

/* Solver::Integrate(t) (interface.hpp:297; interface.cpp:454-475), CV_NORMAL semantics, all internal steps on the
 * device. Callable repeatedly with increasing t_target (the 10 us cadence of tests/src/reactor.cpp:105-126).
 * Returns 0 if every reactor returned CV_SUCCESS, 1 if every live reactor reached the stop time, else the most
 * negative per-reactor code. */
int ct_batch_integrate(ct_batch*, double t_target);
/* Integrate(t_stop) followed by a recovery pass over the reactors that ended with a negative CVODE status: they are re-run
 * as a small batch under other Jacobian / Newton options (Jacobian clip 1, then LU, then difference-quotient Jacobian +
 * LU) and the successful ones merged; what IStrategy::Integrate only sketches (interface.cpp:462-473).  Fresh batches only.
 * n_failed_first / n_failed_final (may be NULL): reactors with a negative status before / after the recovery. */
int ct_batch_integrate_with_retry(ct_batch*, int64_t* n_failed_first, int64_t* n_failed_final);
/* one launch covering n_out equally spaced outputs t_k = (k+1) t_end / n_out; traj (host) n x n_out x N */
int ct_batch_integrate_trajectory(ct_batch*, int n_out, double* traj);
/* Solver::Solution() / Client::State() (client.hpp:110-120): n x N host array */
int ct_batch_get_state(ct_batch*, double* out);
int ct_batch_get_state_device(ct_batch*, const double** dptr);
int ct_batch_get_time(ct_batch*, double* t /*n*/);
int ct_batch_get_ignition(ct_batch*, double* tau /*n, NaN = not ignited*/);
int ct_batch_get_status(ct_batch*, int32_t* status /*n*/);
/* IStrategy::MainSolverStatistics / LinearSolverStatistics (interface.cpp:487-623): per reactor
 * nst, nfe, nsetups, netf, nni, ncfn, nje, nfeLS */
int ct_batch_get_stats(ct_batch*, int64_t* stats /*n x 8*/);
/* the rest of MainSolverStatistics (include/sundials/cvode/types.hpp:175-215, filled by CVodeGetLastOrder, GetCurrentOrder,
 * GetActualInitStep, GetLastStep, GetCurrentStep, GetCurrentTime in interface.cpp:487-560): per reactor
 * qlast, qcur, hinused, hlast, hcur, tcur (n_stability_order_reductions is always 0: detection is off) */
int ct_batch_get_stats_ex(ct_batch*, double* stats_ex /*n x 6*/);
/* device time of the last integrate launch in ms (CUDA events on the launch stream) and launch count so far */
double ct_batch_last_kernel_ms(const ct_batch*);
int64_t ct_batch_launch_count(const ct_batch*);
int ct_batch_warps_per_sm(const ct_batch*);

/* ---- multi-GPU ensembles (no reference counterpart: Base::Solve, src/applications/reactors/
 * base.cpp:123-170, integrates one reactor on one host thread).  Reactors are independent, so n reactors are partitioned
 * over the listed devices of one box: one host thread + one CUDA stream per device, each shard processed as a sequence
 * of ordinary batches of at most `chunk` reactors, no collective on the solve path; every device copies its results with
 * cudaMemcpyAsync D2H into disjoint slices of the caller's host arrays (the final gather for the HDF5 writer,
 * src/results/hdf5_writer.cpp:101-110).  Bit-identical to one batch on one device. */
typedef struct ct_ensemble ct_ensemble;
enum { CT_PARTITION_BLOCK = 0 /* reactor i -> shard floor(i*G/n), contiguous slices, results written in place */,
       CT_PARTITION_CYCLIC = 1 /* reactor i -> shard i % G: every shard sees the same stiffness mix of a sorted sweep */ };
/* called on every internal batch right after its creation: apply any ct_batch_set_* option; return < 0 to abort */
typedef int (*ct_batch_configure_fn)(ct_batch*, void* user);
/* page-locked host memory usable by every device (cudaHostAllocPortable): back the ensemble's input / output arrays with
 * it and the per-device copies are true asynchronous DMA transfers */
int ct_host_alloc(int64_t bytes, void** out);
int ct_host_free(void* p);
/* devices: n_devices ordinals (NULL = 0..n_devices-1; n_devices 0 = every visible device) */
ct_ensemble* ct_ensemble_create(ct_mech*, int reactor_type, int64_t n, int n_devices, const int* devices, int* status);
void ct_ensemble_destroy(ct_ensemble*);
int ct_ensemble_n_shards(const ct_ensemble*);
int ct_ensemble_set_partition(ct_ensemble*, int mode, int64_t chunk /* reactors per launch, 0 = keep (default 65536) */);
int ct_ensemble_set_tolerances(ct_ensemble*, double rtol, double atol);
int ct_ensemble_set_stop_time(ct_ensemble*, double t_end);
int ct_ensemble_set_max_steps(ct_ensemble*, int64_t mxstep);
/* recovery pass, once per shard after its last chunk: the ladder of ct_batch_integrate_with_retry over the reactors that
 * ended with a negative status; default on */
int ct_ensemble_set_retry(ct_ensemble*, int on);
/* step budget of the chunk launches when the recovery pass is on (default 10000; 0 = the full budget): a launch lasts as
 * long as its slowest reactor, so reactors that need more steps are cut off and integrated again from the start with the
 * full budget (ct_ensemble_set_max_steps) in the shard's one recovery batch, before the ladder (results unaffected) */
int ct_ensemble_set_first_pass_steps(ct_ensemble*, int64_t steps);
/* launches in flight per device (host threads + streams per shard, each taking every streams-th chunk): the CTAs of a launch
 * leave when the work ring is empty, so the next chunk starts on the freed SMs while the slowest reactors of this one
 * finish; 1..4, default 2 (results unaffected) */
int ct_ensemble_set_streams(ct_ensemble*, int streams);
int ct_ensemble_set_hot_first(ct_ensemble*, int on);  /* queue each chunk hottest T0 first; default on (results unaffected) */
int ct_ensemble_set_configure(ct_ensemble*, ct_batch_configure_fn fn, void* user);
/* CT_TABLE_TP: common grid t[nt], T[n x nt], p[n x nt] (CT_TABLE_V: V[n x nt], dVdt[n x nt] in their place); the arrays are
 * borrowed until ct_ensemble_integrate returns */
int ct_ensemble_set_tp_table(ct_ensemble*, int nt, const double* t, const double* T, const double* p);
/* global reactor indices of a shard: first, first + stride, ... (count of them).  ct_partition_range is the same rule
 * without a handle (host arithmetic only: callers that run one process per GPU use it to pick their slice) */
int ct_partition_range(int64_t n, int n_shards, int mode, int shard, int64_t* first, int64_t* stride, int64_t* count);
int ct_ensemble_shard_range(const ct_ensemble*, int shard, int64_t* first, int64_t* stride, int64_t* count);
/* T[n], p[n], Y[n x K] -> state[n x N], tau[n] (NaN = not ignited), status[n], stats[n x 8] (any output may be NULL).
 * Returns like ct_batch_integrate: 0 / 1 / the most negative per-reactor code; library errors (<= -100) name the shard */
int ct_ensemble_integrate(ct_ensemble*, const double* T, const double* p, const double* Y, double* state, double* tau,
                          int32_t* status, int64_t* stats);
/* per shard after integrate: counts = reactors, chunks, kernel launches, not finished by the chunk launches (negative status,
 * including those cut off by the first-pass budget), still failed after the recovery pass;
 * times = sum of k_integrate device ms (CUDA events; launches of one device overlap), wall seconds of the call */
int ct_ensemble_shard_info(const ct_ensemble*, int shard, int64_t counts[5], double times[2]);

/* ---- stand-alone kernels (parity tests, ncu evidence) ---- */
/* K1: Base/EnergyEnabled::RightHandSide for n states (host arrays). T0,p0,Y0 define the frozen quantities exactly
 * as Reactor::Create does; states n x N; outputs ydot n x N, wdot n x K [kmol/m^3/s], qf/qr n x R (file order; may
 * be NULL). times may be NULL (only CT_TABLE_TP uses it). */
int ct_rhs_eval(ct_batch*, const double* states, const double* times, double* ydot, double* wdot, double* qf, double* qr);
/* K2: Jacobian d(ydot)/d(state), column-major N x N per reactor; mode CT_JAC_DQ or CT_JAC_ANALYTIC */
int ct_jacobian_eval(ct_batch*, int mode, const double* states, const double* times, double* J);
/* K3: batched dense LU (partial pivoting) + solve, or Gauss-Jordan inverse + mat-vec (linsol); A column-major
 * n x N x N, b n x N -> x */
int ct_lu_solve(int N, int64_t n, const double* A, const double* b, double* x, int32_t* info, int linsol);
/* timing loops for the roofline report: run the kernel `reps` times on resident data, return mean ms per launch */
int ct_rhs_time(ct_batch*, const double* states, int reps, double* ms);
/* K1 in the fused kernel's slab layout with warps_per_block warps per CTA and blocks_per_sm CTAs per SM (occupancy scan) */
int ct_rhs_time_cfg(ct_batch*, const double* states, int reps, int warps_per_block, int blocks_per_sm, double* ms);
int ct_jacobian_time(ct_batch*, int mode, const double* states, int reps, double* ms);
int ct_lu_time(int N, int64_t n, int reps, int solves_per_factor, int linsol, double* ms);
/* measured fp64 FMA peak of the current device in TFLOP/s (register-resident DFMA loop) */
int ct_measure_fp64_peak(double* tflops);
/* CVODE Roberts problem on the device stepper (golden file of the reference:
 * tests/data/sundials_cvode/direct_dense_cvRoberts_dns.dat); y_out n_out x 3 */
int ct_roberts_selftest(int n_out, const double* t_out, double rtol, const double* atol3, double* y_out, int64_t* stats8);

#ifdef __cplusplus
}
#endif
#endif /* CHEMTOOLS_B200_H */
