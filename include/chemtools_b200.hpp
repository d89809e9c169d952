// chemtools_b200.hpp -- C++14 binding over the C ABI (chemtools_b200.h): what a ChemTools maintainer adds as
// include/applications/reactors/gpu_batch.hpp.  Header only; link libchemtools_b200.so.
//
// Apps::Reactor::GpuBatch holds n reactors of one Apps::Reactor::Type and mirrors, for the ensemble, what the reference
// spreads over Reactor::Create (include/applications/reactors/factory.hpp:15-72), Reactor::Base
// (include/applications/reactors/base.hpp) and SUNDIALS::CVODE::Solver (include/sundials/cvode/interface.hpp:205-321):
// same method names, argument meaning and error behaviour (invalid options throw std::invalid_argument like
// IStrategy::Set*, src/sundials/cvode/interface.cpp:51-220; Integrate returns CVODE's codes unchanged).
#ifndef CHEMTOOLS_B200_HPP
#define CHEMTOOLS_B200_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "chemtools_b200.h"

namespace Apps {
namespace Reactor {

enum class Type { Const_Pres = CT_CONST_PRES, Const_Vol = CT_CONST_VOL, Const_Temp_Pres = CT_CONST_TEMP_PRES, Const_Temp_Vol = CT_CONST_TEMP_VOL };

// SUNDIALS::CVODE::Types::MainSolverStatistics + LinearSolverStatistics (include/sundials/cvode/types.hpp:175-228), one reactor
struct SolverStatistics {
  long n_internal_steps, n_rhs_evals, n_linear_solver_setups, n_error_test_fails, n_nonlinear_solver_iters, n_nonlinear_solver_conv_fails;
  int last_order_used, current_order;
  double last_internal_step_size, next_internal_step_size, first_internal_step_size, current_time;
  long n_jac_evals, n_rhs_jac_evals;
};

// Chemistry::ThermoKinetics (include/chemistry.hpp:207-226)
class Mechanism {
 public:
  explicit Mechanism(std::string const& path) {
    int st = 0;
    m_h = ct_mech_load(path.c_str(), nullptr, &st);
    if (!m_h) throw std::runtime_error(ct_last_error());
  }
  ~Mechanism() { ct_mech_destroy(m_h); }
  Mechanism(Mechanism const&) = delete;
  Mechanism& operator=(Mechanism const&) = delete;
  size_t nSpecies() const { return static_cast<size_t>(ct_mech_n_species(m_h)); }
  size_t nReactions() const { return static_cast<size_t>(ct_mech_n_reactions(m_h)); }
  size_t speciesIndex(std::string const& name) const {
    int const k = ct_mech_species_index(m_h, name.c_str());
    if (k < 0) throw std::invalid_argument("unknown species " + name);
    return static_cast<size_t>(k);
  }
  // thermo.setState_TPX + getMassFractions (tests/src/reactor.cpp:57-59), n rows
  std::vector<double> MassFractions(std::vector<double> const& mole_fractions) const {
    std::vector<double> y(mole_fractions.size());
    Check(ct_mech_mole_to_mass(m_h, static_cast<int64_t>(mole_fractions.size() / nSpecies()), mole_fractions.data(), y.data()));
    return y;
  }
  ct_mech* handle() const { return m_h; }
  static void Check(int rc) { if (rc < 0) throw std::invalid_argument(ct_last_error()); }

 private:
  ct_mech* m_h;
};

class GpuBatch {
 public:
  // Reactor::Create(type, thermo, kinetics, T, p, Y, rtol, atol, t_end, write_results) for n = T.size() cases;
  // Y is n x nSpecies row-major.  The batch borrows the mechanism.
  GpuBatch(Type type, Mechanism const& mechanism, std::vector<double> const& temperature, std::vector<double> const& pressure,
           std::vector<double> const& mass_fractions, double relative_solver_tolerance, double absolute_solver_tolerance,
           double total_simulation_time)
      : m_n(temperature.size()) {
    if (pressure.size() != m_n || mass_fractions.size() != m_n * mechanism.nSpecies()) throw std::invalid_argument("temperature / pressure / mass fraction sizes disagree");
    int st = 0;
    m_h = ct_batch_create(mechanism.handle(), static_cast<int>(type), static_cast<int64_t>(m_n), &st);
    if (!m_h) throw std::runtime_error(ct_last_error());
    try {
      Check(ct_batch_set_tolerances(m_h, relative_solver_tolerance, absolute_solver_tolerance));
      Check(ct_batch_set_stop_time(m_h, total_simulation_time));
      Check(ct_batch_set_ic(m_h, temperature.data(), pressure.data(), mass_fractions.data()));
    } catch (...) { ct_batch_destroy(m_h); throw; }
    m_N = static_cast<size_t>(ct_batch_n_state(m_h));
  }
  ~GpuBatch() { ct_batch_destroy(m_h); }
  GpuBatch(GpuBatch const&) = delete;
  GpuBatch& operator=(GpuBatch const&) = delete;

  // IStrategy::Set* (include/sundials/cvode/interface.hpp:77-89)
  void SetMaxNumSteps(long n) { Check(ct_batch_set_max_steps(m_h, n)); }
  void SetMaxOrder(int q) { Check(ct_batch_set_max_order(m_h, q)); }
  void SetInitStepSize(double h) { Check(ct_batch_set_init_step(m_h, h)); }
  void SetMinStepSize(double h) { Check(ct_batch_set_min_step(m_h, h)); }
  void SetMaxStepSize(double h) { Check(ct_batch_set_max_step(m_h, h)); }
  void SetMaxErrorTestFailures(int n) { Check(ct_batch_set_max_err_test_fails(m_h, n)); }
  void SetMaxNonlinearIterations(int n) { Check(ct_batch_set_max_nonlin_iters(m_h, n)); }
  void SetMaxConvergenceFailures(int n) { Check(ct_batch_set_max_conv_fails(m_h, n)); }
  void SetNonlinConvergenceCoefficient(double c) { Check(ct_batch_set_nonlin_conv_coef(m_h, c)); }
  // Client::SetTolerances with a vector of absolute tolerances (include/sundials/cvode/client.hpp:122-134)
  void SetTolerances(double rtol, std::vector<double> const& atol) {
    if (atol.size() != m_N) throw std::invalid_argument("absolute tolerance vector must have one entry per state component");
    Check(ct_batch_set_tolerances_vec(m_h, rtol, atol.data()));
  }
  void SetIgnitionThreshold(double temperature) { Check(ct_batch_set_ignition(m_h, 0, temperature)); }

  // Solver::Integrate(t): CVODE's return code (0 CV_SUCCESS, 1 CV_TSTOP_RETURN, < 0 the worst per-reactor failure)
  int Integrate(double time) {
    int const rc = ct_batch_integrate(m_h, time);
    if (rc <= CT_ERR_INVALID) throw std::runtime_error(ct_last_error());
    return rc;
  }
  // Integrate to the stop time and re-run failed reactors under other Jacobian / Newton options
  int IntegrateWithRetry(int64_t* failed_first = nullptr, int64_t* failed_final = nullptr) {
    int const rc = ct_batch_integrate_with_retry(m_h, failed_first, failed_final);
    if (rc <= CT_ERR_INVALID) throw std::runtime_error(ct_last_error());
    return rc;
  }
  // Solver::Solution() / Client::State(): n x N row-major, N = 1 + nSpecies ([T, Y...]) or nSpecies
  std::vector<double> Solution() const {
    std::vector<double> s(m_n * m_N);
    Check(ct_batch_get_state(m_h, s.data()));
    return s;
  }
  std::vector<double> IgnitionDelay() const { std::vector<double> t(m_n); Check(ct_batch_get_ignition(m_h, t.data())); return t; }
  std::vector<int32_t> Status() const { std::vector<int32_t> s(m_n); Check(ct_batch_get_status(m_h, s.data())); return s; }
  // IStrategy::MainSolverStatistics / LinearSolverStatistics (src/sundials/cvode/interface.cpp:487-623)
  std::vector<SolverStatistics> Statistics() const {
    std::vector<int64_t> c(m_n * 8);
    std::vector<double> e(m_n * 6);
    Check(ct_batch_get_stats(m_h, c.data()));
    Check(ct_batch_get_stats_ex(m_h, e.data()));
    std::vector<SolverStatistics> out(m_n);
    for (size_t i = 0; i < m_n; ++i) {
      SolverStatistics& s = out[i];
      s.n_internal_steps = static_cast<long>(c[8 * i]); s.n_rhs_evals = static_cast<long>(c[8 * i + 1]); s.n_linear_solver_setups = static_cast<long>(c[8 * i + 2]);
      s.n_error_test_fails = static_cast<long>(c[8 * i + 3]); s.n_nonlinear_solver_iters = static_cast<long>(c[8 * i + 4]);
      s.n_nonlinear_solver_conv_fails = static_cast<long>(c[8 * i + 5]); s.n_jac_evals = static_cast<long>(c[8 * i + 6]); s.n_rhs_jac_evals = static_cast<long>(c[8 * i + 7]);
      s.last_order_used = static_cast<int>(e[6 * i]); s.current_order = static_cast<int>(e[6 * i + 1]); s.first_internal_step_size = e[6 * i + 2];
      s.last_internal_step_size = e[6 * i + 3]; s.next_internal_step_size = e[6 * i + 4]; s.current_time = e[6 * i + 5];
    }
    return out;
  }
  size_t n() const { return m_n; }
  size_t N() const { return m_N; }

 private:
  static void Check(int rc) { if (rc < 0) throw std::invalid_argument(ct_last_error()); }
  ct_batch* m_h;
  size_t m_n, m_N = 0;
};

}  // namespace Reactor
}  // namespace Apps

#endif  // CHEMTOOLS_B200_HPP
