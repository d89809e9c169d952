// C++14 caller of libchemtools_b200.so through include/chemtools_b200.hpp: the reference's only reactor scenario
// (/root/reference/tests/src/reactor.cpp:50-140: constant-volume GRI-3.0 reactor, H2:2 O2:1 N2:4 at 1001 K and 1 atm,
// rtol 1e-9, atol 1e-15, 20000 steps, Integrate(t) every 10 us, CV_SUCCESS before the end and CV_TSTOP_RETURN at it,
// ignition = first output sample with T >= 1756 K), with the same control flow and the same assertions, plus the option
// validation of tests for IStrategy::Set*.  Built and run by tests/test_gpu_parity.py::test_cxx_binding (needs a GPU).
// usage: gpu_batch_test <mechanism.ctm|.cti>   -> prints "ignition <t> T_end <T> steps <n>" and exits 0
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "chemtools_b200.hpp"

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "REQUIRE failed: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: gpu_batch_test <mechanism>\n"); return 2; }
  using namespace Apps;
  Reactor::Mechanism chemistry(argv[1]);
  double const pressure = 101325.0;  // Cantera::OneAtm
  double const temperature = 1001.0;
  size_t const n_specs = chemistry.nSpecies();
  REQUIRE(n_specs == 53 && chemistry.nReactions() == 325);
  std::vector<double> mole_fractions(n_specs, 0.0);
  mole_fractions[chemistry.speciesIndex("H2")] = 2.0;
  mole_fractions[chemistry.speciesIndex("O2")] = 1.0;
  mole_fractions[chemistry.speciesIndex("N2")] = 4.0;
  std::vector<double> const mass_fractions = chemistry.MassFractions(mole_fractions);
  double const total_simulation_time = 1.0e-3, relative_solver_tolerance = 1.0e-9, absolute_solver_tolerance = 1.0e-15;

  Reactor::GpuBatch reactor(Reactor::Type::Const_Vol, chemistry, {temperature}, {pressure}, mass_fractions, relative_solver_tolerance,
                            absolute_solver_tolerance, total_simulation_time);
  reactor.SetMaxNumSteps(20000);
  // setters validate like IStrategy::Set* (interface.cpp:61-219): std::invalid_argument
  int thrown = 0;
  try { reactor.SetMaxNumSteps(-1); } catch (std::invalid_argument const&) { ++thrown; }
  try { reactor.SetInitStepSize(-1.0); } catch (std::invalid_argument const&) { ++thrown; }
  try { reactor.SetMaxOrder(6); } catch (std::invalid_argument const&) { ++thrown; }
  try { reactor.SetMaxNonlinearIterations(0); } catch (std::invalid_argument const&) { ++thrown; }
  REQUIRE(thrown == 4);

  double time = 0.0;
  double const time_step = 1.0e-5;
  bool ignited = false;
  double time_ignition = 0.0, temp = 0.0;
  while (time < total_simulation_time) {
    time += time_step;
    int const ret_code = reactor.Integrate(time);
    if (time < total_simulation_time) REQUIRE(0 /* CV_SUCCESS */ == ret_code);
    else REQUIRE(1 /* CV_TSTOP_RETURN */ == ret_code);
    std::vector<double> const sol = reactor.Solution();
    temp = sol[0];
    if (temp >= 1756.0 && !ignited) { time_ignition = time; ignited = true; }
  }
  REQUIRE(ignited);
  REQUIRE(std::fabs(time_ignition - 0.31e-3) < 1e-9);  // the slide's ignition sample
  Reactor::SolverStatistics const st = reactor.Statistics()[0];
  REQUIRE(st.n_internal_steps > 500 && st.n_internal_steps < 5000 && st.last_order_used >= 1 && st.last_order_used <= 5);
  REQUIRE(st.current_time == total_simulation_time && st.first_internal_step_size > 0.0 && st.last_internal_step_size > 0.0);
  std::printf("ignition %.6e T_end %.9f steps %ld qlast %d hlast %.6e\n", time_ignition, temp, st.n_internal_steps, st.last_order_used, st.last_internal_step_size);

  // an ensemble with per-component absolute tolerances (Client::SetTolerances with a vector) and the retry entry point
  std::vector<double> T = {1100.0, 1250.0, 1400.0}, p(3, pressure), Y;
  for (int i = 0; i < 3; ++i) Y.insert(Y.end(), mass_fractions.begin(), mass_fractions.end());
  Reactor::GpuBatch batch(Reactor::Type::Const_Pres, chemistry, T, p, Y, 1e-8, 1e-14, 1e-3);
  batch.SetMaxNumSteps(50000);
  std::vector<double> atol(batch.N(), 1e-14);
  atol[0] = 1e-8;  // temperature
  batch.SetTolerances(1e-8, atol);
  int64_t f0 = -1, f1 = -1;
  int const rc = batch.IntegrateWithRetry(&f0, &f1);
  REQUIRE(rc >= 0 && f0 == 0 && f1 == 0);
  std::vector<double> const tau = batch.IgnitionDelay();
  REQUIRE(tau[0] > tau[1] && tau[1] > tau[2] && tau[2] > 0.0);
  std::printf("ensemble tau %.6e %.6e %.6e\n", tau[0], tau[1], tau[2]);
  return 0;
}
