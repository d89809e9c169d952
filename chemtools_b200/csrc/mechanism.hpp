// Mechanism compiler (host side): Cantera .cti  ->  flat device blob.
//
// Replaces, for the reactor path, what the reference obtains from
// Chemistry::ThermoKinetics(path) -> Cantera::newPhase / newKineticsMgr
// (/root/reference/include/chemistry.hpp:71-78, :166-174, :207-226): the first
// ideal_gas phase of the file, its species in declaration order, all reactions.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ctb {

struct MechError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ---- file-level description (file units: mol-cm-s-cal/mol) ----
struct SpeciesDef {
  std::string name;
  std::vector<std::pair<std::string, double>> atoms;
  double Tlo = 0, Tmid = 0, Thi = 0;
  double a_lo[7] = {0}, a_hi[7] = {0};
};

enum class RxnKind { Elementary = 0, ThreeBody = 1, Falloff = 2 };

struct ReactionDef {
  RxnKind kind = RxnKind::Elementary;
  std::string equation;
  std::vector<std::pair<std::string, int>> reactants, products;  // file order, integer coefficients
  bool reversible = true, duplicate = false;
  double kf[3] = {0, 0, 0};
  double kf0[3] = {0, 0, 0};
  bool troe = false;
  double troe_p[4] = {0, 0, 0, 0};  // A, T3, T1, T2
  std::vector<std::pair<std::string, double>> efficiencies;
};

struct MechanismDef {
  std::string phase;
  std::string u_length = "cm", u_time = "s", u_quantity = "mol", u_act_energy = "cal/mol";
  std::vector<std::string> elements;
  std::vector<SpeciesDef> species;  // phase order
  std::vector<ReactionDef> reactions;
  int n_phases = 1;
};

// Parse a Cantera CTI file (Python-call subset of SURVEY.md appendix B).
MechanismDef parse_cti(const std::string& path, const char* phase_id_or_null);
// Line-oriented compiled-mechanism text (".ctm"): the same description, exact round trip.
MechanismDef read_ctm(const std::string& path);
void write_ctm(const MechanismDef&, const std::string& path);
// Parse "2 O + M <=> O2 + M" style equations.
void parse_equation(const std::string& eq, ReactionDef& out, bool& has_M, bool& has_falloff_M);

// ---- physical constant sets (Cantera release dependent) ----
struct ConstantSet {
  std::string name;
  double gas_constant, one_atm, cal_to_J;
  std::map<std::string, double> weights;
};
const ConstantSet& constant_set(const std::string& name);  // throws MechError
extern const char* const kDefaultConstants;

// ---- compiled mechanism (Cantera kmol-m-s-K-J units, device reaction order) ----
// Device reaction order: [falloff | three-body | elementary]; perm[d] = file index.
struct CompiledMech {
  int K = 0, R = 0, n_fall = 0, n_3b = 0, KP = 64;
  double kc_fast_Tmin = 400.0;  // above this temperature 1/Kc may be formed from per-species factors exp(mu0/RT) (range check)
  int n_const = 0;  // trailing elementary reactions with b = 0 and Ea = 0 (k_f = A exactly: no exponential needed)
  int n_exp_end = 0;  // device index where that tail starts (falloff + three-body + temperature-dependent region, the latter padded to a multiple of 32)
  double gas_constant = 0, p_ref = 0;
  std::string constants;
  std::vector<std::string> species_names, equations /*file order*/, elements;
  std::vector<double> natoms;  // [K][n_elements]
  std::vector<double> W;
  // file-order views (for host-side queries / tests)
  std::vector<int> perm, inv_perm;
  // device blob + offsets (bytes from blob start, all 16-byte aligned)
  std::vector<unsigned char> blob;
  struct Offsets {
    int W, rW, Tmid, nasa, A, b, E, A0, b0, E0, troe_a, troe_rt3, troe_rt1, troe_t2, eff_val, eff_ptr,
        sp_ptr, sp_rxn, sp_nu, eff_o, ftype, rxo, ent, pstart, pnplus, lpp, spp, n_pieces, n_eff, n_nnz, total;
  } off{};
  // roofline bookkeeping: algorithmic flop counts per call (SURVEY.md §8d contract)
  double flops_rhs = 0, flops_jac = 0, flops_lu = 0, flops_solve = 0;
  int n_transcendental = 0;
};

CompiledMech compile_mechanism(const MechanismDef&, const std::string& constants_name);

}  // namespace ctb
