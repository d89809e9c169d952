// Input::Parser of the reference (include/input/parser.hpp:9-28, include/input/types.hpp:245-346), host side.
#pragma once
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "mechanism.hpp"

namespace ctb {

// the exception classes the reference's parser tests assert on (tests/src/input_parser.cpp:27-94)
struct InputError : std::runtime_error {
  enum Kind { BadFile = 1 /* YAML::BadFile */, InvalidNode /* YAML::InvalidNode */, BadConversion /* YAML::BadConversion */,
              Filesystem /* boost::filesystem::filesystem_error */, Runtime /* std::runtime_error */, InvalidArgument /* std::invalid_argument */ };
  Kind kind;
  InputError(Kind k, const std::string& m) : std::runtime_error(m), kind(k) {}
};

struct YamlNode {
  enum Kind { Null, Scalar, Sequence, Mapping } kind = Null;
  std::string scalar;
  std::vector<YamlNode> items;
  std::vector<std::pair<std::string, YamlNode>> entries;  // insertion order, duplicate keys preserved
  const YamlNode* find(const std::string& key) const;     // first entry with that key
};
YamlNode load_yaml(const std::string& path);

struct InputCase {  // Input::Case (types.hpp:245-263): SI after conversion
  enum CompositionType { Invalid = 0, Concentration = 1, MassFraction = 2, MoleFraction = 3 };
  std::string description;
  bool has_description = false;
  double pressure = 0, temperature = 0, time = 0;       // Pa, K, s
  CompositionType composition_type = Invalid;
  std::vector<std::pair<std::string, double>> composition;  // fractions, or mol/m^3 for concentrations
};
struct InputCaseSet {  // Input::CaseSet
  std::string mechanism_path, mechanism_description, description;
  bool has_mechanism_description = false, has_description = false;
  std::vector<InputCase> cases;
};
struct InputParameters {  // Input::Parameters.reactor
  std::vector<InputCaseSet> const_P, const_V, const_TP, const_TV;
};

InputParameters parse_input(const std::string& config_file);  // Parser::Parse
// Parser::Validate for one case (TemperatureIsValid / PressureIsValid / CompositionIsValid)
void validate_case(const InputCase&, const CompiledMech&, double Tmin, double Tmax, std::vector<std::string>* warnings);
// the step the reference lacks: composition of a case as mass fractions in mechanism species order
std::vector<double> case_mass_fractions(const InputCase&, const CompiledMech&);

}  // namespace ctb
