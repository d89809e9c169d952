"""Input::Parser KATs, ported from the reference's tests/src/input_parser.cpp:23-436 (same scenarios, same converted
values, one exception class per failure mode).  Fixtures are generated here with this repository's mechanism path."""
import os

import numpy as np
import pytest

CASE = """    Case:
      Description: {desc}
      Pressure:
        Unit:  {pu}
        {pv}
      Temperature:
        Unit:  {tu}
        Value: {tv}
      Composition:
        Type:  {ctype}
{cunit}        Value: {comp}
      Time:
        Unit:  {timeu}
        Value: {timev}
"""


def case(desc="useful case", pu="Pa", pv="Value: 101325.0", tu="Kelvin", tv=1300.0, ctype="mole fraction", cunit=None,
         comp="[[CH4, 0.1], [O2, 0.5], [N2, 0.4]]", timeu="ms", timev=1.0):
    return CASE.format(desc=desc, pu=pu, pv=pv, tu=tu, tv=tv, ctype=ctype, cunit="" if cunit is None else "        Unit:  %s\n" % cunit,
                       comp=comp, timeu=timeu, timev=timev)


def case_set(reactor, mech, cases, desc="case set"):
    return ("- Reactor:        %s\n  Mechanism:\n    Description:  %s - mechanism\n    Path:         %s\n  Description:    %s\n  Cases:\n"
            % (reactor, reactor, mech, desc)) + "".join(cases)


@pytest.fixture()
def mech(ctb):
    return ctb.mechanism_path("gri30")


@pytest.fixture()
def write(tmp_path):
    def w(text, name="config.yaml"):
        p = tmp_path / name
        p.write_text(text)
        return str(p)
    return w


def test_error_classes(ctb, mech, write):
    from chemtools_b200 import input as I
    with pytest.raises(I.BadFile):                                   # inexistent config -> YAML::BadFile
        I.Parser("../non_existent_config.yaml")
    with pytest.raises(I.InvalidNode):                               # pressure value missing -> YAML::InvalidNode
        I.Parser(write(case_set("CONSTANT_PRESSURE", mech, [case(pv="#Value: 101325.0 # commented mandatory key")])))
    with pytest.raises(I.BadConversion):                             # string instead of double -> YAML::BadConversion
        I.Parser(write(case_set("CONSTANT_PRESSURE", mech, [case(pv="Value: wrong_type_should_be_double")])))
    with pytest.raises(I.FilesystemError):                           # wrong mechanism path -> filesystem_error
        I.Parser(write(case_set("CONSTANT_PRESSURE", "../data/no_such_mech.cti", [case()])))
    for bad in (case(pv="Value: -101325.0"),                          # negative pressure
                case(tv=-5.0),                                        # negative temperature
                case(comp="[[CH4, 0.1], [O2, 0.5], [N2, 0.3], [CH4, 0.1]]"),   # duplicate species
                case(comp="[[CH4, -0.1], [O2, 0.6], [N2, 0.5]]"),     # negative species
                case(comp="[[CH4, 0.1], [O2, 0.5], [XYZ, 0.4]]"),     # species not in the mechanism
                case(comp="[[CH4, 0.1], [O2, 0.5], [N2, 0.5]]")):     # fractions do not sum to one
        with pytest.raises(I.InputRuntimeError):
            I.Parser(write(case_set("CONSTANT_PRESSURE", mech, [bad])))
    with pytest.raises(I.InvalidUnit):
        I.Parser(write(case_set("CONSTANT_PRESSURE", mech, [case(pu="parsec")])))
    with pytest.raises(I.InvalidUnit):                               # convertible check: a time unit for a pressure
        I.Parser(write(case_set("CONSTANT_PRESSURE", mech, [case(pu="ms")])))


def test_multiple_reactors(ctb, mech, write):
    """config_multiple_reactors.yaml: 2/1/1/1 case sets with 3,2 / 2 / 1 / 2 cases (tests/src/input_parser.cpp:97-141)."""
    from chemtools_b200 import input as I
    text = (case_set("CONSTANT_PRESSURE", mech, [case("1.%d" % i) for i in range(3)]) +
            case_set("CONSTANT_PRESSURE", mech, [case("2.%d" % i) for i in range(2)]) +
            case_set("CONSTANT_VOLUME", mech, [case("3.%d" % i) for i in range(2)]) +
            case_set("CONSTANT_TEMPERATURE_PRESSURE", mech, [case("4.0")]) +
            case_set("CONSTANT_TEMPERATURE_VOLUME", mech, [case("5.%d" % i) for i in range(2)]) +
            case_set("SOMETHING_ELSE", mech, [case("ignored")]))
    p = I.Parser(write(text)).GetParameters()
    assert [len(s.Cases()) for s in p.const_P] == [3, 2]
    assert [len(s.Cases()) for s in p.const_V] == [2]
    assert [len(s.Cases()) for s in p.const_TP] == [1]
    assert [len(s.Cases()) for s in p.const_TV] == [2]
    assert p.const_P[0].cases[2].description == "1.2" and p.const_P[0].description == "case set"
    assert p.const_P[0].mechanism_description == "CONSTANT_PRESSURE - mechanism"


def test_converted_values(ctb, chem, mech, write):
    """config.yaml of the reference: exact converted values (tests/src/input_parser.cpp:164-382)."""
    from chemtools_b200 import input as I
    text = case_set("CONSTANT_PRESSURE", mech, [
        case("useful case 1.1", "Pa", "Value: 101325.0", "Kelvin", 1300.0, "mole fraction", None, "[[CH4, 0.1], [O2, 0.5], [N2, 0.4]]", "ms", 1.0),
        case("useful case 1.2", "atm", "Value: 2.0", "degC", 1025.0, "concentration", "mmol/L", "[[CH4, 0.1], [O2, 0.5], [N2, 0.35], [H2, 0.05] ]", "min", 1.0),
        case("useful case 1.3", "bar", "Value: 1.5", "degF", 2003.0, "concentration", "mol/m^3", "[[CH3, 0.03], [CH4, 0.07], [O2, 0.5], [N2, 0.4]]", "h", 1.0),
        case("useful case 1.4", "Pascal", "Value: 201326.0", "K", 1250.0, "mass fraction", None, "[[CH4, 0.15], [O2, 0.45], [N2, 0.4]]", "s", 1.0)],
        desc="case set 1")
    parser = I.Parser(write(text))
    p = parser.GetParameters()
    assert p.const_P and not p.const_V and not p.const_TP and not p.const_TV
    cases = p.const_P[0].Cases()
    assert len(cases) == 4
    c = cases[0]
    assert (c.pressure, c.temperature, c.time) == (101325.0, 1300.0, pytest.approx(1.0e-3, rel=1e-15))
    assert c.composition_type == "MoleFraction" and c.pairs == [("CH4", 0.1), ("O2", 0.5), ("N2", 0.4)]
    c = cases[1]
    assert c.pressure == 202650.0 and c.temperature == pytest.approx(1298.15, rel=1e-15) and c.time == 60.0
    assert c.composition_type == "Concentration" and [n for n, _ in c.pairs] == ["CH4", "O2", "N2", "H2"]
    assert [v for _, v in c.pairs] == pytest.approx([0.1, 0.5, 0.35, 0.05], rel=1e-15)   # mmol/L == mol/m^3
    c = cases[2]
    assert c.pressure == pytest.approx(150000.0, rel=1e-15) and c.temperature == pytest.approx(1368.15, rel=1e-14) and c.time == 3600.0
    c = cases[3]
    assert (c.pressure, c.temperature, c.time) == (201326.0, 1250.0, 1.0) and c.composition_type == "MassFraction"
    assert cases[3].DescriptionIsDefined() and cases[3].description == "useful case 1.4"
    # Case -> (T, p, Y): the step missing in the reference
    ch = chem("gri30")
    Y = parser.mass_fractions(cases[0], ch)
    assert np.abs(Y - ch.mass_fractions(ch.composition({"CH4": 0.1, "O2": 0.5, "N2": 0.4}))[0]).max() < 1e-15
    Y = parser.mass_fractions(cases[1], ch)   # concentrations are proportional to mole numbers
    assert np.abs(Y - ch.mass_fractions(ch.composition({"CH4": 0.1, "O2": 0.5, "N2": 0.35, "H2": 0.05}))[0]).max() < 1e-15
    Y = parser.mass_fractions(cases[3], ch)
    assert Y[ch.speciesIndex("CH4")] == pytest.approx(0.15) and abs(Y.sum() - 1) < 1e-15


def test_temperature_outside_range_warns(ctb, mech, write):
    from chemtools_b200 import input as I
    p = I.Parser(write(case_set("CONSTANT_PRESSURE", mech, [case(tv=150.0)])))
    assert len(p.warnings) == 1 and "outside the mechanism's applicable" in p.warnings[0]


def test_yaml_reader_details(ctb, mech, write):
    """Duplicate `Case` keys are all kept; comments, quotes and flow mappings are understood."""
    from chemtools_b200 import input as I
    text = ("# leading comment\n- Reactor: 'CONSTANT_VOLUME'   # trailing comment\n  Mechanism: {Path: \"%s\"}\n  Cases:\n" % mech +
            case("a") + case("b") + "    NotACase:\n      x: 1\n" + case("c"))
    p = I.Parser(write(text))
    assert [c.description for c in p.const_V[0].cases] == ["a", "b", "c"]
    assert p.const_V[0].mechanism_description is None and p.const_V[0].description is None


# ---- the reference's own fixtures, verbatim (tests/golden/yaml_input/ = /root/reference/tests/data/yaml_input/), and its
#      test tests/src/input_parser.cpp:23-382 ported assertion by assertion.  Only the mechanism path is rewritten (the
#      fixtures point at ../data/gri30.cti relative to the reference's build directory).
GOLDEN_YAML = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "yaml_input")


@pytest.fixture()
def golden(ctb, tmp_path):
    def g(name):
        text = open(os.path.join(GOLDEN_YAML, name)).read()
        assert "../data/gri30.cti" in text or "non_existent_mechanism" in text
        p = tmp_path / name
        p.write_text(text.replace("../data/gri30.cti", ctb.mechanism_path("gri30")))
        return str(p)
    return g


def test_reference_fixtures_error_classes(ctb, golden):
    """tests/src/input_parser.cpp:27-94: one exception class per malformed file."""
    from chemtools_b200 import input as I
    with pytest.raises(I.BadFile):
        I.Parser("../non_existent_config.yaml")
    for name, exc in (("config_invalid_node.yaml", I.InvalidNode), ("config_bad_conversion.yaml", I.BadConversion),
                      ("config_invalid_mech.yaml", I.FilesystemError), ("config_negative_press.yaml", I.InputRuntimeError),
                      ("config_negative_temp.yaml", I.InputRuntimeError), ("config_duplicate_spec.yaml", I.InputRuntimeError),
                      ("config_negative_spec.yaml", I.InputRuntimeError), ("config_invalid_spec.yaml", I.InputRuntimeError),
                      ("config_nonconservative_comp.yaml", I.InputRuntimeError)):
        with pytest.raises(exc):
            I.Parser(golden(name))
        assert exc is I.InputRuntimeError or not issubclass(exc, I.InputRuntimeError)


def test_reference_fixture_multiple_reactors(ctb, golden):
    """tests/src/input_parser.cpp:97-141."""
    from chemtools_b200 import input as I
    p = I.Parser(golden("config_multiple_reactors.yaml")).GetParameters()
    assert len(p.const_P) == 2 and len(p.const_P[0].Cases()) == 3 and len(p.const_P[1].Cases()) == 2
    assert len(p.const_V) == 1 and len(p.const_V[0].Cases()) == 2
    assert len(p.const_TP) == 1 and len(p.const_TP[0].Cases()) == 1
    assert len(p.const_TV) == 1 and len(p.const_TV[0].Cases()) == 2


def test_reference_fixture_config_values(ctb, golden):
    """tests/src/input_parser.cpp:143-382: config.yaml, exact equality as in the reference's REQUIREs."""
    from chemtools_b200 import input as I
    p = I.Parser(golden("config.yaml")).GetParameters()
    assert p.const_P and not p.const_V and not p.const_TP and not p.const_TV
    assert len(p.const_P) == 1
    cases = p.const_P[0].Cases()
    assert len(cases) == 4
    c = cases[0]
    assert c.pressure == 101325.0 and c.temperature == 1300.0 and c.time == 1.0e-3
    assert c.composition_type == "MoleFraction" and c.pairs == [("CH4", 0.1), ("O2", 0.5), ("N2", 0.4)]
    assert sum(v for _, v in c.pairs) == 1.0
    c = cases[1]
    assert c.pressure == 202650.0 and c.temperature == 1298.15 and c.time == 60.0
    assert c.composition_type == "Concentration" and c.pairs == [("CH4", 0.1), ("O2", 0.5), ("N2", 0.35), ("H2", 0.05)]
    c = cases[2]
    assert c.pressure == 150000.0 and c.temperature == 1368.15 and c.time == 3600.0
    assert c.composition_type == "Concentration" and c.pairs == [("CH3", 0.03), ("CH4", 0.07), ("O2", 0.5), ("N2", 0.4)]
    c = cases[3]
    assert c.pressure == 201326.0 and c.temperature == 1250.0 and c.time == 1.0
    assert c.composition_type == "MassFraction" and c.pairs == [("CH4", 0.15), ("O2", 0.45), ("N2", 0.4)]
    assert sum(v for _, v in c.pairs) == 1.0
    assert p.const_P[0].description == "case set 1" and p.const_P[0].mechanism_description == "Constant pressure reactor - mechanism 1"
    assert [k.description for k in cases] == ["useful case 1.%d" % i for i in range(1, 5)]


def test_units_follow_the_llnl_units_library(ctb, mech, write):
    """Bare "C" and "F" are coulomb and farad in the units library the reference uses, not temperature scales: the
    reference rejects them as not convertible to K (std::invalid_argument)."""
    from chemtools_b200 import input as I
    for u in ("C", "F"):
        with pytest.raises(I.InvalidUnit):
            I.Parser(write(case_set("CONSTANT_PRESSURE", mech, [case(tu=u)])))
