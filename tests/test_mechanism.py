"""Mechanism compiler (C++ CTI parser + blob) against the oracle's independent Python front end, the reference's
structure KAT, the shipped .ctm files and the C-ABI surface.  CPU tier: no GPU needed (host-only entry points)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import REFERENCE, ROOT, needs_reference


@pytest.mark.parametrize("name", ["gri12", "gri211", "gri30"])
def test_structure_kat(chem, goldens, name):
    """tests/src/chemistry.cpp:250-322 of the reference."""
    g = goldens["structure"][name]
    ch = chem(name)
    assert ch.nSpecies() == g["n_species"] and ch.nReactions() == g["n_reactions"]
    for sp, idx in g["species_index"].items():
        assert ch.speciesIndex(sp) == idx
    for i, eq in g["reaction"].items():
        assert ch.reactionString(int(i)) == eq
    assert ch.speciesIndex("NOPE") == -1 and ch.reactionString(10 ** 6) is None


@pytest.mark.parametrize("name", ["gri12", "gri211", "gri30"])
def test_compiled_arrays_match_oracle_frontend(chem, omech, name):
    """Unit conversion and ordering: product compiler vs oracle (two independent implementations)."""
    ch, om = chem(name), omech(name)
    perm, A, b, E = ch.compiled_arrays()
    assert sorted(perm.tolist()) == list(range(om.nR))
    kinds = om.rtype[perm]
    assert np.all(np.diff(np.array([{2: 0, 1: 1, 0: 2}[k] for k in kinds])) >= 0)  # falloff | three-body | elementary
    assert np.abs(A / om.A[perm] - 1).max() < 1e-15
    assert np.array_equal(b, om.b[perm])
    assert np.abs(E - om.EaR[perm]).max() <= 1e-12 * np.abs(om.EaR).max()
    assert np.abs(ch.getMolecularWeights() / om.W - 1).max() < 1e-15
    assert ch.species_names == om.species
    assert [ch.reactionString(i) for i in range(om.nR)] == om.equations
    for k in range(om.K):
        for e, el in enumerate(ch.elements()):
            assert ch.nAtoms(k, e) == om.natoms[k, [x.lower() for x in om.elements].index(el.lower())]


@needs_reference
@pytest.mark.parametrize("name", ["gri12", "gri211", "gri30"])
def test_cti_parser_matches_shipped_ctm(ctb, chem, name, tmp_path):
    """The shipped .ctm files are exactly what the C++ parser produces from the reference's .cti files."""
    a = ctb.ThermoKinetics(os.path.join(REFERENCE, "data", name + ".cti"))
    b = chem(name)
    for x, y in zip(a.compiled_arrays(), b.compiled_arrays()):
        assert np.array_equal(x, y)
    out = tmp_path / (name + ".ctm")
    a.save_ctm(out)
    assert out.read_text() == open(ctb.mechanism_path(name)).read()


@needs_reference
def test_phase_selection_and_constants(ctb):
    """newPhase(file) takes the first ideal_gas entry (chemistry.hpp:75); named phases and constant sets are selectable."""
    p = os.path.join(REFERENCE, "data", "gri30.cti")
    a = ctb.ThermoKinetics(p)
    b = ctb.ThermoKinetics(p, phase_id="gri30_mix")
    assert a.species_names == b.species_names
    with pytest.raises(ctb.reactors.CtError):
        ctb.ThermoKinetics(p, phase_id="nope")
    c = ctb.ThermoKinetics(p, constants="cantera-2.4")
    assert c.gas_constant() == 8314.4621 and a.gas_constant() == 8314.46261815324
    with pytest.raises(ctb.reactors.CtError):
        ctb.ThermoKinetics(p, constants="nope")


def test_parser_error_paths(ctb, tmp_path):
    """Unsupported CTI constructs fail loudly, missing files give an error not a crash."""
    with pytest.raises(ctb.reactors.CtError):
        ctb.ThermoKinetics(str(tmp_path / "missing.cti"))
    base = '''units(length="cm", time="s", quantity="mol", act_energy="cal/mol")
ideal_gas(name="g", elements="H", species="H2 H", reactions="all")
species(name="H2", atoms="H:2", thermo=(NASA([200,1000],[1,0,0,0,0,0,0]), NASA([1000,3000],[1,0,0,0,0,0,0])))
species(name="H", atoms="H:1", thermo=(NASA([200,1000],[1,0,0,0,0,0,0]), NASA([1000,3000],[1,0,0,0,0,0,0])))
'''
    ok = tmp_path / "ok.cti"
    ok.write_text(base + 'three_body_reaction("2 H + M <=> H2 + M", [1e18, -1, 0], efficiencies="H2:2.5")\n')
    m = ctb.ThermoKinetics(str(ok))
    assert m.nSpecies() == 2 and m.reactionString(0) == "2 H + M <=> H2 + M"
    for bad in ('reaction("H2 <=> 2 H", [1e10, 0, 0], order="H2:0.5")\n',            # explicit orders
                'falloff_reaction("2 H (+ M) <=> H2 (+ M)", kf=[1,0,0], kf0=[1,0,0], falloff=SRI(A=1,B=2,C=3))\n',
                'chebyshev_reaction("H2 <=> 2 H", Tmin=300)\n',
                'reaction("1.5 H <=> H2", [1, 0, 0])\n',                            # fractional stoichiometry
                'reaction("H2 <=> 2 X", [1, 0, 0])\n',                              # undeclared species
                'reaction("H2 + M <=> 2 H + M", [1, 0, 0])\n'):                     # M in a plain reaction
        f = tmp_path / "bad.cti"
        f.write_text(base + bad)
        with pytest.raises(ctb.reactors.CtError):
            ctb.ThermoKinetics(str(f))


def test_mole_to_mass_and_flop_model(chem, omech):
    ch, om = chem("gri30"), omech("gri30")
    X = np.random.default_rng(0).random((5, om.K))
    assert np.abs(ch.mass_fractions(X) - om.X_to_Y(X)).max() < 1e-15
    fm = ch.flop_model(0)
    assert fm["lu"] == pytest.approx(2 / 3 * 54 ** 3) and fm["solve"] == 2 * 54 ** 2
    assert 2.0e4 < fm["rhs"] < 4.0e4  # SURVEY.md section 8d estimates 29.4 kflop


def test_c_abi_exports_every_declared_symbol(ctb):
    """include/chemtools_b200.h vs the built library: every declared entry point resolves (no compute calls)."""
    hdr = open(os.path.join(ROOT, "include", "chemtools_b200.h")).read()
    names = set(re.findall(r"\b(ct_[a-z0-9_]+)\s*\(", hdr))
    names -= {"ct_mech", "ct_batch"}
    assert len(names) > 45
    lib = C.CDLL(os.path.join(ROOT, "chemtools_b200", "libchemtools_b200.so"))
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    bound = set(ctb._lib.lib()._ct_signatures)
    assert names <= bound, sorted(names - bound)


def test_no_cpu_fallback_without_device(ctb, chem):
    """On a box without a GPU every compute entry point fails loudly with CT_ERR_NO_DEVICE."""
    if ctb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    ch = chem("gri30")
    with pytest.raises(ctb.reactors.CtError) as e:
        ctb.ReactorBatch(0, ch, [1000.0], [1e5], np.ones((1, ch.n_species)) / ch.n_species)
    assert e.value.code == -102
    with pytest.raises(ctb.reactors.CtError):
        ctb.measure_fp64_peak()


def test_cxx_binding_compiles_and_links(ctb, tmp_path):
    """include/chemtools_b200.hpp + tests/cxx/gpu_batch_test.cpp (the C++14 GpuBatch binding of INTEGRATION.md) build with
    g++ -std=c++14 -Wall -Wextra -Werror against the library; without a GPU the program fails loudly (no CPU path)."""
    import subprocess
    exe = str(tmp_path / "gpu_batch_test")
    libdir = os.path.join(ROOT, "chemtools_b200")
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "cxx", "gpu_batch_test.cpp"), "-L" + libdir, "-lchemtools_b200", "-Wl,-rpath," + libdir])
    if ctb.device_count() > 0:
        pytest.skip("a CUDA device is present: the program itself runs in the GPU tier (test_cxx_binding)")
    out = subprocess.run([exe, ctb.mechanism_path("gri30")], capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 and "no CUDA device" in out.stderr


@pytest.mark.parametrize("name,shape", [("gri30", 1), ("gri211", 2), ("gri12", 3)])
def test_static_shapes_cover_the_shipped_mechanisms(chem, name, shape):
    """mech_shapes.inc (generated at build time from the mechanism compiler) matches what the library compiles from the
    shipped mechanisms: the specialised kernels are selected for them."""
    assert chem(name).static_shape() == shape
