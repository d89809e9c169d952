"""ORACLE (test infrastructure, not product code).

Independent Cantera-CTI front end for the CPU oracle.  A ``.cti`` file is
executable Python (reference: /root/reference/data/gri30.cti:7-48 `units`,
`ideal_gas`; :55-72 `species`/`NASA`; :1017-1055 `reaction`,
`three_body_reaction`, `falloff_reaction`, `Troe`), so this front end simply
``exec``s it against stub callables and records what was declared.  The
product's own mechanism compiler is a separate C++ recursive-descent parser
(chemtools_b200/csrc/cti_parser.cpp); tests compare the two.

The reference loads a mechanism through `Cantera::newPhase(file)` with no
phase id (/root/reference/include/chemistry.hpp:75), i.e. the FIRST
`ideal_gas` entry of the file; species order is the order of that entry's
species list, reactions='all' keeps file order.

Output of :func:`load_cti` is a plain dict in FILE units (mol-cm-s-cal), no
conversion: conversion to Cantera's kmol-m-s-K system is done in
oracle/oracle.py so that it can be checked separately.
"""
import json
import re
import sys


class _Rec:
    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)


def _parse_pairs(text):
    """'AR:0.83  C2H6:3' -> [('AR', 0.83), ('C2H6', 3.0)] (file order)."""
    out = []
    for tok in text.replace(",", " ").split():
        name, val = tok.rsplit(":", 1)
        out.append((name.strip(), float(val)))
    return out


_TERM = re.compile(r"^\s*(\d+(?:\.\d*)?)?\s*(\S+)\s*$")


def _parse_side(side):
    """'2 O + M' -> ([('O', 2)], has_M, has_falloff_M)"""
    falloff = False
    s = side
    m = re.search(r"\(\s*\+\s*M\s*\)", s)
    if m:
        falloff = True
        s = s[: m.start()] + s[m.end():]
    terms = []
    third = False
    # species names may contain '+'?  Not in GRI; split on ' + ' style pluses.
    for part in re.split(r"\s\+\s|\+(?=\s)|(?<=\s)\+", s):
        part = part.strip()
        if not part:
            continue
        if part == "M":
            third = True
            continue
        mm = _TERM.match(part)
        coeff = mm.group(1)
        name = mm.group(2)
        # '2O' style (no blank) is not used by the GRI files; integer prefix needs a blank
        terms.append((name, float(coeff) if coeff else 1.0))
    return terms, third, falloff


def parse_equation(eq):
    if "<=>" in eq:
        lhs, rhs = eq.split("<=>")
        rev = True
    elif "=>" in eq:
        lhs, rhs = eq.split("=>")
        rev = False
    elif "=" in eq:
        lhs, rhs = eq.split("=")
        rev = True
    else:
        raise ValueError("bad reaction equation: " + eq)
    r, m1, f1 = _parse_side(lhs)
    p, m2, f2 = _parse_side(rhs)
    if m1 != m2 or f1 != f2:
        raise ValueError("unbalanced third body in: " + eq)
    return r, p, rev, m1, f1


def load_cti(path):
    phases = []
    species = {}
    species_order = []
    reactions = []
    unit_rec = {}

    def units(**kw):
        unit_rec.update(kw)

    def state(**kw):
        return kw

    def ideal_gas(name=None, elements="", species="", reactions="all", **kw):
        phases.append(dict(name=name, elements=elements.split(), species=species.split(),
                           reactions=reactions))

    def NASA(trange, coeffs):
        if len(coeffs) != 7:
            raise ValueError("NASA-7 expected")
        return dict(T=[float(trange[0]), float(trange[1])], a=[float(c) for c in coeffs])

    def gas_transport(**kw):
        return kw

    def _species(name=None, atoms="", thermo=None, transport=None, note=None, **kw):
        name = str(name)
        if not isinstance(thermo, (tuple, list)) or len(thermo) != 2:
            raise ValueError("two NASA ranges expected for " + name)
        lo, hi = sorted(thermo, key=lambda r: r["T"][0])
        species[name] = dict(name=name, atoms=_parse_pairs(atoms), nasa_lo=lo, nasa_hi=hi)
        species_order.append(name)

    def Troe(A=None, T3=None, T1=None, T2=None):
        return dict(type="Troe", A=float(A), T3=float(T3), T1=float(T1),
                    T2=(None if T2 is None else float(T2)))

    def _rxn(kind, equation, kf, kf0=None, falloff=None, efficiencies="", options=None, **kw):
        if kw:
            raise ValueError("unsupported reaction keyword(s): %s" % sorted(kw))
        r, p, rev, third, fall = parse_equation(equation)
        if kind == "three_body" and not third:
            raise ValueError("three_body_reaction without M: " + equation)
        if kind == "falloff" and not fall:
            raise ValueError("falloff_reaction without (+ M): " + equation)
        if kind == "elementary" and (third or fall):
            raise ValueError("reaction() with third body: " + equation)
        reactions.append(dict(kind=kind, equation=equation, reactants=r, products=p,
                              reversible=rev, kf=[float(x) for x in kf],
                              kf0=None if kf0 is None else [float(x) for x in kf0],
                              falloff=falloff, efficiencies=_parse_pairs(efficiencies),
                              duplicate=(options is not None and "duplicate" in options)))

    def reaction(equation, kf, **kw):
        _rxn("elementary", equation, kf, **kw)

    def three_body_reaction(equation, kf, **kw):
        _rxn("three_body", equation, kf, **kw)

    def falloff_reaction(equation, kf=None, kf0=None, **kw):
        _rxn("falloff", equation, kf, kf0=kf0, **kw)

    env = dict(units=units, state=state, ideal_gas=ideal_gas, NASA=NASA,
               gas_transport=gas_transport, species=_species, Troe=Troe,
               reaction=reaction, three_body_reaction=three_body_reaction,
               falloff_reaction=falloff_reaction, OneAtm=101325.0)
    with open(path) as f:
        src = f.read()
    exec(compile(src, path, "exec"), env)

    ph = phases[0]
    if ph["reactions"] != "all":
        raise ValueError("only reactions='all' supported")
    sp = [species[n] for n in ph["species"]]
    return dict(phase=ph["name"], units=unit_rec, elements=ph["elements"], species=sp,
                reactions=reactions, n_phases=len(phases))


if __name__ == "__main__":
    # usage: python oracle/cti_frontend.py in.cti out.json
    mech = load_cti(sys.argv[1])
    with open(sys.argv[2], "w") as f:
        json.dump(mech, f, indent=0, separators=(",", ":"))
    print(len(mech["species"]), "species", len(mech["reactions"]), "reactions")
