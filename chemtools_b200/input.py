"""Input::Parser mirror (reference: include/input/parser.hpp:9-28, include/input/types.hpp:245-346,
src/input_parser.cpp) over the C ABI.  Exceptions map one-to-one onto the classes the reference's tests assert on."""
import ctypes as C

from ._lib import CtError, dp, lib


class BadFile(CtError):            # YAML::BadFile
    pass


class InvalidNode(CtError):        # YAML::InvalidNode
    pass


class BadConversion(CtError):      # YAML::BadConversion
    pass


class FilesystemError(CtError):    # boost::filesystem::filesystem_error
    pass


class InputRuntimeError(CtError):  # std::runtime_error
    pass


class InvalidUnit(CtError):        # std::invalid_argument
    pass


_ERR = {-110: BadFile, -111: InvalidNode, -112: BadConversion, -113: FilesystemError, -114: InputRuntimeError, -115: InvalidUnit}
COMPOSITION_TYPES = {1: "Concentration", 2: "MassFraction", 3: "MoleFraction"}


class Case:
    def __init__(self, description, pressure, temperature, time, composition_type, pairs, _addr):
        self.description, self.pressure, self.temperature, self.time = description, pressure, temperature, time
        self.composition_type, self.pairs, self._addr = composition_type, pairs, _addr

    def DescriptionIsDefined(self):
        return self.description is not None


class CaseSet:
    def __init__(self, mechanism_path, mechanism_description, description, cases):
        self.mechanism_path, self.mechanism_description, self.description, self.cases = mechanism_path, mechanism_description, description, cases

    def Cases(self):
        return self.cases


class Parser:
    """Input::Parser(config_file): parses and validates; GetParameters() gives the four case-set lists."""

    def __init__(self, config_file):
        L = lib()
        st = C.c_int(0)
        self._h = L.ct_input_parse(str(config_file).encode(), C.byref(st))
        if not self._h:
            raise _ERR.get(st.value, CtError)(st.value, L.ct_last_error().decode())
        dec = lambda b: None if b is None else b.decode()
        self.const_P, self.const_V, self.const_TP, self.const_TV = ([] for _ in range(4))
        for t, dst in enumerate((self.const_P, self.const_V, self.const_TP, self.const_TV)):
            for s in range(L.ct_input_n_case_sets(self._h, t)):
                cases = []
                for c in range(L.ct_input_n_cases(self._h, t, s)):
                    p, T, tm = C.c_double(), C.c_double(), C.c_double()
                    ct, npairs = C.c_int(), C.c_int()
                    L.ct_input_case(self._h, t, s, c, C.cast(C.byref(p), dp), C.cast(C.byref(T), dp), C.cast(C.byref(tm), dp), C.byref(ct), C.byref(npairs))
                    pairs = []
                    for i in range(npairs.value):
                        v = C.c_double()
                        pairs.append((L.ct_input_case_species(self._h, t, s, c, i, C.cast(C.byref(v), dp)).decode(), v.value))
                    cases.append(Case(dec(L.ct_input_case_description(self._h, t, s, c)), p.value, T.value, tm.value,
                                      COMPOSITION_TYPES[ct.value], pairs, (t, s, c)))
                dst.append(CaseSet(dec(L.ct_input_mechanism_path(self._h, t, s)), dec(L.ct_input_mechanism_description(self._h, t, s)),
                                   dec(L.ct_input_set_description(self._h, t, s)), cases))
        self.warnings = [L.ct_input_warning(self._h, i).decode() for i in range(L.ct_input_n_warnings(self._h))]

    def GetParameters(self):
        return self

    def mass_fractions(self, case, chemistry):
        import numpy as np
        Y = np.zeros(chemistry.n_species)
        t, s, c = case._addr
        rc = lib().ct_input_case_mass_fractions(self._h, t, s, c, chemistry._h, Y.ctypes.data_as(dp))
        if rc < 0:
            raise _ERR.get(rc, CtError)(rc, lib().ct_last_error().decode())
        return Y

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ct_input_destroy(self._h)
                self._h = None
        except Exception:
            pass
