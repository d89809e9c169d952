"""Synthetic initial-condition sweeps of BASELINE.json's configs.

Pure host-side input generation (numpy); every generator returns
(mechanism name, reactor Type, T0[n], p0[n], X[n, K] mole fractions, t_end, extra).
"""
import enum

import numpy as np

try:
    from .reactors import Type
except ImportError:  # loaded as a plain file (bench.py --impl reference: numpy only, the CUDA library stays unmapped)
    class Type(enum.IntEnum):
        Const_Pres = 0
        Const_Vol = 1
        Const_Temp_Pres = 2
        Const_Temp_Vol = 3
        Table_Temp_Pres = 4
        Table_Vol = 5

ONE_ATM = 101325.0


def _fill(chem, n, spec):
    X = np.zeros((n, chem.n_species))
    for name, v in spec.items():
        X[:, chem.speciesIndex(name)] = v
    return X


def config1(chem):
    """1 ConstPres reactor, GRI-3.0, CH4/air phi=1, 1200 K, 1 atm (CPU reference run)."""
    return "gri30", Type.Const_Pres, np.array([1200.0]), np.array([ONE_ATM]), _fill(chem, 1, {"CH4": 1.0, "O2": 2.0, "N2": 7.52}), 1e-1, {}


def config2(chem, n_side=64, n_phi=None, index=None):
    """ConstVol H2/air ignition-delay sweep, GRI-3.0, n_side x n_phi reactors over T0=900-1500 K x phi=0.5-2.0
    (n_phi = n_side by default); `index` selects reactors of that global sweep (a shard)."""
    n_phi = n_side if n_phi is None else n_phi
    T0 = np.repeat(np.linspace(900.0, 1500.0, n_side), n_phi)
    phi = np.tile(np.linspace(0.5, 2.0, n_phi), n_side)
    if index is not None:
        T0, phi = T0[index], phi[index]
    n = len(T0)
    return "gri30", Type.Const_Vol, T0, np.full(n, ONE_ATM), _fill(chem, n, {"H2": 2.0 * phi, "O2": 1.0, "N2": 3.76}), 1e-3, {"phi": phi}


def config3(chem, mech="gri211", n=65536, seed=20211):
    """ConstPres CH4/air ensemble on GRI-2.11 / GRI-1.2 with randomised T0 / p / phi."""
    rng = np.random.default_rng(seed)
    T0 = rng.uniform(1000.0, 1800.0, n)
    p0 = np.exp(rng.uniform(np.log(0.5), np.log(20.0), n)) * ONE_ATM
    phi = rng.uniform(0.5, 2.0, n)
    return mech, Type.Const_Pres, T0, p0, _fill(chem, n, {"CH4": phi, "O2": 2.0, "N2": 7.52}), 1e-3, {"phi": phi}


def config4(chem, n=262144, nt=101, seed=30):
    """Prescribed T(t), p(t) reactors (linear ramps tabulated on nt points), GRI-3.0."""
    rng = np.random.default_rng(seed)
    phi = rng.uniform(0.5, 2.0, n)
    Ta = rng.uniform(900.0, 1300.0, n)
    Tb = rng.uniform(1500.0, 2200.0, n)
    pa = np.exp(rng.uniform(0.0, np.log(10.0), n)) * ONE_ATM
    t_end = 1e-3
    tt = np.linspace(0.0, t_end, nt)
    tabT = Ta[:, None] + (Tb - Ta)[:, None] * tt[None, :] / t_end
    tabp = pa[:, None] * (1.0 + 0.5 * tt[None, :] / t_end)
    return "gri30", Type.Table_Temp_Pres, Ta, pa, _fill(chem, n, {"CH4": phi, "O2": 2.0, "N2": 7.52}), t_end, {"table": (tt, tabT, tabp), "phi": phi}


def config5_index(chem, idx, n_side=256):
    """Reactors idx[] of the 256^3 ConstPres GRI-3.0 CH4/air ignition-delay map (T0 x p x phi, phi fastest)."""
    idx = np.asarray(idx, dtype=np.int64)
    iT, ip, iphi = idx // (n_side * n_side), (idx // n_side) % n_side, idx % n_side
    T0 = np.linspace(1000.0, 2000.0, n_side)[iT]
    p0 = np.geomspace(1.0, 40.0, n_side)[ip] * ONE_ATM
    phi = np.linspace(0.5, 2.0, n_side)[iphi]
    return "gri30", Type.Const_Pres, T0, p0, _fill(chem, len(idx), {"CH4": phi, "O2": 2.0, "N2": 7.52}), 1e-3, {"phi": phi}


def config5_design(n, n_side=256):
    """n global reactor indices spread over the whole 256^3 map: a 3-D Kronecker (additive-recurrence) sequence, so that
    every prefix and every strided subset covers all of T0, p and phi alike (a plain stride of 2^k would pin phi)."""
    g = 1.22074408460575947536  # real root of x^4 = x + 1: the 3-D analogue of the golden ratio
    al = np.array([1.0 / g, 1.0 / g ** 2, 1.0 / g ** 3])
    k = np.arange(1, n + 1, dtype=np.float64)[:, None]
    c = np.floor(((0.5 + k * al[None, :]) % 1.0) * n_side).astype(np.int64)
    return c[:, 0] * n_side * n_side + c[:, 1] * n_side + c[:, 2]


def config5_slice(chem, lo, hi, n_side=256):
    """Reactors [lo, hi) of the 256^3 ConstPres GRI-3.0 CH4/air ignition-delay map (T0 x p x phi)."""
    return config5_index(chem, np.arange(lo, hi), n_side)


def shard_range(n, rank, world):
    """Contiguous block partition of reactor indices: reactor i -> rank floor(i*world/n)."""
    lo = (n * rank + world - 1) // world
    hi = (n * (rank + 1) + world - 1) // world
    return lo, hi
