"""chemtools_b200 — B200-native batched homogeneous-reactor integrator behind
ChemTools' reactor application API.

Only the reactor hot path lives here: mechanism compiler, fused RHS /
Jacobian / LU kernels and the device-side BDF stepper, all behind the C ABI of
include/chemtools_b200.h.  No CPU fallback exists.
"""
import os

from .reactors import (CV_SUCCESS, CV_TSTOP_RETURN, Create, ReactorBatch, ThermoKinetics, Type,  # noqa: F401
                       device_count, lu_solve, lu_time, measure_fp64_peak, roberts_selftest, set_device)

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def mechanism_path(name):
    """Path of a shipped compiled mechanism (gri30 / gri211 / gri12 .ctm), or `name` itself if it is a file."""
    if os.path.exists(name):
        return name
    p = os.path.join(_DATA, name + ".ctm")
    if not os.path.exists(p):
        raise FileNotFoundError("no mechanism '%s' (looked for %s)" % (name, p))
    return p
