"""Stand-alone kernels K1 (right-hand side), K2 (analytic / DQ Jacobian) and K3 (dense LU / Gauss-Jordan + solve):
time per launch from CUDA events inside the library, achieved fp64 TFLOP/s from the mechanism's flop model, and the
HBM traffic each launch needs (states in, results out) against the measured peaks.  Writes one JSON document.
usage: python tools/kernel_rooflines.py [n_reactors] > profiles/r01_kernel_rooflines.json"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import chemtools_b200 as ctb
from chemtools_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
peaks = {}
try:
    peaks = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm_peak = float(peaks.get("hbm_gbs", 6459.3))
fp64_peak = ctb.measure_fp64_peak()
chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
side = int(round(n ** 0.5))
mech, rt, T0, p0, X, t_end, _ = W.config2(chem, side)
n = len(T0)
Y = chem.mass_fractions(X)
K = chem.n_species; N = K + 1
# states half way through ignition make every reaction active: initial state with 1 % of each radical
st = np.concatenate([np.linspace(1200.0, 2400.0, n)[:, None], Y * 0.9 + 0.1 / K], axis=1)
b = ctb.ReactorBatch(rt, chem, T0, p0, Y, 1e-8, 1e-14, t_end)
fm = chem.flop_model(rt)
out = {"device": peaks.get("gpu_name", "B200"), "n_reactors": n, "mechanism": "gri30", "fp64_peak_tflops_measured": fp64_peak,
       "hbm_peak_gbs": hbm_peak, "flop_model": fm, "kernels": []}
def row(name, ms, flops_per_unit, bytes_per_unit, note):
    tf = flops_per_unit * n / (ms * 1e-3) / 1e12
    gbs = bytes_per_unit * n / (ms * 1e-3) / 1e9
    out["kernels"].append({"kernel": name, "ms_per_launch": ms, "reactors_per_s": n / (ms * 1e-3), "achieved_tflops": tf,
                           "frac_fp64_peak": tf / fp64_peak, "algorithmic_hbm_gbs": gbs, "frac_hbm_peak": gbs / hbm_peak, "note": note})
ms = b.time_rhs(st, reps=20)
row("K1 k_rhs", ms, fm["rhs"], (N + N + K) * 8, "state in, ydot + wdot out; mechanism blob staged once per CTA in shared memory")
ms = b.time_jacobian(st, "analytic", reps=5)
row("K2 k_jacobian (analytic)", ms, fm["jac"], (N + N * N) * 8, "state in, dense N x N Jacobian out")
ms = b.time_jacobian(st, "dq", reps=2)
row("K2 k_jacobian (difference quotients, the reference's way)", ms, (N + 1) * fm["rhs"], (N + N * N) * 8, "N + 1 right-hand sides per Jacobian")
for mode in ("lu", "inverse"):
    for solves in (1, 12):
        ms = ctb.lu_time(N, n, reps=5, solves_per_factor=solves, mode=mode)
        row("K3 k_lu_solve (%s, %d solves per factorisation)" % (mode, solves), ms, fm["lu"] * (3 if mode == "inverse" else 1) + solves * fm["solve"],
            (N * N + 2 * N * solves) * 8, "matrix in, solutions out; %s" % ("Gauss-Jordan inverse = 3 x the LU flops, mat-vec solves" if mode == "inverse" else "partial-pivot LU, triangular solves"))
print(json.dumps(out, indent=1))
