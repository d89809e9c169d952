"""Ad-hoc GPU probe: kernel time of the fused integrator under different batch sizes, work-queue orders,
linear-solver modes and Jacobian clip thresholds (development aid; results go to profiles/)."""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import chemtools_b200 as ctb
from chemtools_b200 import workloads as W

chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
out = []
def run(side, order=None, linsol="inverse", clip=1.0, jac="analytic", label=""):
    mech, rt, T0, p0, X, t_end, extra = W.config2(chem, side)
    Y = chem.mass_fractions(X)
    b = ctb.ReactorBatch(rt, chem, T0, p0, Y, 1e-8, 1e-14, t_end, max_num_steps=50000)
    b.SetLinearSolver(linsol); b.SetJacobianClip(clip); b.SetJacobian(jac)
    if order is not None:
        b.SetOrder(order)
    code = b.Integrate(t_end)
    ms = b.LastKernelMs()
    st = b.Statistics()
    n = len(T0)
    slots = 148 * b.WarpsPerSM()
    rec = dict(label=label, n=n, linsol=linsol, clip=clip, jac=jac, code=int(code), kernel_ms=ms, reactors_per_s=n / ms * 1e3,
               nst_mean=float(st[:, 0].mean()), nst_max=int(st[:, 0].max()), nfe_mean=float(st[:, 1].mean()), nsetups_mean=float(st[:, 2].mean()),
               ncfn_sum=int(st[:, 5].sum()), nje_mean=float(st[:, 6].mean()),
               ideal_ms_from_nfe=None)
    # time per RHS-equivalent if perfectly balanced
    rec["ms_per_kilo_nfe_per_slot"] = ms / (st[:, 1].sum() / slots) * 1e3
    print(json.dumps(rec), flush=True)
    out.append(rec)
    return st

for side in (64, 128):
    n = side * side
    st = run(side, None, label="natural order")
    T0 = np.repeat(np.linspace(900.0, 1500.0, side), side)
    run(side, np.argsort(-T0, kind="stable").astype(np.int32), label="hot first")
    run(side, np.argsort(-st[:, 1], kind="stable").astype(np.int32), label="longest first (from previous run's nfe)")
    run(side, np.random.default_rng(0).permutation(n).astype(np.int32), label="random order")
run(64, None, linsol="lu", label="LU + triangular solves")
run(64, None, jac="dq", label="DQ Jacobian")
run(64, None, clip=0.0, label="clip 0")
run(64, None, clip=1e30, label="clip off")
os.makedirs(os.path.join(os.path.dirname(__file__), "..", "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "perf_probe.json"), "w"), indent=1)
