"""A/B probe of scheduling / linear-algebra options on config 2 (development aid)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import chemtools_b200 as ctb
from chemtools_b200 import workloads as W
chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
mech, rt, T0, p0, X, t_end, extra = W.config2(chem, int(os.environ.get("SIDE", "64")))
Y = chem.mass_fractions(X)
for rep in range(2):
    for gang in (1, 0):
        for linsol in ("lu", "inverse"):
            b = ctb.ReactorBatch(rt, chem, T0, p0, Y, 1e-8, 1e-14, t_end, max_num_steps=50000)
            b.SetLinearSolver(linsol); b.SetPhaseAlignment(gang)
            code = b.Integrate(t_end)
            st = b.Statistics()
            print(json.dumps(dict(rep=rep, gang=gang, linsol=linsol, code=int(code), kernel_ms=round(b.LastKernelMs(), 2), reactors_per_s=round(len(T0) / b.LastKernelMs() * 1e3),
                                  nst_mean=float(st[:, 0].mean()), nfe_mean=float(st[:, 1].mean()))), flush=True)
