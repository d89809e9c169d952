"""A/B probe of scheduling options on config 2 (development aid)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import chemtools_b200 as ctb
from chemtools_b200 import workloads as W
chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
mech, rt, T0, p0, X, t_end, extra = W.config2(chem, int(os.environ.get("SIDE", "64")))
Y = chem.mass_fractions(X)
cases = json.loads(os.environ["CASES"]) if os.environ.get("CASES") else [dict(w=12), dict(period=2), dict(period=3), dict(period=4), dict(w=13), dict(w=14)]
last = None
for rep in range(2):
    for c in cases:
        b = ctb.ReactorBatch(rt, chem, T0, p0, Y, 1e-8, 1e-14, t_end, max_num_steps=50000)
        b.SetPhaseAlignment(c.get("gang", 1)); b.SetMatrixPool(1, c.get("w", 0)); b.SetAlignmentPollNs(c.get("ns", 200)); b.SetAlignmentMaxPolls(c.get("polls", 0)); b.SetAlignmentPeriod(c.get("period", 1))
        if "slice" in c:
            b.SetTimeSlicing(c["slice"])
        if c.get("order") and last is not None:
            b.SetOrder(np.argsort(-last[:, 1], kind="stable").astype(np.int32))
        code = b.Integrate(t_end)
        st = b.Statistics(); last = st
        print(json.dumps(dict(rep=rep, **c, code=int(code), wpb=b.WarpsPerSM(), kernel_ms=round(b.LastKernelMs(), 2), reactors_per_s=round(len(T0) / b.LastKernelMs() * 1e3))), flush=True)
