"""A/B of scheduling options on config 2 (results are unaffected by any of them): phase alignment on / off / period,
time-slice length.  python tools/ab_sched.py"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import chemtools_b200 as ctb
from chemtools_b200 import workloads as W
chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
mech, rt, T0, p0, X, t_end, _ = W.config2(chem, 64)
Y = chem.mass_fractions(X)
order = np.argsort(-T0, kind="stable").astype(np.int32)
out = []
for name, fn in (("default", lambda b: None), ("no alignment", lambda b: b.SetPhaseAlignment(0)), ("period 2", lambda b: b.SetAlignmentPeriod(2)),
                 ("slice 32", lambda b: b.SetTimeSlicing(32)), ("slice 128", lambda b: b.SetTimeSlicing(128)), ("default", lambda b: None)):
    b = ctb.ReactorBatch(rt, chem, T0, p0, Y, 1e-8, 1e-14, t_end, max_num_steps=20000)
    b.SetOrder(order); fn(b)
    b.Integrate(t_end)
    out.append(dict(variant=name, kernel_ms=b.LastKernelMs()))
    print(out[-1], flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "ab_sched.json"), "w"))
