"""A/B of the waiting scheme (hardware mbarrier wait vs __nanosleep polling) on config 2 and a config-5 sample.
python tools/ab_wait.py"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import chemtools_b200 as ctb
from chemtools_b200 import workloads as W
chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
out = {}
for name in ("config2", "config5"):
    if name == "config2":
        mech, rt, T0, p0, X, t_end, _ = W.config2(chem, 64)
    else:
        mech, rt, T0, p0, X, t_end, _ = W.config5_index(chem, W.config5_design(16384))
    Y = chem.mass_fractions(X)
    order = np.argsort(-T0, kind="stable").astype(np.int32)
    ref = None
    for ns in (0, 200, 0, 200):
        b = ctb.ReactorBatch(rt, chem, T0, p0, Y, 1e-8, 1e-14, t_end, max_num_steps=20000)
        b.SetOrder(order); b.SetAlignmentPollNs(ns)
        b.Integrate(t_end)
        st = b.Statistics()
        if ref is None:
            ref = (b.Solution(), st)
        same = bool(np.array_equal(ref[0], b.Solution()) and np.array_equal(ref[1], st))
        out.setdefault(name, []).append(dict(poll_ns=ns, kernel_ms=b.LastKernelMs(), identical=same, failed=int((b.Status() < 0).sum())))
        print(name, out[name][-1], flush=True)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "ab_wait.json"), "w"))
