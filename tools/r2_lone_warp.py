"""Latency of one reactor alone on the GPU (one warp, nothing to hide latencies behind): us per BDF step, the speed of
every tail.  python tools/r2_lone_warp.py [n] [workload: h2 | ch4]"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import chemtools_b200 as ctb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
wk = sys.argv[2] if len(sys.argv) > 2 else "h2"
chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
X = np.zeros((n, chem.n_species))
if wk == "h2":
    X[:, chem.speciesIndex("H2")] = 2.0; X[:, chem.speciesIndex("O2")] = 1.0; X[:, chem.speciesIndex("N2")] = 3.76
    rt, T0, p0 = ctb.Type.Const_Vol, np.full(n, 1100.0), np.full(n, 101325.0)
else:
    X[:, chem.speciesIndex("CH4")] = 1.0; X[:, chem.speciesIndex("O2")] = 2.0; X[:, chem.speciesIndex("N2")] = 7.52
    rt, T0, p0 = ctb.Type.Const_Pres, np.full(n, 1500.0), np.full(n, 10 * 101325.0)
Y = chem.mass_fractions(X)
b = ctb.ReactorBatch(rt, chem, T0, p0, Y, 1e-8, 1e-14, 1e-3, max_num_steps=50000)
for _ in range(3):
    b2 = ctb.ReactorBatch(rt, chem, T0, p0, Y, 1e-8, 1e-14, 1e-3, max_num_steps=50000)
    b2.Integrate(1e-3)
    st = b2.Statistics()
print(json.dumps(dict(n=n, workload=wk, kernel_ms=b2.LastKernelMs(), nst=int(st[0, 0]), nfe=int(st[0, 1]), nsetups=int(st[0, 2]), nje=int(st[0, 6]),
                      us_per_step=1e3 * b2.LastKernelMs() / st[0, 0])))
