import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import chemtools_b200 as ctb
from chemtools_b200 import workloads as W
from oracle import oracle as O
chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
n = 4096
stride = (256 ** 3) // n
idx = (np.arange(n) * stride) % (256 ** 3)
iT, ip, iphi = idx // 65536, (idx // 256) % 256, idx % 256
T0 = np.linspace(1000.0, 2000.0, 256)[iT]; p0 = np.geomspace(1.0, 40.0, 256)[ip] * 101325.0; phi = np.linspace(0.5, 2.0, 256)[iphi]
X = W._fill(chem, n, {"CH4": phi, "O2": 2.0, "N2": 7.52}); Y = chem.mass_fractions(X)
w = np.array([3864, 4053, 4048, 4063, 3849, 4094, 4079, 4046, 4089, 4038])
T0, p0, Y = T0[w].copy(), p0[w].copy(), np.ascontiguousarray(Y[w])
om = O.Mechanism("gri30")
for rtol, atol in ((1e-8, 1e-14), (1e-9, 1e-15)):
    r = om.integrate_batch(0, T0, p0, Y, 1e-3, rtol=rtol, atol=atol, max_steps=2000000, nthreads=10)
    print("tol", rtol, "oracle nst", r["stats"][:, 0].tolist(), flush=True)
    for jac, linsol, clip in (("dq", "lu", 1.0), ("analytic", "lu", 1.0), ("analytic", "inverse", 1.0), ("analytic", "inverse", 0.0), ("analytic", "lu", 0.0), ("analytic", "inverse", 1e30)):
        b = ctb.ReactorBatch(0, chem, T0, p0, Y, rtol, atol, 1e-3, max_num_steps=400000)
        b.SetJacobian(jac); b.SetLinearSolver(linsol); b.SetJacobianClip(clip)
        code = b.Integrate(1e-3); st = b.Statistics()
        print(" gpu", jac, linsol, clip, "code", code, "nst", st[:, 0].tolist(), "netf", st[:, 3].tolist(), "minY", np.round(np.log10(np.maximum(-b.MassFractions().min(1), 1e-30)), 1).tolist(), flush=True)
