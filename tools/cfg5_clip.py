"""Scan of the Jacobian clip threshold on a strided config-5 sample (the measurement behind DESIGN.md section 3, item 5)."""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import chemtools_b200 as ctb
from chemtools_b200 import workloads as W
chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
n = 4096
stride = (256 ** 3) // n
idx = (np.arange(n) * stride) % (256 ** 3)
iT, ip, iphi = idx // 65536, (idx // 256) % 256, idx % 256
T0 = np.linspace(1000.0, 2000.0, 256)[iT]; p0 = np.geomspace(1.0, 40.0, 256)[ip] * 101325.0; phi = np.linspace(0.5, 2.0, 256)[iphi]
X = W._fill(chem, n, {"CH4": phi, "O2": 2.0, "N2": 7.52}); Y = chem.mass_fractions(X)
for rtol, atol in ((1e-8, 1e-14), (1e-9, 1e-15)):
    for clip in (1.0, 3.0, 10.0, 30.0, 100.0, 1000.0):
        b = ctb.ReactorBatch(0, chem, T0, p0, Y, rtol, atol, 1e-3, max_num_steps=200000)
        b.SetJacobianClip(clip)
        code = b.Integrate(1e-3); st = b.Statistics(); nst = st[:, 0]
        print("tol", rtol, "clip", clip, "code", code, "ms", round(b.LastKernelMs()), "nst mean", round(nst.mean()), "p99", round(np.percentile(nst, 99)), "max", nst.max(), "n>20000:", int((nst > 20000).sum()), "failed", int((b.Status() < 0).sum()), flush=True)
