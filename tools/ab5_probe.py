"""A/B of warps per SM (and so pool slabs) on a config-5 slice (development aid)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import chemtools_b200 as ctb
from chemtools_b200 import workloads as W
chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
n = int(os.environ.get("N", "16384"))
stride = (256 ** 3) // n
idx = (np.arange(n) * stride) % (256 ** 3)
iT, ip, iphi = idx // 65536, (idx // 256) % 256, idx % 256
T0 = np.linspace(1000.0, 2000.0, 256)[iT]; p0 = np.geomspace(1.0, 40.0, 256)[ip] * W.ONE_ATM; phi = np.linspace(0.5, 2.0, 256)[iphi]
Y = chem.mass_fractions(W._fill(chem, n, {"CH4": phi, "O2": 2.0, "N2": 7.52}))
for c in json.loads(os.environ.get("CASES", '[{"w":12},{"w":10},{"w":11}]')):
    b = ctb.ReactorBatch(ctb.Type.Const_Pres, chem, T0, p0, Y, 1e-8, 1e-14, 1e-3, max_num_steps=50000)
    b.SetMatrixPool(1, c.get("w", 0))
    if "slice" in c: b.SetTimeSlicing(c["slice"])
    if "clip" in c: b.SetJacobianClip(c["clip"])
    if "ns" in c: b.SetAlignmentPollNs(c["ns"])
    code = b.Integrate(1e-3)
    st = b.Statistics()
    print(json.dumps(dict(**c, nst=int(st[:, 0].mean()), nsetups=int(st[:, 2].mean()), maxnst=int(st[:, 0].max()), code=int(code), wpb=b.WarpsPerSM(), kernel_ms=round(b.LastKernelMs(), 1), reactors_per_s=round(n / b.LastKernelMs() * 1e3))), flush=True)
