import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import chemtools_b200 as ctb
from chemtools_b200 import workloads as W
from oracle import oracle as O
chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
n = int(os.environ.get("N", "4096"))
stride = (256 ** 3) // n
idx = (np.arange(n) * stride) % (256 ** 3)
iT, ip, iphi = idx // 65536, (idx // 256) % 256, idx % 256
T0 = np.linspace(1000.0, 2000.0, 256)[iT]; p0 = np.geomspace(1.0, 40.0, 256)[ip] * 101325.0; phi = np.linspace(0.5, 2.0, 256)[iphi]
X = W._fill(chem, n, {"CH4": phi, "O2": 2.0, "N2": 7.52}); Y = chem.mass_fractions(X)
for clip in (1.0, 0.0):
    b = ctb.ReactorBatch(0, chem, T0, p0, Y, 1e-8, 1e-14, 1e-3, max_num_steps=int(os.environ.get("MX", "300000")))
    b.SetJacobianClip(clip)
    code = b.Integrate(1e-3); st = b.Statistics(); status = b.Status()
    nst = st[:, 0]
    print("clip", clip, "code", code, "kernel_ms", round(b.LastKernelMs(), 1), "reactors/s", round(n / b.LastKernelMs() * 1e3), "failed", int((status < 0).sum()),
          "nst mean", nst.mean(), "median", np.median(nst), "p90", np.percentile(nst, 90), "p99", np.percentile(nst, 99), "max", nst.max(), "sum", nst.sum(), flush=True)
    worst = np.argsort(-nst)[:6]
    print(" worst:", [(int(i), round(T0[i]), round(p0[i] / 101325, 1), round(phi[i], 2), int(nst[i]), int(st[i, 3]), int(st[i, 5])) for i in worst], flush=True)
om = O.Mechanism("gri30")
w = worst[:4]
t0 = time.time()
r = om.integrate_batch(0, T0[w], p0[w], Y[w], 1e-3, rtol=1e-8, atol=1e-14, max_steps=2000000, nthreads=4)
print("oracle on the worst 4:", r["stats"][:, 0].tolist(), r["status"].tolist(), round(time.time() - t0, 1), "s")
sample = np.linspace(0, n - 1, 64).astype(int)
t0 = time.time()
r = om.integrate_batch(0, T0[sample], p0[sample], Y[sample], 1e-3, rtol=1e-8, atol=1e-14, max_steps=2000000, nthreads=16)
print("oracle sample of 64: nst mean", r["stats"][:, 0].mean(), "max", r["stats"][:, 0].max(), round(time.time() - t0, 1), "s; gpu nst on same:", nst[sample].mean(), nst[sample].max())
