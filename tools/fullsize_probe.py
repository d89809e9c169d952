"""Full-size config 3 / 4 runs: status histogram, step-count distribution, initial conditions of failed reactors."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import chemtools_b200 as ctb
from chemtools_b200 import workloads as W
name = sys.argv[1]
mxstep = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if name.startswith("config3"):
    m = name.split("_")[1]
    chem = ctb.ThermoKinetics(ctb.mechanism_path(m))
    mech, rt, T0, p0, X, t_end, extra = W.config3(chem, m, n or 65536, 20211)
else:
    chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
    mech, rt, T0, p0, X, t_end, extra = W.config4(chem, n or 262144, 101, 30)
Y = chem.mass_fractions(X)
b = ctb.ReactorBatch(rt, chem, T0, p0, Y, 1e-8, 1e-14, t_end, max_num_steps=mxstep)
if "table" in extra:
    b.SetTable(*extra["table"])
for a in sys.argv[4:]:
    kk, v = a.split("=")
    if kk == "clip": b.SetJacobianClip(float(v))
    if kk == "linsol": b.SetLinearSolver(v)
    if kk == "jac": b.SetJacobian(v)
code = b.Integrate(t_end)
st = b.Statistics(); status = b.Status()
u, c = np.unique(status, return_counts=True)
print(name, sys.argv[4:], "code", code, "kernel ms", round(b.LastKernelMs(), 1), "status histogram", dict(zip(u.tolist(), c.tolist())))
nst = st[:, 0]
print("nst percentiles 50/90/99/99.9/max:", np.percentile(nst, [50, 90, 99, 99.9, 100]).astype(int).tolist(), "netf max", int(st[:, 3].max()), "ncfn max", int(st[:, 5].max()), "mean nst/nje/nsetups", np.round(st[:, [0, 6, 2]].mean(0), 1).tolist())
bad = np.where(status < 0)[0]
worst = np.argsort(-nst)[:16]
sel = np.unique(np.concatenate([bad[:48], worst]))
out = dict(idx=sel, T0=T0[sel], p0=p0[sel], Y=Y[sel], status=status[sel], stats=st[sel], time=b.Time()[sel])
if "table" in extra:
    tt, tT, tp = extra["table"]
    out.update(tab_t=tt, tab_T=tT[sel], tab_p=tp[sel])
np.savez(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "fullsize_%s.npz" % name), **out)
print("failed:", bad[:20].tolist(), "T0", np.round(T0[bad[:10]], 1).tolist(), "p/atm", np.round(p0[bad[:10]] / 101325, 2).tolist(), "t reached", b.Time()[bad[:10]].tolist())
