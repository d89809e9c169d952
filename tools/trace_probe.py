"""Step trace of one reactor that failed / crawled at full size (development aid)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import chemtools_b200 as ctb
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
name, mech, which = sys.argv[1], sys.argv[2], int(sys.argv[3])
mx = int(sys.argv[4]) if len(sys.argv) > 4 else 3000
d = np.load(os.path.join(root, "scratch", "fullsize_%s.npz" % name))
k = np.where(d["idx"] == which)[0][0]
chem = ctb.ThermoKinetics(ctb.mechanism_path(mech))
rt = 4 if "tab_t" in d else 0
b = ctb.ReactorBatch(rt, chem, d["T0"][k:k+1], d["p0"][k:k+1], d["Y"][k:k+1], 1e-8, 1e-14, 1e-3, max_num_steps=mx)
if rt == 4:
    b.SetTable(d["tab_t"], d["tab_T"][k:k+1], d["tab_p"][k:k+1])
for a in sys.argv[5:]:
    kk, v = a.split("=")
    if kk == "clip": b.SetJacobianClip(float(v))
    if kk == "linsol": b.SetLinearSolver(v)
    if kk == "jac": b.SetJacobian(v)
b.SetTrace(0, 200000)
code = b.Integrate(1e-3)
tr = b.Trace(); st = b.Statistics()[0]
print("code", code, "stats", st.tolist(), "records", len(tr))
names = [chem.speciesName(i) for i in range(chem.n_species)]
def comp(i):
    i = int(i)
    return "T" if (rt <= 1 and i == 0) else names[i - (1 if rt <= 1 else 0)]
np.save(os.path.join(root, "gpurun_out", "trace_%s_%d.npy" % (name, which)), tr)
show = list(range(min(len(tr), 12))) + list(range(max(12, len(tr) - 60), len(tr)))
for i in show:
    r = tr[i]
    print("%6d t=%.9e h=%.3e q=%d kflag=%4d eflag=%3d dsm=%.3e acnrm=%.3e worst=%-6s acor=%+.3e zpred=%+.3e ncf/nef=%d" % (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], comp(r[8]), r[9], r[10], r[11]))
# which components dominate failures
fail = tr[(tr[:, 4] != 0) | (tr[:, 5] != 0)]
if len(fail):
    u, c = np.unique(fail[:, 8].astype(int), return_counts=True)
    print("failed attempts:", len(fail), "of", len(tr), "dominant components:", [(comp(a), int(n)) for a, n in sorted(zip(u, c), key=lambda x: -x[1])[:8]])
