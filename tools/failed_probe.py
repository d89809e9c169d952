"""Re-run the reactors that failed at full size with different Jacobian / linear-solver options (development aid)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import chemtools_b200 as ctb
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
for name, mech, rt in (("config3_gri12", "gri12", 0), ("config3_gri211", "gri211", 0), ("config4", "gri30", 4)):
    d = np.load(os.path.join(root, "scratch", "fullsize_%s.npz" % name))
    chem = ctb.ThermoKinetics(ctb.mechanism_path(mech))
    bad = np.where(d["status"] < 0)[0][:12]
    T0, p0, Y = d["T0"][bad].copy(), d["p0"][bad].copy(), np.ascontiguousarray(d["Y"][bad])
    print(name, "idx", d["idx"][bad].tolist())
    for jac, linsol, clip in (("analytic", "inverse", 10.0), ("analytic", "lu", 10.0), ("dq", "lu", 10.0), ("dq", "inverse", 10.0), ("analytic", "inverse", 1.0), ("analytic", "inverse", 100.0), ("analytic", "lu", 1.0)):
        b = ctb.ReactorBatch(rt, chem, T0, p0, Y, 1e-8, 1e-14, 1e-3, max_num_steps=100000)
        if "tab_t" in d:
            b.SetTable(d["tab_t"], d["tab_T"][bad], d["tab_p"][bad])
        b.SetJacobian(jac); b.SetLinearSolver(linsol); b.SetJacobianClip(clip)
        code = b.Integrate(1e-3); st = b.Statistics()
        print("  %-8s %-7s clip %-5g code %3d status %s nst %s netf %s" % (jac, linsol, clip, code, b.Status().tolist(), st[:, 0].tolist(), st[:, 3].tolist()), flush=True)
