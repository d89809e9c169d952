"""Per-reactor schedule of one config-2 launch (GPU global timer): busy fraction, tail, LPT bound."""
import sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import chemtools_b200 as ct
from chemtools_b200 import workloads
side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
order_mode = sys.argv[2] if len(sys.argv) > 2 else "hot"
slice_steps = int(sys.argv[3]) if len(sys.argv) > 3 else -1
chem = ct.ThermoKinetics(ct.mechanism_path("gri30"))
_, rt, T0, p0, X, t_end, _ = workloads.config2(chem, side)
Y = chem.mass_fractions(X)
n = len(T0)
w = {"T": T0}
for rep in range(2):
    b = ct.ReactorBatch(rt, chem, T0, p0, Y, 1e-8, 1e-14, t_end, max_num_steps=200000)
    if order_mode == "hot":
        b.SetOrder(np.argsort(-T0, kind="stable").astype(np.int32))
    elif order_mode == "lpt" and rep == 1:
        b.SetOrder(np.argsort(-prev, kind="stable").astype(np.int32))  # oracle order: longest measured first
    b.SetTimeline(True)
    if slice_steps >= 0:
        b.SetTimeSlicing(slice_steps)
    b.Integrate(t_end)
    prev = (b.Timeline()[:, 1] - b.Timeline()[:, 0]).astype(np.float64)
tl = b.Timeline(); st = b.Statistics()
t0 = tl[:, 0].min(); s = (tl[:, 0] - t0) * 1e-6; e = (tl[:, 1] - t0) * 1e-6; d = e - s
slots = len(np.unique(tl[:, 2])); T = e.max()
print("kernel ms", b.LastKernelMs())
print("n=%d slots=%d makespan=%.1f ms  sum(d)=%.1f s  busy=%.3f" % (n, slots, T, d.sum() * 1e-3, d.sum() / (slots * T)))
print("duration ms: min %.1f med %.1f mean %.1f p90 %.1f max %.1f" % (d.min(), np.median(d), d.mean(), np.percentile(d, 90), d.max()))
print("last start %.1f ms; queue empty at %.1f ms; durations of the last 10 to finish:" % (s.max(), s.max()), np.round(d[np.argsort(e)[-10:]], 1))
# active warps over time
ts = np.linspace(0, T, 41)
act = [(int(((s <= t) & (e > t)).sum())) for t in ts]
print("active reactors over time:", act)
nst = st[:, 0]
print("corr(duration, nst)=%.3f  corr(duration, T0)=%.3f" % (np.corrcoef(d, nst)[0, 1], np.corrcoef(d, w["T"])[0, 1]))
# speed (steps per ms) early vs late
early = s < 0.2 * T; late = s > 0.6 * T
print("steps/ms of reactors started early %.2f, late %.2f" % ((nst[early] / d[early]).mean(), (nst[late] / d[late]).mean() if late.any() else float("nan")))
np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "timeline_%d.npy" % n), np.column_stack([s, e, tl[:, 2], nst, w["T"]]))
