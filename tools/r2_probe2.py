"""Round-2 probe 2 (GPU): A/B of the fused kernel and of K1 under kernel-selection / scheduling options.
usage: python tools/r2_probe2.py [tag] > gpurun_out/r2_probe2_<tag>.json"""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import chemtools_b200 as ctb
from chemtools_b200 import workloads as W

out = {"tag": sys.argv[1] if len(sys.argv) > 1 else ""}
chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
K = chem.n_species
out["static_shape"] = chem.static_shape()
# ---- K1
n = 65536
mech, rt, T0, p0, X, t_end, _ = W.config2(chem, 256)
Y = chem.mass_fractions(X)
st = np.concatenate([np.linspace(1200.0, 2400.0, n)[:, None], Y * 0.9 + 0.1 / K], axis=1)
b = ctb.ReactorBatch(rt, chem, T0, p0, Y, 1e-8, 1e-14, t_end)
fm = chem.flop_model(rt)
out["fp64_peak"] = ctb.measure_fp64_peak()
rows = []
for static in (1, 0):
    b.SetStaticShape(static)
    for wpb, bps in ((12, 1), (16, 1), (24, 1)):
        ms = b.time_rhs_cfg(st, wpb, bps, reps=20)
        rows.append({"static": static, "warps_per_sm": wpb * bps, "ms": ms, "rhs_per_s": n / (ms * 1e-3), "frac": fm["rhs"] * n / (ms * 1e-3) / 1e12 / out["fp64_peak"]})
out["k1"] = rows
print(json.dumps(out), flush=True)

def run(chem, rt, T0, p0, Y, rtol, atol, opts, reps=2, mxstep=50000):
    bb = ctb.ReactorBatch(rt, chem, T0, p0, Y, rtol, atol, 1e-3, max_num_steps=mxstep)
    bb.SetOrder(np.argsort(-T0, kind="stable").astype(np.int32))
    for k, v in opts.items():
        getattr(bb, k)(*v) if isinstance(v, tuple) else getattr(bb, k)(v)
    ms = []
    for _ in range(reps):
        lib = ctb.reactors.lib()
        ctb.reactors.check(lib.ct_batch_set_ic(bb._h, ctb.reactors._p(T0), ctb.reactors._p(p0), ctb.reactors._p(Y)))
        bb.Integrate(1e-3)
        ms.append(bb.LastKernelMs())
    s = bb.Statistics(); stt = bb.Status()
    return {"ms": min(ms), "reactors_per_s": len(T0) / (min(ms) * 1e-3), "failed": int((stt < 0).sum()), "warps": bb.WarpsPerSM(),
            "mean": dict(zip(ctb.reactors.STAT_NAMES, s.mean(0).tolist())), "max_nst": int(s[:, 0].max())}

mech, rt, T0, p0, X, t_end, _ = W.config2(chem, 64)
Y = chem.mass_fractions(X)
res = {}
for name, opts in (("default", {}), ("runtime_shape", {"SetStaticShape": 0}), ("no_alignment", {"SetPhaseAlignment": 0}),
                   ("no_slicing", {"SetTimeSlicing": 0}), ("lu", {"SetLinearSolver": "lu"})):
    try:
        res[name] = run(chem, rt, T0, p0, Y, 1e-8, 1e-14, opts)
    except Exception as e:
        res[name] = {"error": str(e)}
    print(name, json.dumps(res[name]), file=sys.stderr, flush=True)
out["config2"] = res
print(json.dumps(out), flush=True)
nb = 16384
stride = (256 ** 3) // nb
idx = np.arange(nb) * stride
iT, ip, iphi = idx // 65536, (idx // 256) % 256, idx % 256
T0 = np.linspace(1000.0, 2000.0, 256)[iT]; p0 = np.geomspace(1.0, 40.0, 256)[ip] * W.ONE_ATM; phi = np.linspace(0.5, 2.0, 256)[iphi]
Y = chem.mass_fractions(W._fill(chem, nb, {"CH4": phi, "O2": 2.0, "N2": 7.52}))
res = {}
for name, opts in (("default", {}), ("no_alignment", {"SetPhaseAlignment": 0})):
    res[name] = run(chem, ctb.Type.Const_Pres, T0, p0, Y, 1e-8, 1e-14, opts, reps=1)
    print("cfg5", name, json.dumps(res[name]), file=sys.stderr, flush=True)
out["config5_slice"] = res
print(json.dumps(out))
