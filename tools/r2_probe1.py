"""Round-2 probe 1 (GPU): occupancy scan of the stand-alone right-hand side K1 (warps per SM vs throughput) and the
config-5 slice baseline (time, failures, counters).  usage: python tools/r2_probe1.py > gpurun_out/r2_probe1.json"""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import chemtools_b200 as ctb
from chemtools_b200 import workloads as W

out = {}
chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
K = chem.n_species
n = 65536
mech, rt, T0, p0, X, t_end, _ = W.config2(chem, 256)
Y = chem.mass_fractions(X)
st = np.concatenate([np.linspace(1200.0, 2400.0, n)[:, None], Y * 0.9 + 0.1 / K], axis=1)
b = ctb.ReactorBatch(rt, chem, T0, p0, Y, 1e-8, 1e-14, t_end)
fm = chem.flop_model(rt)
out["fp64_peak"] = ctb.measure_fp64_peak()
rows = [{"cfg": "classic k_rhs (6 warps/SM, slab with matrix)", "ms": b.time_rhs(st, reps=20)}]
for wpb, bps in ((6, 1), (6, 2), (8, 1), (8, 2), (12, 1), (13, 1), (16, 1), (24, 1), (6, 3), (4, 4), (4, 6), (8, 3)):
    try:
        ms = b.time_rhs_cfg(st, wpb, bps, reps=20)
        rows.append({"cfg": "%d warps x %d CTAs/SM" % (wpb, bps), "warps_per_sm": wpb * bps, "ms": ms})
    except Exception as e:
        rows.append({"cfg": "%d warps x %d CTAs/SM" % (wpb, bps), "error": str(e)})
for r in rows:
    if "ms" in r:
        r["rhs_per_s"] = n / (r["ms"] * 1e-3)
        r["tflops"] = fm["rhs"] * n / (r["ms"] * 1e-3) / 1e12
        r["frac"] = r["tflops"] / out["fp64_peak"]
out["k1_scan"] = rows
print(json.dumps(out), flush=True)

# config-5 slice: 16384 reactors strided over the map
nb = 16384
stride = (256 ** 3) // nb
idx = np.arange(nb) * stride
iT, ip, iphi = idx // 65536, (idx // 256) % 256, idx % 256
T0 = np.linspace(1000.0, 2000.0, 256)[iT]; p0 = np.geomspace(1.0, 40.0, 256)[ip] * W.ONE_ATM; phi = np.linspace(0.5, 2.0, 256)[iphi]
Y = chem.mass_fractions(W._fill(chem, nb, {"CH4": phi, "O2": 2.0, "N2": 7.52}))
res = {}
for tag, rtol, atol in (("head", 1e-8, 1e-14), ("ref", 1e-9, 1e-15)):
    bb = ctb.ReactorBatch(ctb.Type.Const_Pres, chem, T0, p0, Y, rtol, atol, 1e-3, max_num_steps=50000)
    bb.SetOrder(np.argsort(-T0, kind="stable").astype(np.int32))
    bb.Integrate(1e-3)
    t0 = time.time(); bb.Integrate  # noqa
    bb2 = ctb.ReactorBatch(ctb.Type.Const_Pres, chem, T0, p0, Y, rtol, atol, 1e-3, max_num_steps=50000)
    bb2.SetOrder(np.argsort(-T0, kind="stable").astype(np.int32))
    bb2.Integrate(1e-3)
    s = bb2.Statistics(); stt = bb2.Status()
    res[tag] = {"ms": bb2.LastKernelMs(), "reactors_per_s": nb / (bb2.LastKernelMs() * 1e-3), "failed": int((stt < 0).sum()),
                "mean": dict(zip(ctb.reactors.STAT_NAMES, s.mean(0).tolist())), "nst_pct": np.percentile(s[:, 0], [50, 90, 99, 99.9, 100]).tolist()}
out["config5_slice"] = res
print(json.dumps(out))
