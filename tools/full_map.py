"""BASELINE config 5 at full size: the 256^3 = 16 777 216 constant-pressure GRI-3.0 CH4/air ignition-delay map
(T0 1000-2000 K x p 1-40 atm x phi 0.5-2, to 1 ms) on every GPU of the box through the product's multi-GPU path
(chemtools_b200.sharded.Ensemble -> ct_ensemble_*: one host thread + stream per GPU, cyclic partition, 65 536-reactor
launches, recovery pass, results gathered in pinned host arrays), then one HDF5 file (tau map, final temperature, status,
step counts).  North-star target: >= 1 M reactors per minute on 8 B200.

  python tools/full_map.py [--planes 256] [--super 32] [--devices 0,1,...] [--out /tmp/config5_map.h5] [--log gpurun_out/full_map.json]

--planes P integrates P of the 256 T0 planes, evenly spread over T0 (P = 256 is the whole map).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--planes", type=int, default=256, help="number of T0 planes (evenly spread over the 256; 256 = the whole map)")
    ap.add_argument("--super", type=int, default=32, help="T0 planes per ensemble call (host memory: 65536 * 0.9 KB per plane)")
    ap.add_argument("--devices", default="")
    ap.add_argument("--chunk", type=int, default=65536)
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--atol", type=float, default=1e-14)
    ap.add_argument("--max-steps", type=int, default=50000)
    ap.add_argument("--first-pass", type=int, default=-1, help="step budget of the chunk launches (-1 = the library's default)")
    ap.add_argument("--streams", type=int, default=0)
    ap.add_argument("--out", default="/tmp/config5_map.h5")
    ap.add_argument("--log", default=os.path.join(ROOT, "gpurun_out", "full_map.json"))
    args = ap.parse_args()

    import chemtools_b200 as ctb
    from chemtools_b200 import sharded
    from chemtools_b200.results import HDF5Writer
    devices = [int(d) for d in args.devices.split(",")] if args.devices else list(range(ctb.device_count()))
    chem = ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
    K, NS = chem.n_species, 256
    T_axis = np.linspace(1000.0, 2000.0, NS)
    p_axis = np.geomspace(1.0, 40.0, NS) * 101325.0
    phi_axis = np.linspace(0.5, 2.0, NS)
    X = np.zeros((NS, K))
    X[:, chem.speciesIndex("CH4")] = phi_axis; X[:, chem.speciesIndex("O2")] = 2.0; X[:, chem.speciesIndex("N2")] = 7.52
    Y_phi = chem.mass_fractions(X)  # composition depends on phi only
    planes = np.unique(np.round(np.linspace(0, NS - 1, args.planes)).astype(int))
    n_total = len(planes) * NS * NS
    tau = np.full(n_total, np.nan); T_end = np.zeros(n_total); status = np.zeros(n_total, dtype=np.int32); nst = np.zeros(n_total, dtype=np.int32)
    per = args.super * NS * NS
    # page-locked host arrays, reused by every ensemble call; each GPU fills its own rows
    pin = {k: sharded.PinnedArray(s, d) for k, s, d in (("T", (per,), np.float64), ("p", (per,), np.float64), ("Y", (per, K), np.float64),
                                                        ("state", (per, K + 1), np.float64), ("tau", (per,), np.float64),
                                                        ("status", (per,), np.int32), ("stats", (per, 8), np.int64))}
    log = dict(workload="BASELINE config 5: %d x 256 x 256 constant-pressure GRI-3.0 CH4/air reactors to 1 ms" % len(planes), n_reactors=n_total,
               devices=devices, rtol=args.rtol, atol=args.atol, max_steps=args.max_steps, chunk=args.chunk, partition="cyclic", calls=[])
    t_integrate = 0.0
    t_wall0 = time.perf_counter()
    done = 0
    for s0 in range(0, len(planes), args.super):
        pl = planes[s0:s0 + args.super]
        m = len(pl) * NS * NS
        iT = np.repeat(pl, NS * NS); ip = np.tile(np.repeat(np.arange(NS), NS), len(pl)); iphi = np.tile(np.arange(NS), len(pl) * NS)
        T, p, Y = pin["T"].array[:m], pin["p"].array[:m], pin["Y"].array[:m]
        T[:] = T_axis[iT]; p[:] = p_axis[ip]; np.take(Y_phi, iphi, axis=0, out=Y)
        e = sharded.Ensemble(chem, ctb.Type.Const_Pres, m, devices, "cyclic", args.chunk)
        e.SetTolerances(args.rtol, args.atol); e.SetStopTime(1e-3); e.SetMaxNumSteps(args.max_steps); e.SetRetry(True)
        if args.first_pass >= 0:
            e.SetFirstPassSteps(args.first_pass)
        if args.streams:
            e.SetStreams(args.streams)
        out = {k: pin[k].array[:m] for k in ("state", "tau", "status", "stats")}
        t0 = time.perf_counter()
        res = e.Integrate(T, p, Y, out=out)
        dt = time.perf_counter() - t0
        t_integrate += dt
        tau[done:done + m] = res["tau"]; T_end[done:done + m] = res["state"][:, 0]; status[done:done + m] = res["status"]; nst[done:done + m] = res["stats"][:, 0]
        done += m
        call = dict(planes=[int(pl[0]), int(pl[-1])], reactors=m, seconds=dt, reactors_per_s=m / dt, code=res["code"],
                    failed_first=sum(s["failed_first"] for s in res["shards"]), failed_final=sum(s["failed_final"] for s in res["shards"]),
                    shard_wall_s=[round(s["wall_s"], 2) for s in res["shards"]], shard_kernel_s=[round(s["kernel_ms"] * 1e-3, 2) for s in res["shards"]],
                    launches=sum(s["launches"] for s in res["shards"]))
        log["calls"].append(call)
        print("planes %3d-%3d: %8d reactors in %6.1f s = %7.0f reactors/s, failed %d -> %d, shard wall %s" %
              (pl[0], pl[-1], m, dt, m / dt, call["failed_first"], call["failed_final"], call["shard_wall_s"]), flush=True)
    wall = time.perf_counter() - t_wall0
    log.update(integrate_seconds=t_integrate, wall_seconds_incl_ic_generation=wall, reactors_per_s=n_total / t_integrate, reactors_per_minute=60.0 * n_total / t_integrate,
               north_star_target_reactors_per_minute=1.0e6, failed_first=sum(c["failed_first"] for c in log["calls"]), failed_final=int((status < 0).sum()),
               ignited=int(np.isfinite(tau).sum()), nst_mean=float(nst.mean()), nst_max=int(nst.max()),
               nst_percentiles={str(q): float(np.percentile(nst, q)) for q in (50, 90, 99, 99.9, 99.99)}, first_pass=args.first_pass, streams=args.streams,
               T_end_min=float(T_end.min()), T_end_max=float(T_end.max()), tau_min=float(np.nanmin(tau)), tau_max=float(np.nanmax(tau)))
    # one HDF5 write of the gathered map (the reference's layout: /<group>/<set> f64 + Name / Notation / Unit)
    t0 = time.perf_counter()
    shape = (len(planes), NS, NS)
    w = HDF5Writer()
    w.add("/Axes/Initial temperature", T_axis[planes], "Initial temperature", "T0", "K")
    w.add("/Axes/Pressure", p_axis, "Pressure", "p", "Pa")
    w.add("/Axes/Equivalence ratio", phi_axis, "Equivalence ratio", "phi", "-")
    w.add("/Ignition delay/Ignition delay", tau.reshape(shape), "Ignition delay", "tau", "s")
    w.add("/Temperature/Temperature", T_end.reshape(shape), "Temperature", "T", "K")
    w.add("/Solver/Status", status.astype(np.float64).reshape(shape), "CVODE return code", "status", "-")
    w.add("/Solver/nst", nst.astype(np.float64).reshape(shape), "nst", "nst", "-")
    w.write(args.out)
    log.update(hdf5_file=args.out, hdf5_bytes=os.path.getsize(args.out), hdf5_write_seconds=time.perf_counter() - t0)
    # physics sanity of the gathered map: at fixed p and phi the ignition delay falls with T0
    tm = tau.reshape(shape)
    mono_viol = 0
    checked = 0
    for j in range(0, NS, 37):
        for k in range(0, NS, 41):
            col = tm[:, j, k]
            ok = np.isfinite(col)
            if ok.sum() > 2:
                checked += 1
                mono_viol += int(np.any(np.diff(col[ok]) > 0))
    log.update(tau_monotone_columns_checked=checked, tau_monotone_violations=mono_viol)
    # a thumbnail of the map for the record (every 8th point per axis)
    thumb = tm[::max(1, len(planes) // 32), ::8, ::8]
    os.makedirs(os.path.dirname(os.path.abspath(args.log)), exist_ok=True)
    np.save(os.path.splitext(args.log)[0] + "_tau_thumb.npy", thumb.astype(np.float32))
    json.dump(log, open(args.log, "w"), indent=1)
    print(json.dumps({k: v for k, v in log.items() if k != "calls"}))


if __name__ == "__main__":
    main()
