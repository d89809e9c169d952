#!/usr/bin/env python
"""Benchmark of the reactor hot path (contract: see the task brief).

A "step" = one pass of the hot path over one batch of synthetic initial
conditions: every reactor of the batch integrated from 0 to t_end by the
device-side BDF stepper.  Metric = GRI-3.0 reactors integrated to 1 ms per
second (BASELINE.json).  Default workload = BASELINE config 2 (4096
constant-volume H2/air reactors, one B200).  Multi-GPU: independent ensembles,
one process per GPU, no collective on the solve path (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|config5|config3_gri211|config3_gri12|config4]
                  [--batch B] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RTOL, ATOL, MAX_STEPS = 1e-8, 1e-14, 50000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--batch", type=int, default=0, help="reactors per GPU per step (0 = the workload's own size)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(name, chem_factory, batch, rank):
    import chemtools_b200 as ctb
    from chemtools_b200 import workloads as W
    if name == "config2":
        chem = chem_factory("gri30")
        side = 64 if not batch else max(1, int(round(batch ** 0.5)))
        mech, rt, T0, p0, X, t_end, extra = W.config2(chem, side)
        desc = "config2: %d constant-volume GRI-3.0 H2/air reactors, T0=900-1500 K x phi=0.5-2.0, 1 atm, to 1 ms" % len(T0)
    elif name == "config5":
        chem = chem_factory("gri30")
        n = batch or 65536
        # a strided sample of the 256^3 map so every slice sees the same stiffness mix
        stride = (256 ** 3) // n
        idx = (np.arange(n) * stride + rank) % (256 ** 3)
        mech, rt, t_end, extra = "gri30", ctb.Type.Const_Pres, 1e-3, {}
        iT, ip, iphi = idx // 65536, (idx // 256) % 256, idx % 256
        T0 = np.linspace(1000.0, 2000.0, 256)[iT]
        p0 = np.geomspace(1.0, 40.0, 256)[ip] * W.ONE_ATM
        phi = np.linspace(0.5, 2.0, 256)[iphi]
        X = W._fill(chem, n, {"CH4": phi, "O2": 2.0, "N2": 7.52})
        desc = "config5 slice: %d constant-pressure GRI-3.0 CH4/air reactors strided over the 256^3 T0 x p x phi map, to 1 ms" % n
    elif name.startswith("config3"):
        m = name.split("_")[1]
        chem = chem_factory(m)
        mech, rt, T0, p0, X, t_end, extra = W.config3(chem, m, batch or 65536, 20211 + rank)
        desc = "config3: %d constant-pressure %s CH4/air reactors, randomised T0/p/phi, to 1 ms" % (len(T0), m)
    elif name == "config4":
        chem = chem_factory("gri30")
        mech, rt, T0, p0, X, t_end, extra = W.config4(chem, batch or 262144, 101, 30 + rank)
        desc = "config4: %d prescribed T(t),p(t) GRI-3.0 reactors (101-point tables), to 1 ms" % len(T0)
    else:
        raise SystemExit("unknown workload " + name)
    Y = chem.mass_fractions(X)
    return dict(mech=mech, chem=chem, rt=rt, T0=T0, p0=p0, Y=Y, t_end=t_end, extra=extra, desc=desc)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(wl, sample_idx, nthreads, native=True):
    """The reference-equivalent CPU path (oracle: Cantera-restated RHS + CVODE-restated BDF with dense DQ
    Jacobian), one reactor per host thread.  Returns (seconds, stats)."""
    from oracle import oracle as O
    if native:
        try:
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "libct_oracle_native.so"])
            O.build = lambda force=False: os.path.join(ROOT, "oracle", "libct_oracle_native.so")
        except Exception:
            pass
    om = O.Mechanism(wl["mech"])
    T0, p0, Y = wl["T0"][sample_idx], wl["p0"][sample_idx], wl["Y"][sample_idx]
    table = None
    if "table" in wl["extra"]:
        tt, tT, tp = wl["extra"]["table"]
        table = (tt, tT[sample_idx], tp[sample_idx])
    rt = {0: "Const_Pres", 1: "Const_Vol", 2: "Const_Temp_Pres", 3: "Const_Temp_Vol", 4: "Table_TP"}[int(wl["rt"])]
    t0 = time.perf_counter()
    r = om.integrate_batch(rt, T0, p0, Y, wl["t_end"], rtol=RTOL, atol=ATOL, max_steps=MAX_STEPS, nthreads=nthreads, table=table)
    return time.perf_counter() - t0, r


def reference_arm(args):
    """--impl reference: the reference's CPU path (its restatement; Cantera/SUNDIALS are not installable offline),
    all host threads, bounded sample per step.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import chemtools_b200 as ctb
    cache = {}

    def chem_factory(m):
        if m not in cache:
            cache[m] = ctb.ThermoKinetics(ctb.mechanism_path(m))
        return cache[m]
    wl = workload(args.workload, chem_factory, args.batch, 0)
    n = len(wl["T0"])
    cores = os.cpu_count() or 1
    # size the per-step sample from a probe so that steps+warmup stays within ~2 minutes
    probe = np.linspace(0, n - 1, min(n, 2 * cores)).astype(int)
    tp, _ = cpu_reference_run(wl, probe, cores)
    per_reactor = tp / len(probe)
    budget = 90.0 / max(1, args.steps + args.warmup)
    ns = int(min(n, max(cores, budget / max(per_reactor, 1e-6))))
    sample = np.linspace(0, n - 1, ns).astype(int)
    for _ in range(args.warmup):
        cpu_reference_run(wl, sample, cores, native=False)
    times = []
    for _ in range(args.steps):
        t, r = cpu_reference_run(wl, sample, cores, native=False)
        times.append(t)
    tot = sum(times)
    value = ns * args.steps / tot
    line = {"impl": "reference", "metric": "GRI-3.0 reactors integrated to 1 ms per sec", "value": value, "unit": "reactors/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "rtol": RTOL, "atol": ATOL},
            "cpu_baseline": {"value": value, "unit": "reactors/s", "cores": cores, "kind": "port",
                             "sample": "%d reactors evenly strided over the workload per step, one reactor per host thread, CVODE-style dense DQ Jacobian" % ns},
            "e2e": {"value": value, "unit": "reactors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)
    import torch
    import chemtools_b200 as ctb
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    ctb.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cache = {}

    def chem_factory(m):
        if m not in cache:
            cache[m] = ctb.ThermoKinetics(ctb.mechanism_path(m))
        return cache[m]
    wl = workload(args.workload, chem_factory, args.batch, rank)
    chem, n, K = wl["chem"], len(wl["T0"]), wl["chem"].n_species
    N = K + 1 if int(wl["rt"]) <= 1 else K
    stream = torch.cuda.current_stream(dev)

    # ---- resident inputs (value) and pinned host inputs (e2e)
    dT = torch.from_numpy(wl["T0"]).to(dev); dp = torch.from_numpy(wl["p0"]).to(dev); dY = torch.from_numpy(wl["Y"]).to(dev)
    hT = torch.from_numpy(wl["T0"]).pin_memory(); hp = torch.from_numpy(wl["p0"]).pin_memory(); hY = torch.from_numpy(wl["Y"]).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2
    table = wl["extra"].get("table")
    # expected-cost order: hottest first (longest integrations start first; shortens the tail of the work queue)
    order = np.argsort(-wl["T0"], kind="stable").astype(np.int32)

    def make_batch(device_resident):
        if device_resident:
            b = ctb.ReactorBatch(wl["rt"], chem, dT, dp, dY, RTOL, ATOL, wl["t_end"], max_num_steps=MAX_STEPS, device_arrays=True)
        else:
            b = ctb.ReactorBatch(wl["rt"], chem, hT.numpy(), hp.numpy(), hY.numpy(), RTOL, ATOL, wl["t_end"], max_num_steps=MAX_STEPS)
        b.SetStream(stream.cuda_stream)
        if table is not None:
            b.SetTable(*table)
        b.SetOrder(order)
        return b

    launches = [0]

    def step_resident(b):
        flush.zero_()  # L2 flush between iterations
        ctb.reactors.check(ctb.reactors.lib().ct_batch_set_ic_device(b._h, dT.data_ptr(), dp.data_ptr(), dY.data_ptr()))
        code = b.Integrate(wl["t_end"])
        return code

    def step_e2e(b):
        flush.zero_()
        ctb.reactors.check(ctb.reactors.lib().ct_batch_set_ic(b._h, ctb.reactors._p(hT.numpy()), ctb.reactors._p(hp.numpy()), ctb.reactors._p(hY.numpy())))
        code = b.Integrate(wl["t_end"])
        st = b.Solution(); tau = b.IgnitionDelay(); stats = b.Statistics(); status = b.Status()
        return code, st, tau, stats, status

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    b = make_batch(True)
    for _ in range(args.warmup):
        code = step_resident(b)
    if code < 0:
        # per-reactor CVODE failure codes (CV_TOO_MUCH_WORK ...) are results, not errors: counted in "status" below; a
        # library error (code <= -100) has already raised
        print("bench: worst per-reactor status %d (see status.failed)" % code, file=sys.stderr)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches[0] = b.LaunchCount()
    kernel_ms = []
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident(b)
        kernel_ms.append(b.LastKernelMs())
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    stats = b.Statistics(); status = b.Status()
    gpu_launches = b.LaunchCount() - launches[0]  # the library's own count: k_init + k_ring_init + k_integrate per step

    # ---- e2e through the public API with host buffers
    b2 = make_batch(False)
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e(b2)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e(b2)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    barrier()
    h2d = n * (2 + K) * 8 + (0 if table is None else (table[1].size + table[2].size) * 0)
    d2h = n * (N + 1) * 8 + n * 8 * 8 + n * 4

    tmax = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = tmax.tolist()

    # ---- roofline of the dominant kernel (k_integrate): algorithmic fp64 flops / measured kernel time
    fm = chem.flop_model(int(wl["rt"]))
    st = stats.astype(np.float64)
    flops = float((st[:, 1] + st[:, 7]).sum() * fm["rhs"] + st[:, 6].sum() * fm["jac"] + st[:, 2].sum() * fm["lu"] + st[:, 4].sum() * fm["solve"])
    k_ms = float(np.mean(kernel_ms))
    peak = ctb.measure_fp64_peak()
    achieved = flops / (k_ms * 1e-3) / 1e12

    if rank == 0:
        ncu_traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            try:
                ncu_traffic = json.load(open(tf)).get(args.workload)
            except Exception:
                pass
        line = {
            "metric": "GRI-3.0 reactors integrated to 1 ms per sec", "value": world * n * args.steps / (ms_total * 1e-3), "unit": "reactors/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "reactors_per_gpu_per_step": n, "rtol": RTOL, "atol": ATOL, "jacobian": "analytic", "newton_solver": "equilibrated Gauss-Jordan inverse + mat-vec",
                       "scheduling": "persistent CTAs, time-sliced work ring (64 steps per slice), phase-aligned warps, Newton matrices in a shared pool",
                       "l2": "256 MiB buffer rewritten between iterations (L2 flush)", "parallelism": "independent ensembles x%d, no collective" % world,
                       "warps_per_sm": b.WarpsPerSM()},
            "e2e": {"value": world * n * args.steps / (e2e_ms * 1e-3), "unit": "reactors/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": gpu_launches,
            "roofline": {"bound": "fp64", "kernel": "k_integrate", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": "measured live: register-resident DFMA loop (MEASURED_PEAKS.json has no fp64 figure)",
                         "flops_per_launch": flops, "kernel_ms": k_ms, "kernel_share_of_step": k_ms * args.steps / ms_total,
                         "traffic": ncu_traffic,
                         "flop_model": fm, "mean_counters": dict(zip(ctb.reactors.STAT_NAMES, st.mean(0).tolist()))},
            "clocks": clocks,
            "status": {"failed": int((status < 0).sum()), "n": int(n), "max_steps": MAX_STEPS},
        }
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            probe = np.linspace(0, n - 1, min(n, 2 * cores)).astype(int)
            tp, _ = cpu_reference_run(wl, probe, cores)
            ns = int(min(n, max(2 * cores, args.cpu_seconds / max(tp / len(probe), 1e-6))))
            sample = np.linspace(0, n - 1, ns).astype(int)
            ts, r = cpu_reference_run(wl, sample, cores, native=False)
            line["cpu_baseline"] = {"value": ns / ts, "unit": "reactors/s", "cores": cores, "kind": "port",
                                    "sample": "%d reactors evenly strided over the same workload, one reactor per host thread (%d threads), oracle = restated Cantera RHS + CVODE BDF with dense DQ Jacobian, %.1f s" % (ns, cores, ts),
                                    "mean_counters": dict(zip(ctb.reactors.STAT_NAMES, r["stats"].astype(float).mean(0).tolist()))}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
