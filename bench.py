#!/usr/bin/env python
"""Benchmark of the reactor hot path (contract: see the task brief).

A "step" = one pass of the hot path over one batch of synthetic initial conditions: every reactor of the batch integrated
from 0 to t_end by the device-side BDF stepper.  Metric = GRI-3.0 reactors integrated to 1 ms per second (BASELINE.json).

Headline workload = BASELINE config 2 (constant-volume H2/air sweep, 4096 reactors per GPU).  Every line also carries
  * `tolerances`: the headline at rtol 1e-8 / atol 1e-14 and at the reference test's 1e-9 / 1e-15, the second with the
    largest ignition-delay deviation from the oracle on a 64-reactor sample (a parity number inside the driver's record);
  * `config5`: the north-star workload (constant-pressure GRI-3.0 CH4/air map, BASELINE config 5) measured in the same
    run on a sample of the 256^3 map, with its own value / e2e / status before and after the recovery pass / roofline /
    cpu_baseline.

Multi-GPU (torchrun, one process per GPU): ONE global index set per workload (4096 x N config-2 reactors: 64 T0 x 64 N
phi; 65536 x N config-5 map points) partitioned over the ranks by the library's own rule (ct_partition_range, cyclic: every
rank sees the same stiffness mix), no collective on the solve path; the e2e region ends with the gather of every rank's
results on rank 0 (torch.distributed.gather over NCCL) — weak scaling.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|config5|config3_gri211|config3_gri12|config4]
                  [--batch B] [--impl ours|reference] [--no-cpu-baseline] [--no-config5] [--no-tight]
"""
import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RTOL, ATOL, MAX_STEPS = 1e-8, 1e-14, 20000  # 20000 steps: the reference reactor test's budget (tests/src/reactor.cpp:93)
RTOL_REF, ATOL_REF = 1e-9, 1e-15  # the reference test's setting (tests/src/reactor.cpp:62-63)
METRIC = "GRI-3.0 reactors integrated to 1 ms per sec"
RT_NAMES = {0: "Const_Pres", 1: "Const_Vol", 2: "Const_Temp_Pres", 3: "Const_Temp_Vol", 4: "Table_TP"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--batch", type=int, default=0, help="reactors per GPU per step (0 = the workload's own size)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of a cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the config5 sub-record")
    ap.add_argument("--no-tight", action="store_true", help="skip the 1e-9 / 1e-15 tolerance entry")
    return ap.parse_args()


def load_workloads():
    """chemtools_b200/workloads.py as a plain numpy module: neither arm needs the package for initial conditions, and
    the reference arm never maps the CUDA library."""
    spec = importlib.util.spec_from_file_location("_bench_workloads", os.path.join(ROOT, "chemtools_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class OracleChem:
    """speciesIndex / mass_fractions over the oracle's own mechanism tables (numpy only)."""

    def __init__(self, name):
        from oracle import oracle as O
        self.om = O.Mechanism(name)
        self.n_species = self.om.K

    def speciesIndex(self, s):
        return self.om.index[s]

    def mass_fractions(self, X):
        return self.om.X_to_Y(np.asarray(X, dtype=np.float64))


def cyclic_indices(n_global, world, rank):
    """The library's cyclic partition rule (ct_partition_range): reactor i -> rank i % world."""
    return np.arange(rank, n_global, world, dtype=np.int64)


def workload(W, name, chem_factory, batch, rank, world):
    """Rank `rank`'s shard of the global index set of a workload."""
    if name == "config2":
        chem = chem_factory("gri30")
        side = 64 if not batch else max(1, int(round(batch ** 0.5)))
        n_global = side * side * world
        idx = cyclic_indices(n_global, world, rank)
        mech, rt, T0, p0, X, t_end, extra = W.config2(chem, side, side * world, idx)
        desc = "config2: %d constant-volume GRI-3.0 H2/air reactors per GPU, T0=900-1500 K x phi=0.5-2.0, 1 atm, to 1 ms" % len(T0)
    elif name == "config5":
        chem = chem_factory("gri30")
        n = batch or 65536
        n_global = n * world
        design = W.config5_design(n_global)
        idx = design[cyclic_indices(n_global, world, rank)]
        mech, rt, T0, p0, X, t_end, extra = W.config5_index(chem, idx)
        extra = dict(extra, map_index=idx)
        desc = "config5 sample: %d constant-pressure GRI-3.0 CH4/air reactors per GPU spread over the 256^3 T0 x p x phi map (3-D Kronecker sequence), to 1 ms" % n
    elif name.startswith("config3"):
        m = name.split("_")[1]
        chem = chem_factory(m)
        n = batch or 65536
        n_global = n * world
        mech, rt, T0, p0, X, t_end, extra = W.config3(chem, m, n_global, 20211)
        idx = cyclic_indices(n_global, world, rank)
        T0, p0, X = T0[idx], p0[idx], X[idx]
        desc = "config3: %d constant-pressure %s CH4/air reactors per GPU, randomised T0/p/phi, to 1 ms" % (n, m)
    elif name == "config4":
        chem = chem_factory("gri30")
        n = batch or 262144
        n_global = n * world
        mech, rt, T0, p0, X, t_end, extra = W.config4(chem, n_global, 101, 30)
        idx = cyclic_indices(n_global, world, rank)
        tt, tT, tp = extra["table"]
        T0, p0, X, extra = T0[idx], p0[idx], X[idx], {"table": (tt, tT[idx], tp[idx])}
        desc = "config4: %d prescribed T(t),p(t) GRI-3.0 reactors per GPU (101-point tables), to 1 ms" % n
    else:
        raise SystemExit("unknown workload " + name)
    Y = chem.mass_fractions(X)
    return dict(name=name, mech=mech, chem=chem, rt=int(rt), T0=np.ascontiguousarray(T0), p0=np.ascontiguousarray(p0), Y=np.ascontiguousarray(Y),
                t_end=t_end, extra=extra, desc=desc, n_global=n_global)


def config_of(wl, world, rtol, atol):
    """The `config` object: the same keys and values in both arms."""
    return {"workload": wl["desc"], "reactors_per_gpu_per_step": len(wl["T0"]), "reactors_global_per_step": wl["n_global"], "n_gpus": world,
            "rtol": rtol, "atol": atol, "max_steps": MAX_STEPS, "t_end": wl["t_end"],
            "partition": "one global index set, cyclic over the ranks (reactor i -> rank i % N), no collective on the solve path",
            "l2": "256 MiB buffer rewritten between iterations (L2 flush)"}


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


# ------------------------------------------------------------------------------------------------ CPU side (oracle)
_native_oracle = [False]


def cpu_reference_run(wl, sample_idx, nthreads, rtol=RTOL, atol=ATOL):
    """The reference-equivalent CPU path (oracle: Cantera-restated RHS + CVODE-restated BDF with dense DQ Jacobian),
    one reactor per host thread.  Returns (seconds, results)."""
    from oracle import oracle as O
    if not _native_oracle[0]:
        _native_oracle[0] = True
        try:  # -march=native build of the same sources for the timing legs
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "libct_oracle_native.so"])
            O.build = lambda force=False: os.path.join(ROOT, "oracle", "libct_oracle_native.so")
        except Exception:
            pass
    om = wl.get("_om") or O.Mechanism(wl["mech"])
    wl["_om"] = om
    T0, p0, Y = wl["T0"][sample_idx], wl["p0"][sample_idx], wl["Y"][sample_idx]
    table = None
    if "table" in wl["extra"]:
        tt, tT, tp = wl["extra"]["table"]
        table = (tt, tT[sample_idx], tp[sample_idx])
    t0 = time.perf_counter()
    r = om.integrate_batch(RT_NAMES[wl["rt"]], T0, p0, Y, wl["t_end"], rtol=rtol, atol=atol, max_steps=MAX_STEPS, nthreads=nthreads, table=table)
    return time.perf_counter() - t0, r


def cpu_baseline_leg(wl, seconds, stat_names):
    """Bounded sample of the same workload on all host cores (probe first, then a sample sized for `seconds`)."""
    n = len(wl["T0"])
    cores = os.cpu_count() or 1
    probe = np.linspace(0, n - 1, min(n, 2 * cores)).astype(int)
    tp, _ = cpu_reference_run(wl, probe, cores)
    ns = int(min(n, max(2 * cores, seconds / max(tp / len(probe), 1e-6))))
    sample = np.linspace(0, n - 1, ns).astype(int)
    ts, r = cpu_reference_run(wl, sample, cores)
    return {"value": ns / ts, "unit": "reactors/s", "cores": cores, "kind": "port",
            "sample": "%d reactors evenly strided over the same workload, one reactor per host thread (%d threads), oracle = restated Cantera RHS + CVODE BDF with dense DQ Jacobian, %.1f s" % (ns, cores, ts),
            "failed": int((r["status"] < 0).sum()),
            "mean_counters": dict(zip(stat_names, r["stats"].astype(float).mean(0).tolist()))}


def reference_arm(args):
    """--impl reference: the reference's CPU path (its restatement; Cantera/SUNDIALS are not installable offline), all host
    threads, a bounded sample of the same workload per step.  Rank 0 only.  Numpy + oracle/ only: the CUDA library is
    neither imported nor mapped here."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    W = load_workloads()
    cache = {}

    def chem_factory(m):
        if m not in cache:
            cache[m] = OracleChem(m)
        return cache[m]
    world = max(1, args.gpus)
    wl = workload(W, args.workload, chem_factory, args.batch, 0, world)
    wl["_om"] = wl["chem"].om
    n = len(wl["T0"])
    cores = os.cpu_count() or 1
    # size the per-step sample from a probe so that steps + warmup stay within ~2 minutes
    probe = np.linspace(0, n - 1, min(n, 2 * cores)).astype(int)
    tp, _ = cpu_reference_run(wl, probe, cores)
    budget = 90.0 / max(1, args.steps + args.warmup)
    ns = int(min(n, max(cores, budget / max(tp / len(probe), 1e-6))))
    sample = np.linspace(0, n - 1, ns).astype(int)
    for _ in range(args.warmup):
        cpu_reference_run(wl, sample, cores)
    tot = 0.0
    for _ in range(args.steps):
        t, r = cpu_reference_run(wl, sample, cores)
        tot += t
    value = ns * args.steps / tot
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "reactors/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wl, world, RTOL, ATOL),
            "cpu_baseline": {"value": value, "unit": "reactors/s", "cores": cores, "kind": "port",
                             "sample": "%d of the %d reactors of rank 0's shard per step (evenly strided), one reactor per host thread on %d threads, CVODE-style dense DQ Jacobian; a rate, so the sample size does not enter the ratio" % (ns, n, cores),
                             "failed": int((r["status"] < 0).sum())},
            "e2e": {"value": value, "unit": "reactors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "native_library_imported": "chemtools_b200" in sys.modules}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU side
class GpuRunner:
    def __init__(self, args):
        import torch
        import chemtools_b200 as ctb
        self.torch, self.ctb, self.args = torch, ctb, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", rank=self.rank, world_size=self.world, device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        torch.cuda.set_device(self.local_rank)
        ctb.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.stream = torch.cuda.current_stream(self.dev)
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)  # 256 MiB > 126 MB L2
        self.W = load_workloads()
        self._chem = {}
        self.peak = None

    def chem_factory(self, m):
        if m not in self._chem:
            self._chem[m] = self.ctb.ThermoKinetics(self.ctb.mechanism_path(m))
        return self._chem[m]

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def gather_to_rank0(self, arrays):
        """The final gather of the sharded run: every rank's result rows to rank 0 (NCCL gather of device tensors, then one
        D2H on rank 0).  Returns the bytes rank 0 received."""
        if self.dist is None:
            return 0
        torch, got = self.torch, 0
        for a in arrays:
            t = torch.from_numpy(np.ascontiguousarray(a)).to(self.dev, non_blocking=True)
            bufs = [torch.empty_like(t) for _ in range(self.world)] if self.rank == 0 else None
            self.dist.gather(t, bufs, dst=0)
            if self.rank == 0:
                full = torch.cat(bufs, 0).cpu()
                got += full.numel() * full.element_size()
        return got

    def run(self, wl, rtol, atol, steps, warmup, e2e=True, sample_clocks=False):
        """value (resident inputs, CUDA events, max over ranks) and e2e (host buffers through the C ABI, copies and the
        final gather inside the timed region) of one workload."""
        torch, ctb = self.torch, self.ctb
        L = ctb.reactors.lib()
        chem, n, K = wl["chem"], len(wl["T0"]), wl["chem"].n_species
        N = K + 1 if wl["rt"] <= 1 else K
        dT = torch.from_numpy(wl["T0"]).to(self.dev); dp = torch.from_numpy(wl["p0"]).to(self.dev); dY = torch.from_numpy(wl["Y"]).to(self.dev)
        hT = torch.from_numpy(wl["T0"]).pin_memory(); hp = torch.from_numpy(wl["p0"]).pin_memory(); hY = torch.from_numpy(wl["Y"]).pin_memory()
        table = wl["extra"].get("table")
        # expected-cost order: hottest first (longest integrations start first; shortens the tail of the work queue)
        order = np.argsort(-wl["T0"], kind="stable").astype(np.int32)

        def make_batch(resident):
            if resident:
                b = ctb.ReactorBatch(wl["rt"], chem, dT, dp, dY, rtol, atol, wl["t_end"], max_num_steps=MAX_STEPS, device_arrays=True)
            else:
                b = ctb.ReactorBatch(wl["rt"], chem, hT.numpy(), hp.numpy(), hY.numpy(), rtol, atol, wl["t_end"], max_num_steps=MAX_STEPS)
            b.SetStream(self.stream.cuda_stream)
            if table is not None:
                b.SetTable(*table)
            b.SetOrder(order)
            return b

        def step_resident(b):
            self.flush.zero_()  # L2 flush between iterations
            ctb.reactors.check(L.ct_batch_set_ic_device(b._h, dT.data_ptr(), dp.data_ptr(), dY.data_ptr()))
            return b.Integrate(wl["t_end"])

        def step_e2e(b):
            self.flush.zero_()
            ctb.reactors.check(L.ct_batch_set_ic(b._h, ctb.reactors._p(hT.numpy()), ctb.reactors._p(hp.numpy()), ctb.reactors._p(hY.numpy())))
            r = b.IntegrateWithRetry()  # Integrate + recovery pass + D2H of state, tau, counters, status
            got = self.gather_to_rank0([r["state"], r["ignition_delay"], r["statistics"], r["status"]])
            return r, got

        b = make_batch(True)
        code = 0
        for _ in range(max(warmup, 1)):
            code = step_resident(b)
        if code < 0:
            # per-reactor CVODE failure codes (CV_TOO_MUCH_WORK ...) are results, not errors: counted in "status"; a library
            # error (code <= -100) has already raised
            print("bench: worst per-reactor status %d (see status.failed_first)" % code, file=sys.stderr)
        sampler = ClockSampler(self.local_rank) if (sample_clocks and self.rank == 0) else None
        if sampler:
            sampler.start()
        l0 = b.LaunchCount()
        kernel_ms = []
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            step_resident(b)
            kernel_ms.append(b.LastKernelMs())
        e1.record(self.stream)
        self.barrier()
        ms_total = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        stats, status = b.Statistics(), b.Status()
        tau, state = b.IgnitionDelay(), b.Solution()
        out = dict(n=n, N=N, K=K, launches=b.LaunchCount() - l0, warps_per_sm=b.WarpsPerSM(), clocks=clocks, stats=stats, status=status, tau=tau, state=state)
        del b
        e2e_ms, gathered, f_first, f_final = 0.0, 0, int((status < 0).sum()), None
        if e2e:
            b2 = make_batch(False)
            for _ in range(max(1, min(warmup, 2))):
                step_e2e(b2)
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                r, gathered = step_e2e(b2)
            torch.cuda.synchronize(self.dev)
            e2e_ms = (time.perf_counter() - t0) * 1e3
            self.barrier()
            f_first, f_final = r["failed_first"], r["failed_final"]
            del b2
        tmax = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=self.dev)
        fsum = torch.tensor([float(f_first), float(f_final or 0), float(n)], dtype=torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(tmax, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(fsum, op=self.dist.ReduceOp.SUM)
        ms_total, e2e_ms = tmax.tolist()
        # ---- roofline of the dominant kernel (k_integrate): algorithmic fp64 flops / measured kernel time (rank 0's shard)
        fm = chem.flop_model(wl["rt"])
        st = stats.astype(np.float64)
        flops = float((st[:, 1] + st[:, 7]).sum() * fm["rhs"] + st[:, 6].sum() * fm["jac"] + st[:, 2].sum() * fm["lu"] + st[:, 4].sum() * fm["solve"])
        k_ms = float(np.mean(kernel_ms))
        if self.peak is None:
            self.peak = ctb.measure_fp64_peak()
        achieved = flops / (k_ms * 1e-3) / 1e12
        world = self.world
        out.update(
            value=world * n * steps / (ms_total * 1e-3), ms_per_step=ms_total / steps,
            e2e=None if not e2e else {"value": world * n * steps / (e2e_ms * 1e-3), "unit": "reactors/s", "h2d_bytes_per_step": n * (2 + K) * 8,
                                      "d2h_bytes_per_step": n * N * 8 + n * 8 + n * 8 * 8 + n * 4 + (gathered if self.rank == 0 else 0),
                                      "includes": "H2D of T, p, Y from pinned host memory; Integrate + recovery pass; D2H of state, tau, counters, status" + ("; NCCL gather of every rank's results on rank 0 + D2H there" if world > 1 else "")},
            roofline={"bound": "fp64", "kernel": "k_integrate", "achieved": achieved, "peak": self.peak, "unit": "TFLOP/s", "frac": achieved / self.peak,
                      "peak_source": "measured live: register-resident DFMA loop (MEASURED_PEAKS.json has no fp64 figure; ncu counters of that loop: profiles/r02_dfma_peak_ncu.csv)",
                      "flops_per_launch": flops, "kernel_ms": k_ms, "kernel_share_of_step": k_ms * steps / ms_total, "flop_model": fm,
                      "mean_counters": dict(zip(ctb.reactors.STAT_NAMES, st.mean(0).tolist()))},
            status={"n_global": int(fsum[2].item()), "failed_first": int(fsum[0].item()), "failed_after_retry": None if f_final is None else int(fsum[1].item()),
                    "max_steps": MAX_STEPS})
        return out


def ncu_numbers(workload_name):
    """DRAM traffic and pipe counters of k_integrate from the committed ncu capture of this command (profiles/traffic.json)."""
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(tf))
        return d.get(workload_name), d.get(workload_name + "_counters")
    except Exception:
        return None, None


def tau_deviation(tau_gpu, r_cpu):
    ign = np.isfinite(r_cpu["tau"])
    if not np.array_equal(ign, np.isfinite(tau_gpu)) or ign.sum() == 0:
        return None
    return float(np.abs(tau_gpu[ign] / r_cpu["tau"][ign] - 1).max())


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)
    G = GpuRunner(args)
    ctb, rank, world = G.ctb, G.rank, G.world
    wl = workload(G.W, args.workload, G.chem_factory, args.batch, rank, world)
    res = G.run(wl, RTOL, ATOL, args.steps, args.warmup, e2e=True, sample_clocks=True)
    tight = None
    if not args.no_tight:
        tight = G.run(wl, RTOL_REF, ATOL_REF, args.steps, args.warmup, e2e=False)
    res5 = wl5 = None
    if not args.no_config5 and args.workload != "config5":
        wl5 = workload(G.W, "config5", G.chem_factory, 0, rank, world)
        # 65 536 reactors per GPU and launch (8 s): half the steps of the headline so that the default run stays within minutes
        steps5 = max(2, (args.steps + 1) // 2)
        res5 = G.run(wl5, RTOL, ATOL, steps5, args.warmup, e2e=True)
    if rank == 0:
        traffic, counters = ncu_numbers(args.workload)
        res["roofline"]["traffic"] = traffic
        if counters:
            res["roofline"]["ncu_counters"] = counters
        line = {
            "metric": METRIC, "value": res["value"], "unit": "reactors/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wl, world, RTOL, ATOL),
            "implementation": {"jacobian": "analytic", "newton_solver": "equilibrated Gauss-Jordan inverse + mat-vec",
                               "scheduling": "persistent CTAs, time-sliced work ring (64 steps per slice), phase-aligned warps, Newton matrices in a shared pool",
                               "warps_per_sm": res["warps_per_sm"]},
            "e2e": res["e2e"], "gpu_launches": res["launches"], "roofline": res["roofline"], "clocks": res["clocks"], "status": res["status"],
        }
        stat_names = ctb.reactors.STAT_NAMES
        cores = os.cpu_count() or 1
        tol = [{"rtol": RTOL, "atol": ATOL, "value": res["value"], "ms_per_step": res["ms_per_step"]}]
        if tight is not None:
            tol.append({"rtol": RTOL_REF, "atol": ATOL_REF, "value": tight["value"], "ms_per_step": tight["ms_per_step"], "roofline_frac": tight["roofline"]["frac"],
                        "failed": tight["status"]["failed_first"]})
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(wl, args.cpu_seconds, stat_names)
            # parity numbers inside the driver's record: 64 reactors of this shard against the oracle at matched tolerances
            sample = np.linspace(0, res["n"] - 1, min(64, res["n"])).astype(int)
            for entry, r_gpu in zip(tol, (res, tight)):
                if r_gpu is None:
                    continue
                _, rc = cpu_reference_run(wl, sample, cores, entry["rtol"], entry["atol"])
                entry["max_tau_dev_vs_oracle_sample"] = tau_deviation(r_gpu["tau"][sample], rc)
                entry["max_T_dev_vs_oracle_sample"] = float(np.abs(r_gpu["state"][sample, 0] / rc["state"][:, 0] - 1).max()) if wl["rt"] <= 1 else None
                entry["oracle_sample"] = len(sample)
        line["tolerances"] = tol
        if res5 is not None:
            t5, c5 = ncu_numbers("config5")
            res5["roofline"]["traffic"] = t5
            rec = {"config": config_of(wl5, world, RTOL, ATOL), "value": res5["value"], "unit": "reactors/s", "ms_per_step": res5["ms_per_step"], "steps": steps5, "warmup": args.warmup,
                   "reactors_per_minute": 60.0 * res5["value"], "north_star_target_reactors_per_minute_on_8_gpus": 1.0e6,
                   "e2e": res5["e2e"], "status": res5["status"], "roofline": res5["roofline"], "gpu_launches": res5["launches"]}
            if not args.no_cpu_baseline:
                rec["cpu_baseline"] = cpu_baseline_leg(wl5, args.cpu_seconds, stat_names)
                sample = np.linspace(0, res5["n"] - 1, 32).astype(int)
                _, rc = cpu_reference_run(wl5, sample, cores)
                good = rc["status"] >= 0
                rec["max_tau_dev_vs_oracle_sample"] = tau_deviation(res5["tau"][sample][good], {"tau": rc["tau"][good]})
                rec["oracle_sample"] = int(good.sum())
            line["config5"] = rec
        print(json.dumps(line))
    if G.dist is not None:
        G.dist.barrier()
        G.dist.destroy_process_group()


if __name__ == "__main__":
    main()
