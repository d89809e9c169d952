"""The driver the reference never wrote (src/main.cpp:5-12 is a stub): Input::Parameters -> reactor batches ->
Integrate -> results.  One batch per (case set, total simulation time); every batch runs on the GPU through the C ABI."""
import os

import numpy as np

from . import input as Input
from .reactors import ReactorBatch, ThermoKinetics, Type

_TYPES = (("const_P", Type.Const_Pres), ("const_V", Type.Const_Vol), ("const_TP", Type.Const_Temp_Pres), ("const_TV", Type.Const_Temp_Vol))


def run(config_file, n_outputs=100, rtol=1e-9, atol=1e-15, max_num_steps=20000, results_dir=None, devices=None):
    """Integrate every case of a YAML configuration.  Defaults are the reference test's solver settings
    (tests/src/reactor.cpp:62-63, :93, :101).  Returns a list of result dicts, one per case; with `results_dir`
    each case is also written as <results_dir>/<set>_<case>.h5 in the reference's HDF5 layout.
    devices: GPU ordinals; the batches (one per case set and simulation time) are dealt out over them, one host thread
    and its own batches per GPU, nothing shared; the results are gathered here for the writer."""
    from concurrent.futures import ThreadPoolExecutor
    from .reactors import set_device
    parser = Input.Parser(config_file)
    jobs = []
    for attr, rtype in _TYPES:
        for si, cs in enumerate(getattr(parser, attr)):
            chem = ThermoKinetics(cs.mechanism_path)
            by_time = {}
            for ci, c in enumerate(cs.cases):
                by_time.setdefault(c.time, []).append((ci, c))
            for t_end, group in by_time.items():
                jobs.append((attr, rtype, si, cs, chem, t_end, group))
    devices = [0] if not devices else list(devices)

    def integrate(job_dev):
        (attr, rtype, si, cs, chem, t_end, group), dev = job_dev
        set_device(dev)  # cudaSetDevice is per host thread
        T = np.array([c.temperature for _, c in group]); p = np.array([c.pressure for _, c in group])
        Y = np.stack([parser.mass_fractions(c, chem) for _, c in group])
        b = ReactorBatch(rtype, chem, T, p, Y, rtol, atol, t_end, max_num_steps=max_num_steps)
        code, traj = b.IntegrateTrajectory(n_outputs)
        tau = b.IgnitionDelay() if rtype <= Type.Const_Vol else np.full(len(group), np.nan)
        return traj, b.Status(), b.Statistics(), tau

    def worker(dev_index):  # device d takes jobs d, d + G, d + 2G, ...
        return [(j, integrate((jobs[j], devices[dev_index]))) for j in range(dev_index, len(jobs), len(devices))]
    with ThreadPoolExecutor(max_workers=len(devices)) as pool:
        done = dict(r for part in pool.map(worker, range(len(devices))) for r in part)
    out = []
    for jn, (attr, rtype, si, cs, chem, t_end, group) in enumerate(jobs):
        traj, status, stats, tau = done[jn]
        times = np.array([(k + 1) * (t_end / n_outputs) if k + 1 < n_outputs else t_end for k in range(n_outputs)])
        for j, (ci, c) in enumerate(group):
            res = dict(reactor=rtype.name, case_set=si, case=ci, description=c.description, mechanism=cs.mechanism_path,
                       species=chem.species_names, time=times, status=int(status[j]), statistics=stats[j], ignition_delay=float(tau[j]))
            if rtype <= Type.Const_Vol:
                res["temperature"] = traj[j, :, 0]; res["mass_fractions"] = traj[j, :, 1:]
            else:
                res["temperature"] = np.full(n_outputs, c.temperature); res["mass_fractions"] = traj[j]
            if results_dir is not None:
                from .results import write_case_hdf5
                os.makedirs(results_dir, exist_ok=True)
                res["file"] = os.path.join(results_dir, "%s_set%d_case%d.h5" % (attr, si, ci))
                write_case_hdf5(res["file"], res, energy_enabled=rtype <= Type.Const_Vol)
            out.append(res)
    return out


def main(argv=None):
    """python -m chemtools_b200.app cases.yaml results_dir [n_outputs] [--devices 0,1,...]"""
    import sys
    args = list(sys.argv[1:] if argv is None else argv)
    devices = None
    if "--devices" in args:
        i = args.index("--devices")
        devices = [int(d) for d in args[i + 1].split(",")]
        del args[i:i + 2]
    if len(args) < 2:
        print(main.__doc__)
        return 2
    res = run(args[0], n_outputs=int(args[2]) if len(args) > 2 else 100, results_dir=args[1], devices=devices)
    for r in res:
        print("%-16s set %d case %d  status %d  steps %d  tau_ign %s" % (r["reactor"], r["case_set"], r["case"], r["status"],
                                                                      int(r["statistics"][0]), "%.6e s" % r["ignition_delay"] if np.isfinite(r["ignition_delay"]) else "-"))
    return 0 if all(r["status"] >= 0 for r in res) else 1


if __name__ == "__main__":
    raise SystemExit(main())
