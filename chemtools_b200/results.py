"""Results writer: the HDF5 layout of the reference's Results::HDF5Writer
(/root/reference/src/results/hdf5_writer.cpp:136-217): groups `Time`, `Temperature`, `Mass fractions`; datasets
/<group>/<set> (f64); attributes Name / Notation / Unit as registered in Base::RegisterBasicResults
(/root/reference/src/applications/reactors/base.cpp:123-142) and EnergyEnabled::RegisterAdditionalResults (:210-218).
/root/reference/scripts/results/results.py reads exactly this (f[group][set][:], .attrs)."""
import numpy as np

from ._lib import check, dp, lib, lp


class HDF5Writer:
    def __init__(self):
        self._h = lib().ct_h5_create()

    def set_layout(self, chunk_elems=0, deflate_level=0, vlen_strings=True):
        """Storage of the datasets added afterwards: (1000, 6) = the reference's unlimited / chunked / deflated rank-1
        datasets (hdf5_writer.cpp:156-181); vlen_strings: variable-length string attributes like the reference's."""
        check(lib().ct_h5_set_layout(self._h, int(chunk_elems), int(deflate_level), int(bool(vlen_strings))))

    def add(self, path, data, name=None, notation=None, unit=None):
        a = np.ascontiguousarray(data, dtype=np.float64)
        dims = np.array(a.shape, dtype=np.int64)
        enc = lambda s: None if s is None else s.encode()
        check(lib().ct_h5_add_dataset(self._h, path.encode(), a.ndim, dims.ctypes.data_as(lp), a.ctypes.data_as(dp), enc(name), enc(notation), enc(unit)))

    def write(self, filename):
        check(lib().ct_h5_write(self._h, str(filename).encode()))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ct_h5_destroy(self._h)
                self._h = None
        except Exception:
            pass


def write_case_hdf5(filename, res, energy_enabled=True, reference_storage=True):
    """One reactor, the reference's layout verbatim: 1-D datasets of n_outputs samples, variable-length string attributes
    and (reference_storage) the reference's unlimited dimension, chunks of 1000 and deflate level 6."""
    w = HDF5Writer()
    if reference_storage:
        w.set_layout(1000, 6, True)
    w.add("/Time/Time", res["time"], "Time", "t", "s")
    if energy_enabled:
        w.add("/Temperature/Temperature", res["temperature"], "Temperature", "T", "K")
    for k, sp in enumerate(res["species"]):
        w.add("/Mass fractions/" + sp, res["mass_fractions"][:, k], "Y_" + sp, "Y_" + sp, "-")
    w.write(filename)


def write_batch_hdf5(filename, species, time, traj, energy_enabled=True, tau=None, status=None, stats=None):
    """A whole batch in one file: same groups and attributes, datasets are [reactor][sample]."""
    w = HDF5Writer()
    w.add("/Time/Time", time, "Time", "t", "s")
    sh = 1 if energy_enabled else 0
    if energy_enabled:
        w.add("/Temperature/Temperature", traj[:, :, 0], "Temperature", "T", "K")
    for k, sp in enumerate(species):
        w.add("/Mass fractions/" + sp, traj[:, :, sh + k], "Y_" + sp, "Y_" + sp, "-")
    if tau is not None:
        w.add("/Ignition delay/Ignition delay", tau, "Ignition delay", "tau", "s")
    if status is not None:
        w.add("/Solver/Status", np.asarray(status, dtype=np.float64), "CVODE return code", "status", "-")
    if stats is not None:
        for j, nm in enumerate(("nst", "nfe", "nsetups", "netf", "nni", "ncfn", "nje", "nfeLS")):
            w.add("/Solver/" + nm, np.asarray(stats)[:, j].astype(np.float64), nm, nm, "-")
    w.write(filename)
