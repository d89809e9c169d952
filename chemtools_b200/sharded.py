"""Multi-GPU ensembles in one process: n independent reactors partitioned over the GPUs of one
box, one host thread + CUDA stream per GPU inside the library (ct_ensemble_*), chunked launches, no collective on the
solve path; every GPU writes its results with cudaMemcpyAsync into disjoint slices of ONE pinned host array per result,
which then goes to the HDF5 writer (the reference's consumer: src/results/hdf5_writer.cpp:101-110; its single-reactor
producer: Base::Solve, src/applications/reactors/base.cpp:123-170).

    res = sharded.integrate(chem, Type.Const_Pres, T, p, Y, rtol=1e-8, atol=1e-14, t_end=1e-3, devices=[0, 1, 2, 3])
    res["state"], res["tau"], res["status"], res["stats"], res["shards"]

A sharded run is bit-identical to a single batch on one GPU: a shard is a sequence of ordinary batches.
"""
import ctypes as C

import numpy as np

from ._lib import CtError, check, dp, ip, lib, lp
from .reactors import Type, has_energy

PARTITION = {"block": 0, "cyclic": 1}
CONFIGURE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)


class PinnedArray:
    """numpy view of page-locked host memory (ct_host_alloc: cudaHostAllocPortable, usable by every device)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        dt = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * dt.itemsize
        self._ptr = C.c_void_p()
        check(lib().ct_host_alloc(nbytes, C.byref(self._ptr)))
        buf = (C.c_char * max(nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dt, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._ptr:
            self.array = None
            lib().ct_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pinned_like(a):
    """Copy of `a` in page-locked memory; returns (PinnedArray, ndarray view)."""
    a = np.ascontiguousarray(a)
    pa = PinnedArray(a.shape, a.dtype)
    pa.array[...] = a
    return pa


def partition_range(n, n_shards, shard, partition="block"):
    """(first, stride, count) of a shard's global reactor indices: the library's rule (ct_partition_range), usable by
    callers that run one process per GPU (bench.py under torchrun)."""
    a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
    check(lib().ct_partition_range(int(n), int(n_shards), PARTITION[partition], int(shard), C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


class Ensemble:
    """ct_ensemble handle: n reactors of one type / mechanism over `devices`."""

    def __init__(self, chemistry, reactor_type, n, devices=None, partition="block", chunk=0):
        self.chem, self.rt, self.n = chemistry, int(reactor_type), int(n)
        self.K = chemistry.n_species
        self.N = self.K + 1 if has_energy(self.rt) else self.K
        st = C.c_int(0)
        if devices is None:
            nd, dv = 0, None
        else:
            devs = np.ascontiguousarray(list(devices), dtype=np.int32)
            nd, dv = len(devs), devs.ctypes.data_as(ip)
            if nd == 0:
                raise ValueError("empty device list")
        self._h = lib().ct_ensemble_create(chemistry._h, self.rt, self.n, nd, dv, C.byref(st))
        if not self._h:
            raise CtError(st.value, lib().ct_last_error().decode())
        self._cb = None
        self.SetPartition(partition, chunk)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ct_ensemble_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def n_shards(self):
        return lib().ct_ensemble_n_shards(self._h)

    def SetPartition(self, partition, chunk=0):
        check(lib().ct_ensemble_set_partition(self._h, PARTITION[partition], int(chunk)))

    def SetTolerances(self, rtol, atol):
        check(lib().ct_ensemble_set_tolerances(self._h, float(rtol), float(atol)))

    def SetStopTime(self, t_end):
        check(lib().ct_ensemble_set_stop_time(self._h, float(t_end)))

    def SetMaxNumSteps(self, n):
        check(lib().ct_ensemble_set_max_steps(self._h, int(n)))

    def SetRetry(self, on):
        check(lib().ct_ensemble_set_retry(self._h, int(bool(on))))

    def SetFirstPassSteps(self, n):
        """Step budget of the chunk launches (the recovery pass re-runs the few longer reactors with the full budget)."""
        check(lib().ct_ensemble_set_first_pass_steps(self._h, int(n)))

    def SetStreams(self, n):
        """Launches in flight per device (the tail of one chunk overlaps the next chunk); default 2."""
        check(lib().ct_ensemble_set_streams(self._h, int(n)))

    def SetHotFirst(self, on):
        check(lib().ct_ensemble_set_hot_first(self._h, int(bool(on))))

    def SetConfigure(self, fn):
        """fn(batch_handle) is called on every internal batch: apply any lib().ct_batch_set_* option there."""
        if fn is None:
            self._cb = None
            check(lib().ct_ensemble_set_configure(self._h, None, None))
            return

        def tramp(handle, _user):
            try:
                r = fn(C.c_void_p(handle))
                return 0 if r is None else int(r)
            except Exception:
                return -100
        self._cb = CONFIGURE_FN(tramp)
        check(lib().ct_ensemble_set_configure(self._h, C.cast(self._cb, C.c_void_p), None))

    def SetTable(self, t, T, p):
        self._tab = (np.ascontiguousarray(t, dtype=np.float64), np.ascontiguousarray(T, dtype=np.float64), np.ascontiguousarray(p, dtype=np.float64))
        assert self._tab[1].shape == (self.n, len(self._tab[0])) and self._tab[2].shape == self._tab[1].shape
        check(lib().ct_ensemble_set_tp_table(self._h, len(self._tab[0]), *[a.ctypes.data_as(dp) for a in self._tab]))

    def ShardRange(self, shard):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib().ct_ensemble_shard_range(self._h, shard, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def Integrate(self, T, p, Y, want_state=True, want_stats=True, out=None):
        """T[n], p[n], Y[n, K] host arrays (pinned or pageable) -> dict of results.  The result arrays are allocated
        page-locked (one array per result, every GPU fills its own slice) unless given in `out`."""
        T = np.ascontiguousarray(T, dtype=np.float64); p = np.ascontiguousarray(p, dtype=np.float64); Y = np.ascontiguousarray(Y, dtype=np.float64)
        assert T.shape == (self.n,) and p.shape == (self.n,) and Y.shape == (self.n, self.K)
        out = dict(out or {})
        keep = []

        def buf(name, shape, dtype, want=True):
            if name in out or not want:
                return out.get(name)
            pa = PinnedArray(shape, dtype)
            keep.append(pa)
            out[name] = pa.array
            return pa.array
        state = buf("state", (self.n, self.N), np.float64, want_state)
        tau = buf("tau", (self.n,), np.float64)
        status = buf("status", (self.n,), np.int32)
        stats = buf("stats", (self.n, 8), np.int64, want_stats)
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        code = lib().ct_ensemble_integrate(self._h, ptr(T), ptr(p), ptr(Y), ptr(state), ptr(tau), ptr(status), ptr(stats))
        if code <= -20:  # call-level error (CT_ILL_INPUT, library errors); -1 ... -6 are per-reactor CVODE codes (the worst of them), results not errors
            check(code)
        shards = []
        for g in range(self.n_shards()):
            counts = np.zeros(5, dtype=np.int64); times = np.zeros(2)
            check(lib().ct_ensemble_shard_info(self._h, g, counts.ctypes.data_as(lp), times.ctypes.data_as(dp)))
            first, stride, count = self.ShardRange(g)
            shards.append(dict(first=first, stride=stride, reactors=int(counts[0]), chunks=int(counts[1]), launches=int(counts[2]),
                               failed_first=int(counts[3]), failed_final=int(counts[4]), kernel_ms=float(times[0]), wall_s=float(times[1])))
        out.update(code=code, shards=shards, _pinned=keep)
        return out


def integrate(chemistry, reactor_type, T, p, Y, rtol=1e-8, atol=1e-14, t_end=1e-3, devices=None, max_num_steps=20000,
              partition="block", chunk=0, retry=True, table=None, configure=None, want_state=True, want_stats=True, streams=None, first_pass_steps=None):
    """One call: partition, integrate on every device, gather.  Returns the Ensemble.Integrate dict."""
    e = Ensemble(chemistry, reactor_type, len(T), devices, partition, chunk)
    e.SetTolerances(rtol, atol); e.SetStopTime(t_end); e.SetMaxNumSteps(max_num_steps); e.SetRetry(retry)
    if configure is not None:
        e.SetConfigure(configure)
    if streams is not None:
        e.SetStreams(streams)
    if first_pass_steps is not None:
        e.SetFirstPassSteps(first_pass_steps)
    if table is not None:
        e.SetTable(*table)
    return e.Integrate(T, p, Y, want_state=want_state, want_stats=want_stats)


def write_hdf5(path, chemistry, reactor_type, res, T0=None, p0=None, extra=None):
    """Final states of an ensemble in the reference's HDF5 layout (groups and Name / Notation / Unit attributes of
    hdf5_writer.cpp:136-217, Base::RegisterBasicResults base.cpp:123-142), one value per reactor instead of one per
    output time: /Temperature/Temperature, /Mass fractions/<species>, /Ignition delay/Ignition delay, /Solver/Status,
    /Solver/nst, /Initial conditions/{Temperature, Pressure}; `extra` = {"/group/set": (array, notation, unit)}."""
    from .results import HDF5Writer
    energy = has_energy(reactor_type)
    w = HDF5Writer()
    if T0 is not None:
        w.add("/Initial conditions/Temperature", T0, "Initial temperature", "T0", "K")
    if p0 is not None:
        w.add("/Initial conditions/Pressure", p0, "Initial pressure", "p0", "Pa")
    st = res.get("state")
    if st is not None:
        if energy:
            w.add("/Temperature/Temperature", st[:, 0], "Temperature", "T", "K")
        Yf = st[:, 1:] if energy else st
        for k, name in enumerate(chemistry.species_names):
            w.add("/Mass fractions/" + name, Yf[:, k], "Y_" + name, "Y_" + name, "-")
    w.add("/Ignition delay/Ignition delay", res["tau"], "Ignition delay", "tau", "s")
    w.add("/Solver/Status", np.asarray(res["status"], dtype=np.float64), "CVODE return code", "status", "-")
    if res.get("stats") is not None:
        w.add("/Solver/nst", np.asarray(res["stats"])[:, 0].astype(np.float64), "nst", "nst", "-")
    for name, (arr, notation, unit) in (extra or {}).items():
        w.add(name, arr, name.rsplit("/", 1)[-1], notation, unit)
    w.write(path)
    return path
