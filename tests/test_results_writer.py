"""Results writer: layout of the reference's HDF5Writer (src/results/hdf5_writer.cpp:136-217, base.cpp:123-142,
:210-218), checked with the structural reader of tests/h5_min_reader.py."""
import numpy as np
import pytest

from h5_min_reader import H5Min


def test_single_reactor_layout(ctb, chem, tmp_path):
    from chemtools_b200.results import write_case_hdf5
    ch = chem("gri30")
    n_out = 100
    rng = np.random.default_rng(0)
    res = dict(time=np.arange(1, n_out + 1) * 1e-5, temperature=1000 + rng.random(n_out), species=ch.species_names,
               mass_fractions=rng.random((n_out, ch.n_species)))
    f = tmp_path / "test.h5"
    write_case_hdf5(str(f), res)
    h = H5Min(str(f)).root
    assert sorted(h) == ["Mass fractions", "Temperature", "Time"]          # scripts/results/demo.py:7-9
    assert np.array_equal(h["Time"]["Time"]["data"], res["time"])
    assert h["Time"]["Time"]["attrs"] == {"Name": "Time", "Notation": "t", "Unit": "s"}
    assert np.array_equal(h["Temperature"]["Temperature"]["data"], res["temperature"])
    assert h["Temperature"]["Temperature"]["attrs"] == {"Name": "Temperature", "Notation": "T", "Unit": "K"}
    mf = h["Mass fractions"]
    assert sorted(mf) == sorted(ch.species_names) and len(mf) == 53      # > 32 entries in one group
    for k, sp in enumerate(ch.species_names):
        assert np.array_equal(mf[sp]["data"], res["mass_fractions"][:, k])
        assert mf[sp]["attrs"] == {"Name": "Y_" + sp, "Notation": "Y_" + sp, "Unit": "-"}


def test_batch_layout_and_many_entries(ctb, tmp_path):
    from chemtools_b200.results import HDF5Writer, write_batch_hdf5
    rng = np.random.default_rng(1)
    n, n_out, K = 5, 7, 3
    traj = rng.random((n, n_out, K + 1))
    f = tmp_path / "batch.h5"
    write_batch_hdf5(str(f), ["A", "B(S)", "C2H6"], np.linspace(0, 1, n_out), traj, tau=rng.random(n), status=[0, 1, -1, 0, 0],
                     stats=rng.integers(0, 100, (n, 8)))
    h = H5Min(str(f)).root
    assert h["Temperature"]["Temperature"]["data"].shape == (n, n_out)
    assert np.array_equal(h["Mass fractions"]["B(S)"]["data"], traj[:, :, 2])
    assert np.array_equal(h["Solver"]["Status"]["data"], [0, 1, -1, 0, 0])
    # a group with more entries than one symbol-table node holds (2K = 64): several SNODs under one B-tree node
    w = HDF5Writer()
    for i in range(150):
        w.add("/big/d%03d" % i, np.full(2, float(i)), "n", "n", "-")
    w.add("/scalar/x", np.float64(3.5))
    w.write(str(tmp_path / "big.h5"))
    h = H5Min(str(tmp_path / "big.h5")).root
    assert len(h["big"]) == 150 and h["big"]["d149"]["data"][1] == 149.0 and h["scalar"]["x"]["data"] == 3.5
    with pytest.raises(ctb.reactors.CtError):
        w.add("/big/d000", np.zeros(2))      # duplicate object


def test_variable_length_string_attributes_and_reference_storage(ctb, tmp_path):
    """The reference writes Name / Notation / Unit as variable-length strings (H5::StrType(0, H5T_VARIABLE),
    src/results/hdf5_writer.cpp:194-208) and its datasets unlimited, chunked by 1000 and deflated at level 6 (:156-181).
    Default here: vlen strings (one global heap collection, values deduplicated); set_layout(1000, 6) = the reference's
    storage for rank-1 datasets, incl. a two-level chunk B-tree beyond 64 chunks; fixed-length strings stay selectable."""
    from chemtools_b200.results import HDF5Writer
    rng = np.random.default_rng(4)
    short, long_ = rng.random(2500), np.cumsum(rng.random(70001))
    w = HDF5Writer()
    w.set_layout(1000, 6, True)
    w.add("/Time/Time", short, "Time", "t", "s")
    w.add("/Temperature/Temperature", long_, "Temperature", "T", "K")
    w.add("/Temperature/Empty", np.zeros(0), "Empty", "e", "-")
    w.set_layout(0, 0, True)
    w.add("/Batch/T", rng.random((3, 5)), "Temperature", "T", "K")   # same strings again: one heap object each
    f = tmp_path / "ref_storage.h5"
    w.write(str(f))
    h = H5Min(str(f)).root
    t = h["Time"]["Time"]
    assert np.array_equal(t["data"], short) and t["attrs"] == {"Name": "Time", "Notation": "t", "Unit": "s"}
    assert t["storage"] == {"layout": "chunked", "chunk": 1000, "deflate": 6, "maxdims": [0xFFFFFFFFFFFFFFFF], "attr_kinds": {"vlen"}}
    T = h["Temperature"]["Temperature"]
    assert np.array_equal(T["data"], long_) and T["storage"]["layout"] == "chunked"      # 71 chunks: root + two leaves
    assert h["Temperature"]["Empty"]["data"].shape == (0,)
    b = h["Batch"]["T"]
    assert b["storage"]["layout"] == "contiguous" and b["attrs"] == {"Name": "Temperature", "Notation": "T", "Unit": "K"}
    raw = open(f, "rb").read()
    assert raw.count(b"GCOL") == 1 and raw.count(b"Temperature\0") >= 1
    # fixed-length strings on request
    w2 = HDF5Writer()
    w2.set_layout(0, 0, False)
    w2.add("/a/b", np.arange(4.0), "Name of b", "b", "m/s")
    w2.write(str(tmp_path / "fixed.h5"))
    g = H5Min(str(tmp_path / "fixed.h5")).root["a"]["b"]
    assert g["attrs"] == {"Name": "Name of b", "Notation": "b", "Unit": "m/s"} and g["storage"]["attr_kinds"] == {"fixed"}
    assert b"GCOL" not in open(tmp_path / "fixed.h5", "rb").read()
    with pytest.raises(ctb.reactors.CtError):
        HDF5Writer().set_layout(1000, 12, True)


LIBHDF5_FILE = None
try:
    import scipy.io.matlab.tests as _smt
    import os as _os
    LIBHDF5_FILE = _os.path.join(_os.path.dirname(_smt.__file__), "data", "testhdf5_7.4_GLNX86.mat")
    if not _os.path.exists(LIBHDF5_FILE):
        LIBHDF5_FILE = None
except Exception:
    pass


@pytest.mark.skipif(LIBHDF5_FILE is None, reason="scipy's MATLAB v7.3 test file (written by the real HDF5 library) not found")
def test_structures_against_a_file_written_by_libhdf5(ctb, tmp_path):
    """The only file in the image that the real HDF5 library wrote (a MATLAB v7.3 file = HDF5 behind a 512-byte user
    block): the structural reader walks its superblock, root symbol table, B-tree, SNOD, local heap and object header, and
    the f64 datatype / dataspace / string-attribute messages the emitter produces are byte-identical in form to
    libhdf5's own."""
    import struct
    from chemtools_b200.results import HDF5Writer

    class Raw(H5Min):
        def _object(self, addr):
            msgs = self._messages(addr)
            if 0x11 in [t for t, _ in msgs]:
                return self._group(*struct.unpack_from("<QQ", dict(msgs)[0x11], 0))
            return msgs
    real = Raw(LIBHDF5_FILE, 512).root
    assert list(real) == ["testdouble"]
    rm = real["testdouble"]
    rd = dict((t, m) for t, m in rm if t != 0x0C)
    assert H5Min._datatype(rd[0x03]) == ("f64", 8) and H5Min._dataspace(rd[0x01]) == [9, 1]
    w = HDF5Writer()
    w.set_layout(0, 0, False)
    w.add("/testdouble", np.zeros((9, 1)), "double")
    f = tmp_path / "mine.h5"
    w.write(str(f))
    mine = Raw(str(f)).root["testdouble"]
    md = dict((t, m) for t, m in mine if t != 0x0C)
    assert md[0x03] == rd[0x03][:len(md[0x03])]          # IEEE f64 little-endian datatype message: identical bytes
    assert md[0x01] == rd[0x01][:len(md[0x01])]          # dataspace message: identical bytes
    ra = [m for t, m in rm if t == 0x0C][0]
    ma = [m for t, m in mine if t == 0x0C][0]
    # attribute message: version, name / datatype / dataspace sizes, fixed-length null-terminated ASCII string type, scalar space
    assert ra[:2] == ma[:2] and struct.unpack_from("<HH", ra, 4) == struct.unpack_from("<HH", ma, 4) == (8, 8)
    assert ra[8:21] == b"MATLAB_class\0" and ra[24:28] == ma[16:20] == bytes([0x13, 0, 0, 0])
    assert ra[32:40] == ma[24:32]                        # scalar dataspace
