"""Independent check of a results file with the real HDF5 library, wherever h5py exists (it is not in the build image):
opens the file the way the reference's consumer does (scripts/results/results.py:14-43: h5py.File, f[group][set],
.attrs, `attrs['Notation'] + ' (' + attrs['Unit'] + ')'`) and compares every dataset with what the structural reader of
tests/h5_min_reader.py sees.

  python tools/verify_h5.py results.h5        exit 0 = verified, 3 = h5py not importable (nothing verified)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))


def main(path):
    try:
        import h5py
    except ImportError:
        print("verify_h5: h5py is not importable here; nothing verified")
        return 3
    from h5_min_reader import H5Min
    mine = H5Min(path).root
    n = 0
    with h5py.File(path, "r") as f:
        def walk(g, m, prefix):
            nonlocal n
            assert sorted(g.keys()) == sorted(m.keys()), (prefix, sorted(g.keys()), sorted(m.keys()))
            for k in g.keys():
                if isinstance(g[k], h5py.Group):
                    walk(g[k], m[k], prefix + "/" + k)
                    continue
                the_set = g[k]                                   # Result.get_set
                attributes = {a: the_set.attrs[a] for a in the_set.attrs.keys()}
                data = np.array(the_set)
                assert np.array_equal(data, m[k]["data"], equal_nan=True), prefix + "/" + k
                if attributes:
                    label = attributes["Notation"] + " (" + attributes["Unit"] + ")"   # Result.plot: needs str, not bytes
                    assert isinstance(label, str) and attributes["Name"] == m[k]["attrs"]["Name"]
                n += 1
        walk(f, mine, "")
    print("verify_h5: %d datasets of %s read by h5py %s agree with the structural reader; attributes are str" % (n, path, h5py.__version__))
    return 0


if __name__ == "__main__":
    raise SystemExit(main(sys.argv[1]))
