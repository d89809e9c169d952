"""Collect the reference's own golden vectors for the reactor path into
tests/golden/reference_goldens.json (run here; needs /root/reference).

Sources (paths under /root/reference):
  tests/data/sundials_cvode/direct_dense_cvRoberts_dns.dat       36 values, CVODE Roberts problem
  tests/src/chemistry.cpp:147-158, :250-322                       equilibrium thermo KAT, mechanism structure KAT
  doc/presentation/figures/output.png (transcribed in SURVEY.md §6) slide trajectory, 7 significant digits
"""
import json
import os

import numpy as np

REF = "/root/reference"
out = {}
rob = np.loadtxt(os.path.join(REF, "tests/data/sundials_cvode/direct_dense_cvRoberts_dns.dat"))
out["roberts"] = {
    "source": "tests/data/sundials_cvode/direct_dense_cvRoberts_dns.dat; set-up tests/src/sundials_cvode_direct_dense.cpp:91-195",
    "rtol": 1e-4, "atol": [1e-8, 1e-14, 1e-6], "y0": [1.0, 0.0, 0.0],
    "t_out": [0.4 * 10 ** k for k in range(12)], "y": rob.reshape(12, 3).tolist(),
    "tolerance": 100 * float(np.finfo(np.float32).eps),
}
out["equilibrium"] = {
    "source": "tests/src/chemistry.cpp:147-152 (HP equilibrium of CH4:1/O2:1/N2:3.76 from 500 K, 2 atm, GRI-3.0)",
    "X0": {"CH4": 1.0, "O2": 1.0, "N2": 3.76}, "T0": 500.0, "p": 202650.0,
    "T": 1706.732, "rho": 0.3239982, "h_mole": -5.626613e6, "s_mole": 2.433695e5, "cp_mole": 3.734765e4,
    "tolerance": 100 * float(np.finfo(np.float32).eps),
}
out["structure"] = {
    "source": "tests/src/chemistry.cpp:250-322",
    "gri12": {"n_species": 32, "n_reactions": 177, "species_index": {"CH4": 13}, "reaction": {"163": "C2H4 + CH3 <=> C2H3 + CH4"}},
    "gri211": {"n_species": 49, "n_reactions": 279, "species_index": {"HO2": 6}, "reaction": {"65": "CH3O + H <=> CH3 + OH"}},
    "gri30": {"n_species": 53, "n_reactions": 325, "species_index": {}, "reaction": {}},
}
out["slide_trajectory"] = {
    "source": "doc/presentation/figures/output.png (HDFView of /Temperature/Temperature), scenario tests/src/reactor.cpp:50-64,101",
    "reactor": "Const_Vol", "mechanism": "gri30", "X": {"H2": 2.0, "O2": 1.0, "N2": 4.0}, "T0": 1001.0, "p0": 101325.0,
    "rtol": 1e-9, "atol": 1e-15, "t_end": 1e-3, "n_out": 100, "max_steps": 20000, "ignition_threshold": 1756.0,
    "ignition_time": 0.31e-3,
    "T": {"13": 1001.006, "14": 1001.010, "15": 1001.016, "16": 1001.025, "17": 1001.039, "18": 1001.062, "19": 1001.099,
          "20": 1001.161, "21": 1001.266, "22": 1001.456, "23": 1001.818, "24": 1002.559, "25": 1004.186, "26": 1007.904,
          "27": 1016.563, "28": 1038.311, "29": 1110.314, "30": 1851.254, "31": 2426.991, "32": 2655.308, "33": 2771.150},
}
dst = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_goldens.json")
with open(dst, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", dst)
