"""Compile the reference's Cantera .cti mechanisms into the ".ctm" text form
shipped under chemtools_b200/data/ (the GPU box has no /root/reference).

Run here (needs /root/reference):  python tools/make_ctm.py
The C++ CTI parser of the product does the work (ct_mech_load + ct_mech_save_ctm);
tests/test_mechanism.py checks the result against the oracle's independent
Python front end (oracle/cti_frontend.py -> oracle/mech/*.json).
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import chemtools_b200 as ctb  # noqa: E402

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(__file__), "..", "chemtools_b200", "data")
os.makedirs(OUT, exist_ok=True)
for name in ("gri30", "gri211", "gri12"):
    tk = ctb.ThermoKinetics(os.path.join(REF, name + ".cti"))
    tk.save_ctm(os.path.join(OUT, name + ".ctm"))
    print(name, tk.n_species, "species", tk.n_reactions, "reactions")
