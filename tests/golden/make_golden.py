"""Generate tests/golden/*.npz from the UNMODIFIED reference build (oracle/_ref/libwf_ref.so).

Run in the container that has /root/reference:   python tests/golden/make_golden.py
Each file holds the reference-layout arrays after 1 step and after N steps of one small case
(single OpenMP thread: the reference's 2D-quad hourglass has a data race, Mechanical.C:1848).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from cases_golden import CONTACT_ARRAYS, FLOAT_ARRAYS, GOLDEN, INT_ARRAYS, THERMAL_ARRAYS  # noqa: E402
from oracle import refdrv  # noqa: E402


def main():
    refdrv.build("ref")
    refdrv.RefDomain.set_threads(1)
    only_missing = "--missing" in sys.argv
    for name, (case, steps) in GOLDEN.items():
        if only_missing and os.path.exists(os.path.join(HERE, name + ".npz")):
            continue
        d = refdrv.RefDomain()
        case.apply(d)
        out = {nm: d.get(nm) for nm in INT_ARRAYS}
        out["x0"] = d.get("x")
        extra = (CONTACT_ARRAYS if case.contact is not None else []) + (THERMAL_ARRAYS if case.thermal is not None else [])
        for nm in extra:
            out[f"s0_{nm}"] = d.get(nm)
        d.step(1)
        for nm in FLOAT_ARRAYS:
            out[f"s1_{nm}"] = d.get(nm)
        d.step(steps - 1)
        for nm in FLOAT_ARRAYS + list(extra):
            out[f"sN_{nm}"] = d.get(nm)
        if case.dim == 2 and not case.tritet:
            out["sN_m_hg_q"] = d.get("m_hg_q")[: 2 * case.n_elems]
        # diagnostics of SURVEY §8f-1: calcNodalPressureFromElemental (Mechanical.C:1187), calcMinEdgeLength (Domain_d.C:2224)
        d.call("calcNodalPressureFromElemental")
        out["sN_p_node"] = d.get("p_node")
        if not (case.dim == 2 and case.tritet):
            d.call("calcMinEdgeLength")
            c = d.consts()
            out["sN_min_edge"] = np.array([c["min_length"], c["min_height"]])
            out["sN_m_elem_length"] = d.get("m_elem_length")
        out["steps"] = np.array([steps])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path), "bytes", "plastic frac", float((out["sN_pl_strain"] > 0).mean()))


if __name__ == "__main__":
    main()
