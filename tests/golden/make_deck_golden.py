"""Generate tests/golden/deck_*.npz: each deck of tests/golden/decks/ is set up by the reference's OWN front-end
(src/explicit/main.C, compiled unmodified into oracle/_ref/libwf_ref.so — see the end of oracle/ref_harness.cpp) and
then stepped by the harness (member-by-member sequence of SolveChungHulbert, fixed time step, one OpenMP thread).

Run in the container that has /root/reference:   python tests/golden/make_deck_golden.py
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refdrv  # noqa: E402

# deck -> steps.  box_pstri is absent: main.C's unconditional SearchExtNodes() (main.C:650) crashes on 2D triangles.
DECKS = {"box_axiquad": 40, "box_psquad": 40, "file_tet_contact": 60, "file_tet_zones": 12, "box_axiquad_contact": 60}
STATE = "x v a u prev_a m_fi m_mdiag vol p pl_strain sigma_y m_sigma m_tau".split()
CONTACT = "contforce ut_prev node_area trimesh.node".split()
THERMAL = "T m_q_plheat".split()


def run(name, steps):
    d, dt, end_t = refdrv.RefDomain.from_deck(os.path.join(HERE, "decks", name + ".json"))
    info = d.info()
    out = {"info": np.array([info[k] for k in "dim nodxelem n_nodes n_elems bcx bcy bcz domtype".split()]),
           "dt": np.array([dt]), "end_t": np.array([end_t]), "steps": np.array([steps]),
           "x0": d.get("x"), "m_elnod": d.get("m_elnod")}
    tm = d.trimesh_counts()
    out["trimesh"] = np.array([tm["dimension"], tm["nodecount"], tm["elemcount"]])
    names = list(STATE)
    if tm["nodecount"]:
        names += CONTACT
        out["ext_nodes"] = d.get("ext_nodes")
    d.init(dt)
    thermal = bool(np.any(d.get("T") != 0.0))
    if thermal:
        names += THERMAL
    d.step(steps)
    for nm in names:
        out["sN_" + nm] = d.get(nm)
    return out


def main():
    refdrv.build("ref")
    refdrv.RefDomain.set_threads(1)
    for name, steps in DECKS.items():
        out = run(name, steps)
        path = os.path.join(HERE, "deck_" + name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path), "bytes; max|v|", float(np.abs(out["sN_v"]).max()), "plastic frac",
              float((out["sN_pl_strain"] > 0).mean()), "keys", len(out))
    for f in glob.glob(os.path.join(HERE, "decks", "*.out")):   # main.C's log file next to the deck
        os.remove(f)


if __name__ == "__main__":
    main()
