"""Transcribe the numbers of /root/reference/validation/*.txt that pin the hexa hourglass + step sequence into
tests/golden/validation_pins.json (the reference tree is absent on the GPU box and in CI).

    python tests/golden/make_validation_golden.py

Blocks taken (file:lines):
  1elem_3d_red_int_f_0.06.txt:5-42    "C++" block, hourglass 0.06: DISPLACEMENTS / VELOCITIES / ACCEL / FORCES, 8 nodes
  1elem_3d_red_int_f_0.06.txt:73-108  "NO HG" block, same four tables
  1elem_3d_red_int_f_0.06.txt:44-68   F90 block after 126 steps (Disp / Vel)
  several dts.txt:1-9, 30-38          F90 blocks after 10 and 100 steps (Disp)
  4elem_red_0.06_f90_1step.txt        F90, 2x2x2 elements, one step of dt = 2e-6: sigma / shear_stress of a loaded element
  1step_red_int_cube3D_hf_c_0.06.txt:6-9  initial dHdx*detJ matrix
  4_el_hg_1e-3.txt:5-58               F90, 2x2x2 elements, hourglass 0.06, 501 steps of dt = 2e-6 (t = 1.002e-3): Disp / Vel, 27 nodes
  4_el_NO_hg_1e-3.txt:3-29            the same run without hourglass forces: Disp
  2step_1elem_red_no_hg_DIV.txt       F90, one element, state after the SECOND step of dt = 0.8e-5: Disp without hourglass (:12-19),
                                      dHxy x detJ rows (:39-40) and the strain-rate tensor (:46-48) of that run; Disp / Vel / Acc
                                      with hourglass 0.06 (:52-77); "C++ with hourglass" block (:93-141): the four tables
                                      after two steps and the hourglass force of every element node (6 decimals = 10 digits)
  cxx/2time_step.txt:61,79-81         C++ log: sound speed CS_0 and the Chung-Hulbert alpha / beta / gamma the solver printed;
                                      :377-433 corrected accelerations (6 decimals = 11 digits) and velocities of the FIRST step
  1step_red_int_cube3D_hf_c_0.06.txt:82-  F90, first step: element stress, pressure, deviatoric stress, global forces, Disp / Vel / Acc
"""
import json
import os
import re

REF = "/root/reference/validation"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "validation_pins.json")


def table(lines, start, n=8):
    return [[float(t) for t in lines[start + i].split()] for i in range(n)]


def f90_nodes(text, kind, n=8):
    out = {}
    for m in re.finditer(r"nod\s+(\d+)\s+%s\s+(\S+)\s+(\S+)\s+(\S+)" % kind, text):
        out.setdefault(int(m.group(1)) - 1, [float(m.group(i)) for i in (2, 3, 4)])
    return [out[i] for i in range(n)]


def main():
    pins = {}
    L = open(os.path.join(REF, "1elem_3d_red_int_f_0.06.txt")).read().split("\n")
    idx = [i for i, s in enumerate(L) if s.strip() == "DISPLACEMENTS"]
    for name, i0 in (("cxx_hg_0.06", idx[0]), ("cxx_no_hg", idx[1])):
        blk = {}
        j = i0
        for key in ("DISPLACEMENTS", "VELOCITIES", "ACCEL", "FORCES"):
            while L[j].strip() != key:
                j += 1
            blk[key] = table(L, j + 1)
        pins[name] = blk
    txt = "\n".join(L)
    f90 = txt[txt.index("F90"):txt.index("NO HG")]
    pins["f90_126_steps"] = {"Disp": f90_nodes(f90, "Disp"), "Vel": f90_nodes(f90, "Vel")}
    sd = open(os.path.join(REF, "several dts.txt")).read()
    pins["f90_10_steps"] = {"Disp": f90_nodes(sd[:sd.index("C++")], "Disp")}
    rest = sd[sd.index("100 dt"):]
    pins["f90_100_steps"] = {"Disp": f90_nodes(rest[:rest.index("Correction")], "Disp")}
    one = open(os.path.join(REF, "4elem_red_0.06_f90_1step.txt")).read()
    m = re.search(r"shear_stress\s+(2112820\S+)(?:\s+\S+){7}\s+(-4225640\S+)\s+elem\s+7 , sigma\s+(\S+)(?:\s+\S+){7}\s+(\S+)", one)
    pins["f90_8elem_1step"] = {"tau_xx": float(m.group(1)), "tau_zz": float(m.group(2)), "sigma_xx": float(m.group(3)),
                               "sigma_zz": float(m.group(4))}
    dh = open(os.path.join(REF, "1step_red_int_cube3D_hf_c_0.06.txt")).read().split("\n")
    k = [i for i, s in enumerate(dh) if "INITIAL DERIVATIVE MATRIX" in s][0]
    pins["dHdx_detJ"] = table(dh, k + 1, 3)
    f8 = open(os.path.join(REF, "4_el_hg_1e-3.txt")).read()
    f8 = f8[:f8.index("C++")]
    pins["f90_8elem_501_steps"] = {"Disp": f90_nodes(f8, "Disp", 27), "Vel": f90_nodes(f8, "Vel", 27)}
    two = open(os.path.join(REF, "2step_1elem_red_no_hg_DIV.txt")).read()
    nohg, hg = two.split("WITH HOURGLASS 0.06")[:2]
    dh2 = [[float(t) for t in m.group(1).split()] for m in re.finditer(r"dHxy x detJ\s+(.*)", nohg)][:2]
    sr = re.search(r"strain rate\s*\n(.*)\n(.*)\n(.*)\n", nohg)
    pins["f90_1elem_2_steps"] = {
        "Disp_no_hg": f90_nodes(nohg, "Disp"), "dHx_detJ": dh2[0], "dHy_detJ": dh2[1],
        "strain_rate": [[float(t) for t in sr.group(i).split()] for i in (1, 2, 3)],
        "Disp": f90_nodes(hg, "Disp"), "Vel": f90_nodes(hg, "Vel"), "Acc": f90_nodes(hg, "Acc")}
    T = two.split("\n")
    c0 = [i for i, q in enumerate(T) if "C++ with hourglass" in q][0]
    blk = {}
    for key in ("DISPLACEMENTS", "VELOCITIES", "ACCEL", "FORCES"):
        j = [i for i in range(c0, len(T)) if T[i].strip() == key][0]
        blk[key] = table(T, j + 1)
    blk["HG_FORCES"] = [[float(t) for t in q.split(":")[1].split()] for q in T[c0:] if q.startswith("hg forces el 0")][:8]
    pins["cxx_hg_0.06_2_steps"] = blk
    lg = open(os.path.join(REF, "cxx", "2time_step.txt")).read()
    pins["cxx_log_constants"] = {k: float(re.search(r"^%s:? (\S+)" % k, lg, re.M).group(1)) for k in ("CS_0", "alpha", "beta", "gamma")}
    acc = [[float(m.group(i)) for i in (1, 2, 3)] for m in re.finditer(r"Corr Acc (\S+) (\S+) (\S+)", lg)][:8]
    vel = [[float(m.group(i)) for i in (1, 2, 3)] for m in re.finditer(r"Node \d Corr Vel (\S+) (\S+) (\S+)", lg)][:8]
    pins["cxx_log_first_step"] = {"Acc": acc, "Vel": vel}
    one_s = "\n".join(dh)
    one_s = one_s[one_s.index("main loop, CHUNG HULBERT", one_s.index("WITH HOURGLASS")):]
    sg = re.search(r"Element stresses.*\n\s*(\S+)\s+\S+\s+\S+\s*\n\s*\S+\s+(\S+)\s+\S+\s*\n\s*\S+\s+\S+\s+(\S+)", one_s)
    sh = re.search(r"Element shear stresses\s*\n\s*(\S+)\s+\S+\s+\S+\s*\n\s*\S+\s+(\S+)\s+\S+\s*\n\s*\S+\s+\S+\s+(\S+)", one_s)
    gf = re.search(r"Global forces\s*\n((?:\s*\S+\s+\S+\s+\S+\s*\n){8})", one_s)
    pins["f90_1elem_first_step"] = {
        "sigma_diag": [float(sg.group(i)) for i in (1, 2, 3)], "tau_diag": [float(sh.group(i)) for i in (1, 2, 3)],
        "pressure": float(re.search(r"Element pressure\s+(\S+)", one_s).group(1)),
        "forces": [[float(t) for t in q.split()] for q in gf.group(1).strip().split("\n")],
        "Disp": f90_nodes(one_s, "Disp"), "Vel": f90_nodes(one_s, "Vel"), "Acc": f90_nodes(one_s, "Acc")}
    n8 = open(os.path.join(REF, "4_el_NO_hg_1e-3.txt")).read()
    pins["f90_8elem_501_steps_no_hg"] = {"Disp": f90_nodes(n8[:n8.index("C++")], "Disp", 27)}
    json.dump(pins, open(OUT, "w"), indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
