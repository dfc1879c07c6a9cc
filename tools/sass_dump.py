#!/usr/bin/env python
"""SASS listings of the four kernels of the shipped hexa step (profiles/r02_sass_*.txt): cuobjdump -sass of the in-tree
library, one file per kernel, preceded by the instruction-class histogram.  No GPU needed.
usage: tools/sass_dump.py [out_dir]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "weldformfem_b200", "libwf_b200.so")
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles")
KERNELS = {"E1_k_elem_vol_brick": r"wf_fast\d+k_elem_vol_brickILi304EE", "N1_k_node_vol": r"wf_fast\d+k_node_volILi8ELi5ELb0EE",
           "E2_k_elem_main_hex_brick": r"wf_fast\d+hexfast\d+k_elem_main_hex_brickILi304ELi176ELi4ELb1EE",
           "N2_k_node_update": r"wf_fast\d+k_node_updateILi3ELb0ELi4ELb1ELb1ELi5ELb0ELin1EE"}
names = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
for tag, pat in KERNELS.items():
    m = re.search(r"Function (\S*" + pat + r"\S*):\n\s*(REG:\S+ STACK:\S+ SHARED:\S+ LOCAL:\S+)", names)
    assert m, tag
    fn, res = m.group(1), m.group(2)
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, LIB], capture_output=True, text=True).stdout
    ins = [l for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
    hist = collections.Counter()
    for l in ins:
        t = l.split("*/", 1)[1].split()
        op = t[1] if t[0].startswith("@") else t[0]
        hist[op.split(".")[0].rstrip(";")] += 1
    with open(os.path.join(OUT, f"r02_sass_{tag}.txt"), "w") as f:
        f.write(f"# {fn}\n# {res}\n# {len(ins)} instructions; classes: " + ", ".join(f"{k} {v}" for k, v in hist.most_common(24)) + "\n")
        f.write("# fp64 arithmetic: DFMA/DADD/DMUL; LDGSTS = cp.async staging; no tensor-core (HMMA/UTC*MMA) or TMA (UTMALDG) "
                "instructions: nothing on this path is a dense contraction, and the staging is an index gather of 8-byte words\n")
        f.write("\n".join(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in ins) + "\n")
    print(tag, res, len(ins), dict(hist.most_common(6)))
