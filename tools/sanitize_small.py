#!/usr/bin/env python
"""Small fast-flavour runs (hex, tet, quad; single engine) for compute-sanitizer:
   compute-sanitizer --tool racecheck|memcheck|synccheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dataclasses
import numpy as np
from weldformfem_b200 import cases
from weldformfem_b200.domain import Domain_d

for case in (dataclasses.replace(cases.c3_hexes(9), top_vel=-200.0), dataclasses.replace(cases.c2_tets(5), top_vel=-200.0),
             dataclasses.replace(cases.c4_axisymm_quads(20), top_vel=-50.0)):
    d = Domain_d(strict=False)
    case.apply(d)
    d.step(3)
    d.step(1)
    print(case.name, "max|v|", float(np.abs(d.get("v")).max()), "nonfinite", d.nonfinite_flag())
    d.close()
