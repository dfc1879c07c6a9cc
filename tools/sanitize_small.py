#!/usr/bin/env python
"""Small fast-flavour runs for compute-sanitizer (hex in brick order incl. ragged bricks, tet, quad; open stepping with
per-step BC uploads; a 2-rank partitioned step with the folded halo exchange):
   compute-sanitizer --tool racecheck|memcheck|synccheck python tools/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dataclasses
import numpy as np
from weldformfem_b200 import cases
from weldformfem_b200.domain import Domain_d

for case in (dataclasses.replace(cases.c3_hexes(13), top_vel=-200.0), dataclasses.replace(cases.c3_hexes(8), top_vel=-200.0),
             dataclasses.replace(cases.c2_tets(5), top_vel=-200.0), dataclasses.replace(cases.c4_axisymm_quads(20), top_vel=-50.0)):
    d = Domain_d(strict=False)
    case.apply(d)
    d.step(3)
    d.step(1)
    _, dims, vals = case.bc_arrays()
    dl = case.dim - 1
    for i in range(3):                      # open stepping: BC upload on the copy stream + patch kernel + monitor
        d.set_bc_values(dl, vals[dims == dl] * (1.0 + 0.1 * i))
        d.step_open(1)
        d.monitor_async()
        d.monitor_wait()
    d.step_close()
    print(case.name, "max|v|", float(np.abs(d.get("v")).max()), "nonfinite", d.nonfinite_flag())
    d.close()

from weldformfem_b200.distributed import LocalCluster
case = dataclasses.replace(cases.c3_hexes(8), n=(6, 5, 9), top_vel=-200.0)
cl = LocalCluster(2, [0, 0])
case.apply(cl)
cl.step(4)
print("2 ranks", float(np.abs(cl.get("v")).max()), cl.nonfinite_flag())
cl.close()
