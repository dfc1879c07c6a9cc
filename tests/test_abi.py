"""The C-ABI library loads on a CPU-only box and exports every symbol include/wf_engine.h declares;
without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "wf_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wf_[A-Za-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from weldformfem_b200 import _lib, build
    build.build()
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) > 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.DECLARED) == names, set(names) ^ set(_lib.DECLARED)
    assert b"sm_100a" in lib.wf_version()


def test_sass_is_sm100a_fp64():
    """The cubin in the library targets sm_100a; the shipped hexa main pass is fp64 arithmetic (DFMA / DADD / DMUL) fed
    by cp.async staging (LDGSTS) and shared-memory accumulation, with no spills and no tensor-core path; the committed
    listings under profiles/ are what the current sources compile to (instruction-class histogram)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "weldformfem_b200", "libwf_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    m = re.search(r"Function (\S*wf_fast\d+hexfast\d+k_elem_main_hex_brickILi304ELi176ELi4ELb1EE\S*):\n\s*REG:(\d+) STACK:(\d+)", res)
    assert m and int(m.group(2)) <= 128 and int(m.group(3)) == 0, "128 registers, no spills: four resident CTAs"
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", m.group(1), lib], capture_output=True, text=True).stdout
    ops = re.findall(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", sass)
    assert ops.count("DFMA") > 200 and ops.count("DADD") > 200 and ops.count("LDGSTS") >= 21 and ops.count("LDS") > 60
    assert not any(o.startswith(("HMMA", "UTC", "UTMALDG")) for o in ops)
    listing = os.path.join(ROOT, "profiles", "r02_sass_E2_k_elem_main_hex_brick.txt")
    head = open(listing).read().split("\n", 3)
    assert m.group(1) in head[0] and f"REG:{m.group(2)} STACK:0" in head[1], "regenerate with tools/sass_dump.py"


def test_register_budgets_of_the_shipped_kernels():
    """Occupancy of the step kernels is decided by a register count at the edge of an allocation step; a harmless-looking
    edit can cost a resident CTA without any test failing (seen in round 2: two extra registers in the quadrilateral main
    pass, 128 -> 130 = 3 CTAs instead of 4, E2 +25 %).  The budgets the measured numbers of DESIGN.md 3 rest on:"""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "weldformfem_b200", "libwf_b200.so")
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    budgets = {   # mangled-name pattern: (registers <=, spill bytes <=)
        r"wf_fast\d+hexfast\d+k_elem_main_hex_brickILi304ELi176ELi4ELb1EE": (128, 0),     # hexa E2: 4 CTAs
        r"wf_fast\d+k_elem_vol_brickILi304EE": (40, 0),                                   # hexa E1
        r"wf_fast\d+k_node_volILi8ELi5ELb0EE": (48, 0),                                   # N1: 5 CTAs of 256
        r"wf_fast\d+k_node_updateILi3ELb0ELi4ELb1ELb1ELi5ELb0ELin1EE": (48, 8),           # hexa / tet N2
        r"wf_fast\d+k_node_updateILi3ELb0ELi4ELb1ELb1ELi5ELb1ELi3EE": (48, 16),           # partitioned: phase 3
        r"wf_fast\d+k_elem_mainILi1ELb0ELb0ELb0ELb1ELi5ELb1EE": (96, 40),                 # tet E2 (tile sums): 5 CTAs
        r"wf_fast\d+k_elem_mainILi2ELb0ELb0ELb0ELb0ELi1ELb0EE": (128, 0),                 # quad E2: 4 CTAs
    }
    for pat, (regs, stack) in budgets.items():
        m = re.search(r"Function (\S*" + pat + r"\S*):\n\s*REG:(\d+) STACK:(\d+)", res)
        assert m, pat
        assert int(m.group(2)) <= regs and int(m.group(3)) <= stack, (m.group(1), m.group(2), m.group(3))


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from weldformfem_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.wf_create(C.byref(h), 3, 8, 3, 0)
    assert rc != 0 and not h
    assert b"no CPU fallback" in lib.wf_last_error(None)
    from weldformfem_b200.domain import Domain_d, WfError
    with pytest.raises(WfError):
        Domain_d().box((0, 0, 0), (1, 1, 1), 0.25)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "weldformfem_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "wf_oracle" not in txt, f


def test_null_engine_handle_is_an_error_not_a_crash():
    """Domain_d.set_material / set_stab before the mesh passed NULL to the C ABI, which dereferenced it."""
    from weldformfem_b200 import _lib
    from weldformfem_b200.domain import Domain_d, WfError
    lib = _lib.load()
    mat = _lib.wf_material()
    assert lib.wf_set_material(None, C.byref(mat)) != 0
    assert b"null engine handle" in lib.wf_last_error(None)
    assert lib.wf_step(None, 1) != 0 and lib.wf_get_array(None, b"x", None, 0) != 0 and lib.wf_SearchExtNodes(None) != 0
    with pytest.raises(WfError, match="null engine handle"):
        Domain_d().set_material(1e9, 0.3, 1000.0)
    with pytest.raises(WfError, match="null engine handle"):
        Domain_d().set_stab(hg_visc=0.1)


def test_bad_arguments():
    from weldformfem_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.wf_create(C.byref(h), 3, 5, 3, 0) != 0
    assert b"unsupported element" in lib.wf_last_error(None)
    assert lib.wf_create(C.byref(h), 2, 4, 3, 0) != 0
