"""configs[1] on its REAL meshes: the reference's cylinder of tetrahedra (examples/input/tetra_cyl.k, node valence
3..38) and the unstructured hexahedral cylinder (cyl_hex.k), loaded the way Domain_d::CreateFromLSDyna does
(/root/reference/src/common/Domain_d.C:1647-1699: coordinates + connectivity + setNodElem), compressed between a clamped
bottom and a moving top until a good part of the mesh is plastic.  The meshes are committed as parsed arrays
(tests/golden/make_mesh_golden.py).  Checkers: the reference compiled here (oracle/_ref, tetrahedra) and the plain-C
port pinned to it (hexahedra with the F90 hourglass).

What these add over the structured boxes: irregular node valence (SELL padding, incidence-table pitch of the pulled
tile sums), element order unrelated to geometry (the Morton reordering really permutes), and hexahedra whose force
tiles are checked for corner conflicts on a mesh that is not a lattice."""
import os

import numpy as np
import pytest

from parity_util import STATE, compare, relerr

HERE = os.path.dirname(os.path.abspath(__file__))
E_, NU, RHO, SY0, KH, MH = 68.9e9, 0.3, 2700.0, 190.4e6, 386.796e6, 0.154   # cases.Case defaults (Hollomon aluminium)
HEIGHT = 0.03


def load_mesh(name):
    g = np.load(os.path.join(HERE, "golden", "meshes", name + ".npz"))
    return g["x"], g["elnod"]


def min_edge(x, el):
    k = el.shape[1]
    pairs = [(a, b) for a in range(k) for b in range(a + 1, k)]
    if k == 8:   # the twelve edges of the hexahedron
        pairs = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    return min(float(np.linalg.norm(x[el[:, a]] - x[el[:, b]], axis=1).min()) for a, b in pairs)


def setup(dom, x, el, hexa_hg, vtop, dt, press=0, init=True):
    from weldformfem_b200.cases import HOLLOMON
    k = el.shape[1]
    dom.set_mesh(3, k, x.ravel(), el.ravel())
    dom.set_material(E_, NU, RHO, HOLLOMON, SY0, KH, MH)
    dom.set_stab()
    dom.set_options(press, 0.0, 0.0, hexa_hg)
    bottom = np.nonzero(x[:, 2] <= 1e-9)[0]
    top = np.nonzero(x[:, 2] >= HEIGHT - 1e-9)[0]
    assert bottom.size > 20 and top.size > 20
    trip = [(int(n), d, 0.0) for n in bottom for d in range(3)] + [(int(n), d, (vtop if d == 2 else 0.0)) for n in top for d in range(3)]
    if hasattr(dom, "add_bcs"):
        dom.add_bcs(np.array([t[0] for t in trip], np.int32), np.array([t[1] for t in trip], np.int32),
                    np.array([t[2] for t in trip]))
    else:
        for n, d, v in trip:
            dom.add_bc(n, d, v)
    dom.allocate_bcs()
    if init:
        dom.init(dt)
    return dom


def case_params(name):
    x, el = load_mesh(name)
    dt = 0.1 * min_edge(x, el) / np.sqrt(E_ / (3 * (1 - 2 * NU)) / RHO)
    return x, el, dt


# ---- CPU: the meshes are what the reference front-end reads, and the port is pinned on them -------------------------
def test_mesh_fixtures_are_the_reference_files(oracle_ref):
    """main.C reading examples/input/Compression_tetra.json (File block -> CreateFromLSDyna) ends with the same
    coordinates and connectivity as the committed arrays (skipped where the reference tree is absent)."""
    src = "/root/reference/examples/input"
    if not os.path.isdir(src):
        pytest.skip("reference tree not present")
    import shutil, tempfile
    with tempfile.TemporaryDirectory() as td:
        for f in ("Compression_tetra.json", "tetra_cyl.k"):
            shutil.copy(os.path.join(src, f), td)
            os.chmod(os.path.join(td, f), 0o644)
        d, dt, end_t = oracle_ref.from_deck(os.path.join(td, "Compression_tetra.json"))
    x, el = load_mesh("tetra_cyl")
    assert np.array_equal(d.get("x").reshape(-1, 3), x)
    assert np.array_equal(d.get("m_elnod").reshape(-1, 4), el)
    assert end_t == 0.02


@pytest.mark.parametrize("name", ["tetra_cyl", "cyl_hex"])
def test_port_matches_compiled_reference_on_real_meshes(name, oracle_port, oracle_ref):
    x, el, dt = case_params(name)
    oracle_ref.set_threads(1)
    hg = 0.06 if name == "cyl_hex" else 0.0
    a = setup(oracle_port(), x, el, hg, -40.0, dt)
    b = setup(oracle_ref(), x, el, hg, -40.0, dt)
    a.step(30)
    b.step(30)
    for nm in STATE + ["m_nodel", "m_nodel_loc", "m_nodel_offset", "m_nodel_count"]:
        assert np.array_equal(a.get(nm), b.get(nm)), nm


# ---- GPU -----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,press", [("tetra_cyl", 0), ("tetra_cyl", 3), ("cyl_hex", 0)])
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_cylinder_meshes_match_reference(name, press, strict, oracle_port, oracle_ref):
    from weldformfem_b200.domain import Domain_d
    x, el, dt = case_params(name)
    hexes = name == "cyl_hex"
    checker = oracle_port if hexes else oracle_ref    # the hexa hourglass of the harness is not upstream code: use the pinned port
    if not hexes:
        oracle_ref.set_threads(0)
    ref = setup(checker(), x, el, 0.06 if hexes else 0.0, -40.0, dt, press)
    # elem_order=1: the Morton reordering really permutes these meshes (default for hexahedra, forced for the tets)
    eng = setup(Domain_d(strict=strict, elem_order=1), x, el, 0.06 if hexes else 0.0, -40.0, dt, press)
    for nm in ("m_elnod", "m_nodel", "m_nodel_loc", "m_nodel_offset", "m_nodel_count"):
        assert np.array_equal(eng.get(nm), ref.get(nm)), nm
    perm = eng.get("elem_perm")
    assert not np.array_equal(perm, np.arange(perm.size))          # the engine really reorders this mesh
    ref.step(1)
    eng.step(1)
    compare(eng, ref, STATE, 1e-12 if strict else 1e-10, f"{name} 1 step")
    ref.step(199)
    eng.step(199)
    plastic = float((ref.get("pl_strain") > 0).mean())
    assert plastic > 0.15, plastic
    compare(eng, ref, STATE, 1e-6, f"{name} 200 steps")
    assert not eng.nonfinite_flag()
    eng.close()
