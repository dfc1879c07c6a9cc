"""The deck front-end: host/wf_weldform + host/wf_deck.hpp run a WeldFormFEM JSON deck (+ LS-Dyna
`.k` mesh) on the engine.  The checker is the reference's OWN front-end: src/explicit/main.C compiled unmodified into
oracle/_ref/libwf_ref.so sets the domain up from the same deck (tests/golden/decks/), the harness steps it, and the
result is committed as tests/golden/deck_*.npz (tests/golden/make_deck_golden.py).

CPU tests: the fixtures still match the reference front-end (when oracle/_ref is built) and `wf_weldform --parse-only`
finds the same mesh / BC / rigid-surface counts.  GPU tests: the state after N steps matches the fixture."""
import json
import os
import subprocess

import numpy as np
import pytest

from parity_util import relerr
from test_host_cpp import _read_dump, host_bins  # noqa: F401  (fixture)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECKS = os.path.join(ROOT, "tests", "golden", "decks")
PINNED = ["box_axiquad", "box_psquad", "file_tet_contact", "file_tet_zones", "box_axiquad_contact"]


def _gold(name):
    return np.load(os.path.join(ROOT, "tests", "golden", f"deck_{name}.npz"))


def _summary(host_bins, deck, *extra):
    r = subprocess.run([os.path.join(host_bins, "wf_weldform"), os.path.join(DECKS, deck + ".json"), *extra],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("name", PINNED)
def test_deck_fixture_is_what_the_reference_front_end_produces(oracle_ref, name):
    """Pins the fixture: re-run main.C on the deck and compare bit for bit (skipped where oracle/_ref is not built)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_deck_golden
    oracle_ref.set_threads(1)
    gold = _gold(name)
    try:
        got = make_deck_golden.run(name, int(gold["steps"][0]))
    finally:
        for f in os.listdir(DECKS):
            if f.endswith(".out"):
                os.remove(os.path.join(DECKS, f))
    assert sorted(got) == sorted(gold.files)
    for k in gold.files:
        assert np.array_equal(np.asarray(got[k]), gold[k]), k


@pytest.mark.parametrize("name", PINNED)
def test_deck_parse_only_counts_match_reference(host_bins, name):
    gold = _gold(name)
    s = _summary(host_bins, name, "--parse-only")
    dim, k, nn, ne, bcx, bcy, bcz, _ = [int(v) for v in gold["info"]]
    assert (s["dim"], s["nodxelem"], s["nodes"], s["elements"]) == (dim, k, nn, ne)
    assert s["bc_count"][:dim] == [bcx, bcy, bcz][:dim]
    assert s["rigid_facets"] == int(gold["trimesh"][2])
    assert s["contact"] == bool(gold["trimesh"][1] > 0)
    assert s["end_time"] == float(gold["end_t"][0])


def test_deck_k_reader_maps_sparse_ids_and_truncates_tets(host_bins):
    """tet_block.k has ids 7, 10, 13, ... and 8-slot solids padded with the last node; hex_block.k is a plain hex file."""
    s = _summary(host_bins, "file_tet_zones", "--parse-only")
    assert (s["nodxelem"], s["nodes"], s["elements"]) == (4, 5 * 5 * 7, 5 * 4 * 4 * 6)
    deck = {"Configuration": {"simTime": 1e-5}, "Materials": [{"type": "Bilinear", "const": [1e9], "density0": 7850.0,
            "youngsModulus": 2e11, "poissonsRatio": 0.3, "yieldStress0": 3e8}],
            "DomainBlocks": [{"type": "File", "fileName": "hex_block.k"}], "BoundaryConditions": []}
    path = os.path.join(DECKS, "_tmp_hex.json")
    with open(path, "w") as f:
        json.dump(deck, f)
    try:
        s = _summary(host_bins, "_tmp_hex", "--parse-only")
    finally:
        os.remove(path)
    assert (s["nodxelem"], s["nodes"], s["elements"]) == (8, 5 * 4 * 6, 4 * 3 * 5)


def test_deck_errors_are_loud(host_bins, tmp_path):
    bad = tmp_path / "bad.json"
    bad.write_text('{"Configuration": {}, "Materials": [{"type": "Hollomon"}], "DomainBlocks": [{"type": "Sphere"}]}')
    r = subprocess.run([os.path.join(host_bins, "wf_weldform"), str(bad), "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 1 and "File or Box" in r.stderr
    bad.write_text('{"Configuration": {')
    r = subprocess.run([os.path.join(host_bins, "wf_weldform"), str(bad), "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.strip()


def test_deck_options_the_engine_does_not_implement_are_refused(host_bins, tmp_path):
    """pressAlgorithm outside {0, 1} (the reference would silently skip the pressure update,
    Solver_explicit.C:733-743) and devElastic = false (calcElemPressureRigid) must not run a different algorithm
    silently — both front-ends refuse them; an explicit "pspg_scale" key is honoured (main.C leaves it indeterminate)."""
    from weldformfem_b200 import deck
    base = json.load(open(os.path.join(DECKS, "box_psquad.json")))
    for key, val, msg in (("pressAlgorithm", 2, "pressAlgorithm"), ("devElastic", False, "devElastic")):
        j = json.loads(json.dumps(base))
        j["Configuration"][key] = val
        path = tmp_path / f"bad_{key}.json"
        path.write_text(json.dumps(j))
        r = subprocess.run([os.path.join(host_bins, "wf_weldform"), str(path), "--parse-only"], capture_output=True, text=True)
        assert r.returncode == 1 and msg in r.stderr, r.stderr
        with pytest.raises(ValueError, match=msg):
            deck.load(str(path))
    j = json.loads(json.dumps(base))
    j["Stabilization"] = {"hg_visc": 0.1, "hg_stiff": 0.1, "pspg_scale": 0.25}
    path = tmp_path / "pspg.json"
    path.write_text(json.dumps(j))
    assert deck.load(str(path)).stab["pspg_scale"] == 0.25
    assert deck.load(os.path.join(DECKS, "box_axiquad.json")).stab["pspg_scale"] == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", PINNED)
def test_deck_run_matches_reference_front_end(host_bins, tmp_path, name):
    gold = _gold(name)
    steps = int(gold["steps"][0])
    dump = str(tmp_path / "d.bin")
    s = _summary(host_bins, name, "--steps", str(steps), "--dump", dump, "--strict")
    assert s["steps"] == steps
    assert abs(s["dt"] - float(gold["dt"][0])) <= 1e-15 * float(gold["dt"][0]) * 4, (s["dt"], gold["dt"])
    got = _read_dump(dump)
    worst = {}
    for key in gold.files:
        if not key.startswith("sN_"):
            continue
        nm = key[3:]
        worst[nm] = relerr(got[nm], gold[key])
    bad = {k: v for k, v in worst.items() if not v <= 1e-8}
    assert not bad, (name, bad, worst)


def test_deck_k_reader_free_format_and_two_line_solids(host_bins, tmp_path):
    """Comma-separated cards, CRLF line ends, comment lines, and the '*ELEMENT_SOLID' form that puts eid/pid on one
    line and the nodes on the next (both appear in LS-PrePost output)."""
    k = tmp_path / "m.k"
    k.write_bytes(("*KEYWORD\r\n*NODE\r\n$ comment\r\n"
                   "10,0.0,0.0,0.0\r\n20,1.0,0.0,0.0\r\n30,0.0,1.0,0.0\r\n40,0.0,0.0,1.0\r\n50,1.0,1.0,1.0\r\n"
                   "*ELEMENT_SOLID\r\n"
                   "1,1\r\n10,20,30,40,40,40,40,40\r\n"
                   "2,1\r\n20,30,40,50,50,50,50,50\r\n*END\r\n").encode())
    deck = tmp_path / "d.json"
    deck.write_text(json.dumps({
        "Configuration": {"simTime": 1e-5},
        "Materials": [{"type": "Hollomon", "const": [386.796e6, 0.154], "density0": 2700.0, "youngsModulus": 68.9e9,
                       "poissonsRatio": 0.3, "yieldStress0": 190.4e6}],
        "DomainBlocks": [{"type": "File", "fileName": "m.k"}],
        "BoundaryConditions": [{"zoneId": 1, "valueType": 0, "value": [0, 0, 0], "start": [-1, -1, -0.1], "end": [2, 2, 0.1]}]}))
    r = subprocess.run([os.path.join(host_bins, "wf_weldform"), str(deck), "--parse-only"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    s = json.loads(r.stdout.strip().splitlines()[-1])
    assert (s["nodxelem"], s["nodes"], s["elements"]) == (4, 5, 2)
    assert s["bc_count"] == [3, 3, 3]          # the three nodes with z = 0


# ---- the Python deck loader (weldformfem_b200/deck.py), same checker -----------------------------------------------------
@pytest.mark.parametrize("name", PINNED)
def test_python_deck_loader_reproduces_reference_front_end_on_the_oracle(oracle_port, name):
    """deck.load + DeckSetup.apply drive the plain-C oracle to the state the reference's own main.C + step loop produce:
    bit for bit, including dt, the BC lists, the rigid surfaces and the thermal settings (CPU only)."""
    from weldformfem_b200 import deck
    gold = _gold(name)
    S = deck.load(os.path.join(DECKS, name + ".json"))
    o = oracle_port()
    S.apply(o)
    assert S.dt == float(gold["dt"][0]) and S.sim_time == float(gold["end_t"][0])
    info = o.info()
    assert [info[k] for k in "dim nodxelem n_nodes n_elems bcx bcy bcz".split()] == [int(v) for v in gold["info"][:7]]
    assert np.array_equal(o.get("m_elnod"), gold["m_elnod"]) and np.array_equal(o.get("x"), gold["x0"])
    o.step(int(gold["steps"][0]))
    for key in gold.files:
        if key.startswith("sN_"):
            assert np.array_equal(o.get(key[3:]), gold[key]), key


def test_python_k_reader_matches_cpp_reader(host_bins, tmp_path):
    from weldformfem_b200 import deck
    x, el = deck.read_k(os.path.join(DECKS, "tet_block.k"))
    s = _summary(host_bins, "file_tet_zones", "--parse-only")
    assert (el.shape[1], len(x), len(el)) == (s["nodxelem"], s["nodes"], s["elements"])
    k = tmp_path / "m.k"
    k.write_text("*NODE\n1,0,0,0\n2,1,0,0\n3,0,1,0\n4,0,0,1\n*ELEMENT_SOLID\n1,1\n1,2,3,4,4,4,4,4\n")
    x, el = deck.read_k(str(k))
    assert el.tolist() == [[0, 1, 2, 3]] and x.shape == (4, 3)
    with pytest.raises(ValueError):
        k.write_text("*NODE\n1,0,0,0\n*ELEMENT_SOLID\n1,1,1,2,3,4,4,4,4,4\n")
        deck.read_k(str(k))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["file_tet_contact", "box_axiquad"])
def test_python_deck_loader_on_the_engine(name):
    from weldformfem_b200 import deck
    from weldformfem_b200.domain import Domain_d
    gold = _gold(name)
    S = deck.load(os.path.join(DECKS, name + ".json"))
    eng = S.apply(Domain_d(strict=True))
    assert abs(S.dt - float(gold["dt"][0])) <= 4e-16 * S.dt
    eng.step(int(gold["steps"][0]))
    for key in gold.files:
        if key.startswith("sN_") and key[3:] not in ("trimesh.node",):
            assert relerr(eng.get(key[3:]), gold[key]) <= 1e-8, key


@pytest.mark.parametrize("name", ["hex_file"])
def test_decks_main_c_cannot_run_still_match_the_compiled_reference_step(oracle_port, oracle_ref, tmp_path, name):
    """Hexahedral `.k` files crash the reference's front-end (its unconditional SearchExtNodes, main.C:650, overruns on
    8-node elements) but not its step: set up by deck.py, the plain-C oracle and the compiled reference sources must
    agree bit for bit.  (Triangle decks also crash main.C; the reference's calcMinEdgeLength does not cover them either,
    so their time step is the engine's own and they are only checked by the GPU parity cases.)"""
    from weldformfem_b200 import deck
    oracle_ref.set_threads(1)
    if name == "hex_file":
        path = str(tmp_path / "hex_file.json")
        import shutil
        shutil.copy(os.path.join(DECKS, "hex_block.k"), str(tmp_path / "hex_block.k"))
        with open(path, "w") as f:
            json.dump({"Configuration": {"simTime": 1e-4, "cflFactor": 0.3, "zSymm": True, "symtol": 1e-6},
                       "Materials": [{"type": "Hollomon", "const": [386.796e6, 0.154], "density0": 2700.0,
                                      "youngsModulus": 68.9e9, "poissonsRatio": 0.3, "yieldStress0": 190.4e6}],
                       "DomainBlocks": [{"type": "File", "fileName": "hex_block.k"}],
                       "BoundaryConditions": [{"zoneId": 2, "valueType": 0, "value": [0.0, 0.0, -80.0],
                                               "start": [-1, -1, 0.00499], "end": [1, 1, 0.00501]}]}, f)
    else:
        path = os.path.join(DECKS, name + ".json")
    doms = []
    for cls in (oracle_port, oracle_ref):
        S = deck.load(path)
        d = cls()
        S.apply(d)
        d.step(40)
        doms.append((S, d))
    (Sa, a), (Sb, b) = doms
    assert Sa.dt == Sb.dt and Sa.dt > 0
    info = a.info()
    assert info["bcx"] + info["bcy"] + info["bcz"] > 0 and np.abs(a.get("v")).max() > 0
    for nm in ("x", "v", "u", "m_fi", "m_sigma", "pl_strain", "p"):
        assert np.array_equal(a.get(nm), b.get(nm)), nm
