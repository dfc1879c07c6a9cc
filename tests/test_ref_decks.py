"""The reference's OWN example decks (examples/input/*.json + tetra_cyl.k) — not decks written for this repository —
through both front-ends: the reference's src/explicit/main.C (/root/reference/src/explicit/main.C:125-995, compiled
unmodified into oracle/_ref, stepped by the harness) and host/wf_weldform on the engine.

  Compression_tetra.json                 cylinder of 12 197 tets; its BC zones lie outside the mesh (0 BC nodes on either
                                         side), so nothing moves: checks the set-up path (dt, mesh, counts)
  Contact_Compression_tetra.json         the same cylinder between two rigid planes of 800 facets, friction 0.3, thermal
                                         coupling with contact heat exchange, Hollomon
  Contact_Compression_axisymm_quad.json  axisymmetric quads pressed by rigid lines, friction

CPU tests: wf_weldform --parse-only and the Python loader see what main.C sees.  GPU tests: states after N steps."""
import gzip
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from parity_util import relerr
from test_host_cpp import _read_dump, host_bins  # noqa: F401  (fixture)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "golden", "ref_decks")
DECKS = {"Compression_tetra": 20, "Contact_Compression_tetra": 60, "Contact_Compression_axisymm_quad": 80}
STATE = "x v u m_sigma pl_strain p sigma_y vol".split()


@pytest.fixture()
def workdir(tmp_path):
    for f in os.listdir(SRC):
        if f.endswith(".json"):
            shutil.copy(os.path.join(SRC, f), tmp_path)
    with gzip.open(os.path.join(SRC, "tetra_cyl.k.gz"), "rb") as fi, open(tmp_path / "tetra_cyl.k", "wb") as fo:
        shutil.copyfileobj(fi, fo)
    return tmp_path


def test_fixtures_are_the_reference_files():
    src = "/root/reference/examples/input"
    if not os.path.isdir(src):
        pytest.skip("reference tree not present")
    for f in os.listdir(SRC):
        if f.endswith(".json"):
            assert open(os.path.join(SRC, f), "rb").read() == open(os.path.join(src, f), "rb").read(), f
    assert gzip.open(os.path.join(SRC, "tetra_cyl.k.gz"), "rb").read() == open(os.path.join(src, "tetra_cyl.k"), "rb").read()


@pytest.mark.parametrize("name", sorted(DECKS))
def test_front_ends_agree_on_the_setup(host_bins, oracle_ref, workdir, name):
    """mesh, BC counts, rigid surfaces and end time: main.C vs wf_weldform --parse-only vs deck.py (the time step needs the
    min edge length, which the engine computes on the GPU: checked by the GPU test)"""
    from weldformfem_b200 import deck
    ref, dt, end_t = oracle_ref.from_deck(str(workdir / (name + ".json")))
    info, tm = ref.info(), ref.trimesh_counts()
    r = subprocess.run([os.path.join(host_bins, "wf_weldform"), str(workdir / (name + ".json")), "--parse-only"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    s = json.loads(r.stdout.strip().splitlines()[-1])
    assert (s["dim"], s["nodxelem"], s["nodes"], s["elements"]) == (info["dim"], info["nodxelem"], info["n_nodes"], info["n_elems"])
    assert s["bc_count"][:info["dim"]] == [info["bcx"], info["bcy"], info["bcz"]][:info["dim"]]
    assert s["rigid_facets"] == tm["elemcount"] and s["end_time"] == end_t
    S = deck.load(str(workdir / (name + ".json")))
    if S.mesh is not None:
        assert np.array_equal(S.mesh[0].reshape(-1), ref.get("x"))
        assert np.array_equal(S.mesh[1].reshape(-1), ref.get("m_elnod"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(DECKS))
def test_reference_decks_run_like_the_reference(host_bins, oracle_ref, workdir, name):
    steps = DECKS[name]
    oracle_ref.set_threads(1)        # the reference's 2D friction reset races under OpenMP
    ref, dt, end_t = oracle_ref.from_deck(str(workdir / (name + ".json")))
    ref.init(dt)
    ref.step(steps)
    dump = str(workdir / "state.bin")
    r = subprocess.run([os.path.join(host_bins, "wf_weldform"), str(workdir / (name + ".json")), "--steps", str(steps),
                        "--dump", dump, "--strict"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    summary = json.loads(r.stdout.strip().splitlines()[-1])
    assert summary["steps"] == steps and abs(summary["dt"] - dt) <= 4e-16 * dt, (summary["dt"], dt)
    got = _read_dump(dump)
    names = list(STATE)
    if ref.trimesh_counts()["nodecount"]:
        names += ["contforce", "trimesh.node"]
        assert np.abs(ref.get("contforce")).max() > 0, "the tools should be touching the part by now"
    if np.any(ref.get("T") != 0.0):
        names += ["T"]
    worst = {nm: relerr(got[nm], ref.get(nm)) for nm in names}
    bad = {k: v for k, v in worst.items() if not v <= 1e-8}
    assert not bad, (name, bad, worst)
