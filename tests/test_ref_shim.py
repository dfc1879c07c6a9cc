"""The reference-side binding, COMPILED AND RUN against the unmodified reference tree.

tests/ref_shim/Solver_b200.C is the file INTEGRATION.md tells a WeldFormFEM maintainer to add: a subclass of
MetFEM::Domain_d whose AttachB200() / SolveChungHulbert_b200() replace the body of Domain_d::SolveChungHulbert()
(/root/reference/src/explicit/Solver_explicit.C:101) by calls through include/wf_engine.h.  `make -C oracle shim`
builds oracle/_ref/wf_ref_shim = the reference's own src/explicit/main.C (unmodified) + its solver sources + that
shim + libwf_b200.so, with the one call at main.C:995 swapped.  The same binary runs a deck either on the B200 engine
or (WF_SHIM_CPU=1) on the reference's own CPU SolveChungHulbert(); both dump the final state.

CPU test: the CPU arm of the binary reproduces the committed main.C fixtures bit for bit (so the binary really is
the reference).  GPU test: B200 arm == CPU arm within the tolerances of BASELINE.json."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from parity_util import relerr
from test_host_cpp import _read_dump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "_ref", "wf_ref_shim")
DECKS = os.path.join(ROOT, "tests", "golden", "decks")


@pytest.fixture(scope="module")
def shim_bin():
    if os.path.isdir("/root/reference"):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "shim"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    if not os.path.exists(SHIM):
        pytest.skip("oracle/_ref/wf_ref_shim not built (needs /root/reference)")
    return SHIM


def run_shim(shim_bin, deck, tmp_path, tag, cpu, extra_env=None, steps=None):
    work = tmp_path / tag
    work.mkdir()
    for f in os.listdir(DECKS):            # deck + the .k meshes it may name (main.C resolves them against the cwd)
        if f.endswith((".json", ".k")):
            shutil.copy(os.path.join(DECKS, f), work)
    if steps is not None:                  # `while (Time < end_t)` with Time += dt: simTime = (steps - 1/2) dt runs `steps` steps
        import json
        dt = float(np.load(os.path.join(ROOT, "tests", "golden", f"deck_{deck}.npz"))["dt"][0])
        j = json.load(open(work / (deck + ".json")))
        j["Configuration"]["simTime"] = (steps - 0.5) * dt
        j["Configuration"]["fixedTS"] = True
        json.dump(j, open(work / (deck + ".json"), "w"))
    env = dict(os.environ, WF_SHIM_DUMP=str(work / "state.bin"), OMP_NUM_THREADS="1")
    env.pop("WF_SHIM_CPU", None)
    if cpu:
        env["WF_SHIM_CPU"] = "1"
    env.update(extra_env or {})
    r = subprocess.run([shim_bin, deck + ".json"], cwd=work, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return _read_dump(str(work / "state.bin"))


def test_shim_source_is_what_integration_md_quotes():
    src = open(os.path.join(ROOT, "tests", "ref_shim", "Solver_b200.C")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    body = src[src.index("namespace MetFEM {"):]
    assert body in doc, "INTEGRATION.md section 2 must quote tests/ref_shim/Solver_b200.C verbatim"


@pytest.mark.parametrize("deck", ["file_tet_zones", "box_psquad", "box_axiquad"])
def test_shim_cpu_arm_is_the_reference(shim_bin, tmp_path, deck):
    """WF_SHIM_CPU=1: main.C + Domain_d::SolveChungHulbert() of the unmodified sources.  Its result equals the
    committed fixture of the same deck (main.C set-up + the harness's member-by-member loop), bit for bit."""
    gold = np.load(os.path.join(ROOT, "tests", "golden", f"deck_{deck}.npz"))
    steps = int(gold["steps"][0])
    got = run_shim(shim_bin, deck, tmp_path, "cpu", cpu=True, steps=steps)
    t, dt = got["time_dt"]
    assert dt == float(gold["dt"][0])
    assert abs(t / dt - steps) < 1e-6
    for nm in ("x", "v", "u", "m_sigma", "pl_strain", "p"):
        assert np.array_equal(got[nm], gold["sN_" + nm]), nm


@pytest.mark.gpu
@pytest.mark.parametrize("deck", ["file_tet_zones", "box_psquad", "box_axiquad"])
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
def test_shim_b200_arm_matches_reference_cpu_solver(shim_bin, tmp_path, deck, strict):
    """main.C:995 swapped: the deck runs to simTime on the B200 engine through the shim and is compared with the
    reference's own CPU solver inside the same binary (tets with the shipped ANP pressure + Hollomon; plane-strain and
    axisymmetric quads with the shipped hourglass)."""
    # the tet deck runs the pressure law the reference ships as algorithm 1, which ACCUMULATES (Mechanical.C:1243-1247):
    # the state grows exponentially and overflows within a few hundred steps on either side, so it is compared early
    steps = 20 if deck == "file_tet_zones" else 300
    cpu = run_shim(shim_bin, deck, tmp_path, "cpu", cpu=True, steps=steps)
    gpu = run_shim(shim_bin, deck, tmp_path, "gpu", cpu=False, extra_env={"WF_SHIM_STRICT": "1"} if strict else None, steps=steps)
    assert gpu["time_dt"][1] == cpu["time_dt"][1]
    assert abs(gpu["time_dt"][0] - cpu["time_dt"][0]) < 1e-9 * cpu["time_dt"][0]
    assert abs(cpu["time_dt"][0] / cpu["time_dt"][1] - steps) < 1e-6
    assert (cpu["pl_strain"] > 0).mean() > 0.2, "the run should reach plasticity"
    worst = {nm: relerr(gpu[nm], cpu[nm]) for nm in ("x", "v", "u", "m_sigma", "m_tau", "pl_strain", "p", "sigma_y", "vol")}
    assert max(worst.values()) < (1e-9 if strict else 1e-7), worst
