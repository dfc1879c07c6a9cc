"""The C++ host side (host/wf_domain.hpp + drivers) over the C ABI: it builds with the image's g++, fails loudly
without a CUDA device, and on the GPU reproduces configs[0] (golden fixture dumped from the unmodified reference)
and the box workloads (against the oracle) when driven purely from C++."""
import os
import re
import subprocess

import numpy as np
import pytest

from parity_util import relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "host", "_bin")


@pytest.fixture(scope="module")
def host_bins():
    from weldformfem_b200 import build
    build.build()
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return BIN


def test_cpp_host_builds_and_has_no_cpu_path(host_bins):
    for exe in ("main_1_elem_3d", "wf_explicit"):
        assert os.access(os.path.join(host_bins, exe), os.X_OK)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is only visible without one")
    r = subprocess.run([os.path.join(host_bins, "main_1_elem_3d")], capture_output=True, text=True)
    assert r.returncode == 1
    assert "no CPU fallback" in r.stderr


def _parse_blocks(text):
    out, cur = {}, None
    for line in text.splitlines():
        if re.fullmatch(r"[a-z_]+", line.strip()):
            cur = line.strip()
            out[cur] = []
        elif cur and re.match(r"^[-+0-9.einfa ]+$", line.strip()) and line.strip():
            out[cur].append([float(t) for t in line.split()])
        else:
            cur = None
    return {k: np.array(v).ravel() for k, v in out.items()}


@pytest.mark.gpu
@pytest.mark.parametrize("hg,fixture", [(0.06, "c1_1hex_hg006"), (0.0, "c1_1hex_nohg")])
@pytest.mark.parametrize("strict", [1, 0])
def test_cpp_main_1_elem_3d_matches_reference_fixture(host_bins, hg, fixture, strict):
    r = subprocess.run([os.path.join(host_bins, "main_1_elem_3d"), str(hg), str(strict)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "steps 126 " in r.stdout
    got = _parse_blocks(r.stdout)
    gold = np.load(os.path.join(ROOT, "tests", "golden", fixture + ".npz"))
    for nm in ("u", "v", "a", "m_fi", "m_sigma"):
        assert relerr(got[nm], gold["sN_" + nm]) < (1e-9 if strict else 1e-7), nm     # 126 steps; 1e-6 allowed


def _read_dump(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            head = f.readline()
            if not head:
                break
            nm, cnt = head.split()
            out[nm.decode()] = np.frombuffer(f.read(8 * int(cnt)), dtype=np.float64)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n", [("hex", 6), ("tet", 5), ("axiquad", 10), ("quad", 10), ("tri", 10)])
def test_cpp_wf_explicit_matches_oracle(host_bins, oracle_port, tmp_path, kind, n):
    import dataclasses
    from weldformfem_b200 import cases
    case = {"hex": cases.c3_hexes, "tet": cases.c2_tets, "axiquad": cases.c4_axisymm_quads,
            "quad": cases.plane_strain_quads, "tri": cases.plane_strain_tris}[kind](n)
    vtop = -200.0 if case.dim == 3 else -50.0
    case = dataclasses.replace(case, top_vel=vtop)
    dump = str(tmp_path / "d.bin")
    r = subprocess.run([os.path.join(host_bins, "wf_explicit"), "--kind", kind, "--n", str(n), "--steps", "20",
                        "--vtop", str(vtop), "--dump", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = _read_dump(dump)
    ref = oracle_port()
    case.apply(ref)
    ref.step(20)
    for nm in ("x", "v", "u", "m_fi", "m_sigma", "pl_strain", "p"):
        assert relerr(got[nm], ref.get(nm)) < 1e-8, (kind, nm)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n", [("tet", 6), ("quad", 12)])
def test_cpp_wf_explicit_contact_matches_oracle(host_bins, oracle_port, tmp_path, kind, n):
    """The C++ TriMesh_d / setTriMesh / setContactOn path (host/wf_domain.hpp) against the oracle's contact step."""
    from weldformfem_b200 import cases
    vt = -200.0 if kind == "tet" else -100.0
    case = cases.contact_tets(n, tool_vel=vt, two_planes=False) if kind == "tet" else cases.contact_quads(n, tool_vel=vt)
    dump = str(tmp_path / "c.bin")
    r = subprocess.run([os.path.join(host_bins, "wf_explicit"), "--kind", kind, "--n", str(n), "--steps", "40", "--vtop", str(vt),
                        "--contact", "--dump", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = _read_dump(dump)
    ref = oracle_port()
    case.apply(ref)
    ref.step(40)
    assert (ref.get("m_mesh_in_contact") >= 0).any()
    for nm in ("x", "v", "u", "m_fi", "m_sigma", "pl_strain", "p", "contforce", "ut_prev", "node_area"):
        assert relerr(got[nm], ref.get(nm)) < 1e-8, (kind, nm)
