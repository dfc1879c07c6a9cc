"""Legacy-VTK writer: same data set and array names as the reference's VTKWriter.C, binary encoding.

CPU: the Python writer on the oracle (a domain with the same accessors as the engine), read back and compared with the
arrays it was written from; nodal stress against a direct restatement of avgScalar + von Mises.  GPU: the C++ twin
(host/wf_vtk.hpp through `wf_weldform --vtk`) against the arrays dumped by the same run, and against the Python writer
fed with those arrays (identical bytes)."""
import os
import subprocess

import numpy as np
import pytest

from weldformfem_b200 import cases
from weldformfem_b200 import vtk

HERE = os.path.dirname(os.path.abspath(__file__))
DECKS = os.path.join(HERE, "golden", "decks")


def _f32(a):
    return np.asarray(a, np.float64).astype(np.float32).astype(np.float64)


def _check_file(dom, path, binary):
    info = dom.info()
    dim, k, nn, ne = info["dim"], info["nodxelem"], info["n_nodes"], info["n_elems"]
    r = vtk.read_vtk(path)
    x = np.asarray(dom.get("x")).reshape(nn, dim)
    assert np.array_equal(r["POINTS"][:, :dim], _f32(x))
    if dim == 2:
        assert not r["POINTS"][:, 2].any()
    cells = r["CELLS"].reshape(ne, k + 1)
    assert (cells[:, 0] == k).all() and np.array_equal(cells[:, 1:], np.asarray(dom.get("m_elnod")).reshape(ne, k))
    assert (r["CELL_TYPES"] == vtk.VTK_CELL_TYPE[(dim, k)]).all()
    P, Cd = r["POINT_DATA"], r["CELL_DATA"]
    for nm, src in (("DISP", "u"), ("Acceleration", "a"), ("Velocity", "v")):
        assert np.array_equal(P[nm][:, :dim], _f32(np.asarray(dom.get(src)).reshape(nn, dim))), nm
    assert np.array_equal(P["nod_mass"], _f32(dom.get("m_mdiag")))
    assert not P["Part_ID"].any()
    for nm, src in (("pressure", "p"), ("pl_strain", "pl_strain"), ("Vol", "vol"), ("Rho", "rho"), ("Vol_0", "vol_0"),
                    ("sigy", "sigma_y")):
        assert np.array_equal(Cd[nm], _f32(dom.get(src))), nm
    assert np.array_equal(Cd["J"], _f32(np.asarray(dom.get("vol")) / np.asarray(dom.get("vol_0"))))
    # nodal stress: plain loops over the node -> element lists (avgScalar, Domain_d.h:77-87), then sqrt(3 J2)
    sig = np.asarray(dom.get("m_sigma")).reshape(ne, 6)
    nodel, off, cnt = dom.get("m_nodel"), dom.get("m_nodel_offset"), dom.get("m_nodel_count")
    avg = np.zeros((nn, 6))
    for n in range(nn):
        for j in range(cnt[n]):
            avg[n] += sig[nodel[off[n] + j]]
        avg[n] /= cnt[n]
    t = P["SIGMAT"]
    assert np.array_equal(t[:, [0, 4, 8, 1, 5, 2]], _f32(avg)) and not t[:, [3, 6, 7]].any()
    dev = avg[:, :3] - avg[:, :3].mean(1, keepdims=True)
    vm = np.sqrt(3.0 * (0.5 * (dev ** 2).sum(1) + (avg[:, 3:] ** 2).sum(1)))
    assert np.allclose(P["stress"], vm, rtol=1e-6, atol=1e-6 * max(1.0, np.abs(vm).max()))
    order = r["order"]
    assert order.index("DISP") < order.index("Velocity") < order.index("nod_mass") < order.index("stress") < \
        order.index("SIGMAT") < order.index("pressure") < order.index("pl_strain") < order.index("J") < order.index("sigy")
    return r


@pytest.mark.parametrize("key", ["hex", "tet", "axiquad", "pstri"])
@pytest.mark.parametrize("binary", [True, False], ids=["binary", "ascii"])
def test_writer_on_the_oracle_round_trips(key, binary, oracle_port, tmp_path):
    case = {"hex": cases.c3_hexes(4), "tet": cases.c2_tets(3), "axiquad": cases.c4_axisymm_quads(6),
            "pstri": cases.plane_strain_tris(6)}[key]
    dom = oracle_port()
    case.apply(dom)
    dom.step(15)
    path = str(tmp_path / ("out_%s.vtk" % key))
    written = vtk.write_vtk(dom, path, binary=binary)
    assert written[:4] == ["DISP", "Acceleration", "Velocity", "Part_ID"]
    head = open(path, "rb").read(120).split(b"\n")
    assert head[0] == b"# vtk DataFile Version 3.0" and head[2] == (b"BINARY" if binary else b"ASCII")
    _check_file(dom, path, binary)
    if binary:   # both encodings carry the same float32 values
        p2 = str(tmp_path / "ascii.vtk")
        vtk.write_vtk(dom, p2, binary=False)
        a, b = vtk.read_vtk(path), vtk.read_vtk(p2)
        for sec in ("POINT_DATA", "CELL_DATA"):
            assert a[sec].keys() == b[sec].keys()
            for nm in a[sec]:
                assert np.array_equal(a[sec][nm], b[sec][nm]), nm


def test_writer_contact_and_thermal_arrays_in_the_reference_order(oracle_port, tmp_path):
    """A tet block pressed by a rigid plane with thermal coupling: the optional arrays appear, in the order the reference
    writes them (VTKWriter.C:366-660: Temp and ContForce after Part_ID, ext_nodes / nod_area / nod_p between the nodal
    stress and SIGMAT, ele_area first among the cell data)."""
    case = cases.with_thermal(cases.contact_tets(4), heat_cond=25000.0, T_die=200.0)
    dom = oracle_port()
    case.apply(dom)
    dom.step(30)
    path = str(tmp_path / "contact.vtk")
    written = vtk.write_vtk(dom, path)
    want = ["DISP", "Acceleration", "Velocity", "Part_ID", "Temp", "ContForce", "nod_mass", "stress", "ext_nodes", "nod_area",
            "nod_p", "SIGMAT"]
    assert [w for w in written if w in want] == want and written.index("ele_area") < written.index("pressure")
    r = _check_file(dom, path, True)
    P, Cd = r["POINT_DATA"], r["CELL_DATA"]
    nn = dom.info()["n_nodes"]
    assert np.array_equal(P["Temp"], _f32(dom.get("T"))) and P["Temp"].max() > 20.0
    assert np.array_equal(P["ContForce"], _f32(np.asarray(dom.get("contforce")).reshape(nn, 3))) and np.abs(P["ContForce"]).max() > 0
    assert np.array_equal(P["ext_nodes"], (np.asarray(dom.get("ext_nodes")) != 0).astype(np.float64)) and 0 < P["ext_nodes"].sum() < nn
    assert np.array_equal(P["nod_area"], _f32(dom.get("node_area"))) and np.array_equal(Cd["ele_area"], _f32(dom.get("m_elem_area")))
    assert np.array_equal(P["nod_p"], _f32(dom.get("p_node")))


class _DumpDomain:
    """Domain-shaped view of a `wf_weldform --dump` file + the connectivity of the VTK file itself."""

    def __init__(self, dump, vtk_file):
        self.a = dict(dump)
        r = vtk.read_vtk(vtk_file)
        nn, ne = r["POINTS"].shape[0], r["CELL_TYPES"].size
        k = int(r["CELLS"][0])
        dim = 3 if int(r["CELL_TYPES"][0]) in (12, 10) else 2
        el = r["CELLS"].reshape(ne, k + 1)[:, 1:]
        self._info = dict(dim=dim, nodxelem=k, n_nodes=nn, n_elems=ne)
        # setNodElem (Domain_d.C): per node its (element, corner) pairs, ascending element id
        flat = el.ravel()
        order = np.argsort(flat, kind="stable")
        self.a["m_elnod"] = flat.astype(np.uint32)
        self.a["m_nodel"] = (order // k).astype(np.int32)
        cnt = np.bincount(flat, minlength=nn).astype(np.int32)
        self.a["m_nodel_count"] = cnt
        self.a["m_nodel_offset"] = (np.cumsum(cnt) - cnt).astype(np.int32)
        if "ext_nodes" in self.a:
            self.a["ext_nodes"] = self.a["ext_nodes"].astype(np.uint8)

    def info(self):
        return self._info

    def get(self, name):
        return self.a[name]


@pytest.mark.gpu
@pytest.mark.parametrize("deck", ["file_tet_contact", "box_axiquad", "box_psquad"])
@pytest.mark.parametrize("binary", [True, False], ids=["binary", "ascii"])
def test_cpp_writer_on_the_engine(deck, binary, tmp_path):
    """`wf_weldform deck --steps N --vtk out.vtk --dump d.bin` (host/wf_vtk.hpp over the C ABI): the file read back matches
    the dumped arrays, and the Python writer fed with the same arrays produces identical bytes."""
    from test_host_cpp import BIN, _read_dump
    path = os.path.join(DECKS, deck + ".json")
    r = subprocess.run(["make", "-C", os.path.join(os.path.dirname(HERE), "host")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out_c, dump = str(tmp_path / "c.vtk"), str(tmp_path / "d.bin")
    cmd = [os.path.join(BIN, "wf_weldform"), path, "--steps", "25", "--strict", "--vtk", out_c, "--dump", dump]
    r = subprocess.run(cmd + ([] if binary else ["--vtk-ascii"]), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    dom = _DumpDomain(_read_dump(dump), out_c)
    _check_file(dom, out_c, binary)
    names = vtk.read_vtk(out_c)["order"]
    if deck == "file_tet_contact":
        assert {"Temp", "ContForce", "nod_area", "ext_nodes", "ele_area"} <= set(names)
    out_p = str(tmp_path / "p.vtk")
    vtk.write_vtk(dom, out_p, binary=binary)
    assert open(out_c, "rb").read() == open(out_p, "rb").read()


@pytest.mark.gpu
def test_wf_weldform_output_cadence(tmp_path):
    """`wf_weldform deck --out BASE` follows the reference loop (Solver_explicit.C:305, 1036-1041, 1155, 1167): a file after
    every step whose START time is >= tout (tout = 0, then += outTime), numbered from 00000 and listed with that start
    time in BASE_res.json; the steps in between run as fused batches."""
    import json
    from test_host_cpp import BIN
    r = subprocess.run(["make", "-C", os.path.join(os.path.dirname(HERE), "host")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    j = json.load(open(os.path.join(DECKS, "box_psquad.json")))
    j["Configuration"]["simTime"] = 3.0e-6
    j["Configuration"]["outTime"] = 1.0e-6
    deck = str(tmp_path / "cadence.json")
    json.dump(j, open(deck, "w"))
    base = str(tmp_path / "run")
    r = subprocess.run([os.path.join(BIN, "wf_weldform"), deck, "--out", base, "--strict"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    summary = json.loads(r.stdout.strip().splitlines()[-1])
    dt, sim, out = summary["dt"], 3.0e-6, 1.0e-6
    want, t, tout, steps = [], 0.0, 0.0, 0
    while t < sim:
        label = t
        t += dt
        steps += 1
        if label >= tout:
            want.append((steps, label))
            tout += out
    assert summary["steps"] == steps and len(want) == 3 and want[0] == (1, 0.0)
    res = json.load(open(base + "_res.json"))["vtk_files"]
    assert [e["file"] for e in res] == [base + "_%05d.vtk" % i for i in range(len(want))]
    assert [e["time"] for e in res] == [w[1] for w in want]
    rows = open(base + "_energy.csv").read().strip().split("\n")
    assert rows[0] == "t,Ekin,dEint" and len(rows) == 1 + len(want)
    assert [float(x.split(",")[0]) for x in rows[1:]] == [w[1] for w in want]
    # file i holds the state after want[i][0] steps: same bytes as a run stopped there
    for i, (nsteps, _) in enumerate(want):
        single = str(tmp_path / ("single_%d.vtk" % i))
        r = subprocess.run([os.path.join(BIN, "wf_weldform"), deck, "--steps", str(nsteps), "--vtk", single, "--strict"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert open(single, "rb").read() == open(base + "_%05d.vtk" % i, "rb").read(), i
