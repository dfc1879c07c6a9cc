"""Legacy-VTK output of a domain ("binary VTK/raw writer replacing 5-digit ASCII").

The reference writes an ASCII legacy VTK file with four to six significant digits per value through an ostringstream
(`src/common/VTKWriter.C:236-660`).  This writer emits the same data set — UNSTRUCTURED_GRID, the same array names in the
same order, so a ParaView state made for the reference's files opens these — as BINARY (big-endian float32 / int32, the
legacy format's binary encoding) or, for diffing, ASCII.  It works on anything with the `Domain_d` accessors (`get`,
`info`): the B200 engine or the test oracle.

Arrays written (reference name <- domain array); an array the domain does not hold is skipped:
  POINTS <- x;  CELLS / CELL_TYPES <- m_elnod (VTK types 12 hexahedron, 10 tetra, 9 quad, 5 triangle, VTKWriter.C:337-352)
  POINT_DATA: VECTORS DISP <- u, Acceleration <- a, Velocity <- v; SCALARS Part_ID = 0; Temp <- T; VECTORS ContForce <-
    contforce; SCALARS nod_mass <- m_mdiag; stress = von Mises of the nodal average of m_sigma (avgScalar,
    Domain_d.h:77-87: sum over the node's elements / count, then sqrt(3 J2), VTKWriter.C:470-490); ext_nodes; nod_area <-
    node_area; nod_p <- p_node; TENSORS SIGMAT = that nodal average, upper triangle as the reference writes it
    (xx xy xz / 0 yy yz / 0 0 zz, VTKWriter.C:524-530); TENSORS EPSR <- nodal average of m_str_rate (when stored)
  CELL_DATA: ele_area <- m_elem_area; pressure <- p; pl_strain; TENSORS DDEVT <- m_str_rate (when stored); J = vol/vol_0;
    Vol; Rho; Vol_0; sigy <- sigma_y
Not written: the duplicate "Position" vector (= POINTS) and the rigid tool surfaces the reference appends as extra cells.
The C++ twin is host/wf_vtk.hpp (`wf_weldform --vtk FILE`); the two produce identical bytes (tests/test_vtk.py)."""
from __future__ import annotations

import numpy as np

VTK_CELL_TYPE = {(3, 8): 12, (3, 4): 10, (2, 4): 9, (2, 3): 5}
TITLE = "WeldFormFEM explicit step, B200 engine"


def _try(dom, name):
    try:
        a = dom.get(name)
    except Exception:
        return None
    if a is None or np.size(a) == 0:
        return None
    return np.asarray(a)


def nodal_average(dom, elem_vals: np.ndarray, ncomp: int) -> np.ndarray:
    """avgScalar (Domain_d.h:77-87): per node, the sum of the element rows in nodel order divided by the count."""
    nodel = np.asarray(dom.get("m_nodel"), np.int64)
    off = np.asarray(dom.get("m_nodel_offset"), np.int64)
    cnt = np.asarray(dom.get("m_nodel_count"), np.int64)
    nn = cnt.size
    ev = np.asarray(elem_vals, np.float64).reshape(-1, ncomp)
    out = np.zeros((nn, ncomp))
    # same association as the reference loop: entries added one list position at a time, ascending position
    for j in range(int(cnt.max()) if nn else 0):
        m = cnt > j
        out[m] += ev[nodel[off[m] + j]]
    return out / cnt[:, None]


def _pad3(a: np.ndarray, dim: int) -> np.ndarray:
    a = np.asarray(a, np.float64).reshape(-1, dim)
    if dim == 3:
        return a
    return np.concatenate([a, np.zeros((a.shape[0], 1))], axis=1)


def _upper(t6: np.ndarray) -> np.ndarray:
    """flat symmetric rows (xx yy zz xy yz xz) -> 3x3 rows with the lower triangle zero, as the reference prints them"""
    t6 = t6.reshape(-1, 6)
    z = np.zeros(t6.shape[0])
    return np.stack([t6[:, 0], t6[:, 3], t6[:, 5], z, t6[:, 1], t6[:, 4], z, z, t6[:, 2]], axis=1)


class _Out:
    def __init__(self, f, binary):
        self.f, self.binary = f, binary

    def text(self, s):
        self.f.write((s + "\n").encode("ascii"))

    def floats(self, a, per_line):
        a = np.asarray(a, np.float64)
        if self.binary:
            self.f.write(a.astype(">f4").tobytes())
            self.f.write(b"\n")
        else:
            a32 = a.astype(np.float32).reshape(-1, per_line)
            for row in a32:
                self.f.write((" ".join("%.9g" % float(v) for v in row) + "\n").encode("ascii"))

    def ints(self, a, per_line):
        a = np.asarray(a, np.int64)
        if self.binary:
            self.f.write(a.astype(">i4").tobytes())
            self.f.write(b"\n")
        else:
            for row in a.reshape(-1, per_line):
                self.f.write((" ".join(str(int(v)) for v in row) + "\n").encode("ascii"))


def write_vtk(dom, path: str, binary: bool = True) -> list[str]:
    """Write the domain's current state; returns the list of data arrays written (in file order)."""
    info = dom.info()
    dim, k, nn, ne = int(info["dim"]), int(info["nodxelem"]), int(info["n_nodes"]), int(info["n_elems"])
    written = []
    with open(path, "wb") as f:
        o = _Out(f, binary)
        o.text("# vtk DataFile Version 3.0")
        o.text(TITLE)
        o.text("BINARY" if binary else "ASCII")
        o.text("DATASET UNSTRUCTURED_GRID")
        o.text(f"POINTS {nn} float")
        o.floats(_pad3(dom.get("x"), dim), 3)
        el = np.asarray(dom.get("m_elnod"), np.int64).reshape(ne, k)
        o.text(f"CELLS {ne} {ne * (k + 1)}")
        o.ints(np.concatenate([np.full((ne, 1), k, np.int64), el], axis=1), k + 1)
        o.text(f"CELL_TYPES {ne}")
        o.ints(np.full(ne, VTK_CELL_TYPE[(dim, k)], np.int64), 1)

        def scalars(name, a):
            o.text(f"SCALARS {name} float 1")
            o.text("LOOKUP_TABLE default")
            o.floats(a, 1)
            written.append(name)

        def vectors(name, a):
            o.text(f"VECTORS {name} float")
            o.floats(_pad3(a, dim), 3)
            written.append(name)

        def tensors(name, t6):
            o.text(f"TENSORS {name} float")
            o.floats(_upper(t6), 9)
            written.append(name)

        o.text(f"POINT_DATA {nn}")
        vectors("DISP", dom.get("u"))
        vectors("Acceleration", dom.get("a"))
        vectors("Velocity", dom.get("v"))
        scalars("Part_ID", np.zeros(nn))
        T = _try(dom, "T")
        if T is not None:
            scalars("Temp", T)
        cf = _try(dom, "contforce")
        if cf is not None:
            vectors("ContForce", cf)
        scalars("nod_mass", dom.get("m_mdiag"))
        sig = _try(dom, "m_sigma")
        sig_n = None
        if sig is not None:
            sig_n = nodal_average(dom, sig, 6)
            tr3 = (sig_n[:, 0] + sig_n[:, 1] + sig_n[:, 2]) * (1.0 / 3.0)
            s0, s1, s2 = sig_n[:, 0] - tr3, sig_n[:, 1] - tr3, sig_n[:, 2] - tr3
            j2 = 0.5 * (s0 * s0 + 2.0 * sig_n[:, 3] ** 2 + 2.0 * sig_n[:, 5] ** 2 + s1 * s1 + 2.0 * sig_n[:, 4] ** 2 + s2 * s2)
            scalars("stress", np.sqrt(3.0 * j2))
        ext = _try(dom, "ext_nodes")
        if ext is not None:
            scalars("ext_nodes", (np.asarray(ext).view(np.uint8)[:nn] != 0).astype(np.float64))
        na = _try(dom, "node_area")
        if na is not None:
            scalars("nod_area", na)
        pn = _try(dom, "p_node")
        if pn is not None:
            scalars("nod_p", pn)
        if sig_n is not None:
            tensors("SIGMAT", sig_n)
        sr = _try(dom, "m_str_rate")
        if sr is not None:
            tensors("EPSR", nodal_average(dom, sr, 6))

        o.text(f"CELL_DATA {ne}")
        ea = _try(dom, "m_elem_area")
        if ea is not None:
            scalars("ele_area", ea)
        scalars("pressure", dom.get("p"))
        scalars("pl_strain", dom.get("pl_strain"))
        if sr is not None:
            tensors("DDEVT", sr)
        vol, vol0 = np.asarray(dom.get("vol")), np.asarray(dom.get("vol_0"))
        scalars("J", vol / vol0)
        scalars("Vol", vol)
        scalars("Rho", dom.get("rho"))
        scalars("Vol_0", vol0)
        scalars("sigy", dom.get("sigma_y"))
    return written


def read_vtk(path: str) -> dict:
    """Minimal reader of the files write_vtk produces (both encodings): {'POINTS', 'CELLS', 'CELL_TYPES',
    'POINT_DATA': {name: array}, 'CELL_DATA': {name: array}, 'order': [names]}.  Used by the tests."""
    data = open(path, "rb").read()
    pos = 0

    def line():
        nonlocal pos
        while True:
            e = data.index(b"\n", pos)
            s = data[pos:e].decode("ascii", "replace").strip()
            pos = e + 1
            if s:
                return s

    assert line().startswith("# vtk DataFile")
    line()
    binary = line() == "BINARY"
    assert line() == "DATASET UNSTRUCTURED_GRID"

    def block(count, dtype):
        nonlocal pos
        if binary:
            nb = count * 4
            a = np.frombuffer(data, dtype=">f4" if dtype == "f" else ">i4", count=count, offset=pos).astype(
                np.float64 if dtype == "f" else np.int64)
            pos += nb
            if data[pos:pos + 1] == b"\n":
                pos += 1
            return a
        vals = []
        while len(vals) < count:
            vals += line().split()
        if dtype == "f":   # nine significant digits identify the float32 that was printed
            return np.array(vals, np.float64).astype(np.float32).astype(np.float64)
        return np.array(vals, np.int64)

    out = {"POINT_DATA": {}, "CELL_DATA": {}, "order": []}
    n = int(line().split()[1])
    out["POINTS"] = block(3 * n, "f").reshape(n, 3)
    _, ne, size = line().split()
    out["CELLS"] = block(int(size), "i")
    assert line().split()[0] == "CELL_TYPES"
    out["CELL_TYPES"] = block(int(ne), "i")
    section, count = None, 0
    while pos < len(data):
        try:
            s = line()
        except ValueError:
            break
        t = s.split()
        if t[0] in ("POINT_DATA", "CELL_DATA"):
            section, count = t[0], int(t[1])
        elif t[0] == "SCALARS":
            assert line() == "LOOKUP_TABLE default"
            out[section][t[1]] = block(count, "f")
            out["order"].append(t[1])
        elif t[0] == "VECTORS":
            out[section][t[1]] = block(3 * count, "f").reshape(count, 3)
            out["order"].append(t[1])
        elif t[0] == "TENSORS":
            out[section][t[1]] = block(9 * count, "f").reshape(count, 9)
            out["order"].append(t[1])
    return out
