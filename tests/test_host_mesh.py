"""Integer work, bit-exact (no GPU): the engine's host-side box mesher, node->element connectivity and
partition / halo lists against the oracle and the committed reference fixtures."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from cases_golden import GOLDEN, INT_ARRAYS  # noqa: E402


def host_box(V, L, r, tritet):
    from weldformfem_b200 import _lib
    lib = _lib.load()
    dim, k, nn, ne = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    Ld, Vd = (C.c_double * 3)(*L), (C.c_double * 3)(*V)
    assert lib.wf_host_box_counts(Ld, r, int(tritet), C.byref(dim), C.byref(k), C.byref(nn), C.byref(ne)) == 0
    x = np.empty(nn.value * dim.value)
    el = np.empty(ne.value * k.value, dtype=np.uint32)
    assert lib.wf_host_gen_box(Vd, Ld, r, int(tritet), x.ctypes.data_as(C.POINTER(C.c_double)),
                               el.ctypes.data_as(C.POINTER(C.c_uint))) == 0
    return dim.value, k.value, x, el


def host_nodel(nn, k, el):
    from weldformfem_b200 import _lib
    lib = _lib.load()
    ne = el.size // k
    off, cnt = np.empty(nn, np.int32), np.empty(nn, np.int32)
    nodel, loc = np.empty(ne * k, np.int32), np.empty(ne * k, np.int32)
    ip = C.POINTER(C.c_int)
    rc = lib.wf_host_nodel(nn, ne, k, el.ctypes.data_as(C.POINTER(C.c_uint)), off.ctypes.data_as(ip),
                           cnt.ctypes.data_as(ip), nodel.ctypes.data_as(ip), loc.ctypes.data_as(ip))
    return rc, off, cnt, nodel, loc


BOXES = [((0, 0, 0), (0.3, 0.2, 0.5), 0.05, False), ((0.1, -0.2, 0.3), (0.31, 0.2, 0.11), 0.05, True),
         ((0, 0, 0), (0.0127, 0.03, 0.0), 0.00025, False), ((1, 2, 0), (0.4, 0.3, 0.0), 0.05, True),
         ((0, 0, 0), (0.1, 0.1, 0.1), 0.05, False), ((0, 0, 0), (0.7, 0.1, 0.1), 0.05, True)]


@pytest.mark.parametrize("V,L,r,tritet", BOXES)
def test_box_and_connectivity_match_oracle(V, L, r, tritet, oracle_port):
    dim, k, x, el = host_box(V, L, r, tritet)
    o = oracle_port()
    o.box(V, L, r, tritet)
    info = o.info()
    assert (dim, k) == (info["dim"], info["nodxelem"])
    assert np.array_equal(x, o.get("x"))            # coordinates by accumulation: bit-identical
    assert np.array_equal(el, o.get("m_elnod"))
    rc, off, cnt, nodel, loc = host_nodel(info["n_nodes"], k, el)
    assert rc == 0
    assert np.array_equal(off, o.get("m_nodel_offset"))
    assert np.array_equal(cnt, o.get("m_nodel_count"))
    assert np.array_equal(nodel, o.get("m_nodel"))
    assert np.array_equal(loc, o.get("m_nodel_loc"))


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_connectivity_matches_reference_fixtures(name):
    """Against arrays dumped from the unmodified reference build (tests/golden/make_golden.py)."""
    case, _ = GOLDEN[name]
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    pad = 1.0 + 1.0e-6
    L = [case.n[0] * case.h * pad, case.n[1] * case.h * pad, (case.n[2] * case.h * pad) if case.dim == 3 else 0.0]
    dim, k, x, el = host_box((0, 0, 0), L, 0.5 * case.h, case.tritet)
    assert np.array_equal(x, g["x0"])
    assert np.array_equal(el, g["m_elnod"])
    rc, off, cnt, nodel, loc = host_nodel(x.size // dim, k, el)
    assert rc == 0
    for nm, arr in zip(("m_nodel_offset", "m_nodel_count", "m_nodel", "m_nodel_loc"), (off, cnt, nodel, loc)):
        assert np.array_equal(arr, g[nm]), nm


def test_nodel_rejects_out_of_range():
    el = np.array([0, 1, 2, 9], dtype=np.uint32)
    rc, *_ = host_nodel(4, 4, el)
    assert rc != 0


def test_nodel_ragged_unstructured():
    """Random tet soup incl. unused nodes: lists sorted by element id, offsets = exclusive prefix sum."""
    rng = np.random.default_rng(7)
    nn, ne, k = 57, 200, 4
    el = np.stack([rng.choice(nn - 5, size=k, replace=False) for _ in range(ne)]).astype(np.uint32).reshape(-1)
    rc, off, cnt, nodel, loc = host_nodel(nn, k, el)
    assert rc == 0
    assert cnt[-5:].sum() == 0 and cnt.sum() == ne * k
    assert np.array_equal(off, np.concatenate([[0], np.cumsum(cnt)[:-1]]))
    for n in range(nn):
        lst = nodel[off[n]:off[n] + cnt[n]]
        assert np.all(np.diff(lst) > 0)
        for e, ln in zip(lst, loc[off[n]:off[n] + cnt[n]]):
            assert el[e * k + ln] == n


# ---- partition / halo lists (canonical definition) -------------------------------------
def py_partition(P, p, k, nn, el):
    """Independent numpy restatement of the canonical partition."""
    ne = el.size // k
    el = el.reshape(ne, k).astype(np.int64)
    beg = [(ne * q) // P for q in range(P + 1)]
    mine = el[beg[p]:beg[p + 1]]
    l2g = np.unique(mine)
    lel = np.searchsorted(l2g, mine).astype(np.uint32)
    neigh, offs, halo = [], [0], []
    for q in range(P):
        if q == p:
            continue
        shared = np.intersect1d(l2g, np.unique(el[beg[q]:beg[q + 1]]))
        if shared.size:
            neigh.append(q)
            halo.extend(np.searchsorted(l2g, shared).tolist())
            offs.append(len(halo))
    return beg[p], beg[p + 1], l2g.astype(np.int32), lel.reshape(-1), neigh, offs, halo


def c_partition(P, p, k, nn, el, box=None):
    from weldformfem_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    if box is None:
        assert lib.wf_partition_build(C.byref(h), P, p, k, nn, el.size // k, el.ctypes.data_as(C.POINTER(C.c_uint))) == 0
    else:
        V, L, r, tritet = box
        assert lib.wf_partition_build_box(C.byref(h), P, p, (C.c_double * 3)(*V), (C.c_double * 3)(*L), r, int(tritet)) == 0
    eb, ee, nl, nng = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    assert lib.wf_partition_info(h, C.byref(eb), C.byref(ee), C.byref(nl), C.byref(nng)) == 0
    l2g = np.ctypeslib.as_array(lib.wf_partition_node_l2g(h), (nl.value,)).copy()
    lel = np.ctypeslib.as_array(lib.wf_partition_local_elnod(h), ((ee.value - eb.value) * k,)).copy()
    neigh = list(np.ctypeslib.as_array(lib.wf_partition_neigh_ranks(h), (max(nng.value, 1),))[:nng.value])
    offs = list(np.ctypeslib.as_array(lib.wf_partition_halo_offset(h), (nng.value + 1,)))
    halo = list(np.ctypeslib.as_array(lib.wf_partition_halo_nodes(h), (max(offs[-1], 1),))[:offs[-1]])
    lib.wf_partition_free(h)
    return eb.value, ee.value, l2g, lel, neigh, offs, halo


@pytest.mark.parametrize("P", [2, 3, 4, 8])
@pytest.mark.parametrize("box", [((0, 0, 0), (0.5, 0.4, 0.9), 0.05, False), ((0, 0, 0), (0.3, 0.3, 0.4), 0.05, True),
                                 ((0, 0, 0), (0.9, 0.7, 0.0), 0.05, False)])
def test_partition_bit_exact(P, box):
    dim, k, x, el = host_box(*box)
    nn = x.size // dim
    all_l2g = []
    halos = {}
    for p in range(P):
        want = py_partition(P, p, k, nn, el)
        got = c_partition(P, p, k, nn, el)
        got_box = c_partition(P, p, k, nn, el, box=box)
        for g in (got, got_box):
            assert g[0] == want[0] and g[1] == want[1]
            assert np.array_equal(g[2], want[2]) and np.array_equal(g[3], want[3])
            assert [int(q) for q in g[4]] == want[4] and [int(q) for q in g[5]] == want[5]
            assert [int(q) for q in g[6]] == want[6]
        all_l2g.append(got[2])
        for i, q in enumerate(got[4]):
            halos[(p, int(q))] = got[2][np.array(got[6][got[5][i]:got[5][i + 1]], dtype=int)]
    # every node is local somewhere; halo lists are identical on both sides (global ids, ascending)
    assert np.array_equal(np.unique(np.concatenate(all_l2g)), np.arange(nn))
    for (p, q), ids in halos.items():
        assert np.array_equal(ids, halos[(q, p)])
        assert np.all(np.diff(ids) > 0)


# ---- contact: integer artefacts of SearchExtNodes and the rigid-plane mesher -----------------
def host_ext_faces(dim, k, nn, el):
    from weldformfem_b200 import _lib
    lib = _lib.load()
    ne = el.size // k
    facenod = 3 if dim == 3 else 2
    ext = np.zeros(nn, np.uint8)
    fn, fe = np.full(ne * 4 * facenod, -1, np.int32), np.full(ne * 4, -1, np.int32)
    tot, nx = C.c_int(), C.c_int()
    ip = C.POINTER(C.c_int)
    rc = lib.wf_host_ext_faces(dim, k, nn, ne, el.ctypes.data_as(C.POINTER(C.c_uint)), ext.ctypes.data_as(C.POINTER(C.c_ubyte)),
                               C.byref(tot), C.byref(nx), fn.ctypes.data_as(ip), fe.ctypes.data_as(ip))
    return rc, ext, tot.value, fn[: nx.value * facenod].reshape(-1, facenod), fe[: nx.value]


def brute_force_faces(dim, k, el):
    """SearchExtNodes as the reference does it (Domain_d.C:74-156): linear search of every new face in the list."""
    tab = [(0, 1, 2), (0, 1, 3), (1, 2, 3), (0, 2, 3)] if dim == 3 else [(0, 1), (1, 2), (2, 3), (3, 0)]
    faces, count, elem = [], [], []
    for e, nodes in enumerate(el.reshape(-1, k)):
        for f in tab:
            fn = [int(nodes[q]) for q in f]
            key = sorted(fn)
            for i, g in enumerate(faces):
                if sorted(g) == key:
                    count[i] += 1
                    break
            else:
                faces.append(fn); count.append(1); elem.append(e)
    ext = [(f, e) for f, c, e in zip(faces, count, elem) if c == 1]
    return len(faces), np.array([f for f, _ in ext], np.int32), np.array([e for _, e in ext], np.int32)


@pytest.mark.parametrize("V,L,r,tritet", [((0, 0, 0), (0.31, 0.2, 0.11), 0.05, True), ((0, 0, 0), (0.2, 0.2, 0.2), 0.05, True),
                                          ((0, 0, 0), (0.4, 0.3, 0.0), 0.05, False), ((0, 0, 0), (0.1, 0.1, 0.0), 0.05, False)])
def test_ext_faces_match_reference_search_and_oracle(V, L, r, tritet, oracle_port):
    dim, k, x, el = host_box(V, L, r, tritet)
    nn = x.size // dim
    rc, ext, tot, fn, fe = host_ext_faces(dim, k, nn, el)
    assert rc == 0
    btot, bfn, bfe = brute_force_faces(dim, k, el)
    assert tot == btot
    assert np.array_equal(fn, bfn) and np.array_equal(fe, bfe)       # same faces, same (faceList) order
    o = oracle_port()
    o.box(V, L, r, tritet)
    o.call("SearchExtNodes")
    assert np.array_equal(ext, o.get("ext_nodes"))
    # a box: every boundary node is external, no interior node is
    X = x.reshape(-1, dim)
    onb = np.zeros(nn, bool)
    for c in range(dim):
        onb |= np.isclose(X[:, c], X[:, c].min()) | np.isclose(X[:, c], X[:, c].max())
    assert np.array_equal(ext.astype(bool), onb)


def test_ext_faces_refuses_elements_the_reference_tables_do_not_cover():
    dim, k, x, el = host_box((0, 0, 0), (0.2, 0.2, 0.2), 0.05, False)   # hexahedra
    rc, *_ = host_ext_faces(dim, k, x.size // dim, el)
    assert rc != 0


@pytest.mark.parametrize("dimension,axis,orient,dens", [(3, 2, False, 4), (3, 2, True, 1), (3, 0, True, 3), (2, 1, False, 5), (2, 0, True, 2)])
def test_axis_plane_mesh_matches_oracle(dimension, axis, orient, dens, oracle_port):
    from weldformfem_b200.domain import axis_plane_mesh
    p1, p2 = (-0.013, 0.02, 0.031), (0.027, 0.06, 0.031)
    node, elnode, normal, mid = axis_plane_mesh(dimension, 7, axis, orient, p1, p2, dens)
    o = oracle_port()
    if dimension == 3:
        o.box((0, 0, 0), (0.2, 0.2, 0.2), 0.05, True)
    else:
        o.set_domtype(0, False)
        o.box((0, 0, 0), (0.2, 0.2, 0.0), 0.05, False)
    o.add_plane(dimension, 7, axis, orient, p1, p2, dens, (0.0, 0.0, -1.0))
    assert np.array_equal(node.reshape(-1), o.get("trimesh.node"))
    assert np.array_equal(elnode.reshape(-1), o.get("trimesh.elnode"))
    assert np.array_equal(normal.reshape(-1), o.get("trimesh.normal"))
    assert np.array_equal(mid, o.get("trimesh.ele_mesh_id"))


def test_axis_plane_mesh_matches_compiled_reference(oracle_ref):
    """TriMesh_d::AxisPlaneMesh + AddMesh of the unmodified reference against the host mirror's add_plane merge."""
    from weldformfem_b200.domain import axis_plane_mesh
    r = oracle_ref()
    r.box((0, 0, 0), (0.2, 0.2, 0.2), 0.05, True)
    specs = [(0, 2, False, (-0.1, -0.1, 0.21), (0.3, 0.3, 0.21), 4, (0.5, 0.0, -2.0)),
             (1, 2, True, (-0.1, -0.1, -0.01), (0.3, 0.3, -0.01), 2, (0.0, 0.0, 0.0))]
    nodes, els, vels, ids, off = [], [], [], [], 0
    for mid, axis, orient, p1, p2, dens, vel in specs:
        r.add_plane(3, mid, axis, orient, p1, p2, dens, vel)
        n, e, _, m = axis_plane_mesh(3, mid, axis, orient, p1, p2, dens)
        nodes.append(n); els.append(e + off); ids.append(m); vels.append(np.tile(vel, (n.shape[0], 1)))
        off += n.shape[0]
    assert np.array_equal(np.concatenate(nodes).reshape(-1), r.get("trimesh.node"))
    assert np.array_equal(np.concatenate(els).reshape(-1), r.get("trimesh.elnode"))
    assert np.array_equal(np.concatenate(ids), r.get("trimesh.ele_mesh_id"))
    assert np.array_equal(np.concatenate(vels).reshape(-1), r.get("trimesh.node_v"))


# ---- force tiles of the tile-reduced force path (csrc/wf_mesh.cpp: wf_force_tiles_build) ----------------
def _force_tiles_lib(nn, el, dim=3):
    import ctypes as C
    from weldformfem_b200 import _lib
    lib = _lib.load()
    ne, k = el.shape
    elc = np.ascontiguousarray(el, dtype=np.uint32)
    up = elc.ctypes.data_as(C.POINTER(C.c_uint))
    info = (C.c_longlong * 7)()
    assert lib.wf_host_force_tiles(nn, ne, k, dim, up, info, None, None, None, None) == 0
    usable, ntile, stride, tpitch, nsl, nslots, rounds = list(info)
    out = {"usable": bool(usable), "n_tiles": ntile, "stride": stride, "tpitch": tpitch, "rounds": bool(rounds)}
    if not usable:
        return out
    tidx = np.zeros(ne * k, dtype=np.uint8)
    ptr = np.zeros(nsl + 1, dtype=np.int64)
    slots = np.zeros(max(nslots, 1), dtype=np.uint32)
    tab = np.zeros(max(ntile * tpitch, 1), dtype=np.uint8)
    assert lib.wf_host_force_tiles(nn, ne, k, dim, up, info, tidx.ctypes.data_as(C.POINTER(C.c_ubyte)),
                                   ptr.ctypes.data_as(C.POINTER(C.c_longlong)), slots.ctypes.data_as(C.POINTER(C.c_uint)),
                                   tab.ctypes.data_as(C.POINTER(C.c_ubyte)) if tpitch else None) == 0
    out.update(tidx=tidx.reshape(ne, k), ptr=ptr, slots=slots[:nslots], tab=tab[:ntile * tpitch].reshape(ntile, -1) if tpitch else None)
    return out


def _force_tiles_numpy(nn, el, dim=3):
    """Restatement: tile w = elements [32w, 32w+32); unique nodes ascending; hexahedra need distinct nodes per corner
    within a tile; node entries in ascending tile order, sliced-ELL over 32-node slices; tet incidence CSR."""
    ne, k = el.shape
    ntile = (ne + 31) // 32
    uniq = [np.unique(el[32 * w:32 * w + 32]) for w in range(ntile)]
    hexa = k == 8
    rounds = all(len(np.unique(el[32 * w:32 * w + 32][:, n])) == len(el[32 * w:32 * w + 32]) for w in range(ntile) for n in range(k))
    if hexa and not rounds:
        return {"usable": False}
    stride = (max(len(u) for u in uniq) + 3) // 4 * 4
    tidx = np.vstack([np.searchsorted(uniq[w], el[32 * w:32 * w + 32]) for w in range(ntile)]).astype(np.uint8)
    entries = [[] for _ in range(nn)]
    for w in range(ntile):
        for i, g in enumerate(uniq[w]):
            entries[g].append(w * dim * stride + i)
    nsl = (nn + 31) // 32
    ptr = np.zeros(nsl + 1, dtype=np.int64)
    for s in range(nsl):
        ptr[s + 1] = ptr[s] + 32 * max(len(entries[n]) for n in range(32 * s, min(nn, 32 * s + 32)))
    slots = np.full(ptr[-1], 0xFFFFFFFF, dtype=np.uint32)
    for n in range(nn):
        for j, off in enumerate(entries[n]):
            slots[ptr[n >> 5] + 32 * j + (n & 31)] = off
    out = {"usable": True, "n_tiles": ntile, "stride": stride, "tidx": tidx, "ptr": ptr, "slots": slots, "tpitch": 0, "tab": None,
           "rounds": rounds}
    if not hexa:
        tpitch = (stride + 1 + 32 * k + 3) // 4 * 4
        tab = np.zeros((ntile, tpitch), dtype=np.uint8)
        for w in range(ntile):
            t = tidx[32 * w:32 * w + 32].ravel()                 # order: ascending element, then corner
            order = np.argsort(t, kind="stable")
            cnt = np.bincount(t, minlength=stride)
            tab[w, 1:stride + 1] = np.cumsum(cnt)[:stride]
            tab[w, stride + 1:stride + 1 + len(t)] = order
        out.update(tpitch=tpitch, tab=tab)
    return out


@pytest.mark.parametrize("kind,n,shuffle", [("hex", (5, 4, 7), False), ("hex", (33, 2, 2), False), ("hex", (4, 4, 4), True),
                                            ("tet", (3, 4, 5), False), ("tet", (4, 3, 3), True),
                                            ("quad", (37, 9), False), ("quad", (8, 8), True), ("tri", (9, 7), False)])
def test_force_tile_tables_bit_exact(kind, n, shuffle):
    h = 0.01
    L = [(q + 1e-6) * h for q in n] + [0.0] * (3 - len(n))
    dim, k, x, el = host_box((0.0, 0.0, 0.0), L, 0.5 * h, kind in ("tet", "tri"))
    nn = len(x) // dim
    el = el.reshape(-1, k).astype(np.int64)
    if shuffle:
        rng = np.random.default_rng(7)
        el = rng.permutation(nn)[el][rng.permutation(len(el))]
    got = _force_tiles_lib(nn, el, dim)
    want = _force_tiles_numpy(nn, el, dim)
    assert got["usable"] == want["usable"]
    if kind == "hex":
        # in a box mesh every node is corner n of exactly ONE element, so any element order is conflict-free ...
        assert got["usable"]
        # ... until elements are relabelled: the same hexahedra with every other one rotated about its axis
        el2 = el.copy()
        el2[1::2] = el2[1::2][:, [1, 2, 3, 0, 5, 6, 7, 4]]
        g2, w2 = _force_tiles_lib(nn, el2), _force_tiles_numpy(nn, el2)
        assert g2["usable"] == w2["usable"] and not g2["usable"]
    if not want["usable"]:
        return
    assert (got["n_tiles"], got["stride"], got["tpitch"], got["rounds"]) == (want["n_tiles"], want["stride"], want["tpitch"], want["rounds"])
    assert got["rounds"] == (kind in ("hex", "quad"))   # box meshes: one element per (node, corner) for hexes / quads only
    for nm in ("tidx", "ptr", "slots"):
        assert np.array_equal(got[nm], want[nm]), nm
    if kind != "hex":
        assert np.array_equal(got["tab"], want["tab"])
    # every element node is reachable: slot -> (tile, position) -> unique list -> node
    ne, k = el.shape
    cover = np.zeros(nn, dtype=np.int64)
    valid = got["slots"][got["slots"] != 0xFFFFFFFF]
    assert len(valid) == len(set(valid.tolist()))
    for w in range(got["n_tiles"]):
        cover[np.unique(el[32 * w:32 * w + 32])] += 1
    per_node = np.zeros(nn, dtype=np.int64)
    for n_ in range(nn):
        base = got["ptr"][n_ >> 5]
        width = (got["ptr"][(n_ >> 5) + 1] - base) // 32
        per_node[n_] = sum(got["slots"][base + 32 * j + (n_ & 31)] != 0xFFFFFFFF for j in range(width))
    assert np.array_equal(per_node, cover)


# ---- internal element order (wf_host_elem_order) -------------------------------------------------
def host_elem_order(dim, k, x, el, mode=1):
    from weldformfem_b200 import _lib
    lib = _lib.load()
    nn, ne = x.size // dim, el.size // k
    perm = np.empty(ne, np.int32)
    rc = lib.wf_host_elem_order(dim, k, nn, ne, np.ascontiguousarray(x).ctypes.data_as(C.POINTER(C.c_double)),
                                np.ascontiguousarray(el).ctypes.data_as(C.POINTER(C.c_uint)), mode,
                                perm.ctypes.data_as(C.POINTER(C.c_int)))
    return rc, perm


def numpy_elem_order(dim, k, x, el):
    """Restatement of the comment above wf_host_elem_order (csrc/wf_mesh.cpp), same operation order."""
    import math
    X = x.reshape(-1, dim)
    EL = el.reshape(-1, k).astype(np.int64)
    ne = EL.shape[0]
    lo, hi = X.min(0), X.max(0)
    ext = hi - lo
    per_cell = 1 if k == 8 else (6 if dim == 3 else (1 if k == 4 else 2))
    cells = max(1, ne // per_cell)
    act = [c for c in range(dim) if ext[c] > 0.0]
    ncell = [1] * dim
    if act:
        vol = 1.0
        for c in act:
            vol *= ext[c]
        h = math.pow(vol / float(cells), 1.0 / float(len(act)))
        for c in act:
            ncell[c] = max(1, int(ext[c] / h + 0.5))
    q = np.zeros((dim, ne), np.uint64)
    for c in act:
        s = np.zeros(ne)
        for a in range(k):                       # same summation order as the C loop
            s = s + X[EL[:, a], c]
        t = (s / float(k) - lo[c]) / ext[c] * float(ncell[c])
        q[c] = np.clip(t.astype(np.int64), 0, ncell[c] - 1).astype(np.uint64)

    def interleave(vals, bits):
        key = np.zeros(ne, np.uint64)
        nd = len(vals)
        for c, v in enumerate(vals):
            for b in range(bits):
                key |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(nd * b + c)
        return key
    if k == 8 and dim == 3:     # tiles of 4x4x2 in Morton order (z lowest), inside a tile layer / row / column
        U = np.uint64
        tkey = interleave([q[2] >> U(1), q[0] >> U(2), q[1] >> U(2)], 21)
        key = (tkey << U(5)) | ((q[2] & U(1)) << U(4)) | ((q[1] & U(3)) << U(2)) | (q[0] & U(3))
    else:
        key = interleave([q[c] for c in range(dim)], 21 if dim == 3 else 31)
    return np.lexsort((np.arange(ne), key)).astype(np.int32)


@pytest.mark.parametrize("V,L,r,tritet", BOXES + [((0, 0, 0), (1.3, 1.1, 0.9), 0.05, False)])
def test_elem_order_bit_exact_and_is_a_permutation(V, L, r, tritet):
    dim, k, x, el = host_box(V, L, r, tritet)
    rc, perm = host_elem_order(dim, k, x, el)
    assert rc == 0
    assert np.array_equal(np.sort(perm), np.arange(perm.size))
    assert np.array_equal(perm, numpy_elem_order(dim, k, x, el))
    rc, ident = host_elem_order(dim, k, x, el, mode=0)
    assert rc == 0 and np.array_equal(ident, np.arange(perm.size))


def test_elem_order_makes_bricks():
    """On a structured hexa box 32 consecutive elements become a 4x4x2 brick (75 distinct nodes instead of 132)
    and 128 an 8x4x4 brick (225 instead of ~516): that is what the staging / force-partial traffic scales with."""
    dim, k, x, el = host_box((0, 0, 0), (3.2, 3.2, 3.2), 0.05, False)   # 32^3 hexes
    EL = el.reshape(-1, 8)
    rc, perm = host_elem_order(dim, k, x, el)
    assert rc == 0

    def distinct(order, chunk):
        a = np.sort(EL[order].reshape(-1, chunk * 8), 1)
        return 1 + (np.diff(a, axis=1) != 0).sum(1)
    assert distinct(np.arange(EL.shape[0]), 32).mean() > 130
    assert (distinct(perm, 32) == 75).all()
    assert (distinct(perm, 128) == 225).all()


def test_elem_order_rejects_bad_connectivity():
    x = np.zeros(9)
    el = np.array([0, 1, 7], np.uint32)
    rc, _ = host_elem_order(3, 3, x, el)   # dim 3 / k 3 is not an element type, but only the range check matters here
    assert rc != 0


def host_run_slots(ids):
    from weldformfem_b200 import _lib
    lib = _lib.load()
    ids = np.ascontiguousarray(ids, np.int32)
    out = np.empty(ids.size, np.int32)
    n = lib.wf_host_run_slots(ids.size, ids.ctypes.data_as(C.POINTER(C.c_int)), out.ctypes.data_as(C.POINTER(C.c_int)))
    return n, out


def numpy_run_slots(ids):
    slots, cur, r, a = [], 0, 0, 0
    while a < len(ids):
        b = a
        while b + 1 < len(ids) and ids[b + 1] == ids[b] + 1:
            b += 1
        while cur % 16 != (12 * r) % 16:
            cur += 1
        slots += list(range(cur, cur + b - a + 1))
        cur += b - a + 1
        r += 1
        a = b + 1
    return cur, np.array(slots, np.int32)


def test_run_slots_bit_exact_and_conflict_free_on_bricks():
    """wf_host_run_slots against its restatement, and the property it exists for: in the engine's hexa element order
    the 16 lanes of a half-warp (one 4x4 layer of a tile) read, for each of the eight corners, 16 nodes whose slots
    fall into 16 different 8-byte bank pairs — for the CTA copy (128 elements) and the tile accumulators (32)."""
    rng = np.random.default_rng(5)
    for _ in range(20):
        ids = np.unique(rng.integers(0, 400, rng.integers(1, 300)))
        n, sl = host_run_slots(ids)
        n2, sl2 = numpy_run_slots(list(ids))
        assert n == n2 and np.array_equal(sl, sl2)
        assert np.all(np.diff(sl) > 0)
    dim, k, x, el = host_box((0, 0, 0), (2.4001, 2.4001, 2.4001), 0.05, False)   # 24^3 hexes
    assert el.size == 8 * 24 ** 3
    EL = el.reshape(-1, 8).astype(np.int64)
    rc, perm = host_elem_order(dim, k, x, el)
    assert rc == 0
    ELi = EL[perm]
    for chunk, limit in ((128, 304), (32, 176)):
        for c0 in range(0, ELi.shape[0], chunk):
            ids = np.unique(ELi[c0:c0 + chunk])
            n, sl = host_run_slots(ids)
            assert n <= limit
            slot = dict(zip(ids.tolist(), sl.tolist()))
            for h0 in range(c0, c0 + chunk, 16):
                for c in range(8):
                    banks = [slot[g] % 16 for g in ELi[h0:h0 + 16, c]]
                    assert len(set(banks)) == 16


# ---- thread slots of the brick passes (wf_host_brick_plan) ----------------------------------------------------
def host_elem_order_keys(dim, k, x, el):
    from weldformfem_b200 import _lib
    lib = _lib.load()
    nn, ne = x.size // dim, el.size // k
    perm = np.empty(ne, np.int32)
    keys = np.empty(ne, np.uint64)
    rc = lib.wf_host_elem_order_keys(dim, k, nn, ne, np.ascontiguousarray(x).ctypes.data_as(C.POINTER(C.c_double)),
                                     np.ascontiguousarray(el).ctypes.data_as(C.POINTER(C.c_uint)), 1,
                                     perm.ctypes.data_as(C.POINTER(C.c_int)), keys.ctypes.data_as(C.POINTER(C.c_ulonglong)))
    return rc, perm, keys


def host_brick_plan(keys):
    from weldformfem_b200 import _lib
    lib = _lib.load()
    keys = np.ascontiguousarray(keys, np.uint64)
    n = C.c_int(0)
    kp = keys.ctypes.data_as(C.POINTER(C.c_ulonglong))
    rc = lib.wf_host_brick_plan(keys.size, kp, C.byref(n), None)
    if rc:
        return rc, 0, None
    slot = np.empty(n.value * 128, np.int32)
    rc = lib.wf_host_brick_plan(keys.size, kp, C.byref(n), slot.ctypes.data_as(C.POINTER(C.c_int)))
    return rc, n.value, slot


def test_brick_plan_bit_exact_and_every_cta_is_a_clipped_brick():
    """A box whose edge counts are no multiples of the brick (23 x 21 x 19 hexes): in the compact numbering the chunks
    of 128 drift across brick boundaries; with the plan every CTA holds the cells of ONE 8x4x4 group, its node list
    fits the bank-aware layout and every half-warp reads 16 different bank pairs per corner."""
    dim, k, x, el = host_box((0, 0, 0), (1.1501, 1.0501, 0.9501), 0.025, False)
    assert el.size == 8 * 23 * 21 * 19
    rc, perm, keys = host_elem_order_keys(dim, k, x, el)
    assert rc == 0 and np.array_equal(perm, numpy_elem_order(dim, k, x, el))
    assert np.all(np.diff(keys.astype(np.int64)) > 0)
    rc, n_cta, slot = host_brick_plan(keys)
    assert rc == 0
    # restatement: CTA = rank of key >> 7, thread = key & 127
    grp = (keys >> np.uint64(7)).astype(np.int64)
    cta = np.concatenate([[0], np.cumsum(np.diff(grp) != 0)])
    want = np.full((cta[-1] + 1) * 128, -1, np.int32)
    want[cta * 128 + (keys & np.uint64(127)).astype(np.int64)] = np.arange(keys.size)
    assert n_cta == cta[-1] + 1 == 3 * 6 * 5 and np.array_equal(slot, want)
    ELi = el.reshape(-1, 8).astype(np.int64)[perm]
    ragged_compact = 0
    for c0 in range(0, ELi.shape[0], 128):
        n, _ = host_run_slots(np.unique(ELi[c0:c0 + 128]))
        ragged_compact += n > 304
    assert ragged_compact > 0                      # what the plan is for
    for b in range(n_cta):
        s = slot[b * 128:(b + 1) * 128]
        ids = np.unique(ELi[s[s >= 0]])
        n, sl = host_run_slots(ids)
        assert n <= 304
        lut = dict(zip(ids.tolist(), sl.tolist()))
        for h0 in range(0, 128, 16):
            lanes = s[h0:h0 + 16]
            lanes = lanes[lanes >= 0]
            for c in range(8):
                banks = [lut[g] % 16 for g in ELi[lanes, c]]
                assert len(set(banks)) == len(banks)
        for w in range(4):                          # tile accumulators
            lanes = s[32 * w:32 * w + 32]
            lanes = lanes[lanes >= 0]
            if lanes.size == 0:
                continue
            ids = np.unique(ELi[lanes])
            n, sl = host_run_slots(ids)
            assert n <= 176 and ids.size <= 75


def test_brick_plan_refuses_two_elements_in_one_cell():
    keys = np.array([5, 9, 9, 300], np.uint64)
    rc, _, _ = host_brick_plan(keys)
    assert rc != 0
    dim, k, x, el = host_box((0, 0, 0), (0.3, 0.3, 0.3), 0.05, True)      # tets: six per cell, no hexa keys at all
    from weldformfem_b200 import _lib
    lib = _lib.load()
    perm = np.empty(el.size // k, np.int32)
    keys = np.empty(el.size // k, np.uint64)
    assert lib.wf_host_elem_order_keys(dim, k, x.size // dim, el.size // k, x.ctypes.data_as(C.POINTER(C.c_double)),
                                       el.ctypes.data_as(C.POINTER(C.c_uint)), 1, perm.ctypes.data_as(C.POINTER(C.c_int)),
                                       keys.ctypes.data_as(C.POINTER(C.c_ulonglong))) == 0
    assert host_brick_plan(keys)[0] != 0
