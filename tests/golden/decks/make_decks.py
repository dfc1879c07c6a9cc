"""Writes the input decks of the front-end parity tests (tests/test_deck.py): small JSON decks in the reference's
format (keys as read by src/explicit/main.C:191-975) plus two LS-Dyna `.k` meshes in the card layout of
examples/input/tetra_cyl.k / cyl_hex.k (8-column ids, 16-column coordinates, tetrahedra padded to 8 slots by
repeating the last node; node ids are 1-based and deliberately NOT contiguous in the tet file).  The meshes are
structured blocks from numpy, so nothing here is derived from reference data files.

    python tests/golden/decks/make_decks.py
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def block(nx, ny, nz, h, tets):
    xs = [(i * h, j * h, k * h) for k in range(nz + 1) for j in range(ny + 1) for i in range(nx + 1)]
    nid = lambda i, j, k: i + (nx + 1) * (j + (ny + 1) * k)
    el = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                c = [nid(i, j, k), nid(i + 1, j, k), nid(i + 1, j + 1, k), nid(i, j + 1, k),
                     nid(i, j, k + 1), nid(i + 1, j, k + 1), nid(i + 1, j + 1, k + 1), nid(i, j + 1, k + 1)]
                if not tets:
                    el.append(c)
                else:   # 5-tet split, alternating parity so faces match
                    if (i + j + k) % 2 == 0:
                        t = [(0, 1, 3, 4), (1, 2, 3, 6), (1, 4, 5, 6), (3, 4, 6, 7), (1, 3, 4, 6)]
                    else:
                        t = [(0, 1, 2, 5), (0, 2, 3, 7), (0, 4, 5, 7), (2, 5, 6, 7), (0, 2, 5, 7)]
                    el += [[c[a] for a in q] for q in t]
    xs, el = np.array(xs), np.array(el)
    if tets:   # positive Jacobian in the reference's convention (rows x_1-x_0, x_2-x_0, x_3-x_0; Domain_d.C:1866-1887)
        q = xs[el]
        neg = np.linalg.det(q[:, 1:] - q[:, :1]) < 0
        el[neg] = el[neg][:, [0, 2, 1, 3]]
    return xs, el


def write_k(path, x, el, id_of):
    with open(path, "w") as f:
        f.write("$# structured block written by tests/golden/decks/make_decks.py\n*KEYWORD\n*NODE\n")
        f.write("$#   nid               x               y               z      tc      rc\n")
        for n, p in enumerate(x):
            f.write("%8d%16.9g%16.9g%16.9g%8d%8d\n" % (id_of(n), p[0], p[1], p[2], 0, 0))
        f.write("*ELEMENT_SOLID\n$#   eid     pid      n1      n2      n3      n4      n5      n6      n7      n8\n")
        for e, c in enumerate(el):
            ids = [id_of(n) for n in c]
            ids += [ids[-1]] * (8 - len(ids))
            f.write("%8d%8d" % (e + 1, 1) + "".join("%8d" % q for q in ids) + "\n")
        f.write("*END\n")


ALU = {"id": "Solid", "type": "Hollomon", "const": [386.796e6, 0.154], "density0": 2700.0, "youngsModulus": 68.9e9,
       "poissonsRatio": 0.3, "yieldStress0": 190.4e6}


def decks():
    out = {}
    # 1. axisymmetric quads from a Box block, BC zones, Stabilization block (Compression_axisymm_quad.json scaled down)
    out["box_axiquad"] = {
        "Configuration": {"Nproc": 1, "cflFactor": 0.3, "simTime": 1.0e-4, "outTime": 1.0e-3, "domType": "AxiSymm",
                          "AxiSymmVol": False, "artifViscCoeffs": [0.0, 0.0, 0.0], "fixedTS": True},
        "Stabilization": {"hg_visc": 0.1, "hg_stiff": 0.1},
        "Materials": [ALU],
        "DomainBlocks": [{"type": "Box", "zoneId": 0, "start": [0.0, 0.0, 0.0], "dim": [0.0081, 0.0121, 0.0],
                          "elemLength": 0.001}],
        "BoundaryConditions": [
            {"zoneId": 1, "valueType": 0, "value": [0.0, 0.0, 0.0], "start": [-1.0, -1.0e-4, -1.0], "end": [1.0, 1.0e-4, 1.0]},
            {"zoneId": 2, "valueType": 0, "value": [0.0, -40.0, 0.0], "start": [-1.0, 0.0119, -1.0], "end": [1.0, 0.0125, 1.0]}],
    }
    # 2. plane-strain triangles from a Box block, bilinear material, no Stabilization block, x symmetry plane
    out["box_pstri"] = {
        "Configuration": {"Nproc": 1, "cflFactor": 0.2, "simTime": 1.0e-4, "outTime": 1.0e-3, "domType": "plStrain",
                          "xSymm": True, "symtol": 1.0e-5, "artifViscCoeffs": [0.0, 0.0, 0.0]},
        "Materials": [{"id": "Solid", "type": "Bilinear", "const": [1.0e9], "density0": 7850.0, "youngsModulus": 200.0e9,
                       "poissonsRatio": 0.3, "yieldStress0": 300.0e6}],
        "DomainBlocks": [{"type": "Box", "zoneId": 0, "start": [0.0, 0.0, 0.0], "dim": [0.0101, 0.0061, 0.0],
                          "elemLength": 0.001, "elemType": "TriTet"}],
        "BoundaryConditions": [
            {"zoneId": 1, "valueType": 0, "value": [0.0, 0.0, 0.0], "start": [-1.0, -1.0e-4, -1.0], "end": [1.0, 1.0e-4, 1.0]},
            {"zoneId": 2, "valueType": 0, "value": [0.0, 30.0, 0.0], "start": [-1.0, 0.0059, -1.0], "end": [1.0, 0.0065, 1.0]}],
    }
    # 3. plane-strain quads, pressure algorithm 1, artificial viscosity
    out["box_psquad"] = {
        "Configuration": {"Nproc": 1, "cflFactor": 0.25, "simTime": 1.0e-4, "outTime": 1.0e-3, "domType": "plStrain",
                          "artifViscCoeffs": [0.3, 0.03, 0.0]},
        "Stabilization": {"hg_visc": 0.05, "hg_stiff": 0.02},
        "Materials": [ALU],
        "DomainBlocks": [{"type": "Box", "zoneId": 0, "start": [0.0, 0.0, 0.0], "dim": [0.0101, 0.0061, 0.0],
                          "elemLength": 0.001}],
        "BoundaryConditions": [
            {"zoneId": 1, "valueType": 0, "value": [0.0, 0.0, 0.0], "start": [-1.0, -1.0e-4, -1.0], "end": [1.0, 1.0e-4, 1.0]},
            {"zoneId": 2, "valueType": 0, "value": [0.0, -30.0, 0.0], "start": [-1.0, 0.0059, -1.0], "end": [1.0, 0.0065, 1.0]}],
    }
    # 4. tetrahedra from a .k file between two rigid planes: contact + friction, thermal coupling, x/y symmetry planes
    out["file_tet_contact"] = {
        "Configuration": {"Nproc": 1, "cflFactor": 0.15, "simTime": 2.0e-5, "outTime": 1.0e-3, "thermal": True,
                          "plHeatFrac": 0.9, "xSymm": True, "ySymm": True, "artifViscCoeffs": [0.0, 0.0, 0.0]},
        "Stabilization": {"alpha_free": 0.0, "alpha_contact": 0.0, "log_factor": 0.0, "J_min": 0.0},
        "Materials": [dict(ALU, thermalHeatCap=875.0, thermalCond=190.0, thermalExp=2.3e-5)],
        "DomainBlocks": [{"type": "File", "fileName": "tet_block.k", "zoneId": 0}],
        "Contact": [{"fricCoeffStatic": 0.3, "fricCoeffDynamic": 0.2, "penaltyFactor": 0.6, "heatCondCoeff": 25000.0,
                     "dieTemp": 40.0}],
        "RigidBodies": [
            {"type": "Plane", "zoneId": 10, "start": [-0.002, -0.002, 0.006], "flipnormals": True, "partSide": 6,
             "dim": [0.010, 0.010, 0.0]},
            {"type": "Plane", "zoneId": 1, "start": [-0.002, -0.002, 0.0], "partSide": 6, "dim": [0.010, 0.010, 0.0]}],
        "BoundaryConditions": [
            {"zoneId": 1, "valueType": 0, "value": [0.0, 0.0, 0.0]},
            {"zoneId": 10, "valueType": 1, "value": [0.0, 0.0, -60.0]}],
        "InitialConditions": [{"Temp": 25.0}],
    }
    # 5. tetrahedra from the same file, velocity zones instead of contact, pressure algorithm 1 (ANP as shipped)
    out["file_tet_zones"] = {
        "Configuration": {"Nproc": 1, "cflFactor": 0.1, "simTime": 2.0e-5, "outTime": 1.0e-3, "pressAlgorithm": 1,
                          "artifViscCoeffs": [0.0, 0.0, 0.0]},
        "Materials": [ALU],
        "DomainBlocks": [{"type": "File", "fileName": "tet_block.k", "zoneId": 0}],
        "BoundaryConditions": [
            {"zoneId": 1, "valueType": 0, "value": [0.0, 0.0, 0.0], "start": [-1.0, -1.0, -1.0e-5], "end": [1.0, 1.0, 1.0e-5]},
            {"zoneId": 2, "valueType": 0, "value": [0.0, 0.0, -50.0], "start": [-1.0, -1.0, 0.00599], "end": [1.0, 1.0, 0.00601]}],
    }
    # 6. axisymmetric quads pressed by a rigid line (Contact_Compression_axisymm_quad.json scaled down)
    out["box_axiquad_contact"] = {
        "Configuration": {"Nproc": 1, "cflFactor": 0.3, "simTime": 5.0e-5, "outTime": 1.0e-3, "domType": "AxiSymm",
                          "artifViscCoeffs": [0.0, 0.0, 0.0]},
        "Stabilization": {"hg_visc": 0.1, "hg_stiff": 0.1},
        "Materials": [ALU],
        "DomainBlocks": [{"type": "Box", "zoneId": 0, "start": [0.0, 0.0, 0.0], "dim": [0.0081, 0.0121, 0.0],
                          "elemLength": 0.001}],
        "Contact": [{"fricCoeffStatic": 0.2, "fricCoeffDynamic": 0.2, "penaltyFactor": 0.5}],
        "RigidBodies": [
            {"type": "Line", "zoneId": 10, "start": [-0.001, 0.012, 0.0], "flipnormals": True, "partSide": 8, "dim": [0.02, 0.0, 0.0]},
            {"type": "Line", "zoneId": 1, "start": [-0.001, 0.0, 0.0], "partSide": 8, "dim": [0.02, 0.0, 0.0]}],
        "BoundaryConditions": [
            {"zoneId": 1, "valueType": 0, "value": [0.0, 0.0, 0.0]},
            {"zoneId": 10, "valueType": 1, "value": [0.0, -30.0, 0.0]}],
    }
    return out


def main():
    x, el = block(4, 4, 6, 0.001, tets=True)
    write_k(os.path.join(HERE, "tet_block.k"), x, el, lambda n: 3 * n + 7)       # sparse, non-contiguous ids
    x, el = block(4, 3, 5, 0.001, tets=False)
    write_k(os.path.join(HERE, "hex_block.k"), x, el, lambda n: n + 1)
    for name, d in decks().items():
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(d, f, indent=1)
            f.write("\n")


if __name__ == "__main__":
    main()
