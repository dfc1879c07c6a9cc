"""WeldFormFEM input decks (examples/input/*.json + LS-Dyna `.k` meshes) for the Python host side: the counterpart of
host/wf_deck.hpp, i.e. everything src/explicit/main.C:191-975 does between reading the deck and calling
SolveChungHulbert, expressed through the Domain_d-shaped interface shared by `weldformfem_b200.domain.Domain_d` (the
engine) and the oracle drivers (tests).

    from weldformfem_b200 import deck
    setup = deck.load("Contact_Compression_tetra.json")
    dom = setup.apply(Domain_d())          # mesh, material, BCs, contact, thermal, time step, init
    dom.step(setup.n_steps())

Deliberate deviations from main.C are the ones listed at the top of host/wf_deck.hpp (Johnson-Cook / GMT constants go
to the fields the step reads; no "File" rigid body, no remeshing; `fileName` is resolved relative to the deck)."""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field

import numpy as np

BILINEAR, HOLLOMON, JOHNSON_COOK, GMT = 0, 1, 2, 3
STAB_KEYS = ("alpha_free", "alpha_contact", "hg_coeff_free", "hg_coeff_contact", "av_coeff_div", "av_coeff_bulk",
             "log_factor", "p_pspg_bulkfac", "J_min", "hg_visc", "hg_stiff", "pspg_scale")
# "pspg_scale": main.C's loadStabilizationParams never assigns it, so the reference hands an INDETERMINATE value to
# calcElemPressure (Mechanical.C:796); deliberate deviation: 0 unless the deck carries the key


def read_k(path, scale=1.0):
    """*NODE / *ELEMENT_SOLID cards (fixed 8/16-column or comma-separated; eid/pid and nodes on one line or two).  Node
    ids become 0-based indices in order of appearance; a solid is cut at its first repeated node (tetrahedra are
    written as degenerate 8-node solids).  Returns (x[n,3], elnod[e,k])."""
    xs, index_of, elems, sect, pending = [], {}, [], None, None

    def fields(line, widths):
        if "," in line:
            return line.split(",")
        out, pos = [], 0
        for w in widths:
            if pos >= len(line):
                break
            out.append(line[pos:pos + w])
            pos += w
        return out

    with open(path) as f:
        for raw in f:
            line = raw.rstrip("\r\n")
            if not line or line[0] == "$":
                continue
            if line[0] == "*":
                if line.startswith("*NODE") and not line.startswith("*NODE_"):
                    sect = "node"
                elif line.startswith("*ELEMENT_SOLID"):
                    sect = "solid"
                else:
                    sect = None
                continue
            if sect == "node":
                t = fields(line, (8, 16, 16, 16, 8, 8))
                if len(t) < 4:
                    continue
                index_of[int(t[0])] = len(xs)
                xs.append([float(t[1]) * scale, float(t[2]) * scale, float(t[3]) * scale])
            elif sect == "solid":
                v = [int(q) for q in fields(line, (8,) * 10) if q.strip()]
                if not v:
                    continue
                if pending is None and len(v) == 2:
                    pending = v
                    continue
                nodes = v if pending is not None else v[2:]
                pending = None
                uniq = []
                for n in nodes:
                    if n in uniq:
                        break
                    uniq.append(n)
                elems.append(uniq)
    if not elems:
        raise ValueError(f"{path}: no *ELEMENT_SOLID records")
    k = len(elems[0])
    if k not in (4, 8) or any(len(e) != k for e in elems):
        raise ValueError(f"{path}: solids must be all tetrahedra or all hexahedra")
    try:
        el = np.array([[index_of[n] for n in e] for e in elems], dtype=np.int32)
    except KeyError:
        raise ValueError(f"{path}: element references an unknown node id") from None
    return np.array(xs, dtype=np.float64), el


@dataclass
class DeckSetup:
    path: str
    cfg: dict
    dim: int = 3
    domtype: str = "3D"
    vol_weight: bool = False
    box: dict | None = None            # start, L, r, tritet
    mesh: tuple | None = None          # (x, elnod) of a File domain
    material: dict = field(default_factory=dict)
    stab: dict = field(default_factory=dict)
    press: int = 0
    av: tuple = (0.0, 0.0)
    thermal: dict | None = None
    bconds: list = field(default_factory=list)
    bodies: list = field(default_factory=list)
    contact: dict | None = None
    sym: tuple = (False, False, False)
    symtol: float = 1.0e-4
    cfl: float = 0.3
    sim_time: float = 0.0
    dt: float = 0.0                    # filled by apply()

    def n_steps(self):
        """`while (Time < end_t)` of SolveChungHulbert with the fixed step."""
        n, t = 0, 0.0
        while t < self.sim_time:
            t += self.dt
            n += 1
        return n

    # ---- main.C:352-975 through a Domain_d-shaped object ------------------------------------------------------------
    def apply(self, dom, init=True, hexa_hg=0.0):
        m = self.material
        if self.dim == 2 or self.box is not None:
            dom.set_domtype({"AxiSymm": 2, "AxiSym": 2}.get(self.domtype, 0 if self.domtype == "plStrain" else 3)
                            if self.box is not None else 3, self.vol_weight)
        if self.box is not None:
            b = self.box
            dom.box(b["start"], b["L"], b["r"], b["tritet"])
        else:
            x, el = self.mesh
            dom.set_mesh(3, el.shape[1], x.ravel(), el.ravel())
        if m["model"] in (JOHNSON_COOK, GMT):
            dom.set_material_ext(m["E"], m["nu"], m["rho"], m["model"], m["sy0"], m["params"], self.cfg.get("T0", 20.0))
        else:
            dom.set_material(m["E"], m["nu"], m["rho"], m["model"], m["sy0"], m["K"], m["m"])
        if self.thermal is not None:
            t = self.thermal
            dom.thermal_on(t["k_T"], t["cp_T"], t["exp_T"], t["plheatfrac"], t["T0"])
        dom.set_stab(**self.stab)
        dom.set_options(self.press, self.av[0], self.av[1], hexa_hg)
        x0 = np.asarray(dom.get("x")).reshape(-1, self.dim)
        nn = len(x0)
        triplets = []
        if self.contact is None:        # Domain_d::AddBCVelZone for every BC block (main.C:737-749)
            for b in self.bconds:
                lo, hi = np.array(b["start"][:self.dim]), np.array(b["end"][:self.dim])
                inside = np.all((x0 >= lo) & (x0 <= hi), axis=1)
                for n in np.nonzero(inside)[0]:
                    triplets += [(int(n), d, float(b["value"][d])) for d in range(self.dim)]
        for n in range(nn):             # symmetry planes (main.C:947-963)
            for d in range(3):
                if self.sym[d] and d < self.dim and x0[n, d] < self.symtol:
                    triplets.append((n, d, 0.0))
        if hasattr(dom, "add_bcs") and triplets:
            dom.add_bcs(np.array([t[0] for t in triplets], dtype=np.int32), np.array([t[1] for t in triplets], dtype=np.int32),
                        np.array([t[2] for t in triplets]))
        else:
            for n, d, v in triplets:
                dom.add_bc(n, d, v)
        dom.allocate_bcs()
        if self.contact is not None:    # main.C:650-848
            dom.call("SearchExtNodes")
            for i, body in enumerate(self.bodies):
                dom.add_plane(body["dimension"], i, body["axis"], not body["flipnormals"], body["start"], body["dim"],
                              body["partSide"], body["vel"])
            c = self.contact
            dom.contact_on(c["mu_sta"], c["mu_dyn"], c["penalty_factor"], self.sim_time)
            if self.thermal is not None:
                dom.set_contact_heat(c["heat_cond"], c["T_die"])
        # time step (main.C:862-883): cflFactor * min edge length / sqrt(K / rho)
        if hasattr(dom, "consts"):
            dom.call("calcMinEdgeLength")
            min_len = dom.consts()["min_length"]
        else:
            min_len = dom.calcMinEdgeLength()[0]
        bulk = m["E"] / (3.0 * (1.0 - 2.0 * m["nu"]))
        self.dt = self.cfl * min_len / math.sqrt(bulk / m["rho"])
        if init:
            dom.init(self.dt)
        return dom


def _vec(j, default=(0.0, 0.0, 0.0)):
    return tuple(float(q) for q in j[:3]) if j is not None else tuple(default)


def load(path) -> DeckSetup:
    with open(path) as f:
        j = json.load(f)
    cfg = j.get("Configuration") or {}
    mats = j.get("Materials") or [{}]
    blocks = j.get("DomainBlocks") or [{}]
    S = DeckSetup(path=path, cfg=cfg)
    st = j.get("Stabilization")
    S.stab = {k: float(st.get(k, 0.0)) for k in STAB_KEYS} if st else {"hg_stiff": 0.1}   # main.C:84-120, Domain_d.h:283-296
    S.sim_time = float(cfg.get("simTime", 0.0))
    S.cfl = float(cfg.get("cflFactor", 0.3))
    if cfg.get("plasticType", "Hardening") != "Hardening":
        raise ValueError("plasticType other than Hardening is not supported")
    av = cfg.get("artifViscCoeffs")
    if av and len(av) >= 2:
        S.av = (float(av[0]), float(av[1]))
    S.domtype = cfg.get("domType", "3D")
    S.vol_weight = bool(cfg.get("AxiSymmVol", False))
    S.sym = (bool(cfg.get("xSymm", False)), bool(cfg.get("ySymm", False)), bool(cfg.get("zSymm", False)))
    S.symtol = float(cfg.get("symtol", 1.0e-4))
    # Solver_explicit.C:733-743 dispatches on 0 and 1 only (any other value would skip the pressure update altogether)
    if int(cfg.get("pressAlgorithm", 0)) not in (0, 1):
        raise ValueError("pressAlgorithm must be 0 or 1")
    if cfg.get("devElastic", True) is False:
        raise ValueError("devElastic = false (calcElemPressureRigid) is not supported")
    S.press = int(cfg.get("pressAlgorithm", 0))
    blk = blocks[0]
    kind = blk.get("type", "Box")
    if kind == "File":
        fn = blk.get("fileName", "")
        if not fn.endswith(".k"):
            raise ValueError("DomainBlocks[0].fileName must be an LS-Dyna .k file")
        S.mesh = read_k(fn if os.path.isabs(fn) else os.path.join(os.path.dirname(os.path.abspath(path)), fn))
        S.dim = 3
    elif kind == "Box":
        L = _vec(blk.get("dim"))
        S.box = dict(start=_vec(blk.get("start")), L=(L[0], L[1], 0.0), r=float(blk.get("elemLength", 0.06)) / 2.0,
                     tritet=blk.get("elemType", "") == "TriTet")      # main.C:418: the z extent is dropped
        S.dim = 2
    else:
        raise ValueError("DomainBlocks[0].type must be File or Box")
    mt = mats[0]
    c = [float(q) for q in mt.get("const", [])] + [0.0] * 10
    E, nu, rho, Fy = float(mt.get("youngsModulus", 0)), float(mt.get("poissonsRatio", 0)), float(mt.get("density0", 0)), float(mt.get("yieldStress0", 0))
    thermal = bool(cfg.get("thermal", False))
    T0 = 20.0
    for ic in j.get("InitialConditions") or []:
        T0 = float(ic.get("Temp", T0))
    cfg["T0"] = T0
    typ = mt.get("type", "Bilinear")
    if typ == "Bilinear":
        S.material = dict(model=BILINEAR, E=E, nu=nu, rho=rho, sy0=Fy, K=0.0, m=1.0)
    elif typ == "Hollomon":
        S.material = dict(model=HOLLOMON, E=E, nu=nu, rho=rho, sy0=Fy, K=c[0], m=c[1])
    elif typ == "JohnsonCook":      # (A = Fy, B, n, C, eps_0, m, T_m, T_t), main.C:539-541
        p = (Fy, c[0], c[1], c[2], c[3], c[4], c[5], c[6]) if thermal else (Fy, c[0], c[1], c[2], c[3], 1.0, 1.0e10, 0.0)
        S.material = dict(model=JOHNSON_COOK, E=E, nu=nu, rho=rho, sy0=Fy, params=p)
    elif typ == "GMT":
        er, sr, tr = mt.get("strRange", [0.0, 1e10]), mt.get("strdotRange", [0.0, 1e10]), mt.get("tempRange", [0.0, 1e10])
        S.material = dict(model=GMT, E=E, nu=nu, rho=rho, sy0=Fy, params=tuple(c[:8]) + (er[0], er[1], sr[0], sr[1], tr[0], tr[1]))
    else:
        raise ValueError(f"material type '{typ}' is not supported")
    if thermal:
        S.thermal = dict(k_T=float(mt.get("thermalCond", 0.0)), cp_T=float(mt.get("thermalHeatCap", 0.0)),
                         exp_T=float(mt.get("thermalExp", 0.0)), plheatfrac=float(cfg.get("plHeatFrac", 0.9)), T0=T0)
    for b in j.get("BoundaryConditions") or []:
        S.bconds.append(dict(zoneId=int(b.get("zoneId", 0)), value=_vec(b.get("value")), start=_vec(b.get("start")), end=_vec(b.get("end"))))
    rbs = j.get("RigidBodies") or []
    if rbs and "type" in rbs[0]:
        if len(rbs) > 2:
            raise ValueError("at most two rigid bodies, like main.C")
        for rb in rbs:
            start, dim_ = _vec(rb.get("start")), _vec(rb.get("dim"))
            typ = rb.get("type")
            if typ == "Plane":
                dimension, axis = 3, 2
            elif typ == "Line":
                dimension = 2
                if dim_[0] > 0.0:
                    axis = 1
                elif dim_[1] > 0.0:
                    axis = 0
                else:
                    raise ValueError("rigid Line has null dimension")
            else:
                raise ValueError(f"rigid body type '{typ}' is not supported (Plane, Line)")
            vel = (0.0, 0.0, 0.0)
            for b in S.bconds:
                if b["zoneId"] == int(rb.get("zoneId", 0)):
                    vel = b["value"]
            S.bodies.append(dict(dimension=dimension, axis=axis, flipnormals=bool(rb.get("flipnormals", False)), start=start,
                                 dim=dim_, partSide=int(rb.get("partSide", 1)), vel=vel))
        ct = (j.get("Contact") or [{}])[0]
        S.contact = dict(mu_sta=float(ct.get("fricCoeffStatic", 0.0)), mu_dyn=float(ct.get("fricCoeffDynamic", 0.0)),
                         penalty_factor=float(ct.get("penaltyFactor", -1.0)), heat_cond=float(ct.get("heatCondCoeff", 0.0)),
                         T_die=float(ct.get("dieTemp", 20.0)))
    return S
