"""Synthetic workloads of BASELINE.json (`configs`) / SURVEY.md §8(d), as data.

A case is applied to any object that offers the Domain_d-shaped host surface
(``set_domtype, box, set_material, set_stab, set_options, add_bc(s), allocate_bcs, init``):
the engine's :class:`weldformfem_b200.domain.Domain_d` and the two CPU checkers under
``oracle/`` share it, so a parity test drives all of them with the same call.

Geometry / material / BC sets follow the reference's own drivers and decks:
  * C1: src/common/main_1_elem_3d.C:60-164 (0.1 m steel cube, one hexa, top v_z = -1)
  * C2/C3: examples/input/Compression_tetra.json, Compression_hexa_hollomon.json
           (Hollomon aluminium, bottom plane clamped, top plane (0,0,v_z))
  * C4: examples/input/Compression_axisymm_quad.json (axisymmetric quads, top v_y)
Node ids of box meshes follow Domain_d::AddBoxLength (src/common/Domain_d.C:1205-1234):
id = i + (nx+1) * (j + (ny+1) * k).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

BILINEAR, HOLLOMON, JOHNSON_COOK, GMT = 0, 1, 2, 3
PLANE_STRAIN, AXISYMM, DOM_3D = 0, 2, 3


@dataclass
class Case:
    name: str
    dim: int
    n: tuple            # elements per side (nx, ny[, nz])
    h: float            # element edge length
    tritet: bool = False
    domtype: int = DOM_3D
    vol_weight: bool = False
    E: float = 68.9e9
    nu: float = 0.3
    rho0: float = 2700.0
    model: int = HOLLOMON
    sy0: float = 190.4e6
    K: float = 386.796e6
    m: float = 0.154
    # Johnson-Cook (A B n C eps_0 m T_m T_t) / GMT (n1 n2 C1 C2 m1 m2 I1 I2 e_min e_max er_min er_max T_min T_max)
    # constants read by the free functions of include/common/Material.cuh:377-483, and the uniform temperature
    mat_params: tuple = ()
    temp: float = 20.0
    cfl: float = 0.3
    dt: float | None = None
    press: int = 0
    hexa_hg: float = 0.0
    stab: dict = field(default_factory=dict)
    av: tuple = (0.0, 0.0)
    top_vel: float = -10.0
    bc_style: str = "clamp"   # "clamp": bottom all dims 0, top (0,..,top_vel); "c1": main_1_elem_3d.C;
                              # "bottom": bottom clamped only (the top is pushed by a rigid surface)
    # contact with rigid surfaces (examples/input/Contact_Compression_*.json): planes = dicts with the arguments of
    # TriMesh_d::AxisPlaneMesh + "vel"; contact = dict(mu_sta, mu_dyn, penalty_factor, end_steps)
    planes: tuple = ()
    contact: dict | None = None
    # thermal coupling (config "thermal", Thermal.C): dict(k_T, cp_T, exp_T, plheatfrac, T0[, heat_cond, T_die])
    thermal: dict | None = None

    # ---- derived -------------------------------------------------------------------------
    @property
    def nodxelem(self):
        if self.dim == 3:
            return 4 if self.tritet else 8
        return 3 if self.tritet else 4

    @property
    def n_elems(self):
        base = int(np.prod(self.n))
        if self.tritet:
            base *= 6 if self.dim == 3 else 2
        return base

    @property
    def n_nodes(self):
        return int(np.prod([q + 1 for q in self.n]))

    @property
    def bulk(self):
        return self.E / (3.0 * (1.0 - 2.0 * self.nu))

    @property
    def timestep(self):
        if self.dt is not None:
            return self.dt
        return self.cfl * self.h / math.sqrt(self.bulk / self.rho0)   # src/explicit/main.C:862-879

    def bc_nodes(self):
        """(node, dim, value) triplets in the order AddBCVelZone would append them
        (ascending node id, dims x,y[,z] per node; src/common/Domain_d.C:432-452)."""
        d = self.dim
        if self.bc_style == "c1":
            out = [(0, 0, 0.0), (0, 1, 0.0), (0, 2, 0.0), (1, 1, 0.0), (1, 2, 0.0), (2, 0, 0.0),
                   (2, 2, 0.0), (3, 2, 0.0)] + [(i + 4, 2, -1.0) for i in range(4)]
            return out
        nplane = (self.n[0] + 1) * ((self.n[1] + 1) if d == 3 else 1)
        nlayers = self.n[d - 1]
        bottom = np.arange(nplane)
        top = np.arange(nplane) + nplane * nlayers
        out = []
        for nd in bottom:
            out += [(int(nd), dd, 0.0) for dd in range(d)]
        if self.bc_style == "bottom":
            return out
        for nd in top:
            out += [(int(nd), dd, (self.top_vel if dd == d - 1 else 0.0)) for dd in range(d)]
        return out

    def bc_arrays(self):
        t = self.bc_nodes()
        nodes = np.array([q[0] for q in t], dtype=np.int32)
        dims = np.array([q[1] for q in t], dtype=np.int32)
        vals = np.array([q[2] for q in t], dtype=np.float64)
        return nodes, dims, vals

    # ---- drive a Domain_d-shaped object ---------------------------------------------------
    def apply(self, dom, init: bool = True):
        if self.dim == 2:
            dom.set_domtype(self.domtype, self.vol_weight)
        pad = 1.0 + 1.0e-6
        L = [self.n[0] * self.h * pad, self.n[1] * self.h * pad,
             (self.n[2] * self.h * pad) if self.dim == 3 else 0.0]
        dom.box((0.0, 0.0, 0.0), L, 0.5 * self.h, self.tritet)
        if self.model in (JOHNSON_COOK, GMT):
            dom.set_material_ext(self.E, self.nu, self.rho0, self.model, self.sy0, self.mat_params, self.temp)
        else:
            dom.set_material(self.E, self.nu, self.rho0, self.model, self.sy0, self.K, self.m)
        if self.thermal is not None:   # main.C:436-441, 567-570
            t = self.thermal
            dom.thermal_on(t["k_T"], t["cp_T"], t.get("exp_T", 0.0), t.get("plheatfrac", 0.9), t.get("T0", 20.0))
        dom.set_stab(**self.stab)
        dom.set_options(self.press, self.av[0], self.av[1], self.hexa_hg)
        if hasattr(dom, "add_bcs"):
            dom.add_bcs(*self.bc_arrays())
        else:
            for nd, dd, val in self.bc_nodes():
                dom.add_bc(nd, dd, val)
        dom.allocate_bcs()
        if self.contact is not None:   # order of src/explicit/main.C:650-862
            dom.call("SearchExtNodes")
            for pl in self.planes:
                dom.add_plane(self.dim, pl["id"], pl["axis"], pl["positaxisorent"], pl["p1"], pl["p2"], pl["dens"], pl["vel"])
            c = self.contact
            dom.contact_on(c["mu_sta"], c["mu_dyn"], c["penalty_factor"], c["end_steps"] * self.timestep)
            if self.thermal is not None and "heat_cond" in self.thermal:   # heatCondCoeff / dieTemp, main.C:718-719
                dom.set_contact_heat(self.thermal["heat_cond"], self.thermal["T_die"])
            dom.call("calcMinEdgeLength")
        if init:
            dom.init(self.timestep)
        return dom


def c1_one_hex(hexa_hg: float = 0.06) -> Case:
    """configs[0]: 1-element reduced-integration hexa compression with hourglass."""
    return Case("c1_1hex", 3, (1, 1, 1), 0.1, E=206e9, nu=0.3, rho0=7850.0, model=BILINEAR,
                sy0=1.0e10, K=0.0, m=1.0, dt=0.8e-5, hexa_hg=hexa_hg, bc_style="c1")


def c2_tets(n: int = 26, press: int = 0) -> Case:
    """configs[1]: constant-stress tets with nodal-averaged pressure, J2 Hollomon; n=26 -> 105 456 tets."""
    return Case(f"c2_tet_n{n}", 3, (n, n, n), 1.0e-3, tritet=True, cfl=0.1, press=press)


def c3_hexes(n: int = 215, hexa_hg: float = 0.06) -> Case:
    """configs[2]: structured hexa cube, reduced integration + viscous hourglass; n=215 -> 9 938 375."""
    return Case(f"c3_hex_n{n}", 3, (n, n, n), 1.0e-3, cfl=0.3, hexa_hg=hexa_hg)


def c4_axisymm_quads(n: int = 1000) -> Case:
    """configs[3]: 2D axisymmetric quads with hourglass; n=1000 -> 1e6 quads."""
    return Case(f"c4_axiquad_n{n}", 2, (n, n), 0.5e-3, domtype=AXISYMM, cfl=0.3,
                stab=dict(hg_visc=0.1, hg_stiff=0.1), top_vel=-1.0)


def c5_block(n: int = 431, tritet: bool = False) -> Case:
    """configs[4]: 80 M-element hexa (n=431) / tet (n=237) block for 2/4/8 GPUs."""
    if tritet:
        return Case(f"c5_tet_n{n}", 3, (n, n, n), 1.0e-3, tritet=True, cfl=0.1)
    return Case(f"c5_hex_n{n}", 3, (n, n, n), 1.0e-3, cfl=0.3, hexa_hg=0.06)


def contact_tets(n: int = 6, tool_vel: float = -200.0, mu=(0.3, 0.2), two_planes: bool = True, stab: dict | None = None) -> Case:
    """Tet block upset between rigid planes (examples/input/Contact_Compression_tetra.json scaled down): bottom clamped,
    a rigid plane with normal -z comes down on the top face (plus, optionally, a resting plane under the bottom)."""
    h, L = 1.0e-3, n * 1.0e-3
    planes = [dict(id=0, axis=2, positaxisorent=False, p1=(-0.5 * L, -0.5 * L, L + 0.01 * h), p2=(1.5 * L, 1.5 * L, L + 0.01 * h),
                   dens=4, vel=(0.5, 0.0, tool_vel))]
    if two_planes:
        planes.append(dict(id=1, axis=2, positaxisorent=True, p1=(-0.5 * L, -0.5 * L, -1.0e-4), p2=(1.5 * L, 1.5 * L, -1.0e-4),
                           dens=2, vel=(0.0, 0.0, 0.0)))
    return Case(f"contact_tet_n{n}", 3, (n, n, n), h, tritet=True, cfl=0.1, top_vel=0.0, bc_style="bottom",
                stab=stab or {}, planes=tuple(planes),
                contact=dict(mu_sta=mu[0], mu_dyn=mu[1], penalty_factor=0.6, end_steps=100))


def contact_quads(n: int = 12, tool_vel: float = -100.0, mu=(0.3, 0.2), domtype: int = PLANE_STRAIN) -> Case:
    """2D counterpart (examples/input/Contact_Compression_axisymm_quad.json): rigid line with normal -y above the top edge."""
    h, L = 0.5e-3, n * 0.5e-3
    planes = (dict(id=0, axis=1, positaxisorent=False, p1=(-0.5 * L, L + 0.01 * h, 0.0), p2=(1.5 * L, L + 0.01 * h, 0.0),
                   dens=4, vel=(0.5, tool_vel, 0.0)),)
    return Case(f"contact_quad_n{n}", 2, (n, n), h, domtype=domtype, cfl=0.3, stab=dict(hg_visc=0.1, hg_stiff=0.1),
                top_vel=0.0, bc_style="bottom", planes=planes,
                contact=dict(mu_sta=mu[0], mu_dyn=mu[1], penalty_factor=0.6, end_steps=100))


JC_AL6061 = (324.0e6, 114.0e6, 0.42, 0.002, 1.0, 1.34, 925.0, 294.0)     # A B n C eps_0 m T_m T_t (Johnson & Cook 1983, Al 6061-T6)
GMT_DEMO = (0.0, 0.15, 400.0e6, -0.002, 0.0, 0.02, 0.0, -0.002, 0.01, 3.0, 1.0e-3, 1.0e5, 20.0, 500.0)


def with_johnson_cook(case: Case, temp: float = 400.0) -> Case:
    """Same workload with the Johnson-Cook flow stress (Material.cuh:377-412) at a uniform temperature."""
    import dataclasses
    return dataclasses.replace(case, name=case.name + "_jc", model=JOHNSON_COOK, sy0=JC_AL6061[0], mat_params=JC_AL6061, temp=temp)


# aluminium of examples/input/Contact_Compression_tetra.json (thermalCond 190, thermalHeatCap 87.5 as shipped), plus a
# thermal expansion coefficient and die heat exchange so every coupling term is exercised
THERMAL_AL = dict(k_T=190.0, cp_T=87.5, exp_T=2.3e-5, plheatfrac=0.9, T0=20.0)


def with_thermal(case: Case, **over) -> Case:
    import dataclasses
    return dataclasses.replace(case, name=case.name + "_th", thermal=dict(THERMAL_AL, **over))


def with_gmt(case: Case, temp: float = 100.0) -> Case:
    """Same workload with the GMT flow stress (Material.cuh:418-483)."""
    import dataclasses
    return dataclasses.replace(case, name=case.name + "_gmt", model=GMT, sy0=100.0e6, mat_params=GMT_DEMO, temp=temp)


def plane_strain_quads(n: int = 16) -> Case:
    return Case(f"ps_quad_n{n}", 2, (n, n), 0.5e-3, domtype=PLANE_STRAIN, cfl=0.3,
                stab=dict(hg_visc=0.1, hg_stiff=0.1), top_vel=-1.0)


def plane_strain_tris(n: int = 16) -> Case:
    return Case(f"ps_tri_n{n}", 2, (n, n), 0.5e-3, tritet=True, domtype=PLANE_STRAIN, cfl=0.1,
                top_vel=-1.0)
