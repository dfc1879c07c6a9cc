"""Shared helpers of the parity tests: drive the engine and a CPU checker with the same case and
compare reference-layout arrays."""
import dataclasses

import numpy as np

STATE = "x v a u prev_a m_fi m_mdiag m_voln vol p pl_strain sigma_y m_sigma m_tau".split()


def relerr(got, want):
    """max |got - want| / max |want|  (array-wise relative error, the metric of BASELINE.json)."""
    want = np.asarray(want, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    if want.size == 0:
        return 0.0
    if not np.all(np.isfinite(got)):
        return np.inf
    scale = np.abs(want).max()
    diff = np.abs(got - want).max()
    if scale == 0.0:
        return 0.0 if diff == 0.0 else np.inf
    return diff / scale


def run_pair(case, checker_cls, nsteps, strict, tracking=None, device=0):
    from weldformfem_b200.domain import Domain_d
    ref = checker_cls()
    case.apply(ref)
    eng = Domain_d(device=device, strict=strict)
    case.apply(eng, init=False)
    if tracking:
        eng.set_tracking(**tracking)
    eng.init(case.timestep)
    if nsteps:
        ref.step(nsteps)
        eng.step(nsteps)
    return eng, ref


def compare(eng, ref, names, tol, label=""):
    worst = {}
    for nm in names:
        worst[nm] = relerr(eng.get(nm), ref.get(nm))
    bad = {k: v for k, v in worst.items() if not (v <= tol)}
    assert not bad, f"{label}: relative error above {tol:g}: {bad} (all: {worst})"
    return worst


def fast(case, **kw):
    return dataclasses.replace(case, **kw)
