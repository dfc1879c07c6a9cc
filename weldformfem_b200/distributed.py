"""Multi-GPU host side: element-block partition + halo exchange of shared-node partial sums.

The reference has no domain decomposition; the canonical partition is defined in SURVEY.md §8e and built
by ``wf_partition_build*`` (include/wf_engine.h).  Two ways to run it:

* :class:`RankDomain` — ONE rank of a ``torchrun`` job (one process per GPU).  ``torch.distributed`` is
  plumbing only: it carries the 64-byte CUDA IPC handles and slot offsets at connect time (and, for the
  optional ``halo="nccl"`` transport, the grouped send/recv between phases).  With the default
  ``halo="peer"`` transport the neighbours' receive regions are mapped into this process and the step
  kernels store their partial sums straight into them over NVLink — ``step(n)`` then enqueues n complete
  distributed steps without any host synchronisation.
* :class:`LocalCluster` — all ranks inside one process, driven by one host thread (``wf_step_all``);
  several ranks may share one GPU, which is how the distributed path is parity-tested on a 1-GPU box.

Both offer the Domain_d-shaped surface that :meth:`weldformfem_b200.cases.Case.apply` drives.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .domain import AXISYMM, Domain_d, WfError

_ELEM_ARRAYS = {"vol", "vol_0", "rho", "rho_0", "p", "pl_strain", "sigma_y", "m_detJ", "m_radius", "m_tau", "m_eps",
                "m_str_rate", "m_rot_rate", "m_sigma", "m_f_elem", "m_f_elem_hg", "m_hg_q", "m_dH_detJ_dx",
                "m_dH_detJ_dy", "m_dH_detJ_dz"}
_NODE_SCALARS = {"m_mdiag", "m_voln", "p_node"}


class Partition:
    """Handle of a ``wf_partition`` (host-side, no GPU needed)."""

    def __init__(self, nranks, rank, *, box=None, mesh=None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.nranks, self.rank = int(nranks), int(rank)
        if box is not None:
            V, L, r, tritet = box
            rc = self._lib.wf_partition_build_box(C.byref(self._h), self.nranks, self.rank, (C.c_double * 3)(*V),
                                                  (C.c_double * 3)(*L), float(r), int(tritet))
        else:
            k, n_nodes, elnod = mesh
            el = np.ascontiguousarray(elnod, dtype=np.uint32)
            rc = self._lib.wf_partition_build(C.byref(self._h), self.nranks, self.rank, int(k), int(n_nodes),
                                              el.size // int(k), el.ctypes.data_as(C.POINTER(C.c_uint)))
        if rc != 0:
            raise WfError("partition build failed")
        eb, ee, nl, nng = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._lib.wf_partition_info(self._h, C.byref(eb), C.byref(ee), C.byref(nl), C.byref(nng))
        self.elem_begin, self.elem_end, self.n_local_nodes, self.n_neigh = eb.value, ee.value, nl.value, nng.value

    def _arr(self, fn, n, dtype=np.int32):
        if n == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(fn(self._h), (n,)).astype(dtype, copy=True)

    @property
    def node_l2g(self):
        return self._arr(self._lib.wf_partition_node_l2g, self.n_local_nodes)

    @property
    def neigh_ranks(self):
        return self._arr(self._lib.wf_partition_neigh_ranks, self.n_neigh)

    @property
    def halo_offset(self):
        return self._arr(self._lib.wf_partition_halo_offset, self.n_neigh + 1)

    @property
    def halo_nodes(self):
        return self._arr(self._lib.wf_partition_halo_nodes, int(self.halo_offset[-1]))

    def free(self):
        if self._h:
            self._lib.wf_partition_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------
# connection planning (pure host logic; unit-tested with gloo, world_size 2, without a GPU)
# ---------------------------------------------------------------------------------------------------
def slot_table(neigh_ranks, halo_offset, flag_bytes=None):
    """Byte offsets inside a rank's comm block of the flag slot and receive region of each neighbour;
    mirrors ``wf_halo_slot_offsets`` (csrc/wf_engine.cu): block = [flags, padded to 256 B | regions],
    region of neighbour i = 2 parities x 3 doubles x count_i, in neighbour order."""
    nng = len(neigh_ranks)
    if flag_bytes is None:
        flag_bytes = ((8 * max(nng, 1) + 255) // 256) * 256
    return {int(q): (8 * i, flag_bytes + 8 * 2 * 3 * int(halo_offset[i])) for i, q in enumerate(neigh_ranks)}


def plan_connections(rank, my_neigh, published):
    """``published[q]`` is what rank q announced: ``{"slots": {peer_rank: (flag_off, region_off)}, ...}``.
    Returns, per neighbour index of this rank, ``(peer_rank, flag_off, region_off)`` = where THIS rank must
    write inside the peer's block.  Raises if the halo lists are not symmetric."""
    plan = []
    for i, q in enumerate(my_neigh):
        q = int(q)
        slots = published[q]["slots"]
        if rank not in slots:
            raise WfError(f"rank {q} does not list rank {rank} as a neighbour: halo lists are not symmetric")
        fo, ro = slots[rank]
        plan.append((q, int(fo), int(ro)))
    return plan


def exchange_descriptors(desc, group=None):
    """all-gather one small python object per rank (torch.distributed, any backend)."""
    import torch.distributed as dist
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, desc, group=group)
    return {int(d["rank"]): d for d in out}


class _CudaBuffer:
    """Expose a raw device pointer through __cuda_array_interface__ so torch can wrap it (fp64)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


# ---------------------------------------------------------------------------------------------------
# one rank
# ---------------------------------------------------------------------------------------------------
class RankDomain(Domain_d):
    """One rank's part of a partitioned domain (one GPU)."""

    def __init__(self, rank, nranks, device=0, strict=False, halo="peer"):
        super().__init__(device=device, strict=strict)
        self.rank, self.nranks = int(rank), int(nranks)
        self.halo = halo
        self.partition = None
        self._peers = {}

    # ---- mesh ---------------------------------------------------------------------------------
    def AddBoxLength(self, V, L, r, red_int=True, tritetra=False):
        if not red_int:
            raise WfError("full integration is not implemented by the reference step")
        dim = 3 if L[2] > 0.0 else 2
        k = (4 if tritetra else 8) if dim == 3 else (3 if tritetra else 4)
        self._create(dim, k)
        self.partition = Partition(self.nranks, self.rank, box=(V, L, r, tritetra))
        if self._domtype == AXISYMM:     # every rank of a box cut into element blocks touches the axis x_r = V[0]
            self._ck(self._lib.wf_set_axis_xmin(self._h, float(V[0])))
        self._ck(self._lib.wf_set_mesh_partition(self._h, self.partition._h, None))
        self._after_mesh()

    def box(self, V, L, r, tritet=False):
        self.AddBoxLength(V, L, r, True, bool(tritet))

    def set_mesh(self, dim, k, x, elnod):
        """GLOBAL mesh in, local part kept (every rank passes the same arrays)."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, dim)
        self._create(dim, k)
        self.partition = Partition(self.nranks, self.rank, mesh=(k, x.shape[0], elnod))
        xl = np.ascontiguousarray(x[self.partition.node_l2g])
        if self._domtype == AXISYMM:
            self._ck(self._lib.wf_set_axis_xmin(self._h, float(x[:, 0].min())))
        self._ck(self._lib.wf_set_mesh_partition(self._h, self.partition._h, xl.ctypes.data_as(C.POINTER(C.c_double))))
        self._after_mesh()

    def _after_mesh(self):
        self.node_l2g = self.partition.node_l2g
        self.neigh = [int(q) for q in self.partition.neigh_ranks]
        if self.halo == "nccl":
            self._ck(self._lib.wf_halo_set_transport(self._h, 1))

    # ---- halo plumbing --------------------------------------------------------------------------
    def slot_offsets(self, i):
        fo, ro, rb = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._ck(self._lib.wf_halo_slot_offsets(self._h, int(i), C.byref(fo), C.byref(ro), C.byref(rb)))
        return fo.value, ro.value, rb.value

    def descriptor(self):
        """What this rank publishes at connect time."""
        d = {"rank": self.rank, "neigh": list(self.neigh),
             "slots": {q: self.slot_offsets(i)[:2] for i, q in enumerate(self.neigh)}}
        if self.halo == "peer":
            h = (C.c_ubyte * 64)()
            self._ck(self._lib.wf_halo_ipc_export(self._h, h))
            d["ipc"] = bytes(h)
        return d

    def connect(self, group=None):
        """Exchange descriptors with every rank and map the neighbours' comm blocks (peer transport)."""
        published = exchange_descriptors(self.descriptor(), group)
        plan = plan_connections(self.rank, self.neigh, published)
        if self.halo == "peer":
            for i, (q, fo, ro) in enumerate(plan):
                hb = (C.c_ubyte * 64).from_buffer_copy(published[q]["ipc"])
                base = C.c_void_p()
                self._ck(self._lib.wf_halo_ipc_open(self._h, hb, C.byref(base)))
                self._peers[q] = base.value
                self._ck(self._lib.wf_halo_connect(self._h, i, base, fo, ro))
        return plan

    def _nccl_exchange(self):
        """Grouped ncclSend/ncclRecv of the exchange just packed, ordered on the ENGINE's stream: NCCL starts
        after the pack kernel and the consumer kernels wait for it — no host synchronisation."""
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", self._device)
        ops, keep = [], []
        with torch.cuda.stream(torch.cuda.ExternalStream(self.get_stream(), device=dev)):
            for i, q in enumerate(self.neigh):
                sp, rp, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
                self._ck(self._lib.wf_halo_exchange_ptrs(self._h, i, C.byref(sp), C.byref(rp), C.byref(n)))
                st = torch.as_tensor(_CudaBuffer(sp.value, n.value), device=dev)
                rt = torch.as_tensor(_CudaBuffer(rp.value, n.value), device=dev)
                keep += [st, rt]
                ops.append(dist.P2POp(dist.isend, st, q))
                ops.append(dist.P2POp(dist.irecv, rt, q))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()

    # ---- solve ----------------------------------------------------------------------------------
    def init(self, dt=None):
        if dt is not None:
            self._dt = float(dt)
        if self._dt is None:
            raise WfError("SetDT first")
        if self.halo == "peer":
            self._ck(self._lib.wf_init(self._h, self._dt))
            return
        for ph in range(3):
            self._ck(self._lib.wf_init_phase(self._h, ph, self._dt))
            if ph < 2:
                self._nccl_exchange()

    def step(self, n=1):
        if self.halo == "peer":
            self._ck(self._lib.wf_step(self._h, int(n)))
            return
        for s in range(int(n)):
            last = 1 if s == n - 1 else 0
            for ph in range(3):
                self._ck(self._lib.wf_step_phase(self._h, ph, last))
                if ph < 2:
                    self._nccl_exchange()

    def halo_status(self):
        e = C.c_int()
        self._ck(self._lib.wf_halo_status(self._h, C.byref(e)))
        return e.value

    def elem_range(self):
        return self.partition.elem_begin, self.partition.elem_end


def assemble_global(name, parts, dim, k, n_nodes, n_elems):
    """Merge per-rank LOCAL arrays ``parts = [(l2g, (elem_begin, elem_end), array), ...]`` (rank order) into
    the reference-layout GLOBAL array.  Shared nodes must agree bit for bit between their sharers."""
    if name in _ELEM_ARRAYS:
        per = parts[0][2].size // max(parts[0][1][1] - parts[0][1][0], 1)
        out = np.empty(n_elems * per)
        for _, (eb, ee), a in parts:
            out[eb * per:ee * per] = a
        return out
    per = 1 if name in _NODE_SCALARS else dim
    out = np.full((n_nodes, per), np.nan)
    seen = np.zeros(n_nodes, dtype=bool)
    for l2g, _, a in parts:
        a = a.reshape(-1, per)
        both = seen[l2g]
        if both.any() and not np.array_equal(out[l2g[both]], a[both]):
            raise WfError(f"copies of shared nodes differ between ranks for '{name}'")
        out[l2g] = a
        seen[l2g] = True
    return out.reshape(-1)


# ---------------------------------------------------------------------------------------------------
# all ranks in one process
# ---------------------------------------------------------------------------------------------------
class LocalCluster:
    """nranks engines in this process (``devices[p]`` = CUDA ordinal of rank p; ordinals may repeat)."""

    def __init__(self, nranks, devices=None, strict=False):
        self.nranks = int(nranks)
        devices = list(devices) if devices is not None else [0] * self.nranks
        self.ranks = [RankDomain(p, self.nranks, device=devices[p], strict=strict, halo="peer") for p in range(self.nranks)]
        self._lib = _lib.load()
        self._connected = False
        self._dt = None

    def _each(self, fn, *a, **kw):
        return [getattr(r, fn)(*a, **kw) for r in self.ranks]

    def _handles(self):
        return (C.c_void_p * self.nranks)(*[r._h for r in self.ranks])

    def _ck(self, rc):
        if rc != 0:
            msgs = [self._lib.wf_last_error(r._h).decode() for r in self.ranks]
            raise WfError("; ".join(m for m in msgs if m) or f"error {rc}")

    # Domain_d-shaped surface used by cases.Case.apply
    def set_domtype(self, *a): self._each("set_domtype", *a)
    def box(self, *a): self._each("box", *a); self._shape()
    def set_mesh(self, *a): self._each("set_mesh", *a); self._shape()
    def set_material(self, *a, **kw): self._each("set_material", *a, **kw)
    def set_stab(self, **kw): self._each("set_stab", **kw)
    def set_options(self, *a, **kw): self._each("set_options", *a, **kw)
    def set_tracking(self, **kw): self._each("set_tracking", **kw)
    def add_bcs(self, *a): self._each("add_bcs", *a)
    def allocate_bcs(self): self._each("allocate_bcs")

    def _shape(self):
        r0 = self.ranks[0]
        self.dim, self.nodxelem = r0.dim, r0.nodxelem
        self.n_elems = self.ranks[-1].partition.elem_end
        self.n_nodes = int(max(int(r.node_l2g.max()) for r in self.ranks)) + 1

    def connect(self):
        self._ck(self._lib.wf_connect_all(self._handles(), self.nranks))
        self._connected = True

    def init(self, dt=None):
        if dt is not None:
            self._dt = float(dt)
        if not self._connected:
            self.connect()
        self._ck(self._lib.wf_init_all(self._handles(), self.nranks, self._dt))

    def step(self, n=1):
        self._ck(self._lib.wf_step_all(self._handles(), self.nranks, int(n)))

    def synchronize(self):
        self._each("synchronize")
        for r in self.ranks:
            r.halo_status()

    def info(self):
        return dict(dim=self.dim, nodxelem=self.nodxelem, n_nodes=self.n_nodes, n_elems=self.n_elems)

    def get(self, name):
        self.synchronize()
        parts = [(r.node_l2g, r.elem_range(), r.get(name)) for r in self.ranks]
        return assemble_global(name, parts, self.dim, self.nodxelem, self.n_nodes, self.n_elems)

    def set(self, name, arr):
        arr = np.asarray(arr, dtype=np.float64)
        for r in self.ranks:
            if name in _ELEM_ARRAYS:
                eb, ee = r.elem_range()
                per = arr.size // self.n_elems
                r.set(name, arr[eb * per:ee * per])
            else:
                per = 1 if name in _NODE_SCALARS else self.dim
                r.set(name, arr.reshape(-1, per)[r.node_l2g].reshape(-1))

    def nonfinite_flag(self):
        return any(self._each("nonfinite_flag"))

    def close(self):
        self._each("close")
