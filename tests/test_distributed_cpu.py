"""Host-side logic of the multi-GPU path on CPU: two gloo ranks build their partitions, publish their
descriptors and plan the connections exactly as a torchrun job does before it maps peer memory."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, box, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from weldformfem_b200.distributed import Partition, exchange_descriptors, plan_connections, slot_table
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        pt = Partition(world, rank, box=box)
        neigh, off = pt.neigh_ranks, pt.halo_offset
        l2g, halo = pt.node_l2g, pt.halo_nodes
        desc = {"rank": rank, "neigh": [int(x) for x in neigh], "slots": slot_table(neigh, off),
                "halo_global": {int(qr): l2g[halo[off[i]:off[i + 1]]].tolist() for i, qr in enumerate(neigh)},
                "elems": (pt.elem_begin, pt.elem_end)}
        pub = exchange_descriptors(desc)
        plan = plan_connections(rank, desc["neigh"], pub)
        q.put((rank, desc, plan, {r: pub[r]["halo_global"] for r in pub}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,box", [(2, ((0, 0, 0), (0.5, 0.4, 0.9), 0.05, False)),
                                        (2, ((0, 0, 0), (0.3, 0.3, 0.4), 0.05, True)),
                                        (3, ((0, 0, 0), (0.4, 0.3, 0.3), 0.05, True))])
def test_gloo_ranks_agree_on_halo_plan(world, box):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, box, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, desc, plan, halos = q.get(timeout=120)
        res[rank] = (desc, plan, halos)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # element blocks tile [0, Ne) in rank order
    edges = [res[r][0]["elems"] for r in range(world)]
    assert edges[0][0] == 0 and all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
    for r in range(world):
        desc, plan, halos = res[r]
        assert [p[0] for p in plan] == desc["neigh"]
        for (qr, fo, ro) in plan:
            # the slot this rank writes at the peer is the one the peer reserved for it
            assert res[qr][0]["slots"][r] == (fo, ro)
            # both sides hold the same shared-node list (global ids, ascending)
            mine, theirs = desc["halo_global"][qr], res[qr][0]["halo_global"][r]
            assert mine == theirs and mine == sorted(mine) and len(mine) > 0
        # receive regions of one rank do not overlap: region i spans 2*3*count_i doubles
        regs = sorted((ro, len(desc["halo_global"][qr])) for qr, (fo, ro) in desc["slots"].items())
        for (a, ca), (b, _) in zip(regs, regs[1:]):
            assert a + 8 * 6 * ca <= b
        flags = sorted(fo for fo, _ in desc["slots"].values())
        assert flags == [8 * i for i in range(len(flags))] and (not regs or regs[0][0] >= 8 * len(flags))


def test_plan_rejects_asymmetric_lists():
    from weldformfem_b200.distributed import plan_connections
    from weldformfem_b200.domain import WfError
    with pytest.raises(WfError):
        plan_connections(0, [1], {1: {"slots": {2: (0, 256)}}})


def test_assemble_global_detects_diverged_copies():
    from weldformfem_b200.distributed import assemble_global
    from weldformfem_b200.domain import WfError
    a = (np.array([0, 1, 2]), (0, 1), np.arange(6.0))
    b = (np.array([2, 3]), (1, 2), np.array([4.0, 5.0, 7.0, 8.0]))
    out = assemble_global("x", [a, b], 2, 3, 4, 2)
    assert np.array_equal(out, [0, 1, 2, 3, 4, 5, 7, 8])
    b_bad = (b[0], b[1], np.array([4.0, 5.5, 7.0, 8.0]))
    with pytest.raises(WfError):
        assemble_global("x", [a, b_bad], 2, 3, 4, 2)
    e = assemble_global("pl_strain", [(a[0], (0, 1), np.array([1.0])), (b[0], (1, 2), np.array([2.0]))], 2, 3, 4, 2)
    assert np.array_equal(e, [1.0, 2.0])
