"""N > 1 host-side logic on CPU: two gloo processes, each owning half of the
subdomains, agree on the connection plan (who stores into whose mailbox, at
which displacement, under which flag slot) and it matches the oracle's
put_displacements / neighbour tables — the information the reference obtains
with its MPI handshake and MPI_Alltoall (source/restricted_schwarz.cpp:400-472,
624-658)."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, P, q):
    sys.path.insert(0, os.path.join(ROOT, "schwarz-lib_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import schwz_b200 as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        part = S.partition_regular2d(n * n, P)
        setup = S.Setup(("laplacian2d", n), P, part=part)
        nl = P // world
        my = list(range(rank * nl, (rank + 1) * nl))
        # what each process publishes about its subdomains (the IPC handle is a
        # stand-in on CPU): mailbox layout from the index-set sizes + in-lists
        mine = {}
        for r in my:
            nin, _ = setup.neighbors(r)
            in_total = sum(len(setup.get_list(r, j)) for j in range(len(nin)))
            lay = S.mailbox_layout(in_total, len(nin), P)
            mine[r] = (b"handle-of-%d" % r, lay.as_tuple(), nin.tolist(), in_total)
        allinfo = [None] * world
        dist.all_gather_object(allinfo, mine)
        info = {}
        for d in allinfo:
            info.update(d)
        assert sorted(info) == list(range(P))
        plan = S.remote_connection_plan(setup, my, {k: v[2] for k, v in info.items()})
        # cross-check with the oracle (one process, all subdomains)
        ob = O.Problem(*O.laplacian2d(n), P, part=part)
        seen = set()
        for r, j, peer, recv_off, slot in plan:
            assert peer not in my and r in my
            nin_o, nout_o = ob.neighbors(r)
            assert int(nout_o[j]) == peer
            pd, _ = ob.displacements(r)
            assert recv_off == int(pd[peer])
            peer_in, _ = ob.neighbors(peer)
            assert int(peer_in[slot]) == r
            # my block fits in the peer's receive buffer and does not overlap others
            cnt = len(ob.put_list(r, j))
            assert recv_off + cnt <= info[peer][3]
            lay = info[peer][1]
            assert lay[1] == 2 * lay[0] and lay[0] >= 8 * info[peer][3]   # flags after 2 buffers
            seen.add((r, peer))
        # every remote out-neighbour is covered exactly once
        want = set()
        for r in my:
            _, nout = ob.neighbors(r)
            want |= {(r, int(p)) for p in nout if int(p) not in my}
        assert seen == want
        # ordered sum of the gathered residual norms is identical on every process
        norms = np.array([1.0 / (3 + r) for r in my])
        gathered = [None] * world
        dist.all_gather_object(gathered, norms.tolist())
        total = 0.0
        for part_ in gathered:
            for v in part_:
                total += v
        q.put((rank, len(plan), repr(total)))
    finally:
        dist.destroy_process_group()


def test_connection_plan_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 16, 4, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert res[0][1] > 0 and res[1][1] > 0
    assert res[0][2] == res[1][2]
