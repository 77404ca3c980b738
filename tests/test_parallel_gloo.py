"""CPU, world_size 2 over gloo: the N > 1 plumbing (router decisions on rank 0 -> per-rank request shares ->
completion barrier) and the start-up weight broadcast."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import router as O
    from rr_b200 import parallel
    from rr_b200.models import SPECS, make_weights, broadcast_weights
    n_req = 10
    dec = None
    if rank == 0:      # decisions come from K1 on the GPU box; here from the oracle with the same record shape
        orc = O.OracleRouter([O.Deployment(0, rpm=3, replica=0), O.Deployment(0, rpm=4, replica=1)], 1, {},
                             O.Settings(strategy=O.STRATEGY_LEAST_BUSY), seed=0)
        dec = [d.as_tuple() for d in orc.process([O.Event(O.EV_ADMIT, 0, 8, 0, 0) for _ in range(n_req)])]
    mine = parallel.scatter_assignments(dec, n_req, world, rank, torch.device("cpu"), replica_of=[0, 1])
    total = parallel.gather_done(len(mine), world, rank, torch.device("cpu"))
    w = make_weights(SPECS["tiny"], seed=5, device="cpu", allocate_only=(rank != 0))
    if rank != 0:
        for t in w.tensors():
            t.zero_()
    broadcast_weights(w, src=0)
    chk = float(sum(t.float().abs().sum() for t in w.tensors()))
    q.put((rank, mine.tolist(), total, chk))
    dist.destroy_process_group()


def test_two_rank_assignment_and_broadcast():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, mine0, tot0, chk0), (r1, mine1, tot1, chk1) = res
    # least-busy with rpm 3 / 4: alternate until rank 0's bucket is empty, then rank 1, then 429
    assert mine0 == [0, 2, 4] and mine1 == [1, 3, 5, 6]
    assert tot0 == tot1 == 7                      # 3 of 10 were rate limited
    assert chk0 == chk1 and chk0 > 0              # weights identical after the broadcast
