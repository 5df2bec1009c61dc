"""CPU: the N>1 exchange logic of fcdiff_b200.dist on a world_size-2 gloo group."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fcdiff_b200.dist import EdgeShards
        sh = EdgeShards()
        (C, N, U) = (45, 10, 7)
        ref_F = torch.arange(C * 3, dtype=torch.float64) * 0.5
        (c0, Cl, u0, Ul) = sh.ranges(C, U)
        lqF = torch.full((C * 3,), -1.0, dtype=torch.float64)
        qF = torch.full((C * 3,), -1.0, dtype=torch.float64)
        lqF[c0 * 3:(c0 + Cl) * 3] = ref_F[c0 * 3:(c0 + Cl) * 3]
        qF[c0 * 3:(c0 + Cl) * 3] = 2 * ref_F[c0 * 3:(c0 + Cl) * 3]
        sh.allgather_edges(lqF, qF, C)
        ok = bool(torch.equal(lqF, ref_F) and torch.equal(qF, 2 * ref_F))
        ref_R = torch.arange(N * U * 2, dtype=torch.float64).view(N, U, 2)
        lqR = torch.full((N, U, 2), -7.0, dtype=torch.float64)
        lqR[:, u0:u0 + Ul] = ref_R[:, u0:u0 + Ul]
        qR = lqR.clone()
        sh.allgather_patients(lqR.view(-1), qR.view(-1), N, U)
        ok = ok and bool(torch.equal(lqR, ref_R) and torch.equal(qR, ref_R))
        out = torch.tensor([1.0 + rank, 10.0, 100.0 * (rank + 1), 5.0], dtype=torch.float64)
        red = sh.fix_replicated(sh.allreduce_terms(out, (0, 2)).numpy().copy(), (0, 2))
        ok = ok and red.tolist() == [3.0, 10.0, 300.0, 5.0]
        q.put((rank, ok))
    except Exception as e:      # surface the failure instead of a queue timeout
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_collectives():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
