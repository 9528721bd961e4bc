"""CPU, world_size 2 over gloo: the data-parallel host logic (yet_another_mobilenet_series_b200/
distributed.py) — parameter broadcast at wrap time, gradient mean, BN-statistics mean — against
the closed form of the reference (utils/distributed.py:131-139: SUM all-reduce then / world) and
the oracle's allreduce_mean."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    from yet_another_mobilenet_series_b200 import distributed as ud
    ud.init_dist(backend="gloo")
    assert ud.get_world_size_fallback() == world and ud.get_rank_fallback() == rank
    assert ud.is_master() == (rank == 0)
    torch.manual_seed(100 + rank)  # different weights per rank before the wrap
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1, bias=False), torch.nn.BatchNorm2d(4),
                              torch.nn.Flatten(), torch.nn.Linear(4 * 4, 5))
    ddp = ud.AllReduceDistributedDataParallel(net)
    w_after = net[0].weight.detach().clone()
    g = torch.Generator().manual_seed(rank)
    x = torch.randn(6, 3, 2, 2, generator=g)
    ddp(x).square().mean().backward()
    local = [p.grad.clone() for p in net.parameters()]
    ud.allreduce_grads(ddp)
    rm_local = net[1].running_mean.clone()
    ud.allreduce_bn(ddp)
    out.put((rank, w_after.numpy(), [t.numpy() for t in local],
             [p.grad.numpy().copy() for p in net.parameters()], rm_local.numpy(),
             net[1].running_mean.numpy().copy()))
    called = []
    ud.master_only(lambda: called.append(1))()
    assert len(called) == (1 if rank == 0 else 0)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_ddp_logic_world2():
    from oracle import optim as oo
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=90) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    (_, w0, loc0, red0, rm0, rmr0), (_, w1, loc1, red1, rm1, rmr1) = res
    np.testing.assert_array_equal(w0, w1)  # rank 0's weights everywhere after the wrap
    for a, b, ra, rb in zip(loc0, loc1, red0, red1):
        want = oo.allreduce_mean([a, b])
        np.testing.assert_allclose(ra, want, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(rb, want, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rmr0, (rm0 + rm1) / 2, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rmr1, rmr0)
