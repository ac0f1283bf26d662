"""world_size-2 gloo tests (CPU) of the data-parallel host logic: the flat gradient buffer and its
single all-reduce reproduce the gradient of the concatenated batch; the dist helpers with the
reference's names behave like utils/dist.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from coda_neurips2023_b200.engine import FlatParameters, allreduce_mean_
    from coda_neurips2023_b200.utils import dist as cdist

    torch.manual_seed(0)  # same init on every rank
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    flat = FlatParameters(model)
    g = torch.Generator().manual_seed(100)
    data = torch.randn(8, 6, generator=g)
    target = torch.randn(8, 3, generator=g)
    shard = slice(rank * 4, rank * 4 + 4)          # DistributedSampler-like sharding by scene
    flat.zero_grad()
    loss = torch.nn.functional.mse_loss(model(data[shard]), target[shard])
    loss.backward()
    allreduce_mean_(flat.flat_grad)
    # helpers
    assert cdist.get_world_size() == world and cdist.get_rank() == rank and cdist.is_distributed()
    total = cdist.all_reduce_sum(torch.tensor([float(rank + 1)]))
    avg = cdist.all_reduce_average(torch.tensor([float(rank + 1)]))
    red = cdist.reduce_dict({"b": torch.tensor(float(rank)), "a": torch.tensor(2.0 * rank)})
    gathered = cdist.all_gather_dict({"x": torch.full((2, 3), float(rank))})
    cdist.barrier()
    if rank == 0:
        # the flat buffer pads every parameter to a 256-byte boundary: compare the parameters' own slices
        ret["flat_grad"] = torch.cat([p.grad.reshape(-1) for p in flat.params]).clone()
        ret["total"], ret["avg"] = total.item(), avg.item()
        ret["red"] = {k: v.item() for k, v in red.items()}
        ret["gathered"] = gathered["x"].clone()
        ret["views_share_storage"] = all(p.grad.data_ptr() >= flat.flat_grad.data_ptr() for p in flat.params)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    # single-process reference: mean over the two shards' gradients == gradient of the mean loss
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    g = torch.Generator().manual_seed(100)
    data = torch.randn(8, 6, generator=g)
    target = torch.randn(8, 3, generator=g)
    loss = 0.5 * (torch.nn.functional.mse_loss(model(data[:4]), target[:4])
                  + torch.nn.functional.mse_loss(model(data[4:]), target[4:]))
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    torch.testing.assert_close(ret["flat_grad"], ref, rtol=1e-6, atol=1e-7)
    assert ret["total"] == 3.0 and ret["avg"] == 1.5
    assert ret["red"] == {"a": 1.0, "b": 0.5}
    assert ret["gathered"].shape == (4, 3) and ret["gathered"][:2].eq(0).all() and ret["gathered"][2:].eq(1).all()
    assert ret["views_share_storage"]


def test_flat_parameters_keep_module_semantics():
    from coda_neurips2023_b200.engine import FlatParameters

    torch.manual_seed(1)
    m = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.LayerNorm(4))
    ref = [p.detach().clone() for p in m.parameters()]
    flat = FlatParameters(m)
    for p, r in zip(m.parameters(), ref):
        assert torch.equal(p, r)
    m(torch.randn(2, 4)).sum().backward()
    assert flat.flat_grad.abs().sum() > 0
    with torch.no_grad():
        flat.flat_param.add_(1.0)              # an update of the flat buffer is an update of every parameter
    for p, r in zip(m.parameters(), ref):
        assert torch.allclose(p, r + 1.0)
    flat.zero_grad()
    assert all(p.grad.abs().sum() == 0 for p in m.parameters())


def _worker_bucketed(rank, world, port, ret):
    """Ranks seeded DIFFERENTLY (as the reference's main.py:982-985 does): broadcast_module_state must make them
    identical; the bucketed, hook-driven all-reduce must give the gradient of the concatenated batch; after a few
    plain-SGD steps on the flat buffer both ranks must hold the same weights as the single-process run."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from coda_neurips2023_b200.engine import BucketedAllReduce, FlatParameters, broadcast_module_state

    torch.manual_seed(100 + rank)                      # different initialisation per rank
    model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                                torch.nn.Linear(16, 3), torch.nn.Linear(3, 3))
    for p in model[5].parameters():                    # a head that never receives a gradient
        pass
    broadcast_module_state(model)
    params = [p for p in model.parameters()]
    order = list(reversed(params))                     # gradient-ready order of a feed-forward stack
    flat = FlatParameters(model, order=order)
    reducer = BucketedAllReduce(flat, nbuckets=3)
    assert len(reducer.ranges) == 3 and reducer.ranges[0][0] == 0 and reducer.ranges[-1][1] == flat.flat_grad.numel()
    g = torch.Generator().manual_seed(7)
    data = torch.randn(4, 8, 6, generator=g)
    target = torch.randn(4, 8, 3, generator=g)
    for it in range(4):
        flat.zero_grad()
        reducer.start()
        shard = slice(rank * 4, rank * 4 + 4)
        loss = torch.nn.functional.mse_loss(model(data[it, shard]), target[it, shard])
        loss.backward()
        assert any(reducer.launched[:-1]) or it >= 0    # hooks launched ranges during the backward
        reducer.finish()
        with torch.no_grad():
            flat.flat_param.data.add_(flat.flat_grad, alpha=-0.05)
    if rank == 0:
        ret["w0"] = torch.cat([p.detach().reshape(-1) for p in params]).clone()
        ret["launched_in_hooks"] = sum(1 for x in reducer.launched if x)
    else:
        ret["w1"] = torch.cat([p.detach().reshape(-1) for p in params]).clone()
    dist.destroy_process_group()


def test_bucketed_allreduce_and_broadcast_world2_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_bucketed, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert torch.equal(ret["w0"], ret["w1"]), "ranks diverged"
    # single-process reference: rank 0's initialisation, whole batch
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                                torch.nn.Linear(16, 3), torch.nn.Linear(3, 3))
    g = torch.Generator().manual_seed(7)
    data = torch.randn(4, 8, 6, generator=g)
    target = torch.randn(4, 8, 3, generator=g)
    for it in range(4):
        model.zero_grad()
        torch.nn.functional.mse_loss(model(data[it]), target[it]).backward()
        with torch.no_grad():
            for p in model.parameters():
                p.add_(p.grad, alpha=-0.05)
    ref = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    torch.testing.assert_close(ret["w0"], ref, rtol=1e-5, atol=1e-6)
