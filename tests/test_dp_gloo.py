"""World-size-2 data-parallel gradient bucketing on CPU with the gloo backend (no GPU, no NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from transmf_ad_b200.dp import FlatGradReducer, shard_slice


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.ReLU(),
                               torch.nn.Linear(8, 2))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _make_model()
        red = FlatGradReducer(model.parameters())
        g = torch.Generator().manual_seed(1)
        x = torch.randn(8, 16, generator=g)
        y = torch.randint(0, 2, (8,), generator=g)
        sl = shard_slice(8, rank, world)
        for step in range(2):                                               # two steps: bucket state must reset
            model.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(model(x[sl]), y[sl])
            loss.backward()
            red.finish()
        # plain lists: tensors would travel as shared-memory handles served by this (soon exiting) process
        grads = [p.grad.tolist() for p in model.parameters()]
        q.put((rank, grads, red.layout(), red.allreduce_launches))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_equals_mean_of_shard_gradients():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process expectation: mean over ranks of the per-shard gradients
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 16, generator=g)
    y = torch.randint(0, 2, (8,), generator=g)
    expect = None
    for r in range(world):
        model = _make_model()
        sl = shard_slice(8, r, world)
        torch.nn.functional.cross_entropy(model(x[sl]), y[sl]).backward()
        gr = [p.grad.clone() for p in model.parameters()]
        expect = gr if expect is None else [a + b for a, b in zip(expect, gr)]
    expect = [e / world for e in expect]
    for rank, grads, layout, launches in results:
        assert layout[1] == 6 and layout[0] % 64 == 0 and launches == 2        # one all-reduce per step
        for a, b in zip(grads, expect):
            assert torch.allclose(torch.tensor(a), b, atol=1e-6), rank


def test_shard_slice_partitions_the_batch():
    assert [shard_slice(64, r, 8) for r in range(8)][3] == slice(24, 32)
    with pytest.raises(ValueError):
        shard_slice(10, 0, 4)


def test_reducer_is_a_noop_without_process_group():
    model = _make_model()
    red = FlatGradReducer(model.parameters())
    model(torch.randn(4, 16)).sum().backward()
    red.finish()
    assert red.allreduce_launches == 0
