"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: gene sharding without a
collective, the single gradient-bucket all-reduce of data-parallel training, result gathering."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chromoformer_b200 import parallel
from oracle import chromoformer_oracle as oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    r, _, w = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 18955, 18955 * 44):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    units = parallel.sweep_units(44, 18955, 4096)
    assert len(units) == 44 * 5 and units[4] == (0, 16384, 18955) and units[5] == (1, 0, 4096)


def _dp_step(rank, world):
    """Each rank holds a different micro-batch gradient; after the bucket all-reduce + the fused
    update's 1/world scale every rank must hold the same parameters as a single-process run on
    the concatenated batch (mean loss => mean of per-rank gradients)."""
    g = torch.Generator().manual_seed(0)
    n_total, n_active = 1000, 992
    p0 = torch.randn(n_total, generator=g)
    grads = [torch.randn(n_total, generator=g) for _ in range(world)]
    flat_grad = grads[rank].clone()
    scale = parallel.allreduce_gradients(flat_grad, n_active)
    p, m, v = oracle.adamw_step(p0[:n_active], flat_grad[:n_active] * scale, torch.zeros(n_active),
                                torch.zeros(n_active), 1)
    assert torch.equal(flat_grad[n_active:], grads[rank][n_active:])     # tail (grad-less tensors) untouched
    return p


def test_data_parallel_gradient_bucket_allreduce():
    outs = _spawn(_dp_step, 2)
    assert torch.equal(outs[0], outs[1])
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(1000, generator=g)
    grads = [torch.randn(1000, generator=g) for _ in range(2)]
    mean = (grads[0] + grads[1])[:992] * 0.5
    want, _, _ = oracle.adamw_step(p0[:992], mean, torch.zeros(992), torch.zeros(992), 1)
    assert torch.allclose(outs[0], want, rtol=0, atol=1e-7)


def _shard_and_gather(rank, world):
    n = 101
    lo, hi = parallel.shard_range(n, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32).unsqueeze(1).repeat(1, 2)     # "logits" of my genes
    return parallel.gather_sharded(local, n)


def test_sharded_inference_gather():
    outs = _spawn(_shard_and_gather, 2)
    want = torch.arange(101, dtype=torch.float32).unsqueeze(1).repeat(1, 2)
    assert torch.equal(outs[0], want) and torch.equal(outs[1], want)
