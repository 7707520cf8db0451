"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: iteration sharding and the per-iteration framebuffer sum."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imbatracer_b200 import distributed as D


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h, iters = 16, 8, 5
    total = torch.zeros(w * h * 3)
    seen = []
    for k in range(iters):
        g = D.global_iteration(k, rank, world)
        seen.append(g)
        fb = torch.full((w * h * 3,), float(g + 1))          # stands for the iteration's device framebuffer
        D.accumulate_iteration(fb, total)
        assert float(fb.abs().sum()) == 0.0                   # cleared for the next iteration
    gathered = [None] * world
    dist.all_gather_object(gathered, seen)
    if rank == 0:
        out.put((total.numpy().copy(), gathered))
    dist.barrier(); dist.destroy_process_group()


def test_iterations_are_sharded_and_framebuffers_summed_gloo_world2():
    world, iters = 2, 5
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    total, gathered = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60); assert p.exitcode == 0
    all_iters = sorted(i for s in gathered for i in s)
    assert all_iters == list(range(world * iters))            # every progressive iteration exactly once
    expect = sum(i + 1 for i in range(world * iters))          # every rank holds the sum over ALL iterations
    assert np.allclose(total, expect)


def test_single_process_accumulate_is_plain_sum():
    total = torch.zeros(12); fb = torch.arange(12, dtype=torch.float32)
    D.accumulate_iteration(fb, total)
    assert torch.equal(total, torch.arange(12, dtype=torch.float32)) and float(fb.sum()) == 0.0
