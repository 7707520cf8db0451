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


def _stop_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the render loop of imbatracer_b200/render.py in -t mode: one framebuffer sum per step, then the stop decision;
    # rank 1's clock runs out two steps before rank 0's
    deadline = 6 if rank == 0 else 4
    total = torch.zeros(8)
    steps = 0
    while True:
        D.accumulate_iteration(torch.ones(8), total)          # a collective every step: both ranks must take part
        steps += 1
        if D.agree_to_stop(steps >= deadline):
            break
        assert steps < 50
    out.put((rank, steps, float(total[0])))
    dist.barrier(); dist.destroy_process_group()


def test_time_limited_loop_stops_on_all_ranks_after_the_same_step_gloo_world2():
    """render.py -t: each rank decides from its own wall clock; without a collective decision the slower rank would issue an
    all-reduce its peer never joins (hang).  The loop ends on every rank with the step at which the FIRST clock ran out."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_stop_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60); assert p.exitcode == 0
    assert [g[1] for g in got] == [4, 4]                      # both stop after step 4
    assert [g[2] for g in got] == [8.0, 8.0]                  # 4 steps x 2 ranks summed on both


def test_agree_to_stop_without_process_group():
    assert D.agree_to_stop(True) is True and D.agree_to_stop(False) is False
