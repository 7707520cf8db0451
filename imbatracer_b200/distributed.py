"""Multi-GPU plumbing for progressive rendering: one process per GPU, scene replicated, progressive iterations sharded
over ranks, one sum of the W*H*3 float framebuffer per iteration.  torch.distributed is plumbing only; the
reduced tensor is the integrator's own device framebuffer (imba_integrator_set_framebuffer)."""
import torch
import torch.distributed as dist


def global_iteration(local_iteration: int, rank: int, world: int) -> int:
    """Rank r of R renders progressive iterations r, r+R, r+2R, ... (the index the C side derives seeds and the PPM/VCM
    radius schedule from, csrc/integrator.cu, csrc/vcm_host.cu)."""
    return local_iteration * world + rank


def accumulate_iteration(fb: torch.Tensor, total: torch.Tensor, group=None, stream=None):
    """Sums this iteration's per-rank framebuffers over all ranks and adds the result to `total`; clears `fb`.
    With `stream` (a torch.cuda.Stream wrapping the integrator's stream) the collective is ordered after the kernels."""
    ctx = torch.cuda.stream(stream) if stream is not None else _null()
    with ctx:
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(fb, op=dist.ReduceOp.SUM, group=group)
        total.add_(fb)
        fb.zero_()


def agree_to_stop(local_stop: bool, group=None, device=None) -> bool:
    """Collective stop decision for time-limited rendering (-t, render_window.cpp:71-74): every rank measures its own wall clock,
    and every step of the render loop contains a framebuffer all-reduce -- so ranks must leave the loop after the SAME step or
    the slower one waits for a collective nobody joins.  All ranks call this once per step; the loop stops as soon as any rank's
    clock has run out (MAX over the flags)."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return bool(local_stop)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    flag = torch.tensor([1 if local_stop else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    return bool(int(flag.item()))


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class ProgressiveRenderer:
    """Render loop of frontend/render_window.cpp:50-95 for N ranks: integrator.render(1) + framebuffer sum per iteration.
    `image()` = total / (iterations * world * spp), the normalisation of render_window.cpp:105,178.

    With more than one rank the sum is taken off the critical path: the integrator renders alternately into two framebuffers,
    and the NCCL sum + accumulation of iteration k runs on a second stream while iteration k+1 is being rendered (events order
    the hand-over both ways), so the collective costs no device time as long as it is shorter than one iteration."""

    def __init__(self, integrator, width, height, spp=1, group=None, overlap=None, iterations_per_step=1):
        """iterations_per_step: progressive iterations rendered (and summed over the ranks) together; pair it with the
        integrator's `iterations_per_pass` so that PT renders them as one large wavefront pass."""
        self.integ, self.group, self.spp, self.per_step = integrator, group, spp, max(1, int(iterations_per_step))
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.overlap = (self.world > 1) if overlap is None else bool(overlap)
        self.fbs = [torch.zeros(width * height * 3, dtype=torch.float32, device="cuda") for _ in range(2 if self.overlap else 1)]
        self.fb = self.fbs[0]
        self.total = torch.zeros_like(self.fb)
        self.integ.set_framebuffer(self.fb)
        self.stream = torch.cuda.ExternalStream(self.integ.stream_ptr())
        if self.overlap:
            self.comm = torch.cuda.Stream()
            self.rendered = [torch.cuda.Event() for _ in range(2)]
            self.consumed = [torch.cuda.Event() for _ in range(2)]
            for e in self.consumed:
                e.record(self.comm)
        self.iterations = 0; self.steps = 0
        self.shape = (height, width, 3)

    def step(self, iterations=1):
        done = 0
        while done < iterations:
            n = min(self.per_step, iterations - done)
            if not self.overlap:
                self.integ.render(n)
                accumulate_iteration(self.fb, self.total, self.group, self.stream)
            else:
                k = self.steps % 2
                self.stream.wait_event(self.consumed[k])          # buffer k was summed and cleared two steps ago
                self.integ.set_framebuffer(self.fbs[k])
                self.integ.render(n)
                self.rendered[k].record(self.stream)
                self.comm.wait_event(self.rendered[k])
                accumulate_iteration(self.fbs[k], self.total, self.group, self.comm)
                self.consumed[k].record(self.comm)
            self.iterations += n; self.steps += 1; done += n

    def finish(self):
        """Orders everything outstanding on the reduction stream before later work on the integrator's stream."""
        if self.overlap:
            for e in self.consumed:
                self.stream.wait_event(e)

    def image(self):
        self.finish()
        self.integ.sync()
        return (self.total / float(self.iterations * self.world * self.spp)).reshape(self.shape)
