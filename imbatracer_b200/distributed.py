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


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class ProgressiveRenderer:
    """Render loop of frontend/render_window.cpp:50-95 for N ranks: integrator.render(1) + framebuffer sum per iteration.
    `image()` = total / (iterations * world * spp), the normalisation of render_window.cpp:105,178."""

    def __init__(self, integrator, width, height, spp=1, group=None):
        self.integ, self.group, self.spp = integrator, group, spp
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.fb = torch.zeros(width * height * 3, dtype=torch.float32, device="cuda")
        self.total = torch.zeros_like(self.fb)
        self.integ.set_framebuffer(self.fb)
        self.stream = torch.cuda.ExternalStream(self.integ.stream_ptr())
        self.iterations = 0
        self.shape = (height, width, 3)

    def step(self, iterations=1):
        for _ in range(iterations):
            self.integ.render(1)
            accumulate_iteration(self.fb, self.total, self.group, self.stream)
            self.iterations += 1

    def image(self):
        self.integ.sync()
        return (self.total / float(self.iterations * self.world * self.spp)).reshape(self.shape)
