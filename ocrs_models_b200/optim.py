"""Step glue on the device: flat parameter/gradient buffers, fused grad-norm + clip + Adam, and the
data-parallel gradient all-reduce (one NCCL call on the flat bucket).

Mirrors what the reference scripts do per step with ``torch.optim.Adam(model.parameters())``
(train_detection.py:378, train_rec.py:381-382) and ``clip_grad_norm_(..., 4.0)`` (train_rec.py:148).
The reference has no multi-GPU path; the all-reduce follows standard DDP semantics (averaged
gradients, per-replica BatchNorm statistics), SURVEY section 8e.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """The one data-path collective of the DDP step: sum the flat gradient bucket over the ranks in
    place (NCCL over NVLink on GPUs; gloo in the CPU tests). The 1/world averaging is folded into the
    clip/Adam kernels as `grad_scale`, so no extra pass touches the bucket."""
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM, group=group)
    return flat


def shard_seed(base_seed: int, rank: int) -> int:
    """Per-rank data seed (generator seed 1234 + rank); weights use the un-offset seed."""
    return base_seed + rank


class FusedAdam:
    """Adam (+ optional global-norm clipping) over a model whose parameters and gradients are
    re-homed into two flat fp32 buffers. ``step()`` = [all-reduce] -> norm -> clip+Adam, 3 launches."""

    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: float | None = None, process_group=None, world_size: int = 1):
        params = [p for p in model.parameters() if p.requires_grad]
        if not params or not params[0].is_cuda:
            raise RuntimeError("FusedAdam needs CUDA parameters (no CPU path)")
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4  # keep every tensor 16-byte aligned
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self.flat_p[o : o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_g[o : o + p.numel()].view_as(p)
                # the backward passes of this package accumulate straight into this view (grads.deliver) instead of
                # handing autograd one tensor per parameter to add
                p._ocrs_grad_sink = p.grad
        self.params, self.n = params, total
        self.lr, self.betas, self.eps = lr, betas, eps
        self.max_grad_norm = max_grad_norm
        self.group, self.world = process_group, world_size
        self.t = 0
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)  # steps taken, for the CUDA-graph path
        lib = _lib.lib()
        self.partials = torch.empty(lib.ocrs_optim_blocks(), dtype=torch.float32, device=dev)
        self.norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.dev = dev

    def zero_grad(self):
        self.flat_g.zero_()
        self._synced = False

    def sync_gradients(self) -> torch.Tensor:
        """The data-path collective of the step: SUM the flat gradient bucket over the ranks (once per step; `step()`
        calls it if the caller did not). Returns the bucket, which then holds world * (mean gradient): the 1/world
        factor is applied inside the clip/Adam kernels (`grad_scale`), not by another pass over the bucket."""
        if self.world > 1 and not getattr(self, "_synced", False):
            allreduce_sum_(self.flat_g, self.group)
        self._synced = True
        return self.flat_g

    def averaged_gradients(self) -> dict:
        """name-less view for tests: list of (parameter, averaged gradient) after `sync_gradients()`."""
        self.sync_gradients()
        return [(p, p.grad / self.world) for p in self.params]

    def step(self):
        """Returns the (device) gradient-norm tensor when clipping is on, else None."""
        self.sync_gradients()
        return self.apply_update()

    def apply_update(self):
        """Norm + clip + Adam on the (already reduced) flat gradient bucket: the part of `step()` after the collective."""
        st = _lib.stream_ptr(self.dev)
        scale = 1.0 / self.world
        self.t += 1
        clip = self.max_grad_norm if self.max_grad_norm is not None else -1.0
        with torch.cuda.device(self.dev):
            if clip > 0:
                call("ocrs_grad_norm", ptr(self.flat_g), self.n, scale, ptr(self.partials), ptr(self.norm), st)
            # the step count lives on the device (bias corrections are computed in the kernel), so the same recorded
            # launch is valid for every replay of a captured step; self.t mirrors it on the host for eager use
            call("ocrs_adam_step_dev", ptr(self.flat_p), ptr(self.flat_g), ptr(self.m), ptr(self.v), self.n, self.lr,
                 self.betas[0], self.betas[1], self.eps, ptr(self.step_dev), scale, clip, ptr(self.norm), st)
        return self.norm if clip > 0 else None


class GraphedTrainStep:
    """One whole training step (zero_grad -> forward -> loss -> backward -> [all-reduce] -> clip + Adam) captured ONCE into
    a CUDA graph and replayed: the recognition step is ~280 kernel launches of ~30 us each, and the Python / ctypes /
    autograd work that issues them costs about as much host time as the GPU needs to run them, so eager execution is
    launch-bound as soon as the kernels get faster. Replay costs one cudaGraphLaunch.

    Requirements (checked or documented): fixed batch shapes; the batch lives in the static device buffers this object
    owns (`load(batch)` copies into them, host tensors should be pinned); lengths of the CTC loss are passed as DEVICE
    tensors (a pageable host->device copy cannot be recorded). BatchNorm running statistics, `num_batches_tracked`,
    the Adam moments and the device step counter advance on every replay exactly as in eager mode.
    """

    def __init__(self, model: torch.nn.Module, opt: FusedAdam, loss_from_batch, example_batch: dict, warmup: int = 3):
        """loss_from_batch(model, batch_dict_of_static_device_tensors) -> scalar loss tensor."""
        self.model, self.opt, self.loss_fn = model, opt, loss_from_batch
        dev = opt.dev
        self.static = {k: (v.to(dev).clone() if isinstance(v, torch.Tensor) else v) for k, v in example_batch.items()}
        lib = _lib.lib()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside the capture: lazy initialisation, allocator pools, TMA descriptors
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # Multi-GPU: the NCCL all-reduce stays OUTSIDE the graphs (graph 1 = zero_grad + forward + loss + backward, eager
        # all-reduce of the flat bucket, graph 2 = norm + clip + Adam): three host calls per step instead of ~300, and no
        # collective is ever recorded (capturing NCCL inside the step graph hung the 2-GPU bench).
        self.split = opt.world > 1
        self.graph = torch.cuda.CUDAGraph()
        l0 = lib.ocrs_launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = self._fwd_bwd()
            if not self.split:
                self.opt.step()
        self.graph2 = None
        if self.split:
            self.opt.sync_gradients()
            self.graph2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph2):
                self.opt.apply_update()
        self.launches_per_step = int(lib.ocrs_launch_count() - l0)  # kernels of this library inside one replay

    def _fwd_bwd(self):
        self.opt.zero_grad()
        loss = self.loss_fn(self.model, self.static)
        loss.backward()
        return loss.detach()

    def _eager(self):
        loss = self._fwd_bwd()
        self.opt.step()
        return loss

    def load(self, batch: dict):
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                self.static[k].copy_(v, non_blocking=True)

    # ---- input prefetch: the host->device copy of the NEXT batch runs on its own stream while the current step replays ----
    _copy_stream = None
    _has_staged = False

    def prefetch(self, batch: dict):
        """Start copying `batch` (pinned host tensors) into device staging buffers on a copy stream. The next `step()`
        call without a batch waits for it, moves it into the graph's static buffers (device-to-device) and replays."""
        dev = self.opt.dev
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staging = {k: torch.empty_like(v) for k, v in self.static.items() if isinstance(v, torch.Tensor)}
            self._staged_ev = torch.cuda.Event()
            self._free_ev = torch.cuda.Event()
            self._free_ev.record(torch.cuda.current_stream(dev))
        cs = self._copy_stream
        cs.wait_event(self._free_ev)  # the previous batch has left the staging buffers
        with torch.cuda.stream(cs):
            for k, v in batch.items():
                if isinstance(v, torch.Tensor):
                    self._staging[k].copy_(v, non_blocking=True)
            self._staged_ev.record(cs)
        self._staged_keys = [k for k, v in batch.items() if isinstance(v, torch.Tensor)]
        self._has_staged = True

    def _consume_staged(self):
        main = torch.cuda.current_stream(self.opt.dev)
        main.wait_event(self._staged_ev)
        for k in self._staged_keys:
            self.static[k].copy_(self._staging[k], non_blocking=True)
        self._free_ev.record(main)
        self._has_staged = False

    def __call__(self, batch: dict | None = None) -> torch.Tensor:
        """Replay the step (after copying `batch` into the static buffers when given). Returns the device loss tensor of
        this step (overwritten by the next replay)."""
        if batch is not None:
            self.load(batch)
        elif self._has_staged:
            self._consume_staged()
        self.graph.replay()
        if self.split:
            self.opt._synced = False
            self.opt.sync_gradients()
            self.graph2.replay()
        self.opt.t += 1
        return self.loss
