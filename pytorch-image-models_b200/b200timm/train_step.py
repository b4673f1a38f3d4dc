"""One training step -- zero_grad -> forward -> loss -> backward (+ DDP gradient all-reduce) -> optimizer step -- captured
into ONE CUDA graph and replayed.

The step is what benchmark.py:396-420 / train.py:1236-1269 run, wrapped the way timm/task/classification.py:64-65 wraps
the model in DistributedDataParallel.  Eagerly it is ~3 300 kernel launches, every one a ctypes call plus (for the GEMM /
attention kernels) tensor-map encodes on the host, plus autograd's Python frames and DDP's per-parameter hooks: round 1
measured 23 ms of host time for a 32 ms step on one GPU and 30 of 35 ms under DDP -- the step was about to become
host-bound and every stall of the Python thread showed up as an idle GPU.  Captured, the whole step is one
cudaGraphLaunch: the kernels (ours, NCCL's all-reduce kernels on DDP's communication stream, the optimizer) replay with
their dependencies as graph edges, the host is free, and the all-reduce of bucket k overlaps the backward GEMMs of the
layers below it exactly as it does eagerly.

Static-shape contract: inputs are copied into the captured input / target tensors (`__call__`) -- or the graph is captured
on caller-owned resident buffers (`capture_on`).  Hyper-parameters the optimizer kernels take by value (lr, momentum,
weight decay) are baked into the graph: after changing them (an lr scheduler step) call `recapture()`; timm's schedulers
step once per epoch / per update group, and `lr_changed()` tells whether that is needed.

DDP + capture follows the PyTorch CUDA-graphs notes: NCCL async error handling off before the process group is created
(`prepare_ddp_env()`), the DDP wrapper constructed on a side stream, at least 11 eager iterations before capture (DDP's
logger reads CUDA event timings during the first ten).
"""
import os

import torch

from . import ops
from ._lib import B200Error

DDP_WARMUP = 11


def prepare_ddp_env():
    """Call before torch.distributed.init_process_group when a DDP step is going to be captured."""
    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
    os.environ.setdefault("NCCL_ASYNC_ERROR_HANDLING", "0")


def wrap_ddp(model, device, **kw):
    """DistributedDataParallel as timm wraps it (timm/task/classification.py:64-65: device_ids=[device],
    broadcast_buffers=True), constructed on a side stream (required for capturing its backward hooks), with the gradients
    living inside the all-reduce buckets (gradient_as_bucket_view) so that no copy separates the backward kernels'
    output from NCCL's input."""
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        kw.setdefault("gradient_as_bucket_view", True)
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index], broadcast_buffers=True, **kw)
    torch.cuda.current_stream(device).wait_stream(side)
    return net


class TrainStep:
    """step = TrainStep(net, loss_fn, optimizer, example_x, example_y); loss = step(x, y)

    net: the model or its DistributedDataParallel wrapper.  graph=False keeps the eager step (same call signature)."""

    def __init__(self, net, loss_fn, optimizer, example_x, example_y, graph=True, warmup=None, grad_sync=None):
        if not example_x.is_cuda:
            raise B200Error("TrainStep needs CUDA tensors (there is no CPU path)")
        self.net, self.loss_fn, self.opt = net, loss_fn, optimizer
        self.is_ddp = isinstance(net, torch.nn.parallel.DistributedDataParallel)
        self.graph_enabled = graph
        self.grad_sync = grad_sync      # b200timm.grad_sync.PeerGradSync: fused gradient exchange instead of DDP's all-reduce
        if grad_sync is not None and self.is_ddp:
            raise B200Error("TrainStep: pass either a DistributedDataParallel wrapper or grad_sync, not both")
        self.x = example_x.clone()
        self.y = example_y.clone()
        self.launches = 0           # kernels of libb200timm inside one step (counted on the capture pass)
        self._graphs = {}           # (x address, y address) -> (graph, static loss); all graphs share one memory pool: they
        self._pool = None           # are never replayed concurrently and nothing but parameters / optimizer state / the
        self._lr_key = None         # loss scalars has to survive from one replay to the next
        self._warm = False
        self.warmup = warmup if warmup is not None else (DDP_WARMUP if self.is_ddp else 3)

    # ---------------------------------------------------------------------------------------------- the step itself
    def _eager(self, x, y):
        sync = self.grad_sync
        if sync is None:
            self.opt.zero_grad(set_to_none=True)
            loss = self.loss_fn(self.net(x), y)
            loss.backward()
        else:
            sync.begin()                       # zero the exchange buffer, peer barrier, tail grads -> local staging
            try:
                loss = self.loss_fn(self.net(x), y)
                loss.backward()                # BlockFn pushes every block's gradients to all ranks as it finishes
            except BaseException:
                sync.abort()
                raise
            sync.finish()                      # tail push, peer barrier: every p.grad now holds the mean over ranks
        self.opt.step()
        return loss

    def _hyper(self):
        return tuple(tuple((k, v) for k, v in sorted(g.items()) if k != "params" and isinstance(v, (int, float, bool, tuple)))
                     for g in self.opt.param_groups)

    def lr_changed(self):
        return self.graph_enabled and self._hyper() != self._lr_key

    def _capture(self, x, y, warmup=None):
        dev = x.device
        if warmup is None:
            warmup = 1 if self._warm else self.warmup
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):       # eager iterations: lazy kernel attributes, allocator growth, optimizer state, DDP
                self._eager(x, y)         # bucket rebuild + its first-iterations logging
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        if self.grad_sync is None:
            self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(g, pool=self._pool):
            loss = self._eager(x, y)
        if self._pool is None:
            self._pool = g.pool()
        self.launches = ops.launch_count() - n0
        self._lr_key = self._hyper()
        self._warm = True
        return g, loss

    def recapture(self):
        """After an optimizer hyper-parameter change (lr schedule): the kernels take lr / momentum / weight decay by value."""
        for key in list(self._graphs):
            self._graphs[key] = None
        self._graphs.clear()

    def capture_on(self, x, y):
        """A graph that reads caller-owned resident buffers in place (no staging copy).  Returns a zero-argument callable
        that replays it and returns the (static) loss tensor."""
        if not self.graph_enabled:
            return lambda: self._eager(x, y)
        key = (x.data_ptr(), y.data_ptr())
        if self._graphs.get(key) is None:
            self._graphs[key] = self._capture(x, y)
        g, loss = self._graphs[key]

        def replay():
            g.replay()
            ops.note_params_written()       # the optimizer kernels inside the graph wrote the parameters
            return loss
        return replay

    def __call__(self, x, y):
        if not self.graph_enabled:
            return self._eager(x, y)
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        return self.capture_on(self.x, self.y)()
