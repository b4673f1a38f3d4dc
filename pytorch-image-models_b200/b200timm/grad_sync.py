"""Data-parallel gradient exchange without an all-reduce kernel: the mean over ranks of every parameter gradient is formed
by `multimem.red.add` pushes over the NVLink multicast address of one flat, symmetric gradient buffer, issued by the
kernel that finalises a transformer block's gradients (csrc/grad_push.cu, b200_grad_push) while the backward GEMMs of the
next block run.  Replaces, for models this package's BlockFn drives, the bucketed NCCL all-reduce that
DistributedDataParallel performs for the reference (timm/task/classification.py:64-65) -- torch DDP
stays supported (train.py wraps the model itself) and is what `b200timm.train_step.wrap_ddp` builds.

    sync = PeerGradSync(model)                         # after torch.distributed.init_process_group("nccl")
    step = TrainStep(model, loss_fn, optimizer, x, y, grad_sync=sync)

Layout: every transformer block owns one contiguous slot of the flat bf16 buffer, its 14 parameters in BlockFn's workspace
order at 8-element (16-byte) aligned offsets -- so ONE kernel turns the block's fp32 gradient workspace into bf16 and
pushes it; all remaining parameters (patch embedding, tokens, final norm, head) follow ("tail").  `p.grad` of every
parameter is a persistent view into the buffer: the optimizer kernels read the reduced gradients in place.

Per step:  begin()   zero the buffer (and the tail staging buffer), peer barrier A (nobody pushes into a buffer that is
                     still being read or not yet zeroed)
           backward  BlockFn pushes each block's slot; autograd accumulates the tail gradients into a LOCAL staging buffer
           finish()  push the tail, peer barrier B (every rank's pushes have landed); with more than two ranks every rank
                     re-publishes its 1 / world slice to all (the adds land in arrival order and bf16 addition is not
                     associative: this makes the copies bit-identical, as DDP's all-reduce does), barrier C;
                     p.grad -> views of the flat buffer

Memory / streams / symmetric allocation / the barrier come from torch.distributed._symmetric_memory (plumbing, like
torch.distributed itself); the exchange is the kernel.  Needs NVLink multicast (NVSwitch) -- raises B200Error otherwise.
"""
import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import B200Error, DTYPE_BF16, DTYPE_F32

BLOCK_PARAM_ORDER = ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
                     "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias",
                     "ls1.gamma", "ls2.gamma")

_ACTIVE = None      # the PeerGradSync whose backward is running (read by autograd.BlockFn.backward)


def active_sink():
    return _ACTIVE


def _pad8(n):
    return (n + 7) // 8 * 8


def _get(module, dotted):
    obj = module
    for part in dotted.split("."):
        obj = getattr(obj, part, None)
        if obj is None:
            return None
    return obj if isinstance(obj, torch.nn.Parameter) else None


def _fused_block(m):
    """True when functional.block_forward drives module `m` through autograd.BlockFn (whose backward pushes the slot)."""
    from . import functional as BF
    try:
        if BF.is_parallel_block(m) or BF.is_glu(m.mlp) or m.attn.head_dim != 64:
            return False
        qk, sn = BF._attn_norms(m.attn)
        if qk is not None or sn is not None or not BF._is_identity(getattr(m.mlp, "norm", None)):
            return False
        BF.check_attention(m.attn)
        BF.check_mlp(m.mlp)
        BF._ln_params(m.norm1), BF._ln_params(m.norm2)
        return True
    except (B200Error, AttributeError):
        return False


def plan_layout(model):
    """(slots, tail, tail_offset, total): the flat gradient buffer's layout.  slots: id(first block parameter) -> (element
    offset, the block's 14 parameters in BlockFn's workspace order (None = absent), padded size); every BlockFn-driven
    block first, in module order, each parameter at an 8-element (16-byte) aligned offset; then the remaining trainable
    parameters ("tail") from `tail_offset`.  Pure host logic (no CUDA, no process group)."""
    params = [p for p in model.parameters() if p.requires_grad]
    slots, taken, off = {}, set(), 0
    for m in model.modules():
        ps = [_get(m, name) for name in BLOCK_PARAM_ORDER]
        if ps[2] is None or ps[4] is None or ps[8] is None or ps[10] is None or ps[0] is None or ps[6] is None:
            continue
        if any(p is not None and (id(p) in taken or not p.requires_grad) for p in ps):
            continue
        if not _fused_block(m):     # variants that do not run autograd.BlockFn keep their gradients on the tail path
            continue
        size = sum(_pad8(p.numel()) for p in ps if p is not None)
        slots[id(ps[0])] = (off, ps, size)
        taken.update(id(p) for p in ps if p is not None)
        off += size
    tail = [p for p in params if id(p) not in taken]
    return slots, tail, off, off + sum(_pad8(p.numel()) for p in tail)


def grad_push(src, dst_multicast_ptr, n, scale):
    """dst[i] += bf16(src[i] * scale) on every rank (src fp32 / bf16 CUDA tensor, n a multiple of 8)."""
    if not src.is_cuda or not src.is_contiguous() or src.dtype not in (torch.float32, torch.bfloat16):
        raise B200Error("grad_push: source must be a contiguous fp32 / bf16 CUDA tensor")
    rc = _lib.load().b200_grad_push(src.data_ptr(), DTYPE_F32 if src.dtype == torch.float32 else DTYPE_BF16,
                                    dst_multicast_ptr, n, float(scale), ops._stream())
    _lib.check(rc, "b200_grad_push")
    ops._count()


class PeerGradSync:
    def __init__(self, model, group=None):
        if not dist.is_initialized():
            raise B200Error("PeerGradSync needs an initialised torch.distributed process group")
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        params = [p for p in model.parameters() if p.requires_grad]
        if not params or any(p.dtype != torch.bfloat16 or not p.is_cuda for p in params):
            raise B200Error("PeerGradSync covers CUDA models with bf16 parameters (benchmark.py --precision bfloat16)")
        self.device = params[0].device
        self.slots, self.tail, self.tail_offset, self.total = plan_layout(model)
        off = self.tail_offset
        tail_size = self.total - off
        # ---- the symmetric flat buffer and its multicast address
        self.flat = symm.empty(self.total, dtype=torch.bfloat16, device=self.device)
        self.handle = symm.rendezvous(self.flat, self.group.group_name if hasattr(self.group, "group_name") else self.group)
        self.mc_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        if self.mc_ptr == 0:
            raise B200Error("PeerGradSync: this node offers no NVLink multicast address for symmetric memory (NVSwitch / "
                            "fabric support missing); use torch DDP (train_step.wrap_ddp)")
        self.flat.zero_()
        self.tail_local = torch.zeros(max(tail_size, 8), dtype=torch.bfloat16, device=self.device)
        # ---- persistent .grad views
        self._flat_views, self._tail_views = {}, {}
        for off0, ps, _ in self.slots.values():
            o = off0
            for p in ps:
                if p is not None:
                    self._flat_views[id(p)] = self.flat[o:o + p.numel()].view(p.shape)
                    o += _pad8(p.numel())
        o = 0
        for p in self.tail:
            self._flat_views[id(p)] = self.flat[self.tail_offset + o:self.tail_offset + o + p.numel()].view(p.shape)
            self._tail_views[id(p)] = self.tail_local[o:o + p.numel()].view(p.shape)
            o += _pad8(p.numel())
        self.block_params = [p for _, ps, _ in self.slots.values() for p in ps if p is not None]
        self._by_ptr = {ps[0].data_ptr(): slot for slot in self.slots.values() for ps in [slot[1]]}
        self._bind(flat=True)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)

    # ------------------------------------------------------------------------------------------------ step protocol
    def _bind(self, flat):
        for p in self.block_params:
            p.grad = self._flat_views[id(p)]
        for p in self.tail:
            p.grad = (self._flat_views if flat else self._tail_views)[id(p)]

    def begin(self):
        """Start of a step (replaces optimizer.zero_grad): zero the exchange buffer, barrier A, tail grads -> local staging."""
        global _ACTIVE
        self.flat.zero_()
        self.tail_local.zero_()
        self.handle.barrier(channel=0)
        self._bind(flat=False)
        _ACTIVE = self

    def slot_of(self, params):
        """(element offset, padded size) of the block slot whose parameters are exactly `params` (BlockFn's order), else None."""
        first = params[0]
        if first is None:
            return None
        slot = self._by_ptr.get(first.data_ptr())
        if slot is None:
            return None
        off, ps, size = slot
        if len(ps) != len(params) or any((a is None) != (b is None) or (a is not None and a.data_ptr() != b.data_ptr())
                                         for a, b in zip(ps, params)):
            return None
        return off, size

    def push_block(self, workspace_f32, offset, size):
        """One block's fp32 gradient workspace -> mean-reduced bf16 gradients on every rank."""
        if workspace_f32.numel() != size:
            raise B200Error("PeerGradSync: workspace layout does not match the block slot")
        grad_push(workspace_f32, self.mc_ptr + 2 * offset, size, 1.0 / self.world)

    def finish(self):
        """End of the backward: push the tail gradients, barrier B (all ranks' pushes have landed), grads -> reduced views."""
        global _ACTIVE
        _ACTIVE = None
        if self.tail:
            grad_push(self.tail_local, self.mc_ptr + 2 * self.tail_offset, self.total - self.tail_offset, 1.0 / self.world)
        self.handle.barrier(channel=1)
        if self.world > 2:
            # slice r of every copy := rank r's sum (multimem.st), so that all replicas apply the SAME gradient
            per = -(-self.total // self.world)
            per = (per + 7) // 8 * 8
            rank = dist.get_rank(self.group)
            lo = min(rank * per, self.total)
            hi = min(lo + per, self.total)
            if hi > lo:
                rc = _lib.load().b200_grad_publish(self.flat.data_ptr() + 2 * lo, self.mc_ptr + 2 * lo, 2 * (hi - lo), ops._stream())
                _lib.check(rc, "b200_grad_publish")
                ops._count()
            self.handle.barrier(channel=2)
        self._bind(flat=True)

    def abort(self):
        global _ACTIVE
        _ACTIVE = None
