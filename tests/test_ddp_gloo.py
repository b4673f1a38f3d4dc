"""The N > 1 path on CPU: world_size-2 `gloo` runs of the host-side data-parallel logic.

The hot path shards only along the batch: inference = independent replicas, training = torch DDP with
a gradient all-reduce.  The CUDA kernels cannot run here, so a small torch module stands in for the model; what is
tested is the harness the GPU runs use: rank / seed handling, DDP wrapping as timm/task/classification.py:64-65 does,
gradient averaging, replicas staying bit-identical after optimizer steps, max-over-ranks timing reduction, and
that `bench.py --impl reference` prints exactly one JSON line (rank 0) under a 2-rank launch."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(42)                       # identical init on every rank (DDP also broadcasts rank 0's)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 4))
    net = torch.nn.parallel.DistributedDataParallel(model, broadcast_buffers=True)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    torch.manual_seed(42 + rank)                # per-rank data shard, as train.py:498 seeds (seed + rank)
    x, y = torch.randn(8, 16), torch.randint(0, 4, (8,))
    # reference: the gradient of the mean loss over the GLOBAL batch, computed without DDP
    xs = [torch.empty_like(x) for _ in range(world)]
    ys = [torch.empty_like(y) for _ in range(world)]
    dist.all_gather(xs, x)
    dist.all_gather(ys, y)
    ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 4))
    ref.load_state_dict(model.state_dict())
    torch.nn.functional.cross_entropy(ref(torch.cat(xs)), torch.cat(ys)).backward()
    opt.zero_grad()
    torch.nn.functional.cross_entropy(net(x), y).backward()
    ok_grad = all(torch.allclose(p.grad, q.grad, atol=1e-6) for p, q in zip(model.parameters(), ref.parameters()))
    for _ in range(3):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(net(x), y).backward()
        opt.step()
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    in_sync = all(torch.equal(gathered[0], g) for g in gathered)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)      # max-over-ranks timing reduction used by bench.py
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        with open(out, "w") as f:
            json.dump({"ok_grad": ok_grad, "in_sync": in_sync, "tmax": t.item()}, f)
    dist.destroy_process_group()


def test_ddp_gradients_average_and_replicas_stay_in_sync(tmp_path):
    out = str(tmp_path / "r.json")
    mp.spawn(_worker, args=(2, 29541, out), nprocs=2, join=True)
    with open(out) as f:
        r = json.load(f)
    assert r == {"ok_grad": True, "in_sync": True, "tmax": 11.0}


def test_reference_arm_prints_once_under_two_ranks():
    """bench.py --impl reference under a 2-rank launch: rank 0 alone runs and prints; rank 1 exits 0 silently."""
    outs = []
    for rank in (0, 1):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT="29542", B200_REF_TINY="1")
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                            "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
        assert p.returncode == 0, p.stderr[-500:]
        outs.append([l for l in p.stdout.splitlines() if l.startswith("{")])
    assert len(outs[0]) == 1 and outs[1] == []
    line = json.loads(outs[0][0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["unit"] == "img/s"
    # oracle/_ref installed (this container, the GPU box): the unmodified benchmark.py is the reference arm
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "infer" in line or line["cpu_baseline"]["kind"] == "port"


def test_peer_grad_sync_layout_covers_every_parameter_once():
    """b200timm.grad_sync.plan_layout on the product model (host logic, no GPU): every transformer block that BlockFn drives
    owns one slot whose parameters sit at 16-byte aligned offsets in BlockFn's workspace order (so that ONE kernel converts
    and pushes a block's gradient workspace), every other trainable parameter is in the tail, nothing is counted twice;
    variant blocks (qk-norm: not BlockFn) fall back to the tail."""
    sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
    from b200timm.grad_sync import BLOCK_PARAM_ORDER, plan_layout
    from b200timm.vision_transformer import VisionTransformer
    model = VisionTransformer(img_size=32, patch_size=16, embed_dim=128, depth=3, num_heads=2, num_classes=10, init_values=0.1)
    slots, tail, tail_off, total = plan_layout(model)
    assert len(slots) == 3
    seen, expect = set(), 0
    for off, ps, size in sorted(slots.values(), key=lambda s: s[0]):
        assert off == expect and size % 8 == 0 and len(ps) == len(BLOCK_PARAM_ORDER)
        o = off
        for p in ps:
            assert p is not None and o % 8 == 0 and id(p) not in seen
            seen.add(id(p))
            o += (p.numel() + 7) // 8 * 8
        assert o == off + size
        expect = off + size
    assert tail_off == expect and not (seen & {id(p) for p in tail})
    assert seen | {id(p) for p in tail} == {id(p) for p in model.parameters()}
    assert total == tail_off + sum((p.numel() + 7) // 8 * 8 for p in tail)
    qk = VisionTransformer(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_classes=10, qk_norm=True)
    slots, tail, _, _ = plan_layout(qk)
    assert not slots and len(tail) == len(list(qk.parameters()))
