"""CPU test (world_size 2, gloo) of the multi-GPU host logic bench.py uses: images shard by rank, every rank runs the
same per-image path on its own shard, one all_gather of padded descriptors / LAFs / counts at the end, and the
max-over-ranks timing reduction.  The per-image compute is stubbed by the oracle's level-selection (cheap)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard(rank, world, B):
    """bench.py convention: rank r owns global images [r*B, (r+1)*B) (seeds 1234 + r*B + i)."""
    return list(range(rank * B, (rank + 1) * B))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, K = 3, 8
    ids = shard(rank, world, B)
    g = torch.Generator().manual_seed(0)
    all_desc = torch.rand(world * B, K, 128, generator=g)           # what a single process would produce
    all_cnt = torch.randint(1, K + 1, (world * B,), generator=g, dtype=torch.int32)
    desc, cnt = all_desc[ids].clone(), all_cnt[ids].clone()          # this rank's shard
    for b in range(B):
        desc[b, cnt[b]:] = float("nan")                               # rows >= count are unspecified
    gd = torch.empty(world * B, K, 128); gc = torch.empty(world * B, dtype=torch.int32)
    dist.all_gather_into_tensor(gd, desc); dist.all_gather_into_tensor(gc, cnt)
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = torch.equal(gc, all_cnt) and float(t) == 10.0 + world - 1
    for i in range(world * B):
        ok = ok and torch.equal(gd[i, :gc[i]], all_desc[i, :gc[i]])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_gather_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert res == [(0, True), (1, True)]


def _exchange_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from affnet_b200.exchange import DescriptorExchange
    B, K = 3, 5
    x = DescriptorExchange(world, B, K, torch.device("cpu"))
    outs = x.outputs()                                                # per slot: the (lafs, desc, count) views a producer writes directly

    def step_data(r, step):
        g = torch.Generator().manual_seed(100 * step + r)
        return torch.rand(B, K, 2, 3, generator=g), torch.rand(B, K, 128, generator=g), torch.randint(1, K + 1, (B,), generator=g, dtype=torch.int32)

    for step in range(5):                                             # more steps than blocks: blocks are reused
        slot = step & 1
        x.wait_slot(slot)                                             # the gather that last read this block has finished
        l, d, c = step_data(rank, step)
        outs[slot][0].copy_(l); outs[slot][1].copy_(d); outs[slot][2].copy_(c)     # "the kernels write the block"
        x.submit(slot)
    x.drain()
    gd, gl, gc = x.last()
    ok = True
    for r in range(world):
        l, d, c = step_data(r, 4)
        ok = ok and torch.equal(gd[r * B:(r + 1) * B], d) and torch.equal(gl[r * B:(r + 1) * B], l) and torch.equal(gc[r * B:(r + 1) * B], c)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_pipelined_descriptor_exchange():
    """affnet_b200.exchange.DescriptorExchange (what bench.py runs per step at N > 1) on two gloo ranks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 500)
    ps = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert res == [(0, True), (1, True)]


def test_shards_are_disjoint_and_cover():
    seen = []
    for r in range(8):
        seen += shard(r, 8, 64)
    assert seen == list(range(512))                                   # BASELINE config 4: 512 images over 8 GPUs
