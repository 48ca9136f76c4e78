"""N>1 host logic on CPU: world_size-2 gloo run of the block sharding / in-order merge
used for multi-GPU.  The per-rank engine is the CPU oracle here; on
the GPU box the same plumbing drives one gzpb context per rank."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from gzp_b200 import shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = synth.text(65280 * 11 + 99)
    nblocks = (len(data) + 65279) // 65280
    batch = 3
    mine = []
    for first, cnt in shard.my_ranges(nblocks, batch, world, rank):
        enc = b"".join(oracle.encode_block(oracle.BGZF, 6, data[(first + i) * 65280:(first + i + 1) * 65280], None, first + i == nblocks - 1)
                       for i in range(cnt))
        mine.append(enc)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    dist.barrier()
    if rank == 0:
        merged = shard.merge_in_order(nblocks, batch, world, gathered)
        q.put(merged == oracle.compress_stream(oracle.BGZF, 6, 65280, [data]))
    dist.destroy_process_group()


def test_two_rank_sharding_reassembles_the_stream():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
