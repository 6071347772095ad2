"""CPU, world_size 2, gloo: the N>1 host logic - round-robin channel shards, broadcast of every raw IQ chunk
from the ingest rank, independent per-rank results whose union equals the single-process result.  (On GPUs the
same code runs over NCCL; the per-rank demodulator here is the oracle, standing in for libvdl2gpu.)"""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from dumpvdl2_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    from tests import cases, util
    c = cases.case_cfg2(0.5)
    mine = shard.shard_freqs(c["freqs"], rank, world)
    assert list(mine) == [c["freqs"][i] for i in shard.channels_for_rank(len(c["freqs"]), rank, world)]
    o = po.Oracle(c["fs"], c["oversample"], util.fmt_code(c), c["centerfreq"], mine)
    data = util.case_bytes(c)
    n_chunks = -(-data.size // c["chunk"])
    for i in range(n_chunks):
        if rank == 0:                                  # only the ingest rank has the stream
            part = data[i * c["chunk"]:(i + 1) * c["chunk"]]
            n = torch.tensor([part.size])
        else:
            n = torch.tensor([0])
        dist.broadcast(n, src=0)
        buf = torch.from_numpy(part.copy()) if rank == 0 else torch.empty(int(n.item()), dtype=torch.uint8)
        shard.broadcast_chunk(buf, src=0)
        o.process(buf.numpy())
    keys = [(shard.global_channel(f.channel, rank, world), f.burst_seq, f.idx, f.data.hex()) for f in o.frames()]
    allk = shard.gather_frame_keys(keys, dst=0)
    if rank == 0:
        q.put(allk)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_channel_shards_reproduce_single_process():
    from tests import cases, util
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = cases.case_cfg2(0.5)
    want = sorted((f.channel, f.burst_seq, f.idx, f.data.hex()) for f in util.run_oracle(c).frames())
    assert got == want and len(want) >= 4


def test_round_robin_mapping():
    for world in (1, 2, 4, 8):
        seen = sorted(g for r in range(world) for g in shard.channels_for_rank(37, r, world))
        assert seen == list(range(37))
        for r in range(world):
            for li, g in enumerate(shard.channels_for_rank(37, r, world)):
                assert shard.global_channel(li, r, world) == g
