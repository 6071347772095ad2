"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`): the N>1 path on hardware.  Two processes,
one per GPU, channels sharded k mod N, every chunk fanned out from rank 0 by the library's own multi-GPU ingest
helper (vdl2gpu_mg_*) in both of its modes; EVERY rank's frames, metadata and counters are compared with the oracle
run on that rank's channel shard."""
import os
import socket
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, mode, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        import dumpvdl2_b200 as vd
        from dumpvdl2_b200 import shard
        from oracle import pyoracle as po
        from tests import cases, util
        c = cases.case_replicas(n_slots=16, n_rep=4, duration=1.0)
        chunk, per_step = 131072, 4
        data = util.case_bytes(c)
        n_steps = data.size // (chunk * per_step)
        data = data[:n_steps * per_step * chunk]
        mine = shard.my_channels(c["freqs"], rank, world)
        g = vd.Vdl2Channels(c["fs"], c["oversample"], util.fmt_code(c), c["centerfreq"], mine, max_chunk_bytes=chunk, device=rank)
        mg = shard.MultiGpuIngest(g, rank, world, chunk * per_step, mode=mode)
        src = torch.from_numpy(data.copy()).cuda() if rank == 0 else None
        host_src = torch.from_numpy(data.copy()).pin_memory() if rank == 0 else None

        def runs(step, host):
            if rank != 0:
                return None
            t = host_src if host else src
            return [(t.data_ptr() + step * per_step * chunk, per_step * chunk)]
        frames = []
        half = mg.stage(runs(0, False), False)
        for s in range(n_steps):
            nxt = mg.stage(runs(s + 1, s % 2 == 0), s % 2 == 0) if s + 1 < n_steps else None     # alternate device / pinned-host sources
            mg.submit_staged(half, per_step, chunk)
            frames += g.poll()
            half = nxt
        frames += g.flush()
        o = po.Oracle(c["fs"], c["oversample"], util.fmt_code(c), c["centerfreq"], mine)
        o.process_chunked(data, chunk)
        util.assert_frames_equal(frames, o.frames(), f"rank {rank} ({mg.mode_name})")
        assert np.array_equal(g.channel_counters(), o.counters())
        n = len(frames)
        st = g.stats()
        assert st["pool_overflows"] == 0 and st["out_overflows"] == 0 and st["graph_launches"] > 0
        dist.barrier()
        mg.close(); g.close()
        dist.destroy_process_group()
        q.put((rank, "ok", n, mg.mode_name))
    except Exception as e:           # noqa: BLE001 - reported to the parent
        import traceback
        q.put((rank, "fail", traceback.format_exc(), str(e)))


@pytest.mark.parametrize("mode", ["nccl", "ce"])
def test_two_gpu_shards_match_the_oracle_on_every_rank(mode):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", f"rank {r[0]}: {r[2]}"
    assert sum(r[2] for r in res) > 20
