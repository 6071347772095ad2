#!/usr/bin/env python
"""tools/block_trace.py — where do the blocks of K1 and K2 run when two chunks are in flight?  Per-block SM id and start/end
times (VDL2GPU_BLOCK_TRACE=1), summarised per kernel launch: SMs used, blocks per SM, start spread, duration."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["VDL2GPU_BLOCK_TRACE"] = "1"
import numpy as np
import torch
import bench
import dumpvdl2_b200 as vd

order = sys.argv[1] if len(sys.argv) > 1 else "interleaved"
chunks, offs, _ = bench.make_stream(2.0)
d = torch.from_numpy(chunks).cuda()
freqs = bench.channel_freqs(offs, 16384, order)
st = torch.cuda.current_stream()
g = vd.Vdl2Channels(bench.FS, 20, vd.FMT_U8, bench.CENTER, freqs, max_chunk_bytes=bench.CHUNK_BYTES)
for i in range(24):
    g.submit_device(d[i % d.shape[0]].data_ptr(), bench.CHUNK_BYTES, st.cuda_stream)
    g.poll_count()
g.flush_count()
t = g.block_trace()
g.close()
t0 = int(t[:, 4].min())
# group into launches: same kernel, consecutive records with block ids restarting
launches = []
for kern in (1, 2):
    r = t[t[:, 0] == kern]
    r = r[np.argsort(r[:, 4])]
    n = int(r[:, 1].max()) + 1 if len(r) else 1
    for i in range(0, len(r) - n + 1, n):
        blk = r[i:i + n]
        sm = blk[:, 2].astype(int)
        per_sm = np.bincount(sm, minlength=148)
        launches.append(dict(kernel="K1" if kern == 1 else "K2", start_ms=(int(blk[:, 4].min()) - t0) / 1e6, first_end_ms=(int(blk[:, 5].min()) - t0) / 1e6,
                             end_ms=(int(blk[:, 5].max()) - t0) / 1e6, start_spread_ms=(int(blk[:, 4].max()) - int(blk[:, 4].min())) / 1e6,
                             blocks=n, sms_used=int((per_sm > 0).sum()), max_blocks_per_sm=int(per_sm.max()),
                             mean_block_ms=float((blk[:, 5] - blk[:, 4]).mean()) / 1e6))
launches.sort(key=lambda x: x["start_ms"])
for l in launches[8:40]:
    print(json.dumps(l))
