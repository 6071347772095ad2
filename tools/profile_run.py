"""tools/profile_run.py — the bench workload (same stream, channel count and chunk size as bench.py) for a few
chunks, for use under ncu:  ncu ... python tools/profile_run.py --chunks 4 [--channels 16384] [--k1-scalar]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import dumpvdl2_b200 as vd

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=4)
ap.add_argument("--channels", type=int, default=16384)
ap.add_argument("--k1-scalar", action="store_true")
ap.add_argument("--order", default="interleaved", choices=["interleaved", "replica"])
a = ap.parse_args()
chunks, offs, _ = bench.make_stream(1.0)
freqs = bench.channel_freqs(offs, a.channels, a.order)
g = vd.Vdl2Channels(bench.FS, bench.OVERSAMPLE, vd.FMT_U8, bench.CENTER, freqs, max_chunk_bytes=bench.CHUNK_BYTES,
                    flags=(vd.FLAG_K1_SCALAR if a.k1_scalar else 0) | vd.FLAG_NO_OVERLAP)
g.enable_timing(True)
n = 0
for i in range(a.chunks):
    g.process_buf_uchar(chunks[i % chunks.shape[0]])
n = g.flush_count()
print("frames", n, "kernel ms", g.kernel_ms(), "stats", {k: v for k, v in g.stats().items() if v})
