#!/usr/bin/env python
"""tools/bench_streams.py — the mix+filter kernel against the HBM roofline where the SURVEY §8d byte model is REAL traffic:
one IQ stream per channel (the reference's traffic model: every channel thread streams its own float buffer,
src/demod.c:302-310).  16384 streams x 1 channel, streams resident in HBM (raw cu8, 8.6 GB per chunk), K0 converts
them into the time-major float2 plane (34 GB) and K1 (k1_mix_iir_decimate_lanes) reads 8 B per channel-sample from it.

    python tools/bench_streams.py [--streams 16384] [--chunks 6] > gpurun_out/streams.json
    ncu --set full -k regex:k1_mix_iir_decimate_lanes -s 1 -c 1 ... python tools/bench_streams.py --chunks 3   # dram__bytes
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import dumpvdl2_b200 as vd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=16384)
    ap.add_argument("--chunks", type=int, default=6)
    ap.add_argument("--chunk-pairs", type=int, default=bench.CHUNK_PAIRS)
    a = ap.parse_args()
    S, P = a.streams, a.chunk_pairs
    nbytes = 2 * P
    chunks, offs, _ = bench.make_stream(2.0)
    base = torch.from_numpy(chunks.reshape(-1)).cuda()
    L = base.numel() // 2
    base2 = base.view(L, 2)
    raw = torch.empty(S, P, 2, dtype=torch.uint8, device="cuda")
    ar = torch.arange(P, device="cuda", dtype=torch.int64)
    for s0 in range(0, S, 64):                      # stream s = the base stream started 7919 s samples later (wrapping)
        sh = (torch.arange(s0, min(s0 + 64, S), device="cuda", dtype=torch.int64) * 7919) % L
        idx = (ar[None, :] + sh[:, None]) % L
        raw[s0:s0 + idx.shape[0]] = base2[idx]
    del idx
    torch.cuda.synchronize()
    freqs = bench.channel_freqs(offs, S, "interleaved")
    stream = torch.cuda.current_stream()
    res = dict(streams=S, chunk_pairs=P)
    for overlap in (False, True):
        g = vd.Vdl2Channels(bench.FS, 20, vd.FMT_U8, bench.CENTER, freqs, max_chunk_bytes=nbytes, n_streams=S,
                            flags=0 if overlap else vd.FLAG_NO_OVERLAP)
        for _ in range(2):
            g.submit_device(raw.data_ptr(), nbytes, stream.cuda_stream)
        g.flush_count()
        if not overlap:
            g.enable_timing(True)
        k0 = g.kernel_ms()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        frames = 0
        for _ in range(a.chunks):
            g.submit_device(raw.data_ptr(), nbytes, stream.cuda_stream)
            frames += g.poll_count()
        frames += g.flush_count()
        g.stream_wait(stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.chunks
        if not overlap:
            k1 = g.kernel_ms()
            kms = {k: (k1[k][0] - k0[k][0]) / max(k1[k][1] - k0[k][1], 1) for k in k1}
            res["kernel_ms"] = kms
            res["serial_ms_per_chunk"] = ms
        else:
            res["pipelined_ms_per_chunk"] = ms
            res["pipelined_m_chsamples_per_s"] = S * P / ms / 1e3
            res["frames_per_chunk"] = frames / a.chunks
        g.close()
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    k1ms = res["kernel_ms"]["K1"]
    alg = S * P * bench.B_CS
    res["roofline"] = dict(bound="hbm", kernel="k1_mix_iir_decimate_lanes", algorithmic_bytes_per_launch=alg, ms_per_launch=k1ms,
                           achieved=alg / (k1ms * 1e-3) / 1e9, peak=peak, unit="GB/s", frac=alg / (k1ms * 1e-3) / 1e9 / peak,
                           note="one stream per channel: the 8 B read + 0.4 B written per channel-sample of SURVEY §8d are actual DRAM traffic "
                                "(dram__bytes from ncu are committed beside this file)")
    k0ms = res["kernel_ms"]["K0"]
    res["k0_convert"] = dict(ms_per_launch=k0ms, bytes=S * P * (2 + 8), gbs=S * P * 10 / (k0ms * 1e-3) / 1e9)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
