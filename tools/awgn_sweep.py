#!/usr/bin/env python
"""BASELINE.json config 4 at full size: 4096 channels, AWGN sweep Eb/N0 5..15 dB.

4096 channels = 64 frequency slots (25 kHz spacing) x 64 replicas fanned out from ONE 2.1 Msps cu8 stream
(channel = slot * 64 + replica).  Per SNR point: >= 200 injected bursts, seed 0x56444C34 + Eb/N0, Es/N0 = Eb/N0 +
10 log10(3), full-band noise power Pn = Ps (fs/10500) / (Es/N0).  Reported per point: frame-decode rate
(FCS-good frames whose octets equal an injected frame / injected frames) of
  * the GPU path (libvdl2gpu.so, all 4096 channels; the 64 replicas of a slot must agree with each other),
  * the strict oracle (oracle/liboracle.so, one channel per slot),
  * the reference built -O2 -ffast-math as shipped (oracle/_ref/vdl2_ref_fast) and -O2 (vdl2_ref_strict),
and the symmetric difference of the (slot, frame octets) sets GPU vs each of them.

Usage (GPU box):  python tools/awgn_sweep.py --out gpurun_out/awgn_sweep.json
       (no GPU):  python tools/awgn_sweep.py --no-gpu --points 9,12      # oracle / reference legs only
"""
import argparse
import json
import math
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dumpvdl2_b200 import synth           # noqa: E402
from oracle import pyoracle as po         # noqa: E402

FS, CENTER, CHUNK = 2100000, 136975000, 524288


def frame_set(frames, chan_of):
    """{(slot, octets)} of the FCS-good frames + the multiset size (duplicates would be a bug worth seeing)."""
    s, n = set(), 0
    for f in frames:
        data, ch = (f["data"], f["channel"]) if isinstance(f, dict) else (f.data, f.channel)
        if po.crc16(data) != 0xF0B8 or len(data) < 3:
            continue
        s.add((chan_of(ch), data))
        n += 1
    return s, n


def cpu_legs(job):
    """One SNR point: synthesise the stream (kept in /tmp for the GPU leg), run the oracle and the reference builds."""
    ebn0, a = job
    es_n0 = ebn0 + 10.0 * math.log10(3.0)
    seed = 0x56444C34 + ebn0
    t0 = time.time()
    iq, offs, bursts = synth.traffic_stream(FS, a.seconds, a.slots, a.bursts_per_s, es_n0, -20.0, seed, "u8")
    slot_freqs = [CENTER + int(o) for o in offs]
    slot_of_off = {int(o): k for k, o in enumerate(offs)}
    injected = {(slot_of_off[int(b.offset_hz)], fr) for b in bursts for fr in b.frames}
    row = dict(eb_n0_db=ebn0, es_n0_db=round(es_n0, 3), seed=seed, bursts=len(bursts), injected_frames=len(injected),
               stream_s=a.seconds, iq_pairs=int(iq.size // 2))
    path = os.path.join(tempfile.gettempdir(), f"awgn_sweep_{os.getpid()}_{ebn0}.cu8")
    iq.tofile(path)
    o = po.Oracle(FS, 20, po.FMT_U8, CENTER, slot_freqs)          # strict oracle, one channel per slot
    o.process_chunked(iq, CHUNK)
    sets = {"oracle_strict": frame_set(o.frames(), lambda ch: ch)}
    o.close()
    for flavour in ("fast", "strict"):                            # the reference itself, both builds
        if po.ref_binary(flavour) is not None:
            fr, _ = po.run_ref(path, po.FMT_U8, 20, CENTER, slot_freqs, flavour=flavour, chunk=CHUNK)
            sets["reference_" + flavour] = frame_set(fr, lambda ch: ch)
    row["cpu_wall_s"] = round(time.time() - t0, 1)
    return row, sets, injected, slot_freqs, path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", default="5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20",
                    help="Eb/N0 values in dB (config 4 names 5..15; 16..20 show the top of the curve)")
    ap.add_argument("--slots", type=int, default=64)
    ap.add_argument("--replicas", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--bursts-per-s", type=float, default=1.4)
    ap.add_argument("--workers", type=int, default=max(1, min(16, (os.cpu_count() or 2) // 4)))
    ap.add_argument("--no-gpu", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    points = [int(x) for x in a.points.split(",")]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(a.workers) as pool:           # CPU legs first: CUDA is not initialised before the fork
        done = pool.map(cpu_legs, [(p, a) for p in points], chunksize=1)
    if not a.no_gpu:
        import dumpvdl2_b200 as vd
    rows = []
    for row, sets, injected, slot_freqs, path in done:
        if not a.no_gpu:
            iq = np.fromfile(path, np.uint8)
            freqs = [f for f in slot_freqs for _ in range(a.replicas)]
            g = vd.Vdl2Channels(FS, 20, vd.FMT_U8, CENTER, freqs, max_chunk_bytes=CHUNK)
            g.process_chunked(iq, CHUNK)
            fr = g.flush()
            st = g.stats()
            per_rep = [[] for _ in range(a.replicas)]
            for f in fr:
                per_rep[f.channel % a.replicas].append(f)
            per_rep = [frame_set(l, lambda ch: ch // a.replicas) for l in per_rep]
            row["gpu_replicas_identical"] = all(p == per_rep[0] for p in per_rep)
            row["gpu_channels"] = len(freqs)
            row["gpu_overflows"] = int(st["pool_overflows"]) + int(st["out_overflows"])
            sets["gpu"] = per_rep[0]
            g.close()
        os.unlink(path)
        for name, (s, n) in sets.items():
            row[name] = dict(fcs_good_frames=n, decoded_injected=len(s & injected),
                             decode_rate=round(len(s & injected) / len(injected), 4), not_injected=len(s - injected))
        if "gpu" in sets:
            for name in sets:
                if name != "gpu":
                    row["symdiff_gpu_vs_" + name] = len(sets["gpu"][0] ^ sets[name][0])
        rows.append(row)
        print(json.dumps(row), file=sys.stderr, flush=True)
    rep = dict(config="BASELINE.json configs[3]: synthetic 2.1 Msps IQ, 4096 channels, AWGN sweep Eb/N0 5-15 dB",
               layout=f"{a.slots} slots x {a.replicas} replicas, one cu8 stream, fs {FS}, oversample 20", points=rows)
    txt = json.dumps(rep, indent=1)
    if a.out:
        with open(a.out, "w") as f:
            f.write(txt + "\n")
    print("Eb/N0  bursts frames |  " + "  ".join(f"{k:>16s}" for k in ("gpu", "oracle_strict", "reference_fast", "reference_strict"))
          + " | symdiff gpu vs strict/fast/refstrict")
    for r in rows:
        cols = "  ".join(f"{r[k]['decode_rate']:16.4f}" if k in r else " " * 16
                         for k in ("gpu", "oracle_strict", "reference_fast", "reference_strict"))
        sd = "/".join(str(r.get("symdiff_gpu_vs_" + k, "-")) for k in ("oracle_strict", "reference_fast", "reference_strict"))
        print(f"{r['eb_n0_db']:5d} {r['bursts']:7d} {r['injected_frames']:6d} |  {cols} | {sd}")


if __name__ == "__main__":
    main()
