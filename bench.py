#!/usr/bin/env python
"""bench.py — headline benchmark of the dumpvdl2 per-channel DSP hot path on B200.

Metric (BASELINE.json): VDL2 channels demodulated in real time @ 2.1 Msps, reported as whole-job
M channel-samples/s (one channel-sample = one complex input sample processed for one channel);
channels_at_realtime = value / 2.1.

Workload (BASELINE.json config 5 at one GPU, the largest configuration that fits a single B200):
16384 channels per GPU = 64 slots x 256 replicas fanned out from ONE synthetic 2.1 Msps cu8 stream
(Poisson bursts 2/s/slot, Es/N0 20 dB).  One "step" = one pass of the whole path
(K0 convert, K1 mix+IIR+decimate, K2 sync/slice/header, K3 FEC/unstuff/FCS, frames back on the host) over
`chunks_per_step` chunks of 262144 IQ pairs (0.125 s of signal each) for every channel.
With N GPUs the channel count scales with N (weak scaling): global channel k lives on GPU k mod N, rank 0
owns the stream and every chunk is broadcast with NCCL before each rank demodulates its own shard.

  value : chunks resident in HBM before the timed region (N>1: resident on rank 0, NCCL broadcast timed).
  e2e   : same metric through the public host-buffer API (Vdl2Channels.process_buf_uchar): host->device copy
          of every chunk and device->host frame records inside the timed region.
  --impl reference : the unmodified reference (oracle/_ref/vdl2_ref_fast, one pthread per channel, all host
          threads) on a bounded sample of the same stream; falls back to the oracle port if _ref is absent.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 2100000
OVERSAMPLE = 20
CENTER = 136975000
CHUNK_PAIRS = 262144
CHUNK_BYTES = 2 * CHUNK_PAIRS
B_CS = 8.0 + 8.0 / OVERSAMPLE        # algorithmic bytes per channel-sample for K1: 8 B read + 8/os B written


def make_stream(seconds):
    from dumpvdl2_b200 import synth
    iq, offs, bursts = synth.traffic_stream(FS, seconds, 64, 2.0, 20.0, -20.0, 0x56444C33, "u8")
    n_chunks = iq.size // CHUNK_BYTES
    return iq[:n_chunks * CHUNK_BYTES].reshape(n_chunks, CHUNK_BYTES), offs, bursts


def channel_freqs(offs, n_total):
    reps = -(-n_total // len(offs))
    return np.array([CENTER + int(o) for o in offs for _ in range(reps)][:n_total], dtype=np.uint32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2] or [r for (_, r) in self.rows]
        if not rows:
            return None
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 5 + k and r[5 + k].lower().startswith("active") for r in rows)]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=float(rows[0][2]) if rows[0][2].replace(".", "").isdigit() else None,
                    reasons=reasons, samples=len(rows))


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on this host's cores."""
    if rank != 0:
        return
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    n_ch = max(1, min(cores - 1 if cores > 1 else 1, 256))
    chunks, offs, _ = make_stream(1.0)
    freqs = channel_freqs(offs, n_ch)
    exe = po.ref_binary("fast")
    sample_pairs = chunks.size // 2
    with tempfile.NamedTemporaryFile(suffix=".cu8", delete=False) as tf:
        tf.write(chunks.tobytes())
        path = tf.name
    per = []
    kind = "reference" if exe else "port"
    try:
        for step in range(args.warmup + args.steps):
            if exe:
                _, st = po.run_ref(path, po.FMT_U8, OVERSAMPLE, CENTER, freqs, flavour="fast", chunk=CHUNK_BYTES, quiet=True)
                dt = st["wall_s"]
            else:
                n_ch = 4
                o = po.Oracle(FS, OVERSAMPLE, po.FMT_U8, CENTER, freqs[:n_ch])
                t0 = time.time(); o.process_chunked(chunks.reshape(-1), CHUNK_BYTES); dt = time.time() - t0
            if step >= args.warmup:
                per.append(dt)
    finally:
        os.unlink(path)
    t = float(np.mean(per))
    value = n_ch * sample_pairs / t / 1e6
    sample = f"{n_ch} channels x {sample_pairs / FS:.2f} s of the bench stream per step ({'one pthread per channel + producer' if exe else 'single-thread oracle port'})"
    line = dict(impl="reference", metric="VDL2 channel-samples demodulated per second (2.1 Msps channels; real-time channels = value/2.1)",
                value=value, unit="Msamples/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=t * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=f"reference CPU path on a bounded sample: {sample}", fs=FS, oversample=OVERSAMPLE,
                            channels=n_ch, sample_fmt="cu8", l2="n/a (CPU)"),
                channels_at_realtime=value / 2.1,
                cpu_baseline=dict(value=value, unit="Msamples/s", cores=(n_ch + 1 if exe else 1), kind=kind, sample=sample,
                                  host_cpus=cores),
                e2e=dict(value=value, unit="Msamples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def cpu_baseline_leg():
    """Bounded (~10-30 s) timing of the reference CPU path on this host, reported beside the GPU numbers."""
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    exe = po.ref_binary("fast")
    chunks, offs, _ = make_stream(1.0)
    if exe:
        n_ch = max(1, min(cores - 1 if cores > 1 else 1, 256))
        freqs = channel_freqs(offs, n_ch)
        with tempfile.NamedTemporaryFile(suffix=".cu8", delete=False) as tf:
            tf.write(chunks.tobytes()); path = tf.name
        try:
            loops = 4
            _, st = po.run_ref(path, po.FMT_U8, OVERSAMPLE, CENTER, freqs, flavour="fast", chunk=CHUNK_BYTES, quiet=True, loop=loops)
        finally:
            os.unlink(path)
        v = st["ch_msamples_per_s"]
        return dict(value=v, unit="Msamples/s", cores=n_ch + 1, kind="reference", host_cpus=cores,
                    sample=f"{n_ch} channels (one pthread each + 1 producer) x {loops} x {chunks.size / 2 / FS:.2f} s of the bench stream, "
                           f"unmodified reference built -O2 -ffast-math", channels_at_realtime=v / 2.1)
    n_ch = 4
    o = po.Oracle(FS, OVERSAMPLE, po.FMT_U8, CENTER, channel_freqs(offs, n_ch))
    t0 = time.time(); o.process_chunked(chunks.reshape(-1), CHUNK_BYTES); dt = time.time() - t0
    v = n_ch * chunks.size / 2 / dt / 1e6
    return dict(value=v, unit="Msamples/s", cores=1, kind="port", host_cpus=cores,
                sample=f"{n_ch} channels x {chunks.size / 2 / FS:.2f} s, single-thread oracle port", channels_at_realtime=v / 2.1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--channels", type=int, default=16384, help="channels per GPU")
    ap.add_argument("--chunks-per-step", type=int, default=8)
    ap.add_argument("--stream-seconds", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--k1-scalar", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # keep stdout clean for the single JSON line: libraries (NCCL prints its version banner there) write to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import dumpvdl2_b200 as vd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the demodulator path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W, K, CPS = max(args.warmup, 3), args.steps, args.chunks_per_step
    n_total = args.channels * world
    chunks, offs, bursts = make_stream(args.stream_seconds)       # same seed on every rank
    n_chunks = chunks.shape[0]
    all_freqs = channel_freqs(offs, n_total)
    my_freqs = all_freqs[rank::world]                             # channel k -> GPU k mod N
    flags = vd.FLAG_K1_SCALAR if args.k1_scalar else 0
    g = vd.Vdl2Channels(FS, OVERSAMPLE, vd.FMT_U8, CENTER, my_freqs, max_chunk_bytes=CHUNK_BYTES, device=local_rank, flags=flags)
    stream = torch.cuda.current_stream()

    # chunks resident in HBM on the ingest rank (rank 0).  With N > 1 every step's CPS chunks are broadcast in ONE
    # NCCL call (bucket sized for launch latency: 4 MiB instead of 8 x 0.5 MiB) into a double-buffered receive
    # area, then each rank demodulates its channel shard straight out of that buffer.
    if rank == 0:
        d_chunks = torch.from_numpy(chunks).cuda()
    h_chunks = torch.from_numpy(chunks).pin_memory()
    d_recv = [torch.empty(CPS * CHUNK_BYTES, dtype=torch.uint8, device="cuda") for _ in range(2)] if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = dict(chunk=0, bcast=0)

    def next_indices():
        idx = [(state["chunk"] + k) % n_chunks for k in range(CPS)]
        state["chunk"] += CPS
        return idx

    def gather_rows(src, idx, dst):
        """rows idx of src -> dst (flat), contiguous runs copied in one go"""
        k = 0
        while k < len(idx):
            j = k
            while j + 1 < len(idx) and idx[j + 1] == idx[j] + 1:
                j += 1
            dst[k * CHUNK_BYTES:(j + 1) * CHUNK_BYTES].copy_(src[idx[k]:idx[j] + 1].reshape(-1), non_blocking=True)
            k = j + 1

    # N > 1: the broadcast of step s+1 is issued BEFORE the chunks of step s are submitted (the receive area is double
    # buffered), so no rank ever waits for the ingest rank at a step boundary: the collective overlaps the kernels.
    def stage_bcast(src_rows):
        """rank 0 gathers the next step's chunks (device- or pinned-host-resident) into the free half of the receive
        area and everybody joins the broadcast; returns the buffer"""
        idx = next_indices()
        buf = d_recv[state["bcast"] % 2]
        state["bcast"] += 1
        g.wait_input_consumed(stream.cuda_stream)              # every K0 that read this half has run (2 steps ago)
        if rank == 0:
            gather_rows(src_rows, idx, buf)
        dist.broadcast(buf, src=0)
        return buf

    def step_multi(src_rows):
        if state.get("staged") is None or state.get("staged_src") is not src_rows:
            state["staged"] = stage_bcast(src_rows)            # first step of a pass
            state["staged_src"] = src_rows
        buf = state["staged"]
        nxt = stage_bcast(src_rows)                             # next step's chunks travel while this step computes
        for k in range(CPS):
            g.submit_device(buf.data_ptr() + k * CHUNK_BYTES, CHUNK_BYTES, stream.cuda_stream)
        state["staged"] = nxt
        return g.poll_count()

    def step_device():
        """one step, chunks already in HBM (N>1: one NCCL broadcast per step from rank 0, one step ahead)"""
        if world > 1:
            return step_multi(d_chunks if rank == 0 else None)
        for i in next_indices():
            g.submit_device(d_chunks[i].data_ptr(), CHUNK_BYTES, stream.cuda_stream)
        return g.poll_count()

    def step_host():
        """one step through the public host-buffer entry point (process_buf_uchar); with N > 1 the ingest rank copies
        the step's chunks host->device and broadcasts them, the other ranks receive"""
        if world > 1:
            return step_multi(h_chunks)
        for i in next_indices():
            g.process_buf_uchar(h_chunks[i].numpy())
        return g.poll_count()

    def timed(step_fn, steps, sample_clocks=False):
        # one sampler for the job (rank 0's GPU): eight concurrent nvidia-smi pollers contend for the driver lock.
        # It is started BEFORE the barrier so that no rank enters the timed region late.
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        if sample_clocks:
            time.sleep(0.3)
        barrier()
        s0 = g.stats()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        frames = 0
        for _ in range(steps):
            frames += step_fn()          # submits + non-blocking harvest: the pipeline stays full across steps
        frames += g.flush_count()       # everything submitted is processed and its frames are on the host
        g.stream_wait(stream.cuda_stream)
        e1.record(stream)
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        clocks = sampler.stop(t0, t1) if sampler else None
        s1 = g.stats()
        return ms, frames, {k: s1[k] - s0[k] for k in s1}, clocks, (t1 - t0) * 1e3

    for _ in range(W):
        step_device()
    g.flush_count()
    ms_dev, frames_dev, d_dev, clocks, wall_dev = timed(step_device, K, sample_clocks=True)
    for _ in range(2):
        step_host()
    g.flush_count()
    ms_e2e, frames_e2e, d_e2e, _, wall_e2e = timed(step_host, K)

    # Per-kernel durations for the roofline.  The production pipeline runs K0/K1 of chunk c+1 beside K2/K3 of chunk c
    # on two streams, which stretches every kernel's wall time; the kernel's OWN launch duration is therefore measured
    # here, live, with CUDA events on the library's stream, same workload, same process, with that overlap switched off
    # (VDL2GPU_FLAG_NO_OVERLAP) - the condition the committed ncu launch list is taken under as well.
    g_main = g
    g = vd.Vdl2Channels(FS, OVERSAMPLE, vd.FMT_U8, CENTER, my_freqs, max_chunk_bytes=CHUNK_BYTES, device=local_rank,
                        flags=flags | vd.FLAG_NO_OVERLAP)
    for _ in range(2):
        step_device()
    g.flush_count()
    g.enable_timing(True)
    k0 = g.kernel_ms()
    ms_serial, _, _, _, _ = timed(step_device, max(2, K // 2))
    k1 = g.kernel_ms()
    g.close()
    g = g_main

    pairs_per_step = CPS * CHUNK_PAIRS
    cs_per_step = float(n_total) * pairs_per_step
    value = cs_per_step * K / (ms_dev * 1e-3) / 1e6
    e2e_value = cs_per_step * K / (ms_e2e * 1e-3) / 1e6
    kms = {k: (k1[k][0] - k0[k][0], k1[k][1] - k0[k][1]) for k in k1}
    k1_ms_per_launch = kms["K1"][0] / max(kms["K1"][1], 1)
    achieved = len(my_freqs) * CHUNK_PAIRS * B_CS / (k1_ms_per_launch * 1e-3) / 1e9 if k1_ms_per_launch > 0 else None
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        pass
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "k1_ncu_traffic.json")) as f:
            tj = json.load(f)
            if tj.get("channels") == len(my_freqs) and tj.get("chunk_pairs") == CHUNK_PAIRS:
                traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    tot_k = sum(v[0] for v in kms.values()) or 1.0

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_baseline_leg()
            except Exception as ex:          # the baseline is a report, never a dependency of the GPU path
                cpu = dict(error=str(ex))
        line = dict(
            metric="VDL2 channel-samples demodulated per second (2.1 Msps channels; real-time channels = value/2.1)",
            value=value, unit="Msamples/s", n_gpus=world, steps=K, warmup=W, ms_per_step=ms_dev / K,
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
            config=dict(workload=f"{args.channels} channels/GPU ({n_total} total) = 64 slots x {n_total // 64} replicas from one synthetic "
                                 f"2.1 Msps cu8 stream (BASELINE config 5 shape; Poisson bursts 2/s/slot, Es/N0 20 dB); "
                                 f"step = {CPS} chunks x {CHUNK_PAIRS} IQ pairs",
                        fs=FS, oversample=OVERSAMPLE, channels_per_gpu=args.channels, channels_total=n_total, sample_fmt="cu8",
                        chunk_pairs=CHUNK_PAIRS, chunks_per_step=CPS, parallelism=f"channel-shard x{world} (k mod N), NCCL broadcast of IQ" if world > 1 else "single GPU",
                        l2=f"per-chunk working set (decimated buffer {len(my_freqs) * (CHUNK_PAIRS // OVERSAMPLE) * 8 / 1e6:.0f} MB written by K1, read by K2) exceeds the 126 MB L2",
                        k1_impl="scalar" if args.k1_scalar else "pipelined f32x2",
                        pipeline="two streams: K0/K1 of chunk c+1 beside K2a/K2/K3 of chunk c"),
            channels_at_realtime=value / 2.1,
            frames_per_step=frames_dev / K,
            e2e=dict(value=e2e_value, unit="Msamples/s", h2d_bytes_per_step=CPS * CHUNK_BYTES if True else 0,
                     d2h_bytes_per_step=d_e2e["out_bytes"] / K, ms_per_step=ms_e2e / K, channels_at_realtime=e2e_value / 2.1,
                     frames_per_step=frames_e2e / K, api="Vdl2Channels.process_buf_uchar (vdl2gpu_submit) + flush"),
            gpu_launches=int(d_dev["kernel_launches"]),
            roofline=dict(bound="hbm", kernel="k1_mix_iir_decimate", achieved=achieved, peak=peak, unit="GB/s",
                          frac=(achieved / peak) if achieved else None, traffic=traffic, peak_source=peak_src,
                          bytes_per_channel_sample=B_CS, ms_per_launch=k1_ms_per_launch,
                          note="effective-bandwidth model of SURVEY.md §8d (each channel streams the float IQ buffer); the shared stream is "
                               "served from shared memory so DRAM traffic is far lower and the kernel is FP32-issue bound",
                          kernel_share_of_step={k: v[0] / tot_k for k, v in kms.items()},
                          kernel_ms_per_launch={k: v[0] / max(v[1], 1) for k, v in kms.items()},
                          measured="CUDA events on the library's stream, kernels serialised (VDL2GPU_FLAG_NO_OVERLAP), same "
                                   "workload and process; value/e2e are measured with the two-stage stream overlap on",
                          serial_ms_per_step=ms_serial / max(2, K // 2),
                          fp32_pipe=dict(
                              achieved=(len(my_freqs) * CHUNK_PAIRS * 24.0 / (k1_ms_per_launch * 1e-3) / 1e12) if k1_ms_per_launch > 0 else None,
                              peak=148 * 128 * 1.965e9 / 1e12, unit="T lane-ops/s",
                              frac=(len(my_freqs) * CHUNK_PAIRS * 24.0 / (k1_ms_per_launch * 1e-3) / (148 * 128 * 1.965e9)) if k1_ms_per_launch > 0 else None,
                              note="the bound that actually binds K1: 24 individually rounded FP32 operations per channel-sample "
                                   "(12 packed FMUL2/FFMA2; contraction to FMA is not allowed by the bit-exactness contract) "
                                   "against 148 SMs x 128 FP32 lanes x 1.965 GHz")),
            clocks=clocks,
            parity=dict(pool_overflows=int(d_dev["pool_overflows"] + d_e2e["pool_overflows"]), out_overflows=int(d_dev["out_overflows"] + d_e2e["out_overflows"]),
                        bursts_per_step=d_dev["bursts"] / K, fcs_good_per_step=d_dev["fcs_good"] / K, fcs_bad_per_step=d_dev["fcs_bad"] / K),
            wall_ms_per_step=wall_dev / K,
        )
        if cpu is not None:
            line["cpu_baseline"] = cpu
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    barrier()
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
