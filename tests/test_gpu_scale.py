"""-m gpu: BASELINE.json's full-size configurations through size-independent properties.

The oracle cannot run 4096 / 16384 channels in seconds, but the channels of these configurations are replicas of
64 frequency slots fanned out from one stream: the oracle decodes one channel per slot, and every
replica of a slot must produce exactly that channel's frames, metadata and counters (the per-channel state machines
are independent, so any cross-channel interference in the kernels - ring columns, burst pool, output records,
warp-divergent paths - would break the equality)."""
import numpy as np
import pytest
import dumpvdl2_b200 as vd
from dumpvdl2_b200 import synth
from oracle import pyoracle as po
from tests import cases, util

pytestmark = pytest.mark.gpu
FS, CENTER = 2100000, cases.CENTER


def _run(n_slots, n_rep, seconds, es_n0, seed, chunk=524288):
    iq, offs, bursts = synth.traffic_stream(FS, seconds, n_slots, 4.0, es_n0, -20.0, seed, "u8")
    slot_freqs = [CENTER + int(o) for o in offs]
    o = po.Oracle(FS, 20, po.FMT_U8, CENTER, slot_freqs)
    o.process_chunked(iq, chunk)
    want = {}
    for f in o.frames():
        want.setdefault(f.channel, []).append((f.burst_seq, f.idx, f.data, f.num_fec_corrections, f.synd_weight,
                                               np.float32(f.frame_pwr).tobytes(), np.float32(f.ppm_error).tobytes(), f.sync_dec_index))
    freqs = [f for f in slot_freqs for _ in range(n_rep)]          # channel = slot * n_rep + replica
    g = vd.Vdl2Channels(FS, 20, vd.FMT_U8, CENTER, freqs, max_chunk_bytes=chunk)
    g.process_chunked(iq, chunk)
    got = {}
    for f in g.flush():
        got.setdefault(f.channel, []).append((f.burst_seq, f.idx, f.data, f.num_fec_corrections, f.synd_weight,
                                              np.float32(f.frame_pwr).tobytes(), np.float32(f.ppm_error).tobytes(), f.sync_dec_index))
    st = g.stats()
    cnt = g.channel_counters()
    return want, got, st, cnt, o.counters(), bursts


def test_config4_4096_channels_every_replica_equals_the_oracle_channel():
    """BASELINE config 4 shape: 4096 channels = 64 slots x 64 replicas, Es/N0 20 dB (mid-sweep: corrected, failed and
    clean bursts all occur)."""
    want, got, st, cnt, ocnt, bursts = _run(64, 64, 1.0, 20.0, 0x56444C34)
    assert st["pool_overflows"] == 0 and st["out_overflows"] == 0
    n_frames = sum(len(v) for v in want.values())
    assert n_frames > 60
    for ch in range(4096):
        assert sorted(got.get(ch, [])) == sorted(want.get(ch // 64, [])), f"channel {ch} (slot {ch // 64}) differs from the oracle"
        assert np.array_equal(cnt[ch], ocnt[ch // 64])
    assert st["msg_good"] == 64 * n_frames


def test_config5_16384_channels_replica_consistency():
    """BASELINE config 5 shape on one GPU: 16384 channels = 64 slots x 256 replicas, 0.5 s."""
    want, got, st, cnt, ocnt, bursts = _run(64, 256, 0.5, 22.0, 0x56444C35)
    assert st["pool_overflows"] == 0 and st["out_overflows"] == 0
    n_frames = sum(len(v) for v in want.values())
    assert n_frames > 30 and st["msg_good"] == 256 * n_frames
    for ch in range(0, 16384, 37):          # a spread of replicas against the oracle ...
        assert sorted(got.get(ch, [])) == sorted(want.get(ch // 256, []))
    for slot in range(64):                  # ... and all replicas of every slot against each other (counters)
        block = cnt[slot * 256:(slot + 1) * 256]
        assert (block == block[0]).all() and np.array_equal(block[0], ocnt[slot])


def test_config3_256_channels_full_10_seconds():
    """BASELINE config 3 as specified: 10 s of 2.1 Msps cu8, 256 channels = 64 slots x 4 replicas, Poisson
    bursts 2/s/slot, Es/N0 20 dB, seed 0x56444C33.  The oracle decodes the 64 slot channels; every one of the 256 GPU
    channels must carry exactly its slot's frames, metadata and counters."""
    c = cases.case_replicas(n_slots=64, n_rep=1, duration=10.0, seed=0x56444C33, rate_hz=2.0)
    o = util.run_oracle(c)
    want = {}
    for f in o.frames():
        want.setdefault(f.channel, []).append(f)
    n_rep = 4
    freqs = [f for f in c["freqs"] for _ in range(n_rep)]
    g = vd.Vdl2Channels(FS, 20, vd.FMT_U8, CENTER, freqs, max_chunk_bytes=c["chunk"])
    g.process_chunked(util.case_bytes(c), c["chunk"])
    got = {}
    for f in g.flush():
        got.setdefault(f.channel, []).append(f)
    st = g.stats()
    assert st["pool_overflows"] == 0 and st["out_overflows"] == 0
    n = 0
    for ch in range(64 * n_rep):
        mine = got.get(ch, [])
        for f in mine:
            f.channel = ch // n_rep
        util.assert_frames_equal(mine, want.get(ch // n_rep, []), f"channel {ch}")
        n += len(mine)
    assert n == n_rep * len(o.frames()) and n > 1500
    cnt, ocnt = g.channel_counters(), o.counters()
    for ch in range(64 * n_rep):
        assert np.array_equal(cnt[ch], ocnt[ch // n_rep])


def test_balanced_slot_mapping_10048_channels():
    """More than half a machine's worth of channels are spread over all SM sub-partitions (ceil(n / 592) channels per
    warp instead of 32; here 17): channel numbers in frames and counters must be unaffected by the slot layout."""
    want, got, st, cnt, ocnt, bursts = _run(64, 157, 0.4, 22.0, 0x56444C36)
    assert st["pool_overflows"] == 0 and st["out_overflows"] == 0
    n_frames = sum(len(v) for v in want.values())
    assert n_frames > 20 and st["msg_good"] == 157 * n_frames
    for ch in range(64 * 157):
        assert sorted(got.get(ch, [])) == sorted(want.get(ch // 157, [])), f"channel {ch}"
        assert np.array_equal(cnt[ch], ocnt[ch // 157])
