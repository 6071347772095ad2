"""-m gpu: behaviour of the C-ABI surface beyond plain parity: device-buffer submission, incremental polling,
back-pressure, max_ppm veto, AWGN sweep (BASELINE config 4 shape, scaled), empty / ragged input."""
import numpy as np
import pytest
import dumpvdl2_b200 as vd
from dumpvdl2_b200 import synth
from oracle import pyoracle as po
from tests import cases, util

pytestmark = pytest.mark.gpu


def _gpu(case, **kw):
    return vd.Vdl2Channels(case["fs"], case["oversample"], util.fmt_code(case), case["centerfreq"], case["freqs"],
                           max_ppm=case.get("max_ppm", 0.0), max_chunk_bytes=max(case["chunk"], 1 << 16), **kw)


def test_submit_device_and_poll_incrementally():
    import torch
    c = cases.case_cfg2()
    o = util.run_oracle(c)
    g = _gpu(c, n_inflight=2)
    data = util.case_bytes(c)
    d = torch.from_numpy(data.copy()).cuda()
    st = torch.cuda.current_stream()
    got = []
    for off in range(0, data.size, c["chunk"]):
        n = min(c["chunk"], data.size - off)
        g.submit_device(d.data_ptr() + off, n, st.cuda_stream)
        got += g.poll()                      # whatever has finished so far
    got += g.flush()
    util.assert_frames_equal(got, o.frames(), "submit_device + poll")
    s = g.stats()
    assert s["chunks_submitted"] == s["chunks_completed"] == -(-data.size // c["chunk"])
    assert s["iq_samples"] == data.size // 2


def test_frames_are_delivered_in_channel_burst_order_and_timestamps_are_sane():
    import time
    c = cases.case_cfg2()
    t0 = time.time()
    g = _gpu(c)
    g.process_chunked(util.case_bytes(c), c["chunk"])
    fr = g.flush()
    assert len(fr) == 12
    for f in fr:
        assert t0 - 5 < f.burst_timestamp < time.time() + 1


def test_max_ppm_veto_matches_oracle():
    """Config.max_ppm (src/demod.c:192): bursts with a carrier offset beyond the limit are ignored."""
    c = cases.case_mixed_s16()
    c = dict(c); c["max_ppm"] = 1.0          # the case has +-300 Hz (up to ~2.2 ppm) offsets: some bursts vetoed
    o = util.run_oracle(c, trace=True)
    g = _gpu(c, flags=vd.FLAG_TRACE)
    g.process_chunked(util.case_bytes(c), c["chunk"])
    fr = g.flush()
    util.assert_frames_equal(fr, o.frames(), "max_ppm=1.0")
    util.assert_events_equal(g.read_events(), o.events(), "max_ppm events")
    vetoed = [e for e in o.events() if e["kind"] == 1 and e["i"][3] == 0]
    assert 0 < len(vetoed) and len(fr) < 14


@pytest.mark.parametrize("es_n0", [14.0, 17.0, 20.0, 23.0])
def test_awgn_sweep_decode_rate_parity(es_n0):
    """BASELINE config 4 (scaled): same frame set as the oracle at every SNR point, incl. the RS-corrected and the
    failed bursts; the decode rate rises with SNR."""
    fs = 2100000
    offs = synth.slot_offsets(16, 50e3)
    rng = np.random.default_rng(0x56444C34 + int(es_n0))
    bursts, t = [], 0.01
    for k in range(48):
        fr = synth.random_frames(rng)
        bursts.append(synth.BurstSpec(t, offs[k % 16], fr, power_dbfs=-18.0))
        t += synth.burst_duration_s(fr) / 6 + 0.004       # up to ~6 bursts on air at once, on different channels
    iq = synth.synth_stream(fs, t + 0.2, bursts, es_n0_db=es_n0, fmt="u8", seed=int(es_n0) + 1)
    c = cases._mk("awgn", fs, "u8", [cases.CENTER + o for o in offs], iq, chunk=524288, bursts=bursts)
    o = util.run_oracle(c)
    g = _gpu(c)
    g.process_chunked(util.case_bytes(c), c["chunk"])
    fr = g.flush()
    util.assert_frames_equal(fr, o.frames(), f"Es/N0 {es_n0} dB")
    assert np.array_equal(g.channel_counters(), o.counters())
    injected = {f for b in bursts for f in b.frames}
    good = sum(1 for f in fr if f.fcs_ok and f.data in injected)
    rate = good / len(injected)
    print(f"Es/N0 {es_n0:.0f} dB: {good}/{len(injected)} frames decoded (rate {rate:.2f}), identical to the oracle")
    if es_n0 >= 23:
        assert rate > 0.8
    if es_n0 <= 14:
        assert rate < 0.5


def test_empty_ragged_and_oversize_input():
    c = cases.case_cfg2(0.05)
    g = _gpu(c)
    g.submit(np.zeros(0, np.uint8))                    # len == 0 is ignored (src/demod.c:341)
    g.submit(np.full(1, 128, np.uint8))                # less than one IQ pair
    g.submit(np.full(2 * 33 + 1, 127, np.uint8))       # trailing unpaired byte ignored
    assert g.flush() == []
    with pytest.raises(vd.Vdl2GpuError, match="larger than max_chunk_bytes"):
        g.submit(np.zeros((1 << 20) + 2, np.uint8))
    s = g.stats()
    assert s["iq_samples"] == 33 and s["pool_overflows"] == 0


def test_two_contexts_are_independent():
    c1, c2 = cases.case_cfg2(0.4), cases.case_mixed_s16()
    g1, g2 = _gpu(c1), _gpu(c2)
    b1, b2 = util.case_bytes(c1), util.case_bytes(c2)
    for k in range(max(-(-b1.size // c1["chunk"]), -(-b2.size // c2["chunk"]))):
        g1.submit(b1[k * c1["chunk"]:(k + 1) * c1["chunk"]])
        g2.submit(b2[k * c2["chunk"]:(k + 1) * c2["chunk"]])
    util.assert_frames_equal(g1.flush(), util.run_oracle(c1).frames(), "ctx 1")
    util.assert_frames_equal(g2.flush(), util.run_oracle(c2).frames(), "ctx 2")


def test_planar_s16_ingest_equals_interleaved():
    """SDRplay hand-off (separate I / Q arrays, src/sdrplay.c:72-134) through vdl2gpu_submit_planar_s16."""
    c = cases.case_mixed_s16()
    o = util.run_oracle(c)
    g = _gpu(c)
    iq = c["iq"].reshape(-1, 2)
    step = c["chunk"] // 4
    for k in range(0, iq.shape[0], step):
        part = iq[k:k + step]
        g.submit_planar_s16(part[:, 0].copy(), part[:, 1].copy())
    util.assert_frames_equal(g.flush(), o.frames(), "planar cs16 ingest")
