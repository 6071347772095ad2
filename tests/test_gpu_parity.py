"""-m gpu: the CUDA path (through the C-ABI of libvdl2gpu.so) against the oracle and the reference's goldens.

Bar: bit-exact.  K1's decimated samples are compared bit for bit with the oracle's; K2's sync/header events
bit for bit (integers and floats); frames, metadata and counters identical to the oracle, and identical to the
unmodified reference's output stored in tests/golden/*.json.
"""
import numpy as np
import pytest
import dumpvdl2_b200 as vd
from oracle import pyoracle as po
from tests import cases, util

pytestmark = pytest.mark.gpu


def make_gpu(case, flags=0, chunk=None, n_inflight=4):
    return vd.Vdl2Channels(case["fs"], case["oversample"], util.fmt_code(case), case["centerfreq"], case["freqs"],
                           max_ppm=case.get("max_ppm", 0.0), max_chunk_bytes=max(chunk or case["chunk"], 1 << 16),
                           flags=flags, n_inflight=n_inflight)


def run_gpu(case, flags=0, chunk=None):
    g = make_gpu(case, flags, chunk)
    g.process_chunked(util.case_bytes(case), chunk or case["chunk"])
    frames = g.flush()
    return g, frames


@pytest.fixture(scope="module")
def oracle_runs():
    cache = {}

    def get(name):
        if name not in cache:
            c = cases.ALL_GOLDEN[name]()
            cache[name] = (c, util.run_oracle(c, trace=True, dec_tap=True))
        return cache[name]
    return get


def test_tables_match_oracle():
    c = cases.case_cfg2(0.01)
    g = make_gpu(c)
    t = g.tables()
    s, co = po.sincos_lut()
    a, b = po.lpf_design(c["fs"])
    x, d, p = po.sync_consts()
    for got, want in ((t["levels"], po.levels_u8()), (t["sin_lut"], s), (t["cos_lut"], co), (t["A"], a), (t["B"], b),
                      (t["lr_X"], x), (t["pr_phase"], p)):
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert t["lr_denom"][0] == d


@pytest.mark.parametrize("name", ["cfg2", "mixed_s16", "wav", "mirics_os13"])
@pytest.mark.parametrize("scalar", [False, True])
def test_k1_decimated_samples_bit_exact(name, scalar, oracle_runs):
    c, o = oracle_runs(name)
    odec = o.dec_samples()
    g = make_gpu(c, flags=vd.FLAG_KEEP_DEC | (vd.FLAG_K1_SCALAR if scalar else 0))
    b = util.case_bytes(c)
    pos = 0
    max_dec = c["chunk"] // 2 // c["oversample"] + 2
    for off in range(0, b.size, c["chunk"]):
        g.submit(b[off:off + c["chunk"]])
        d = g.read_dec(max_dec)
        want = odec[pos:pos + d.shape[0]]
        assert d.shape == want.shape
        assert np.array_equal(d.view(np.uint32), want.view(np.uint32)), f"{name}: K1 output differs in chunk at byte {off}"
        pos += d.shape[0]
    assert pos == odec.shape[0]
    g.flush()


@pytest.mark.parametrize("name", list(cases.ALL_GOLDEN))
def test_frames_match_oracle_and_reference(name, oracle_runs):
    c, o = oracle_runs(name)
    g, frames = run_gpu(c, flags=vd.FLAG_TRACE)
    st = g.stats()
    assert st["pool_overflows"] == 0 and st["out_overflows"] == 0
    util.assert_frames_equal(frames, o.frames(), f"gpu vs oracle [{name}]")
    gold = util.load_golden(name)
    assert cases.iq_sha256(c) == gold["iq_sha256"], "regenerated stream differs from the one the golden was made from"
    util.assert_matches_golden(frames, gold, "strict", f"gpu vs reference(strict) [{name}]")
    util.assert_matches_golden(frames, gold, "fast", f"gpu vs reference(-ffast-math) [{name}]")
    util.assert_events_equal(g.read_events(), o.events(), f"gpu vs oracle events [{name}]")
    assert np.array_equal(g.channel_counters(), o.counters()), f"per-channel counters differ [{name}]"
    assert st["msg_good"] == len(frames) and st["fcs_good"] == sum(f.fcs_ok for f in frames)


def test_trace_events_can_be_drained_between_chunks(oracle_runs):
    """the trace buffer restarts whenever it has been drained: events read chunk by chunk, concatenated, are the oracle's"""
    c, o = oracle_runs("cfg2")
    g = make_gpu(c, vd.FLAG_TRACE)
    b = util.case_bytes(c)
    got = []
    for off in range(0, b.size, c["chunk"]):
        g.submit(b[off:off + c["chunk"]])
        got += g.read_events()
    g.flush()
    got += g.read_events()
    assert g.read_events() == []
    util.assert_events_equal(got, o.events(), "events drained per chunk")


def test_golden_wav_sha256(oracle_runs):
    import hashlib
    c, _ = oracle_runs("wav")
    _, frames = run_gpu(c)
    h = "".join(f.data.hex() + "\n" for f in sorted(frames, key=lambda f: f.key()))
    assert hashlib.sha256(h.encode()).hexdigest() == "ee98da5344ee2b0167508c2f4e7efc95eb931217ce86b47472a0297c27415477"
    assert [len(f.data) for f in frames] == [314, 186] and all(f.fcs_ok for f in frames)


@pytest.mark.parametrize("chunk", [2 * 7, 2 * 12345, 2 * 100000, 1 << 20])
def test_chunking_does_not_change_results(chunk, oracle_runs):
    """src/demod.c:289-298: filter, NCO and decimation state persist across buffers."""
    c, o = oracle_runs("cfg2")
    if chunk < 1000:       # tiny chunks: only a prefix of the stream, enough to cross many decimation groups
        c = dict(c); c["iq"] = c["iq"][:2 * 6000]
        o = util.run_oracle(c)
    _, frames = run_gpu(c, chunk=chunk)
    util.assert_frames_equal(frames, o.frames(), f"chunk={chunk}")


def test_scalar_and_pipelined_k1_agree_end_to_end(oracle_runs):
    c, o = oracle_runs("noisy")
    _, f1 = run_gpu(c, flags=vd.FLAG_K1_SCALAR)
    util.assert_frames_equal(f1, o.frames(), "scalar K1")


def test_replicas_and_many_channels():
    """BASELINE config 3 shape: replicas sharing a slot must produce identical frames; 256 channels."""
    c = cases.case_replicas(n_slots=64, n_rep=4, duration=0.6)
    o = util.run_oracle(c)
    g, frames = run_gpu(c)
    util.assert_frames_equal(frames, o.frames(), "replicas 64x4")
    by = {}
    for f in frames:
        by.setdefault(f.channel, []).append((f.burst_seq, f.idx, f.data))
    for s in range(64):
        for r in range(1, 4):
            assert by.get(4 * s, []) == by.get(4 * s + r, []), f"replicas of slot {s} disagree"
    assert len(frames) > 100
    st = g.stats()
    assert st["pool_overflows"] == 0 and st["out_overflows"] == 0


def test_raw_launch_stubs():
    """K4 / RS through the plain-pointer stubs, against the oracle's stage functions (K0, K1, K2, K3: test_gpu_stage.py)."""
    import ctypes as C
    import torch
    L = vd.load_library()
    rng = np.random.default_rng(3)
    # RS: codewords with 0..4 errors, shortened blocks
    n = 512
    blocks = np.zeros((n, 255), np.uint8); fec = np.zeros(n, np.int32); want_ret = np.zeros(n, np.int32); want = np.zeros_like(blocks)
    for i in range(n):
        b = po.rs_encode(rng.integers(0, 256, 249, dtype=np.uint8))
        nf = [6, 6, 6, 4, 2, 0][i % 6]
        if nf < 6:
            b[60:249] = 0; b = po.rs_encode(b[:249]); b[249 + nf:] = 0
        for _ in range(int(rng.integers(0, 5))):
            b[int(rng.integers(0, 249 + nf))] ^= int(rng.integers(1, 256))
        blocks[i] = b; fec[i] = nf
        want_ret[i], want[i] = po.rs_verify(b, nf)
    d_b = torch.from_numpy(blocks).cuda(); d_f = torch.from_numpy(fec).cuda(); d_r = torch.zeros(n, dtype=torch.int32, device="cuda")
    assert L.vdl2gpu_launch_rs_verify(d_b.data_ptr(), d_f.data_ptr(), n, d_r.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_r.cpu().numpy(), want_ret)
    assert np.array_equal(d_b.cpu().numpy(), want)
    # K4
    frames = [rng.integers(0, 256, int(rng.integers(0, 400)), dtype=np.uint8).tobytes() for _ in range(100)]
    blob = np.frombuffer(b"".join(frames) + b"\0", np.uint8).copy()
    lens = np.array([len(f) for f in frames], np.uint32); offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint32)
    d_blob = torch.from_numpy(blob).cuda(); d_o = torch.from_numpy(offs.view(np.int32)).cuda(); d_l = torch.from_numpy(lens.view(np.int32)).cuda()
    d_c = torch.zeros(100, dtype=torch.int16, device="cuda")
    assert L.vdl2gpu_launch_fcs_crc16(d_blob.data_ptr(), d_o.data_ptr(), d_l.data_ptr(), 100, d_c.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert [int(x) & 0xFFFF for x in d_c.cpu().numpy()] == [po.crc16(f) for f in frames]
