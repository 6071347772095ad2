"""CPU: the oracle restatement against (a) the reference's own fixture and the outputs of the unmodified
reference stored in tests/golden (made by tests/golden/make_golden.py), (b) the live oracle/_ref binaries
when they are present, (c) the reference's table literals when /root/reference is mounted."""
import hashlib
import os
import re
import tempfile
import numpy as np
import pytest
from oracle import pyoracle as po
from tests import cases, util


@pytest.mark.parametrize("name", list(cases.ALL_GOLDEN))
def test_oracle_matches_reference_golden(name):
    c = cases.ALL_GOLDEN[name]()
    gold = util.load_golden(name)
    assert cases.iq_sha256(c) == gold["iq_sha256"], "stream regenerated from the seed differs from the golden's"
    o = util.run_oracle(c)
    util.assert_matches_golden(o.frames(), gold, "strict", f"oracle vs reference(strict) [{name}]")
    util.assert_matches_golden(o.frames(), gold, "fast", f"oracle vs reference(-ffast-math) [{name}]")


def test_golden_wav_is_the_reference_fixture():
    """1 burst -> 2 frames (314 + 186 octets), both FCS-good, sha256 ee98da53...; the CI of the
    reference greps for the two tails below (.github/workflows/build.yml:17)."""
    c = cases.case_wav()
    assert len(c["raw_bytes"]) == 1582220
    fr = util.run_oracle(c).frames()
    h = "".join(f.data.hex() + "\n" for f in fr)
    assert hashlib.sha256(h.encode()).hexdigest() == "ee98da5344ee2b0167508c2f4e7efc95eb931217ce86b47472a0297c27415477"
    assert [len(f.data) for f in fr] == [314, 186] and all(f.fcs_ok for f in fr)
    assert fr[0].data.endswith(bytes.fromhex("202d5241204252204f56433030350a44bf"))
    assert fr[1].data.endswith(bytes.fromhex("20534c503133350a3ef9"))
    assert fr[0].sync_dec_index == 11975 and fr[0].datalen_octets == 504 and fr[0].num_fec_corrections == 0


def test_chunking_invariance_of_oracle():
    c = cases.case_cfg2(0.4)
    a = util.run_oracle(c, chunk=320000).frames()
    b = util.run_oracle(c, chunk=2 * 9973).frames()
    util.assert_frames_equal(a, b, "oracle chunk 320000 vs 19946")


def test_injected_frames_come_back():
    """generator <-> decoder round trip at high SNR"""
    c = cases.case_cfg2()
    got = {f.data for f in util.run_oracle(c).frames()}
    want = {fr for b in c["bursts"] for fr in b.frames}
    assert want <= got and len(want) == 12


@pytest.mark.skipif(po.ref_binary("strict") is None, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("flavour", ["strict", "fast"])
def test_oracle_matches_live_reference_on_replicas(flavour):
    c = cases.case_replicas(n_slots=8, n_rep=2, duration=0.4)
    with tempfile.NamedTemporaryFile(suffix=".cu8") as tf:
        tf.write(c["iq"].tobytes()); tf.flush()
        ref, _ = po.run_ref(tf.name, po.FMT_U8, c["oversample"], c["centerfreq"], c["freqs"], flavour=flavour, chunk=c["chunk"])
    fr = sorted(util.run_oracle(c).frames(), key=lambda f: f.key())
    assert len(fr) == len(ref) and len(fr) > 4
    for a, b in zip(fr, ref):
        assert (a.channel, a.idx, a.data) == (b["channel"], b["idx"], b["data"])
        if flavour == "strict":     # the -ffast-math build may slice a marginal symbol differently (RS then repairs it)
            assert a.num_fec_corrections == b["num_fec_corrections"]


@pytest.mark.skipif(po.ref_binary("strict") is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_live_reference_on_stress_stream():
    """Demodulator-side differential: carrier offsets, clipping / very weak bursts, back-to-back and overlapping
    bursts (cases.case_stress) - frames and printed metadata identical to the strictly compiled reference."""
    c = cases.case_stress()
    with tempfile.NamedTemporaryFile(suffix=".cu8") as tf:
        tf.write(c["iq"].tobytes()); tf.flush()
        ref, _ = po.run_ref(tf.name, po.FMT_U8, c["oversample"], c["centerfreq"], c["freqs"], flavour="strict", chunk=c["chunk"])
    fr = sorted(util.run_oracle(c).frames(), key=lambda f: f.key())
    assert len(fr) == len(ref) and len(fr) > 15, (len(fr), len(ref))
    for a, b in zip(fr, ref):
        assert (a.channel, a.idx, a.data, a.num_fec_corrections, a.synd_weight) == (b["channel"], b["idx"], b["data"], b["num_fec_corrections"], b["synd_weight"])
        assert abs(a.frame_pwr_dbfs - b["frame_pwr_dbfs"]) < 2e-3 and abs(a.nf_pwr_dbfs - b["nf_pwr_dbfs"]) < 2e-3
        assert abs(a.ppm_error - b["ppm_error"]) < 2e-3


@pytest.mark.skipif(po.ref_binary("strict") is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_live_reference_on_fuzzed_payloads():
    """Differential fuzz of the burst decoder against the reference itself: transmissions whose payloads are random
    bits, or stuffed frames cut at a random bit, some with corrupted code words and header bits, modulated and run
    through oracle/_ref (unmodified src/decode.c, bitstream.c, rs.c) and through the restatement."""
    from dumpvdl2_b200 import synth
    rng = np.random.default_rng(0x7E7E)
    F = "01111110"

    def stuffed(n):
        return "".join(format(b, "08b")[::-1] for b in rng.integers(0, 256, n, dtype=np.uint8)).replace("11111", "111110")
    fs, offs = 2100000, [100e3, -100e3]
    bursts, t = [], [0.02, 0.02]
    for k in range(18):
        nbits = int(rng.integers(40, 5000))
        if k % 3 == 0:
            pl = "".join("01"[b] for b in rng.integers(0, 2, nbits, dtype=np.uint8))
        else:
            pl = F
            while len(pl) < nbits:
                pl += stuffed(int(rng.integers(0, 200))) + F * int(rng.integers(1, 3))
            if k % 3 == 1:
                pl = pl[:nbits]
        n_oct = -(-len(pl) // 8)
        nblk = -(-n_oct // 249)
        corrupt = [(int(rng.integers(0, nblk)), int(rng.integers(0, min(249, n_oct - (nblk - 1) * 249))), int(rng.integers(1, 256)))
                   for _ in range(int(rng.integers(0, 5)) if k % 2 else 0)]
        hdr = (int(rng.integers(0, 25)),) if k % 5 == 4 else ()
        ch = k % 2
        try:
            synth.burst_bits_from_payload(pl, corrupt)
        except ValueError:
            continue
        bursts.append(synth.BurstSpec(t[ch], offs[ch], [], power_dbfs=-12.0, payload=pl, corrupt_octets=corrupt, header_bit_errors=hdr))
        t[ch] += synth.burst_duration_s(None, payload=pl) + 0.012
    iq = synth.synth_stream(fs, max(t) + 0.02, bursts, es_n0_db=32, fmt="u8", seed=0x7E7F)
    c = cases._mk("fuzz", fs, "u8", [cases.CENTER + o for o in offs], iq, chunk=524288, bursts=bursts)
    with tempfile.NamedTemporaryFile(suffix=".cu8") as tf:
        tf.write(c["iq"].tobytes()); tf.flush()
        ref, _ = po.run_ref(tf.name, po.FMT_U8, c["oversample"], c["centerfreq"], c["freqs"], flavour="strict", chunk=c["chunk"])
    fr = sorted(util.run_oracle(c).frames(), key=lambda f: f.key())
    assert len(fr) == len(ref) and len(fr) > 20
    for a, b in zip(fr, ref):
        assert (a.channel, a.idx, a.data, a.num_fec_corrections, a.synd_weight) == (b["channel"], b["idx"], b["data"], b["num_fec_corrections"], b["synd_weight"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources not mounted")
def test_tables_against_reference_literals():
    src = open("/root/reference/src/decode.c").read()
    rows = [int(x, 2) for x in re.findall(r"0b([01]{25})", src.split("syndtable")[0])]
    pats = [int(x, 2) for x in re.findall(r"0b([01]{25})", src.split("syndtable")[1].split("synd_weight")[0])]
    weights = [int(x) for x in re.findall(r"\d+", src.split("synd_weight[1<<HDRFECLEN] = {")[1].split("}")[0])]
    import ctypes as C
    L = po.lib()
    for syn in range(32):
        assert L.vo_synd_weight(syn) == weights[syn]
    for bit in range(25):                   # single-bit errors decode back to the clean word
        w = C.c_uint32(po.lib().vo_header_encode(4029) ^ (1 << bit))
        L.vo_header_decode(C.byref(w))
        assert w.value == 0x2f7c0e
    # every syndrome maps to the reference's error pattern
    for pat in pats:
        w = C.c_uint32(pat)
        s = L.vo_header_decode(C.byref(w))
        assert w.value == 0 and pats[s] == pat
    assert len(rows) == 5
    crc_tab = [int(x, 16) for x in re.findall(r"0x([0-9A-F]{4})", open("/root/reference/src/crc.c").read().split("crctable[256]")[1])][:256]
    for b in range(256):
        assert po.crc16(bytes([b]), init=0) == crc_tab[b]


def test_rs_codec_properties():
    rng = np.random.default_rng(0)
    for trial in range(300):
        msg = rng.integers(0, 256, 249, dtype=np.uint8)
        cw = po.rs_encode(msg)
        assert po.rs_verify(cw, 6)[0] == 0
        nerr = int(rng.integers(0, 6))
        bad = cw.copy()
        pos = rng.choice(255, nerr, replace=False)
        for p in pos:
            bad[p] ^= int(rng.integers(1, 256))
        ret, fixed = po.rs_verify(bad, 6)
        if nerr <= 3:
            assert ret == nerr and np.array_equal(fixed, cw)
        else:
            assert ret == -1 or not np.array_equal(fixed, cw)      # beyond capacity: failure or miscorrection, never silent success
    # shortened last block: trailing zero fill, untransmitted parity declared erased (src/rs.c:36-47)
    for nfec, nerr_ok in ((4, 2), (2, 1)):
        msg = np.zeros(249, np.uint8); msg[:40] = rng.integers(0, 256, 40, dtype=np.uint8)
        cw = po.rs_encode(msg); rx = cw.copy(); rx[249 + nfec:] = 0
        ret, fixed = po.rs_verify(rx, nfec)
        assert ret == 6 - nfec and np.array_equal(fixed, cw)
        rx[3] ^= 0x40
        if nerr_ok >= 1:
            ret, fixed = po.rs_verify(rx, nfec)
            assert ret == 6 - nfec + 1 and np.array_equal(fixed, cw)


def test_empty_and_ragged_inputs():
    o = po.Oracle(2100000, 20, po.FMT_U8, cases.CENTER, [cases.CENTER + 25000])
    o.process(np.zeros(0, np.uint8))
    o.process(np.full(2 * 7, 128, np.uint8))
    o.process(np.full(2 * 13 + 1, 127, np.uint8))      # trailing unpaired component ignored
    assert o.frames() == [] and o.counters().sum() == 0
