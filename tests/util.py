"""Shared helpers for the parity tests."""
import json
import os
import numpy as np
from oracle import pyoracle as po
from tests import cases


def load_golden(name):
    with open(os.path.join(cases.GOLDEN, f"{name}.json")) as f:
        return json.load(f)


def fmt_code(case):
    return po.FMT_S16 if case["fmt"] == "s16" else po.FMT_U8


def case_bytes(case):
    """The byte stream handed to process_buf_*: the WAV case feeds the raw file, header included."""
    if "raw_bytes" in case:
        return np.frombuffer(case["raw_bytes"], np.uint8)
    return np.ascontiguousarray(case["iq"]).view(np.uint8).reshape(-1)


def run_oracle(case, trace=False, dec_tap=False, chunk=None):
    o = po.Oracle(case["fs"], case["oversample"], fmt_code(case), case["centerfreq"], case["freqs"],
                  max_ppm=case.get("max_ppm", 0.0), trace=trace, dec_tap=dec_tap)
    o.process_chunked(case_bytes(case), chunk or case["chunk"])
    return o


def frame_tuple(f):
    """Everything that must be bit-identical between two implementations of the path."""
    return (f.channel, f.burst_seq, f.idx, f.data, f.synd_weight, f.datalen_octets, f.num_fec_corrections)


def assert_frames_equal(got, want, what, meta_exact=True):
    got = sorted(got, key=lambda f: f.key())
    want = sorted(want, key=lambda f: f.key())
    assert len(got) == len(want), f"{what}: {len(got)} frames, expected {len(want)}"
    for a, b in zip(got, want):
        assert frame_tuple(a) == frame_tuple(b), f"{what}: frame mismatch\n got  {a}\n want {b}"
        assert a.fcs_ok == b.fcs_ok
        assert a.sync_dec_index == b.sync_dec_index, f"{what}: sync index {a.sync_dec_index} != {b.sync_dec_index} for {a}"
        if meta_exact:
            for fld in ("frame_pwr", "mag_nf", "ppm_error", "frame_pwr_dbfs", "nf_pwr_dbfs"):
                x, y = np.float32(getattr(a, fld)), np.float32(getattr(b, fld))
                assert x.view(np.uint32) == y.view(np.uint32) or (np.isnan(x) and np.isnan(y)), f"{what}: {fld} {x!r} != {y!r} for {a}"
        else:
            assert abs(a.frame_pwr_dbfs - b.frame_pwr_dbfs) < 0.01 and abs(a.ppm_error - b.ppm_error) < 0.01


def assert_matches_golden(frames, golden, flavour, what):
    """frames (objects with channel/idx/data/metadata) against the reference's own output stored in tests/golden."""
    want = golden[flavour]
    got = sorted(frames, key=lambda f: f.key())
    # the reference harness prints per channel in push order == (burst, idx) order
    want_sorted = sorted(range(len(want)), key=lambda i: (want[i]["channel"], i))
    assert len(got) == len(want), f"{what}: {len(got)} frames, reference has {len(want)}"
    for a, i in zip(got, want_sorted):
        b = want[i]
        assert a.channel == b["channel"] and a.idx == b["idx"], f"{what}: order mismatch {a} vs {b['channel']}/{b['idx']}"
        assert a.data.hex() == b["hex"], f"{what}: frame bytes differ for {a}"
        assert a.synd_weight == b["synd_weight"] and a.datalen_octets == b["datalen_octets"]
        if flavour == "strict":     # -ffast-math float differences may flip a marginal symbol; RS repairs it, the count differs
            assert a.num_fec_corrections == b["num_fec_corrections"]
        tol = 0.0 if flavour == "strict" else 0.01           # 0.01 dB / 0.01 ppm vs the -ffast-math build
        for fld in ("frame_pwr_dbfs", "nf_pwr_dbfs", "ppm_error"):
            x, y = float(np.float32(getattr(a, fld))), float(np.float32(b[fld]))
            if tol == 0.0:
                assert float(f"{x:.9g}") == float(f"{y:.9g}"), f"{what}: {fld} {x!r} != reference {y!r}"
            else:
                assert abs(x - y) <= tol, f"{what}: {fld} {x} vs reference {y}"


def events_key(e):
    return (e["channel"], e["dec_index"], e["kind"])


def assert_events_equal(got, want, what, kinds=(1, 2)):
    g = sorted([e for e in got if e["kind"] in kinds], key=events_key)
    w = sorted([e for e in want if e["kind"] in kinds], key=events_key)
    assert len(g) == len(w), f"{what}: {len(g)} events, expected {len(w)}"
    for a, b in zip(g, w):
        assert events_key(a) == events_key(b) and a["i"] == b["i"], f"{what}: event mismatch {a} vs {b}"
        assert np.array_equal(np.asarray(a["f"], np.float32).view(np.uint32), np.asarray(b["f"], np.float32).view(np.uint32)), \
            f"{what}: event floats differ {a} vs {b}"
