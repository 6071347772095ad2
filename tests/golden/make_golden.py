"""Generate tests/golden/*.json by running the UNMODIFIED reference (oracle/_ref/vdl2_ref_{strict,fast},
built from /root/reference by oracle/Makefile) on the deterministic streams of tests/cases.py.
Run in the build container only:  python tests/golden/make_golden.py
"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import pyoracle as po
from tests import cases


def main():
    po.build()
    for name, fn in cases.ALL_GOLDEN.items():
        c = fn()
        with tempfile.NamedTemporaryFile(suffix=".iq", delete=False) as tf:
            if name == "wav":
                tf.write(c["raw_bytes"])      # the reference feeds the whole file, header included
            else:
                tf.write(np.ascontiguousarray(c["iq"]).view(np.uint8).tobytes())
            path = tf.name
        out = dict(name=name, fs=c["fs"], oversample=c["oversample"], fmt=c["fmt"], centerfreq=c["centerfreq"],
                   freqs=c["freqs"], chunk=c["chunk"], iq_sha256=cases.iq_sha256(c),
                   injected=[[f.hex() for f in b.frames] for b in c["bursts"]])
        for fl in ("strict", "fast"):
            fr, _ = po.run_ref(path, po.FMT_S16 if c["fmt"] == "s16" else po.FMT_U8, c["oversample"], c["centerfreq"],
                               c["freqs"], flavour=fl, chunk=c["chunk"])
            out[fl] = [dict(channel=f["channel"], idx=f["idx"], hex=f["data"].hex(), synd_weight=f["synd_weight"],
                            datalen_octets=f["datalen_octets"], num_fec_corrections=f["num_fec_corrections"],
                            frame_pwr_dbfs=f["frame_pwr_dbfs"], nf_pwr_dbfs=f["nf_pwr_dbfs"], ppm_error=f["ppm_error"]) for f in fr]
        os.unlink(path)
        with open(os.path.join(cases.GOLDEN, f"{name}.json"), "w") as f:
            json.dump(out, f, indent=1)
        same = [x["hex"] for x in out["strict"]] == [x["hex"] for x in out["fast"]]
        ninj = sum(len(b) for b in out["injected"])
        print(f"{name}: strict {len(out['strict'])} frames, fast {len(out['fast'])} frames, injected {ninj}, strict==fast bytes: {same}")


if __name__ == "__main__":
    main()
