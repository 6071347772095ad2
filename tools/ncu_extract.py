#!/usr/bin/env python
"""tools/ncu_extract.py — condense `ncu --set full` reports (gpurun_out/*.ncu-rep) into the small tracked files under profiles/:

    python tools/ncu_extract.py --channels 16384 --chunk-pairs 262144 --tag r02 K1=gpurun_out/x_k1.ncu-rep K2=... K2a=... K3=...

writes profiles/<tag>_ncu_<stage>.csv (the metrics the judge greps: duration, DRAM bytes, issue, pipes, stall reasons) and
updates profiles/kernel_counters.json (DRAM bytes and warp-instructions per launch, read by bench.py for `roofline.stages`).
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_allocated",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def scale(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * m.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, required=True)
    ap.add_argument("--chunk-pairs", type=int, default=262144)
    ap.add_argument("--tag", default="r02")
    ap.add_argument("--note", default="")
    ap.add_argument("reports", nargs="+", help="STAGE=path.ncu-rep")
    a = ap.parse_args()
    kc_path = os.path.join(ROOT, "profiles", "kernel_counters.json")
    kc = json.load(open(kc_path)) if os.path.exists(kc_path) else {}
    for item in a.reports:
        stage, rep = item.split("=", 1)
        m = raw(rep)
        name = m.get("Kernel Name", ("", ""))[1]
        path = os.path.join(ROOT, "profiles", f"{a.tag}_ncu_{stage.lower()}.csv")
        with open(path, "w") as f:
            f.write(f"Kernel Name,,{name}\n")
            if a.note:
                f.write(f"note,,{a.note}\n")
            for k in KEEP:
                if k in m:
                    f.write(f"{k},{m[k][0]},{m[k][1]}\n")
            for k in sorted(m):
                if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                    f.write(f"{k},{m[k][0]},{m[k][1]}\n")
        rd = scale(num(m["dram__bytes_read.sum"][1]), m["dram__bytes_read.sum"][0])
        wr = scale(num(m["dram__bytes_write.sum"][1]), m["dram__bytes_write.sum"][0])
        kc[stage] = dict(kernel=name, channels=a.channels, chunk_pairs=a.chunk_pairs, dram_bytes_read=rd, dram_bytes_write=wr,
                         dram_bytes_per_launch=rd + wr, warp_instructions_per_launch=num(m["smsp__inst_executed.sum"][1]),
                         ncu_duration_ms=num(m["gpu__time_duration.sum"][1]) * ({"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(m["gpu__time_duration.sum"][0], 1.0)),
                         issue_active_pct=num(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][1]), source=os.path.basename(path))
        print(stage, kc[stage])
    with open(kc_path, "w") as f:
        json.dump(kc, f, indent=1)


if __name__ == "__main__":
    main()
