"""CPU check of the Ziv-guarded fast atan2 / hypot (dumpvdl2_b200/csrc/vdl2_fastmath.cuh) that K2a and the K2 walk use on
the device: tools/check_fastmath.cpp compiles the same header with g++ and compares every variant (early-exit and
straight-line) with glibc's (float)atan2((double)im, (double)re) and (float)sqrt((double)re*re + (double)im*im)
(src/demod.c:232,238,256) bit for bit, over eight classes of inputs with the reciprocal / rsqrt seeds perturbed by
+-2^-18 to cover the device's MUFU seeds.  A short run here; the full 1.9e9-sample run is quoted in DESIGN.md."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_atan2_and_hypot_equal_glibc_on_the_cpu(tmp_path):
    exe = str(tmp_path / "check_fastmath")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-I", os.path.join(ROOT, "dumpvdl2_b200", "csrc"),
                           os.path.join(ROOT, "tools", "check_fastmath.cpp"), "-o", exe])
    r = subprocess.run([exe, "2", "4"], capture_output=True, text=True, timeout=600)      # 4 threads x 2 M samples
    assert r.returncode == 0, r.stdout[-2000:]
    m = re.search(r"samples (\d+)\s+slow (\d+) \(([\d.e+-]+)\)\s+mismatches (\d+)\s+special mismatches (\d+)", r.stdout)
    assert m, r.stdout
    assert int(m.group(1)) == 8000000 and int(m.group(4)) == 0 and int(m.group(5)) == 0
    h = re.search(r"hypot: slow \d+ \(([\d.e+-]+)\) mismatches (\d+)", r.stdout)
    assert h and int(h.group(2)) == 0
    # signal-like and noise-like inputs (modes 1, 2) must almost never need the slow path
    for mode in (1, 2):
        mm = re.search(rf"mode {mode} slow rate ([\d.e+-]+)", r.stdout)
        assert mm and float(mm.group(1)) < 1e-4, r.stdout
