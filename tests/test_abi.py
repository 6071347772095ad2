"""CPU: the C-ABI library loads, exports every symbol include/*.h declares, and refuses to run without a GPU
(no CPU fallback on the demodulator path)."""
import ctypes as C
import os
import re
import numpy as np
import pytest
import dumpvdl2_b200 as vd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:vdl2gpu_|process_buf_|vdl2_channel_|sincosf_lut_|input_lpf_|demod_sync_|process_samples|rs_init|rs_verify|decode_vdl2_burst|vdl2gpu_dropin_)\w*)\s*\(", txt)))


@pytest.mark.parametrize("header", ["vdl2gpu.h", "vdl2_dropin.h"])
def test_every_declared_symbol_is_exported(header):
    if not os.path.exists(os.path.join(ROOT, "include", header)):
        pytest.skip(f"{header} not present")
    L = vd.load_library()
    names = _declared(header)
    assert len(names) >= 8
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/{header} but not exported by libvdl2gpu.so"


def test_abi_version_and_struct_sizes():
    L = vd.load_library()
    assert L.vdl2gpu_abi_version() == 1
    from dumpvdl2_b200 import api
    assert C.sizeof(api._Config) == 72 and C.sizeof(api._Stats) == 160 and C.sizeof(api._Event) == 80


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(vd.Vdl2GpuError, match="no usable CUDA device"):
        vd.Vdl2Channels(2100000, 20, vd.FMT_U8, 136975000, [136975000 + 25000])
    L = vd.load_library()
    assert L.vdl2gpu_device_count() == 0


def test_product_does_not_reach_into_oracle():
    """The oracle is the checker only: nothing under dumpvdl2_b200/ may import, link or load it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dumpvdl2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for token in ("pyoracle", "liboracle", "vdl2_oracle", "libhostsim", "pyhostsim", "from oracle", "import oracle"):
                    assert token not in txt, f"{f} references {token}"
    out = os.popen(f"ldd {vd.LIB_PATH}").read()
    assert "oracle" not in out


def test_bad_config_is_rejected():
    L = vd.load_library()
    from dumpvdl2_b200 import api
    cfg = api._Config()
    h = C.c_void_p()
    assert L.vdl2gpu_create(C.byref(cfg), C.byref(h)) == -1          # VDL2GPU_EINVAL
    f = np.array([136975000], np.uint32)
    cfg.sample_rate, cfg.oversample, cfg.n_channels = 2100000, 10, 1    # rate != 105000 * oversample
    cfg.freqs = f.ctypes.data_as(C.POINTER(C.c_uint32))
    assert L.vdl2gpu_create(C.byref(cfg), C.byref(h)) == -1
