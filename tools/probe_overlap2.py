import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import dumpvdl2_b200 as vd
chunks, offs, _ = bench.make_stream(2.0)
freqs = bench.channel_freqs(offs, 16384)
d_chunks = torch.from_numpy(chunks).cuda()
st = torch.cuda.current_stream()
g = vd.Vdl2Channels(bench.FS, bench.OVERSAMPLE, vd.FMT_U8, bench.CENTER, freqs, max_chunk_bytes=bench.CHUNK_BYTES)
def run(nsteps, per, dev, own_stream=None):
    for s in range(nsteps):
        for i in range(per):
            if dev: g.submit_device(d_chunks[(s * per + i) % 16].data_ptr(), bench.CHUNK_BYTES, (own_stream or st).cuda_stream)
            else: g.process_buf_uchar(chunks[(s * per + i) % 16])
        g.flush_count()
def t(name, *a, **k):
    run(2, 8, a[2] if len(a) > 2 else False)
    torch.cuda.synchronize(); t0 = time.time(); run(*a, **k); torch.cuda.synchronize(); dt = time.time() - t0
    print(f"{name:40s} {1e3 * dt / (a[0] * a[1]):6.2f} ms/chunk", flush=True)
t("host, 4 steps x 8", 4, 8, False)
t("device, 4 steps x 8", 4, 8, True)
t("device, 1 step x 32", 1, 32, True)
s2 = torch.cuda.Stream()
t("device, 4x8, producer = side stream", 4, 8, True, own_stream=s2)
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader", "-lms", "200"], stdout=subprocess.DEVNULL)
time.sleep(0.5)
t("host, 4x8, nvidia-smi polling", 4, 8, False)
t("device, 4x8, nvidia-smi polling", 4, 8, True)
p.terminate()
