import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import dumpvdl2_b200 as vd
chunks, offs, _ = bench.make_stream(2.0)
freqs = bench.channel_freqs(offs, 16384)
d_chunks = torch.from_numpy(chunks).cuda()
st = torch.cuda.current_stream()
g = vd.Vdl2Channels(bench.FS, bench.OVERSAMPLE, vd.FMT_U8, bench.CENTER, freqs, max_chunk_bytes=bench.CHUNK_BYTES)
def run(nsteps, per, dev):
    for s in range(nsteps):
        for i in range(per):
            if dev: g.submit_device(d_chunks[(s * per + i) % 16].data_ptr(), bench.CHUNK_BYTES, st.cuda_stream)
            else: g.process_buf_uchar(chunks[(s * per + i) % 16])
        g.flush_count()
for name, dev in (("host", False), ("device", True)):
    run(2, 8, dev)
    torch.cuda.synchronize(); t0 = time.time(); run(4, 8, dev); torch.cuda.synchronize(); dt = time.time() - t0
    print(f"PRIO={os.environ.get('VDL2GPU_PRIO','0')} H2D={os.environ.get('VDL2GPU_EXP_H2D','0')} {name:8s} {1e3 * dt / 32:6.2f} ms/chunk", flush=True)
