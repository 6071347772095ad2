import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import dumpvdl2_b200 as vd
chunks, offs, _ = bench.make_stream(2.0)
freqs = bench.channel_freqs(offs, 16384)
d_chunks = torch.from_numpy(chunks).cuda()
st = torch.cuda.current_stream()
for name, flags, dev in (("serial host", vd.FLAG_NO_OVERLAP, False), ("overlap host", 0, False), ("overlap device", 0, True), ("serial device", vd.FLAG_NO_OVERLAP, True)):
    g = vd.Vdl2Channels(bench.FS, bench.OVERSAMPLE, vd.FMT_U8, bench.CENTER, freqs, max_chunk_bytes=bench.CHUNK_BYTES, flags=flags)
    def run(n):
        for i in range(n):
            if dev: g.submit_device(d_chunks[i % 16].data_ptr(), bench.CHUNK_BYTES, st.cuda_stream)
            else: g.process_buf_uchar(chunks[i % 16])
        return g.flush_count()
    run(8)
    g.enable_timing(True)
    torch.cuda.synchronize(); t0 = time.time()
    n = run(32)
    torch.cuda.synchronize(); t1 = time.time()
    print(f"{name:16s} {1e3*(t1-t0)/32:7.2f} ms/chunk  frames {n}  kernels " + " ".join(f"{k}={v[0]/max(v[1],1):.2f}" for k, v in g.kernel_ms().items()), flush=True)
    g.close()
