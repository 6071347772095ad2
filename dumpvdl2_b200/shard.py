"""Multi-GPU plumbing for the path: channels are independent (reference: one pthread per channel, shared
read-only sample buffer - src/dumpvdl2.c:117-135, src/demod.c:300-301), so the path shards by channel.

  * channel k of the global list lives on rank k mod world (round-robin, BASELINE.json north_star);
  * the rank that owns the SDR / file (`src`) hands every raw IQ chunk to all ranks - the one transfer on the path.
    On GPUs that is the C helper of the library (vdl2gpu_mg_*, csrc/vdl2_mg.cu: NCCL broadcast or copy-engine peer
    copies into IPC-mapped receive buffers), wrapped here as MultiGpuIngest; `broadcast_chunk` is the plain
    torch.distributed form the gloo CPU tests use;
  * frames leave each rank independently; `gather_frame_keys` collects (channel, burst, idx, crc) tuples on the
    ingest rank for accounting.

torch.distributed is only plumbing here; the demodulation itself is libvdl2gpu.so.
"""
import numpy as np


def channels_for_rank(n_channels, rank, world):
    """Indices (into the global channel list) demodulated by `rank`."""
    return list(range(rank, n_channels, world))


def shard_freqs(freqs, rank, world):
    f = np.asarray(freqs)
    return f[rank::world]


def global_channel(local_index, rank, world):
    return local_index * world + rank


def broadcast_chunk(buf, src=0, group=None):
    """Broadcast one raw IQ chunk (a uint8 torch tensor, on the device the backend needs) from `src`."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(buf, src=src, group=group)
    return buf


def gather_frame_keys(keys, dst=0, group=None):
    """keys: list of tuples from this rank; returns the concatenated, sorted list on `dst` (None elsewhere)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sorted(keys)
    world = dist.get_world_size(group)
    out = [None] * world if dist.get_rank(group) == dst else None
    dist.gather_object(keys, out, dst=dst, group=group)
    if out is None:
        return None
    return sorted(k for part in out for k in part)


my_channels = shard_freqs


class MultiGpuIngest:
    """ctypes face of vdl2gpu_mg_* (include/vdl2gpu.h): the double-buffered fan-out of every step's chunks from rank 0.
    torch.distributed only carries the set-up blobs (NCCL unique id / CUDA IPC handles) between the processes.

    mode: "nccl" (default), "ce" (copy engines + stream memory operations, no kernel), or "auto" (ce if every rank can set
    it up, else nccl)."""

    def __init__(self, channels, rank, world, stage_bytes, mode="nccl", share_from=None):
        import ctypes as C
        import torch.distributed as dist
        from . import api
        self.g, self.rank, self.world, self.stage_bytes = channels, rank, world, int(stage_bytes)
        self.L = L = api.load_library()
        L.vdl2gpu_mg_unique_id.argtypes = [C.c_void_p, C.c_size_t]
        L.vdl2gpu_mg_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]
        L.vdl2gpu_mg_blob_bytes.restype = C.c_size_t
        L.vdl2gpu_mg_export.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.vdl2gpu_mg_import.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.vdl2gpu_mg_stage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32]
        L.vdl2gpu_mg_submit.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32]
        L.vdl2gpu_mg_destroy.argtypes = [C.c_void_p]
        self.h = C.c_void_p()
        self.mode_name = None

        def all_ok(ok):
            flags = [None] * world
            dist.all_gather_object(flags, bool(ok))
            return all(flags)

        if mode in ("auto", "ce"):
            ok = L.vdl2gpu_mg_create(channels.h, rank, world, 1, None, self.stage_bytes, C.byref(self.h)) == 0
            blob = b""
            if ok:
                n = L.vdl2gpu_mg_blob_bytes()
                buf = (C.c_uint8 * n)()
                ok = L.vdl2gpu_mg_export(self.h, buf, n) == 0
                blob = bytes(buf)
            blobs = [None] * world
            dist.all_gather_object(blobs, blob)
            if ok and all(len(b) == len(blob) and b for b in blobs):
                allb = b"".join(blobs)
                ok = L.vdl2gpu_mg_import(self.h, allb, len(allb)) == 0
            else:
                ok = False
            if all_ok(ok):
                self.mode_name = "copy-engine (CUDA IPC peer copies + stream memory operations, no kernel)"
            else:
                err = L.vdl2gpu_last_error().decode()
                if self.h:
                    L.vdl2gpu_mg_destroy(self.h)
                    self.h = C.c_void_p()
                if mode == "ce":
                    raise api.Vdl2GpuError(f"copy-engine fan-out could not be set up on every rank (rank {rank}: {err})")
        if self.mode_name is None:
            idb = (C.c_uint8 * 128)()
            if rank == 0:
                api._check(L, L.vdl2gpu_mg_unique_id(idb, 128), "vdl2gpu_mg_unique_id")
            box = [bytes(idb)]
            dist.broadcast_object_list(box, src=0)
            idb = (C.c_uint8 * 128).from_buffer_copy(box[0])
            api._check(L, L.vdl2gpu_mg_create(channels.h, rank, world, 0, idb, self.stage_bytes, C.byref(self.h)), "vdl2gpu_mg_create")
            self.mode_name = "NCCL broadcast (ncclBroadcast on the helper's own stream)"

    def stage(self, runs, src_is_host):
        """runs: [(pointer, nbytes)] on rank 0 (None elsewhere) making up one step; returns the buffer half"""
        import ctypes as C
        from . import api
        if self.rank == 0:
            n = len(runs)
            ptrs = (C.c_void_p * n)(*[r[0] for r in runs])
            sizes = (C.c_uint32 * n)(*[r[1] for r in runs])
            total = sum(r[1] for r in runs)
            return api._check(self.L, self.L.vdl2gpu_mg_stage(self.h, ptrs, sizes, n, 1 if src_is_host else 0, total), "vdl2gpu_mg_stage")
        return api._check(self.L, self.L.vdl2gpu_mg_stage(self.h, None, None, 0, 0, self.stage_bytes), "vdl2gpu_mg_stage")

    def submit_staged(self, half, n_chunks, chunk_bytes):
        from . import api
        api._check(self.L, self.L.vdl2gpu_mg_submit(self.h, half, n_chunks, chunk_bytes), "vdl2gpu_mg_submit")

    def close(self):
        if self.h:
            self.L.vdl2gpu_mg_destroy(self.h)
            self.h = None
