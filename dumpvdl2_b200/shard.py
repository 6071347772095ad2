"""Multi-GPU plumbing for the path: channels are independent (reference: one pthread per channel, shared
read-only sample buffer - src/dumpvdl2.c:117-135, src/demod.c:300-301), so the path shards by channel.

  * channel k of the global list lives on rank k mod world (round-robin, BASELINE.json north_star);
  * the rank that owns the SDR / file (`src`) broadcasts every raw IQ chunk - the one collective on the path
    (NCCL over NVLink on GPUs; the same code runs on gloo for the CPU tests);
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
