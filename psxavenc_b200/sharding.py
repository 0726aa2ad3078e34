"""Multi-GPU partitioning of the encode path.

Video frames are independent (every per-frame state is reset, mdec.c:676-686), so a batch is
split into contiguous frame ranges — rank r owns [r*N/G, (r+1)*N/G) — and the output order is
the concatenation of the ranks' outputs. No data-path collective is needed during the
encode; the only exchange is the gather of the per-frame results {bytes_used, blocks_used,
quant_scale, uncomp_hwords_used} to the rank that muxes sectors.

An ADPCM channel is a strictly sequential chain (adpcm.c:135-136,186-190): it cannot be cut
in time, so whole chains (channels / streams) are dealt round-robin to ranks.

One process per GPU, torch.distributed for the plumbing (NCCL on GPUs; gloo in the CPU tests).
"""
import numpy as np


def frame_range(n_frames, rank, world):
    """Contiguous [first, last) frame range of `rank`; ranges differ by at most one frame."""
    base, extra = divmod(n_frames, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def frame_budgets(n_frames, sectors_num, sectors_den, first_frame=0):
    """Per-frame byte budgets of encode_sector_str (mdec.c:772-774): frame k (0-based) gets
    2016 * (floor((k+1)*num/den) - floor(k*num/den)) bytes — a closed form of the running
    overflow accumulator, so any rank can compute the budgets of its own frame range."""
    k = np.arange(first_frame, first_frame + n_frames, dtype=np.int64)
    return (2016 * ((k + 1) * sectors_num // sectors_den - k * sectors_num // sectors_den)).astype(np.int32)


def stream_owner(stream, world):
    """Rank that encodes ADPCM chain `stream`."""
    return stream % world


def streams_of(rank, n_streams, world):
    return list(range(rank, n_streams, world))


def gather_results(local_results, n_frames, dist=None):
    """All-gather of the per-frame result rows (int32 [n_local, 4] torch tensor) into frame
    order [n_frames, 4] on every rank. With dist None (single process) returns the input."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_results
    import torch
    world = dist.get_world_size()
    counts = [frame_range(n_frames, r, world) for r in range(world)]
    longest = max(b - a for a, b in counts)
    padded = torch.zeros((longest, 4), dtype=torch.int32, device=local_results.device)
    padded[:local_results.shape[0]] = local_results
    out = torch.empty((world * longest, 4), dtype=torch.int32, device=local_results.device)
    dist.all_gather_into_tensor(out, padded)
    rows = [out[r * longest:r * longest + (b - a)] for r, (a, b) in enumerate(counts)]
    return torch.cat(rows)
