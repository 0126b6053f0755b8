"""Multi-GPU plumbing of the sweep (north star item 4): one process per GPU, tracks sharded by
contiguous range, ONE all-reduce of the tally deltas at the end.  torch.distributed is plumbing
only; the per-rank compute is the CUDA kernel behind the C ABI.

The reference has no multi-device path (only -d <id> -> cudaSetDevice,
/root/reference/src/cuda/io.cu:148-158); SURVEY.md section 8(e) defines this one:
  * every rank holds a full replica of fine_source / sigT (21 MB) and zero-initialised tallies
  * rank k sweeps tracks [k*T/P, (k+1)*T/P)
  * all_reduce(sum) over the padded tally array, then flux = flux0 + tallies
"""
from __future__ import annotations


def shard_tracks(n_tracks: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, exhaustive, non-overlapping track range of `rank` (sizes differ by <= 1)."""
    if world < 1 or not (0 <= rank < world) or n_tracks < 0:
        raise ValueError("bad shard request")
    return rank * n_tracks // world, (rank + 1) * n_tracks // world


def shard_segments(n_tracks: int, seg_per_track: int, segments: int, rank: int, world: int) -> int:
    """Number of segments in the shard (the last track of the stream may be short)."""
    tb, te = shard_tracks(n_tracks, rank, world)
    return max(0, min(te * seg_per_track, segments) - tb * seg_per_track)


class DevicePointer:
    """Expose a raw device pointer (e.g. smk_device_tally) to torch through
    __cuda_array_interface__, so NCCL reduces the library's buffer in place."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4",
                                         "data": (ptr, False), "version": 3}


def all_reduce_tallies(tally, group=None):
    """In-place sum of the per-rank tally deltas (a torch tensor on any backend)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tally, op=dist.ReduceOp.SUM, group=group)
    return tally


# ---------------------------------------------------------------------------------------------
# end-to-end path with the host<->device traffic split over the ranks: every rank uploads 1/P of
# the rows, the replicas are completed over NVLink, the tallies are reduced slice-wise to their
# owners and every rank reads back 1/P of the flux.  Total PCIe traffic = 1x the data instead of Px.
# ---------------------------------------------------------------------------------------------
def shard_rows(n_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous row range of `rank` (same partition rule as shard_tracks)."""
    return shard_tracks(n_rows, rank, world)


def gather_row_slices(full, n_rows: int, row_elems: int, world: int, group=None):
    """all-gather with ragged slices: rank k broadcasts rows shard_rows(n_rows, k, world) of the flat
    tensor `full` (n_rows * row_elems elements) to everybody, in place."""
    import torch.distributed as dist
    works = []
    for k in range(world):
        b, e = shard_rows(n_rows, k, world)
        if e > b:
            works.append(dist.broadcast(full[b * row_elems:e * row_elems], src=k, group=group, async_op=True))
    for w in works:
        w.wait()


def reduce_row_slices(full, n_rows: int, row_elems: int, world: int, group=None):
    """reduce-scatter with ragged, row-aligned slices: rows shard_rows(n_rows, k, world) of `full` are
    summed onto rank k, in place (the other rows of a rank's tensor are left partially reduced)."""
    import torch.distributed as dist
    works = []
    for k in range(world):
        b, e = shard_rows(n_rows, k, world)
        if e > b:
            works.append(dist.reduce(full[b * row_elems:e * row_elems], dst=k, op=dist.ReduceOp.SUM, group=group,
                                     async_op=True))
    for w in works:
        w.wait()
