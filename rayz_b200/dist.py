"""One-process-per-GPU plumbing (torch.distributed): which rows a rank owns and the slab gather.

The render itself needs no collective: pixels are independent and the random stream is keyed by the
global pixel index, so each rank renders its own band-interleaved rows (RzRenderParams.shard_*).
The only exchange step of the path is the gather of the finished slabs to rank 0
(BASELINE.json north_star: "slabs are gathered to GPU0 via NCCL gather or P2P copies over NVLink").
Works with any backend: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_row_indices(height: int, shard_index: int, shard_count: int, band_rows: int = 4) -> list[int]:
    """Global rows owned by a shard: row j belongs to shard (j // band_rows) % shard_count (include/rayz_cuda.h)."""
    if shard_count <= 1:
        return list(range(height))
    band_rows = band_rows or 4
    return [j for j in range(height) if (j // band_rows) % shard_count == shard_index]


class SlabGather:
    """Pre-allocates rank 0's receive slabs and row maps; `run(slab)` gathers and interleaves one frame."""

    def __init__(self, height: int, row_shape: tuple, dtype: torch.dtype, device, band_rows: int = 4, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.height, self.band = height, band_rows or 4
        self.rows = [shard_row_indices(height, r, self.world, self.band) for r in range(self.world)]
        self.final = None
        if self.rank == 0:
            self.final = torch.empty((height,) + tuple(row_shape), dtype=dtype, device=device)
            self.slabs = [None] + [torch.empty((len(self.rows[r]),) + tuple(row_shape), dtype=dtype, device=device)
                                   for r in range(1, self.world)]
            self.index = [torch.tensor(self.rows[r], dtype=torch.long, device=device) for r in range(self.world)]

    def my_rows(self) -> int:
        return len(self.rows[self.rank])

    def run(self, slab: torch.Tensor):
        """slab: this rank's compact rows [my_rows, *row_shape]. Returns the full frame on rank 0, None elsewhere."""
        assert slab.shape[0] == self.my_rows(), (slab.shape, self.my_rows())
        if self.world == 1:
            self.final.copy_(slab)
            return self.final
        if self.rank == 0:
            ops = [dist.P2POp(dist.irecv, self.slabs[r], r, self.group) for r in range(1, self.world) if len(self.rows[r])]
            reqs = dist.batch_isend_irecv(ops) if ops else []
            self.final[self.index[0]] = slab
            for q in reqs:
                q.wait()
            for r in range(1, self.world):
                if len(self.rows[r]):
                    self.final[self.index[r]] = self.slabs[r]
            return self.final
        if self.my_rows():
            for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, slab.contiguous(), 0, self.group)]):
                q.wait()
        return None
