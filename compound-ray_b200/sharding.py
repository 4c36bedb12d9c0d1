"""Pose sharding and ommatidium-range sharding across GPUs -- host-side logic only, no device code.

The path shards by camera pose: rank r of R renders the contiguous pose block
[lo, hi) of a P-pose run; scene, BVH and eye are replicated per GPU; the only exchange is the
final allgather of the per-pose uchar4 rows.  Results must not depend on the sharding: the
reference semantics are "pose k is frame k of every sample stream", so a rank whose block starts at
pose lo positions its RNG streams at frame lo (crSetFirstFrame) before rendering.
"""
from __future__ import annotations


def pose_block(rank: int, world: int, n_poses: int) -> tuple[int, int]:
    """Contiguous block of poses for `rank`; blocks differ by at most one pose."""
    base, extra = divmod(n_poses, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def padded_block_size(world: int, n_poses: int) -> int:
    """Rows per rank in the allgather buffer (equal counts; short blocks are padded)."""
    return -(-n_poses // world)


def draws_before_frame(k: int) -> int:
    """Raw XORWOW draws one stream has consumed before frame k (3 on even frames, 1 on odd)."""
    return 3 * ((k + 1) // 2) + (k // 2)


def gathered_row(rank: int, local_index: int, world: int, n_poses: int) -> int:
    """Row of pose (rank, local_index) in the padded allgather result."""
    return rank * padded_block_size(world, n_poses) + local_index


def unpad(gathered, world: int, n_poses: int):
    """Drop the padding rows of an allgathered [world*block, ...] array -> [n_poses, ...]."""
    blk = padded_block_size(world, n_poses)
    parts = []
    for r in range(world):
        lo, hi = pose_block(r, world, n_poses)
        parts.append(gathered[r * blk: r * blk + (hi - lo)])
    import numpy as np
    if isinstance(gathered, np.ndarray):
        return np.concatenate(parts, axis=0)
    import torch
    return torch.cat(parts, dim=0)


class ChunkedPoseGather:
    """The allgather of a pose-sharded run, issued in chunks while the run is still rendering.

    Every rank renders its pose block chunk by chunk (crRenderPoseBatch writes the rows of chunk c straight into
    `send_rows(c)`, a slice of this rank's send block) and calls `issue(c)` after each: one asynchronous allgather
    per chunk, on the process group's own stream, whose output views ARE the final [rank][row] places -- no staging
    and no reshuffle.  The collective of chunk c runs while chunk c+1 is traced; `finish()` waits for all of them and
    returns the rows in pose order.  Chunks are cut on the PADDED block, so every rank issues the same collectives
    whatever its share of the poses.  Consecutive crRenderPoseBatch calls continue the sample streams, so chunking
    does not change a byte of the result."""

    def __init__(self, rank: int, world: int, n_poses: int, row_shape, dtype, device, chunk: int, dist=None):
        import torch
        if dist is None:
            import torch.distributed as dist
        self.rank, self.world, self.n_poses, self.dist = rank, world, n_poses, dist
        self.blk = padded_block_size(world, n_poses)
        self.lo, self.hi = pose_block(rank, world, n_poses)
        row_shape = tuple(row_shape)
        self.send = torch.zeros((self.blk,) + row_shape, dtype=dtype, device=device)
        self.gathered = torch.zeros((world, self.blk) + row_shape, dtype=dtype, device=device)
        chunk = max(1, int(chunk))
        self.chunks = [(c0, min(c0 + chunk, self.blk)) for c0 in range(0, self.blk, chunk)]
        self.handles = []

    def local_poses(self, c: int) -> tuple[int, int]:
        """Local pose indices [a, b) of this rank that fall into chunk c (empty for pure padding)."""
        c0, c1 = self.chunks[c]
        n = self.hi - self.lo
        return min(c0, n), min(c1, n)

    def send_rows(self, c: int):
        """Slice of the send block that receives the rows of chunk c."""
        c0, c1 = self.chunks[c]
        return self.send[c0:c1]

    def issue(self, c: int):
        c0, c1 = self.chunks[c]
        if self.world == 1:
            self.gathered[0, c0:c1] = self.send[c0:c1]
            return
        outs = [self.gathered[r, c0:c1] for r in range(self.world)]          # contiguous views of the final buffer
        self.handles.append(self.dist.all_gather(outs, self.send[c0:c1], async_op=True))

    def finish(self):
        for h in self.handles:
            h.wait()
        self.handles = []
        flat = self.gathered.view((self.world * self.blk,) + tuple(self.gathered.shape[2:]))
        return unpad(flat, self.world, self.n_poses)


# ---------------------------------------------------------------------------------------------
# Secondary partition: one pose, large N*S -- rank r owns the ommatidium rows [lo, hi) of the eye.
# Stream ids keep the GLOBAL indices (crSetOmmatidialShard), so the gathered per-ommatidium RGB
# equals the unsharded frame bit for bit; projection runs after the gather.
# ---------------------------------------------------------------------------------------------
def ommatidia_block(rank: int, world: int, n_ommatidia: int) -> tuple[int, int]:
    """Contiguous block of ommatidium rows for `rank` (same rule as pose_block)."""
    return pose_block(rank, world, n_ommatidia)


def configure_ommatidia_shard(lib, er, ommatidia, rank: int, world: int):
    """Give this rank its rows of `ommatidia` (float32[N][8]) and declare the shard to the library.
    Returns (lo, hi).  Call setCurrentEyeSamplesPerOmmatidium before or after; both reset the streams."""
    lo, hi = ommatidia_block(rank, world, len(ommatidia))
    er.setOmmatidiaFromArray(lib, ommatidia[lo:hi])
    lib.crSetOmmatidialShard(len(ommatidia), lo)
    return lo, hi


def allgather_rows(local_rows, world: int, n_total: int, dist=None):
    """Allgather equally padded per-rank row blocks ([rows, ...] torch tensor) -> [n_total, ...]."""
    import torch
    blk = padded_block_size(world, n_total)
    send = torch.zeros((blk,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    send[: local_rows.shape[0]] = local_rows
    if world == 1:
        return send[:n_total]
    if dist is None:
        import torch.distributed as dist
    gathered = torch.zeros((world * blk,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(gathered.view(-1), send.view(-1))
    return unpad(gathered, world, n_total)
