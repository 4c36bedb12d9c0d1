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
        # The library writes rows into `send` from its OWN non-blocking stream, which is not ordered against torch's
        # current stream: the zero-fills above must have finished before the first crRenderPoseBatch is issued.
        if self.send.is_cuda:
            torch.cuda.current_stream(self.send.device).synchronize()
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


# ---------------------------------------------------------------------------------------------
# The library's own data plane (csrc/cr_comm.cpp): ncclAllGather / grouped ncclBroadcast issued by the C++ host.
# Python only carries the 128-byte NCCL unique id from rank 0 to the other ranks.
# ---------------------------------------------------------------------------------------------
def init_library_comm(lib, rank: int, world: int, dist=None, id_path: str | None = None) -> None:
    """crCommInit on every rank.  The unique id travels over an existing torch.distributed group (`dist`, any backend:
    the bytes are broadcast as a CPU or CUDA tensor) or, without torch, through the file `id_path`."""
    import ctypes as C
    import numpy as np
    buf = np.zeros(128, np.uint8)
    if world == 1:                       # one rank gathers nothing: leave NCCL unloaded (crCommInit with one rank creates no communicator;
        if lib.crCommInit(buf.ctypes.data_as(C.c_void_p), 1, 0) != 0:    # merely asking NCCL for a unique id -- its bootstrap thread -- made a
            raise RuntimeError("crCommInit failed")                      # long single-GPU job 40 % slower: profiles/r03l vs r03k)
        return
    if rank == 0:
        if lib.crCommGetUniqueId(buf.ctypes.data_as(C.c_void_p)) != 0:
            raise RuntimeError("crCommGetUniqueId failed (NCCL not loadable?)")
    if world > 1:
        if dist is not None:
            import torch
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            t = torch.from_numpy(buf).to(dev)
            dist.broadcast(t, src=0)
            buf = t.cpu().numpy().copy()
        elif id_path is not None:
            import os
            import time
            if rank == 0:
                with open(id_path + ".tmp", "wb") as f:
                    f.write(buf.tobytes())
                os.replace(id_path + ".tmp", id_path)
            else:
                for _ in range(6000):
                    if os.path.exists(id_path):
                        break
                    time.sleep(0.01)
                buf = np.frombuffer(open(id_path, "rb").read(), np.uint8).copy()
        else:
            raise ValueError("world > 1 needs a torch.distributed group or an id file")
    if lib.crCommInit(buf.ctypes.data_as(C.c_void_p), int(world), int(rank)) != 0:
        raise RuntimeError("crCommInit failed")


def render_pose_batch_sharded(lib, poses, chunk: int = 0, first_frame: int = 0, out_device_ptr=None):
    """crRenderPoseBatchSharded: every rank passes the same [P][12] poses and gets all P rows (uint8[P][N][4]) back."""
    import ctypes as C
    import numpy as np
    poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, 12)
    n = lib.getCurrentEyeOmmatidialCount()
    if out_device_ptr is not None:
        ms = lib.crRenderPoseBatchSharded(poses.ctypes.data_as(C.c_void_p), len(poses), None, C.c_void_p(out_device_ptr), int(chunk), int(first_frame))
        return None, ms
    out = np.zeros((len(poses), n, 4), dtype=np.uint8)
    ms = lib.crRenderPoseBatchSharded(poses.ctypes.data_as(C.c_void_p), len(poses), out.ctypes.data_as(C.c_void_p), None, int(chunk), int(first_frame))
    if ms < 0:
        raise RuntimeError("crRenderPoseBatchSharded failed")
    return out, ms
