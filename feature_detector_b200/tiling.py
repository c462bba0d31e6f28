"""One large frame, row-tiled across the GPUs of a box (BASELINE.json configs[3]).

Dense stages shard by rows with a 3-row halo (gradient 1 + box window 1 + NMS 1 for Harris / Shi-Tomasi; the radius-3
ring for FAST); the greedy selection (feature_point_detector.cpp:54-74) is global per frame, so the tiles' candidate
keys are gathered to one rank, which selects.  Plumbing only:

* ``plan_tiles``       -- contiguous row ranges and their halo buffers;
* ``exchange_halos``   -- each rank owns its rows; the halo rows travel rank <-> rank +/- 1 as point-to-point sends
                          (NCCL over NVLink on GPUs, gloo on CPU tensors in the tests) -- 3 x cols bytes per seam;
* ``gather_keys``      -- counts first, then the padded key payload (the only collective on the path);
* ``detect_tiled``     -- the multi-rank driver; ``detect_tiled_local`` runs the same tiles one after another on ONE
                          GPU (the seam logic is identical), which is what the single-GPU test-suite checks.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

HALO = 3


@dataclass(frozen=True)
class Tile:
    own_lo: int   # absolute rows [own_lo, own_hi) belong to this tile
    own_hi: int
    buf_lo: int   # absolute rows [buf_lo, buf_hi) are resident (own rows + halos, clipped to the frame)
    buf_hi: int

    @property
    def row_offset(self) -> int:
        return self.buf_lo

    @property
    def own_first(self) -> int:
        return self.own_lo - self.buf_lo

    @property
    def own_count(self) -> int:
        return self.own_hi - self.own_lo

    @property
    def buf_rows(self) -> int:
        return self.buf_hi - self.buf_lo


def plan_tiles(rows: int, n_tiles: int, halo: int = HALO) -> list[Tile]:
    """Split ``rows`` into ``n_tiles`` contiguous blocks (sizes differ by at most one row; trailing tiles may be empty
    when there are more tiles than rows)."""
    base, rem = divmod(rows, n_tiles)
    tiles, lo = [], 0
    for t in range(n_tiles):
        hi = lo + base + (1 if t < rem else 0)
        tiles.append(Tile(lo, hi, max(0, lo - halo), min(rows, hi + halo)) if hi > lo else Tile(lo, lo, lo, lo))
        lo = hi
    return tiles


def exchange_halos(owned, rank: int, world: int, tiles: list[Tile], group=None):
    """``owned``: this rank's own rows, a (own_count, cols) uint8 torch tensor.  Returns the (buf_rows, cols) tile buffer
    with the neighbours' halo rows in place.  Point-to-point sends only; a seam moves 2 x halo x cols bytes."""
    import torch
    import torch.distributed as dist
    me = tiles[rank]
    buf = torch.empty((me.buf_rows, owned.shape[1]), dtype=owned.dtype, device=owned.device)
    buf[me.own_first:me.own_first + me.own_count] = owned
    ops, keep = [], []
    for other in range(world):
        if other == rank or tiles[other].own_count == 0 or me.own_count == 0:
            continue
        o = tiles[other]
        # rows of mine that the other tile's buffer needs, and rows of theirs that mine needs
        s_lo, s_hi = max(me.own_lo, o.buf_lo), min(me.own_hi, o.buf_hi)
        r_lo, r_hi = max(o.own_lo, me.buf_lo), min(o.own_hi, me.buf_hi)
        if s_hi > s_lo:
            t = owned[s_lo - me.own_lo:s_hi - me.own_lo].contiguous()
            keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, other, group=group))
        if r_hi > r_lo:
            ops.append(dist.P2POp(dist.irecv, buf[r_lo - me.buf_lo:r_hi - me.buf_lo], other, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return buf


def gather_keys(keys, rank: int, world: int, dst: int = 0, group=None):
    """``keys``: this rank's candidate keys, a 1-D int64 torch tensor.  On ``dst`` returns all ranks' keys concatenated
    (rank order), elsewhere None.  Counts travel first, then one padded all_gather of the payload."""
    import torch
    import torch.distributed as dist
    n = torch.tensor([keys.numel()], dtype=torch.int64, device=keys.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    width = max(max(counts), 1)
    padded = torch.zeros(width, dtype=torch.int64, device=keys.device)
    padded[:keys.numel()] = keys
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    if rank != dst:
        return None
    return torch.cat([out[r][:counts[r]] for r in range(world)])


def _tile_candidates(ctx, buf, tile: Tile, full_rows: int, params, cand_capacity: int = 0):
    """Candidate keys (int64 torch tensor on buf's device) of one tile buffer."""
    import torch
    cols = buf.shape[1]
    if tile.own_count == 0:
        return torch.zeros(0, dtype=torch.int64, device=buf.device)
    ctx.bind_device(buf.data_ptr(), tile.buf_rows, cols, 1)
    ctx.set_tile(tile.row_offset, tile.own_first, tile.own_count, full_rows)
    try:
        ctx.compute_candidates(params, cand_capacity)
        n = int(ctx.candidate_counts()[0])
        keys = torch.empty(max(n, 1), dtype=torch.int64, device=buf.device)
        ctx.export_candidates(keys.data_ptr(), keys.numel())
    finally:
        ctx.set_tile(0, 0, 0, 0)
    return keys[:n]


def _select(ctx, keys, full_rows: int, cols: int, params):
    import torch
    counts = torch.tensor([keys.numel()], dtype=torch.int32, device=keys.device)
    work = keys.clone() if keys.numel() else torch.zeros(1, dtype=torch.int64, device=keys.device)
    ctx.select_candidates(params, work.data_ptr(), counts.data_ptr(), max(int(keys.numel()), 1), full_rows, cols, 1)
    kp, cnt = ctx.keypoints(max(int(params.needed_feature_num), 1))
    return kp[0, :cnt[0]].copy()


def detect_tiled_local(ctx, frame, n_tiles: int, params, halo: int = HALO):
    """All tiles of ``frame`` (a (rows, cols) uint8 CUDA tensor) on one GPU, one after another.  Returns
    (keypoints structured array, all candidate keys as an int64 tensor)."""
    import torch
    rows, cols = frame.shape
    keys = []
    for t in plan_tiles(rows, n_tiles, halo):
        if t.own_count:
            keys.append(_tile_candidates(ctx, frame[t.buf_lo:t.buf_hi].contiguous(), t, rows, params))
    allk = torch.cat(keys) if keys else torch.zeros(0, dtype=torch.int64, device=frame.device)
    return _select(ctx, allk, rows, cols, params), allk


def detect_tiled(ctx, owned, full_rows: int, params, rank: int, world: int, dst: int = 0, group=None, halo: int = HALO):
    """One frame sharded by rows over ``world`` ranks: ``owned`` is this rank's (own_count, cols) uint8 CUDA tensor.
    Halo exchange -> candidates per tile -> key gather -> selection on ``dst``.  Returns the keypoints there, else None."""
    import torch
    tiles = plan_tiles(full_rows, world, halo)
    buf = exchange_halos(owned, rank, world, tiles, group)
    torch.cuda.current_stream().synchronize()   # the context runs on its own stream
    keys = _tile_candidates(ctx, buf, tiles[rank], full_rows, params)
    allk = gather_keys(keys, rank, world, dst, group)
    if rank != dst:
        return None
    torch.cuda.current_stream().synchronize()
    return _select(ctx, allk, full_rows, owned.shape[1], params)


# ---- the key format, for callers that build or read keys on the host (include/fd_b200.h, fd_set_tile) -------------
def make_keys(response: np.ndarray, rows: np.ndarray, cols: np.ndarray) -> np.ndarray:
    """64-bit candidate keys: high word = ~ordered(response), low word = row << 16 | col (ascending = best first)."""
    b = np.ascontiguousarray(response, np.float32).view(np.uint32).astype(np.uint64)
    ordered = np.where(b & 0x80000000, (~b) & 0xFFFFFFFF, b | 0x80000000)
    hi = (~ordered) & 0xFFFFFFFF
    return ((hi << np.uint64(32)) | (np.asarray(rows, np.uint64) << np.uint64(16)) | np.asarray(cols, np.uint64)).astype(np.uint64)


def split_keys(keys: np.ndarray):
    """Inverse of make_keys: (response float32, rows, cols)."""
    k = np.asarray(keys).astype(np.uint64)
    ordered = (~(k >> np.uint64(32))) & np.uint64(0xFFFFFFFF)
    b = np.where(ordered & np.uint64(0x80000000), ordered & np.uint64(0x7FFFFFFF), (~ordered) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return b.view(np.float32), ((k >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int32), (k & np.uint64(0xFFFF)).astype(np.int32)
