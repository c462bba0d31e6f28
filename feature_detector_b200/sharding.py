"""A batch of frames sharded across the GPUs of a box (BASELINE.json configs[1], [2], [4]; SURVEY.md 8e, first row).

Frames are independent -- nothing crosses frames once each has its own candidate and keypoint slots -- so the batch is cut
into contiguous blocks, one per rank, and the data path needs no collective at all.  The only exchange is the OPTIONAL final
gather of the results, which are fixed-capacity slots per frame (``kp_capacity`` x 16-byte keypoints, ``kp_capacity`` x 32-byte
descriptors, one count), hence one ``all_gather`` over equal-sized blocks (the last ranks' blocks are padded by at most one
frame).  Plumbing only:

* ``frame_range``            -- the block of frames a rank owns;
* ``gather_frame_slots``     -- all_gather of one per-frame slot array (torch tensors: CUDA over NCCL, CPU over gloo in the tests);
* ``gather_frame_records``   -- several slot arrays packed into one byte record per frame: one all_gather for all of them;
* ``detect_sharded``         -- upload / bind this rank's block, fd_detect (+ fd_describe_selected), optional gather straight
                                from the context's device buffers (no host round trip before the collective).
"""
from __future__ import annotations

import numpy as np

from . import KEYPOINT_DTYPE


def frame_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Frames [lo, hi) of rank ``rank``: contiguous blocks whose sizes differ by at most one (the first ``n % world`` ranks
    hold the longer ones; trailing ranks are empty when there are more ranks than frames)."""
    base, rem = divmod(int(n_frames), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_frame_slots(local, n_frames: int, rank: int, world: int, dst: int | None = None, group=None):
    """``local``: this rank's per-frame slots, a torch tensor whose first dimension is its block of frames (possibly empty).
    Returns the (n_frames, ...) tensor of all ranks' slots in frame order -- on every rank, or only on ``dst`` (None elsewhere).
    One all_gather over blocks padded to the longest block."""
    import torch
    import torch.distributed as dist
    lo, hi = frame_range(n_frames, rank, world)
    assert local.shape[0] == hi - lo, (tuple(local.shape), lo, hi)
    if world == 1:
        return local
    longest = frame_range(n_frames, 0, world)[1]
    block = torch.zeros((max(longest, 1),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    block[:hi - lo] = local
    out = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(out, block, group=group)
    if dst is not None and rank != dst:
        return None
    parts = []
    for r in range(world):
        a, b = frame_range(n_frames, r, world)
        parts.append(out[r][:b - a])
    return torch.cat(parts)


def gather_frame_records(tensors, n_frames: int, rank: int, world: int, dst: int | None = None, group=None):
    """Several per-frame slot arrays at once (keypoints, counts, descriptors ...): the arrays of one frame are packed into one
    byte record, so the whole exchange is ONE all_gather (``all_gather_into_tensor``: no per-rank output copies) whatever the
    number of arrays.  ``tensors``: torch tensors with this rank's block of frames as first dimension.  Returns the list of
    (n_frames, ...) tensors in frame order -- on every rank, or only on ``dst`` (None elsewhere)."""
    import torch
    import torch.distributed as dist
    lo, hi = frame_range(n_frames, rank, world)
    n_local = hi - lo
    assert all(t.shape[0] == n_local for t in tensors), ([tuple(t.shape) for t in tensors], lo, hi)
    if world == 1:
        return list(tensors)
    widths = [int(np.prod(t.shape[1:], dtype=np.int64)) * t.element_size() for t in tensors]   # bytes per frame of each array
    longest = max(frame_range(n_frames, 0, world)[1], 1)
    record = torch.zeros((longest, sum(widths)), dtype=torch.uint8, device=tensors[0].device)
    col = 0
    for t, w in zip(tensors, widths):
        if n_local:
            record[:n_local, col:col + w] = t.contiguous().reshape(n_local, -1).view(torch.uint8)
        col += w
    out = torch.empty((world * longest, sum(widths)), dtype=torch.uint8, device=record.device)
    dist.all_gather_into_tensor(out, record, group=group)
    if dst is not None and rank != dst:
        return None
    blocks = []
    for r in range(world):
        a, b = frame_range(n_frames, r, world)
        blocks.append(out[r * longest:r * longest + (b - a)])
    rows = torch.cat(blocks)                                            # (n_frames, record bytes), frame order
    result, col = [], 0
    for t, w in zip(tensors, widths):
        part = rows[:, col:col + w].contiguous()
        result.append(part.view(t.dtype).reshape((n_frames,) + tuple(t.shape[1:])))
        col += w
    return result


class _DeviceArray:
    """Lets torch view context-owned device memory without a copy (``torch.as_tensor`` reads __cuda_array_interface__)."""

    def __init__(self, ptr: int, shape: tuple, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_results(ctx, n_local: int, device, with_descriptors: bool):
    """Torch views (no copy) of the context's keypoint slots, counts and descriptors after fd_detect / fd_describe_selected:
    keypoints (n_local, capacity, 4) float32 as (x, y, response, 0), counts (n_local,) int32, descriptors (n_local, capacity, 32)
    uint8 or None.  Valid until the context's next detect call."""
    import torch
    kp_ptr, cnt_ptr, cap = ctx.device_keypoints()
    if n_local == 0:
        return (torch.zeros((0, cap, 4), dtype=torch.float32, device=device), torch.zeros((0,), dtype=torch.int32, device=device),
                torch.zeros((0, cap, 32), dtype=torch.uint8, device=device) if with_descriptors else None)
    kp = torch.as_tensor(_DeviceArray(kp_ptr, (n_local, cap, 4), "<f4"), device=device)
    cnt = torch.as_tensor(_DeviceArray(cnt_ptr, (n_local,), "<i4"), device=device)
    desc = None
    if with_descriptors:
        import ctypes as C
        d_ptr, d_cap = C.c_void_p(), C.c_int(0)
        ctx._ck(ctx._lib.fd_device_descriptors(ctx._h, C.byref(d_ptr), C.byref(d_cap)))
        assert d_cap.value == cap
        desc = torch.as_tensor(_DeviceArray(d_ptr.value, (n_local, cap, 32), "|u1"), device=device)
    return kp, cnt, desc


def detect_sharded(ctx, frames, n_frames: int, rank: int, world: int, detect, brief=None, gather: bool = True, dst: int | None = None,
                   group=None, cand_capacity: int = 0):
    """This rank's block of a batch of ``n_frames`` frames: ``frames`` is its (hi - lo, rows, cols) uint8 block -- a CUDA tensor
    (bound in place) or a host array (uploaded).  Runs fd_detect (+ fd_describe_selected) and, if ``gather``, collects all ranks'
    slots.  Returns dict(keypoints (n, cap) KEYPOINT_DTYPE, counts (n,) int32, descriptors (n, cap, 32) uint8 or None) as host
    arrays -- for the whole batch after a gather (None on ranks other than ``dst``), else for this rank's block."""
    import torch
    lo, hi = frame_range(n_frames, rank, world)
    n_local = hi - lo
    assert len(frames) == n_local
    device = torch.device("cuda", torch.cuda.current_device())
    if n_local:
        if isinstance(frames, torch.Tensor) and frames.is_cuda:
            frames = frames.contiguous()
            torch.cuda.current_stream().synchronize()      # the context runs on its own stream
            ctx.bind_device(frames.data_ptr(), frames.shape[1], frames.shape[2], n_local)
        else:
            ctx.upload(np.ascontiguousarray(frames, np.uint8))
        ctx.detect(detect, cand_capacity)
        if brief is not None:
            ctx.describe_selected(brief)
        ctx.sync()
    kp, cnt, desc = device_results(ctx, n_local, device, brief is not None) if n_local else (None, None, None)
    if n_local == 0:   # an empty block still takes part in the gather, with slots of the capacity the other ranks use
        cap = max(int(detect.needed_feature_num), 1)
        kp = torch.zeros((0, cap, 4), dtype=torch.float32, device=device)
        cnt = torch.zeros((0,), dtype=torch.int32, device=device)
        desc = torch.zeros((0, cap, 32), dtype=torch.uint8, device=device) if brief is not None else None
    if gather:
        got = gather_frame_records([kp, cnt] + ([desc] if desc is not None else []), n_frames, rank, world, dst, group)
        if got is None:
            return None
        kp, cnt = got[0], got[1]
        desc = got[2] if desc is not None else None
    out_kp = np.ascontiguousarray(kp.cpu().numpy()).view(np.float32)
    rec = np.zeros(out_kp.shape[:2], KEYPOINT_DTYPE)
    rec["x"], rec["y"], rec["response"] = out_kp[..., 0], out_kp[..., 1], out_kp[..., 2]
    return {"keypoints": rec, "counts": cnt.cpu().numpy(), "descriptors": None if desc is None else desc.cpu().numpy()}
