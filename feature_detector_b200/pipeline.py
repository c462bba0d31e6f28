"""Host-resident batches through the C ABI with transfers hidden behind compute.

A batch that starts in host memory is bound by the PCIe copy (370 MB for 1024 frames of 752x480 against ~1 ms of
kernels), so the batch is cut into chunks that alternate between two contexts -- each context owns a stream and its
device buffers: while chunk i is being uploaded and queued on one stream, chunk i-1 computes on the other and its
keypoints / descriptors are copied back.  Results are identical to one fd_detect over the whole batch because frames
are independent.  Plumbing only: every pixel is still touched by the CUDA kernels alone.
"""
from __future__ import annotations

import numpy as np

from . import BriefParams, Context, DetectParams, KEYPOINT_DTYPE


class HostPipeline:
    """detect (+ describe) for frames that live in (preferably pinned) host memory."""

    def __init__(self, device: int = 0, chunk_frames: int = 128, n_contexts: int = 2):
        self.chunk_frames = int(chunk_frames)
        self.contexts = [Context(device) for _ in range(max(2, int(n_contexts)))]

    def close(self):
        for c in self.contexts:
            c.close()
        self.contexts = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def launch_count(self) -> int:
        return sum(c.launch_count for c in self.contexts)

    def run(self, host_ptr: int, rows: int, cols: int, n_frames: int, detect: DetectParams, brief: BriefParams | None,
            out_kp: np.ndarray, out_counts: np.ndarray, out_desc: np.ndarray | None = None, cand_capacity: int = 0):
        """host_ptr: address of n_frames contiguous rows x cols uint8 frames.  out_kp (n_frames, cap) KEYPOINT_DTYPE,
        out_counts (n_frames,) int32, out_desc (n_frames, cap, 32) uint8 -- caller-owned (pin them for full speed)."""
        assert out_kp.shape[0] == n_frames and out_counts.shape[0] == n_frames and out_kp.dtype == KEYPOINT_DTYPE
        frame_bytes = rows * cols
        chunks = [(s, min(s + self.chunk_frames, n_frames)) for s in range(0, n_frames, self.chunk_frames)]

        def fetch(ci):
            s, e = chunks[ci]
            c = self.contexts[ci % len(self.contexts)]
            c.keypoints_into(out_kp[s:e], out_counts[s:e])          # blocks on that context's stream only
            if brief is not None and out_desc is not None:
                c.descriptors_into(out_desc[s:e])

        depth = len(self.contexts) - 1                                # chunks in flight behind the one being queued
        for ci, (s, e) in enumerate(chunks):
            c = self.contexts[ci % len(self.contexts)]
            c.upload_ptr(host_ptr + s * frame_bytes, rows, cols, e - s)  # async when the source is pinned
            c.detect(detect, cand_capacity)
            if brief is not None:
                c.describe_selected(brief)
            if ci >= depth:
                fetch(ci - depth)
        for ci in range(max(0, len(chunks) - depth), len(chunks)):
            fetch(ci)
