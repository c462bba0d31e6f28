#!/usr/bin/env python
"""Phase timeline of one frame's selection (GPU box, debug build: make -C feature_detector_b200/csrc clean all EXTRA=-DFD_SELECT_TRACE):
clock64 stamps of frame 0's CTA at the phase boundaries of select_kernel.  Tags: 1 entry, 2 cell state ready, 40 .. 45 inside pass 1
(iteration start, key loaded, kill test done, minimum posted, pushed, loop done), 10 / 11 / 12 after pass 1 / winners / wipe of a
round, 30 rounds done, 31 ordered, 32 written.  profiles/r2_select_one_frame_timeline.json holds the round-2 timelines."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_b200 as fd  # noqa: E402
from bench import make_frames  # noqa: E402

one = make_frames(1, 0, 752, 480)
det = fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=int(os.environ.get("FD_EXP_KN", "9")))
ctx = fd.Context(0)
ctx.upload(one)
lib = ctx._lib
buf = (C.c_longlong * 512)()
for rep in range(3):
    ctx.detect(det, 65536)
    ctx.sync()
    n = lib.fd_debug_select_trace(buf, 256)
st = [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]
t0 = st[0][0]
print(json.dumps({"candidates": int(ctx.candidate_counts()[0]), "keypoints": int(ctx.keypoint_counts()[0]),
                  "timeline_cycles": [[int(tag), int(t - t0)] for t, tag in st]}))
