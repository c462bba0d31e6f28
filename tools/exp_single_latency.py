#!/usr/bin/env python
"""Where one frame's host-to-host latency goes (GPU box): every piece of the single-frame path timed alone with the wall clock
(call + stream synchronisation), median of 300 after 30 warm-up calls.  The frame is bench.py's frame 0 (752 x 480, FAST thr 10)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_b200 as fd  # noqa: E402
from bench import make_frames  # noqa: E402


def med(fn, n=300, warm=30):
    lat = []
    for i in range(n + warm):
        t0 = time.perf_counter()
        fn()
        if i >= warm:
            lat.append(time.perf_counter() - t0)
    return round(float(np.median(lat)) * 1e6, 1)


def main():
    W, H = 752, 480
    one = make_frames(1, 0, W, H)
    many = make_frames(1024, 0, W, H) if os.environ.get("FD_EXP_BATCH") else None   # before CUDA starts: the generator forks workers
    pinned = torch.from_numpy(one).pin_memory()
    det = fd.DetectParams(fd.FAST, 10.0, 20, int(os.environ.get("FD_EXP_N", "200")), fast_n=int(os.environ.get("FD_EXP_KN", "9")))
    brief = fd.BriefParams()
    ctx = fd.Context(0)
    out = {}
    out["stream synchronise alone"] = med(ctx.sync)

    def up_pageable():
        ctx.upload(one)
        ctx.sync()

    def up_pinned():
        ctx.upload_ptr(pinned.data_ptr(), H, W, 1)
        ctx.sync()

    out["upload, pageable + sync"] = med(up_pageable)
    out["upload, pinned + sync"] = med(up_pinned)

    def cand():
        ctx.compute_candidates(det, 65536)
        ctx.sync()

    def detect():
        ctx.detect(det, 65536)
        ctx.sync()

    def detect_describe():
        ctx.detect(det, 65536)
        ctx.describe_selected(brief)
        ctx.sync()

    out["candidates + sync (frame resident)"] = med(cand)
    out["candidates + selection + sync"] = med(detect)
    out["candidates + selection + BRIEF + sync"] = med(detect_describe)

    def down():
        ctx.keypoints(200)
        ctx.descriptors(200)

    out["keypoints + descriptors download (two calls)"] = med(down)
    out["one call, pageable frame"] = med(lambda: ctx.detect_describe_host(one, det, brief, 200, 65536))
    out["one call, pinned frame"] = med(lambda: ctx.detect_describe_host(pinned.numpy(), det, brief, 200, 65536))
    out["one call, pinned frame, no BRIEF"] = med(lambda: ctx.detect_describe_host(pinned.numpy(), det, None, 200, 65536))
    # the same call on a context that has served a 1024-frame batch before (bench.py's situation)
    if many is not None:
        det2 = det
        out["one call, fresh context, N 200"] = med(lambda: ctx.detect_describe_host(one, det2, brief, 200, 65536))
        host = torch.from_numpy(many).pin_memory()
        dev = host.to("cuda:0")
        big = fd.Context(0)
        big.bind_device(dev.data_ptr(), H, W, 1024)
        big.detect(det2, 65536)
        big.describe_selected(brief)
        big.sync()
        out["one call, context that served 1024 frames"] = med(lambda: big.detect_describe_host(one, det2, brief, 200, 65536))
        out["one call, fresh context again"] = med(lambda: ctx.detect_describe_host(one, det2, brief, 200, 65536))
        f0 = many[0]
        out["one call, frame is a view into the 370 MB pageable batch"] = med(lambda: ctx.detect_describe_host(f0, det2, brief, 200, 65536))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
