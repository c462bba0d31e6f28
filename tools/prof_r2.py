"""Round-2 profiling driver for ncu: one warm + one profiled launch sequence of every kernel a bench line's roofline names, at the bench's
batch sizes -- the headline step (1024 x 752x480: fast_sparse, select, brief), configs[2] (512 x 1280x720 Shi-Tomasi), configs[3]
(64 x 3840x2160 Harris, untiled, and row-tiled over two tiles on this GPU: gather_tiles_kernel), configs[4] (256 x 1920x1080 LSD field +
seed order), Hamming matching.     ncu ... python tools/prof_r2.py     (frames come from bench.make_frames' cache when present)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_b200 as fd
from bench import make_frames
which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)
ctx = fd.Context(0)
if which in ("all", "headline"):
    d = torch.from_numpy(make_frames(1024, 0)).to(dev)
    ctx.bind_device(d.data_ptr(), 480, 752, 1024)
    det, brief = fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9), fd.BriefParams(256, 8)
    for _ in range(2):
        ctx.detect(det, 65536)
        ctx.describe_selected(brief)
        ctx.match_selected()
    ctx.sync()
    for prm in (fd.DetectParams(fd.HARRIS, 30.0, 20, 200), fd.DetectParams(fd.SHI_TOMAS, 40.0, 20, 200)):
        for _ in range(2):
            ctx.detect(prm, 65536)
    ctx.sync()
    del d
if which in ("all", "c2"):
    d = torch.from_numpy(make_frames(512, 0, 1280, 720)).to(dev)
    ctx.bind_device(d.data_ptr(), 720, 1280, 512)
    for _ in range(2):
        ctx.detect(fd.DetectParams(fd.SHI_TOMAS, 40.0, 20, 1000), 131072)
    ctx.sync()
    del d
if which in ("all", "c3"):
    d = torch.from_numpy(make_frames(64, 0, 3840, 2160)).to(dev)
    ctx.bind_device(d.data_ptr(), 2160, 3840, 64)
    har = fd.DetectParams(fd.HARRIS, 30.0, 20, 200)
    for _ in range(2):
        ctx.detect(har, 1 << 20)
    ctx.sync()
    with fd.TiledDetector([0, 0]) as td:
        td.scatter_device(d.data_ptr(), 2160, 3840, 64)
        for _ in range(2):
            td.exchange_halos()
            td.detect(har, 1 << 20)
        td.sync()
    del d
if which in ("all", "c4"):
    d = torch.from_numpy(make_frames(256, 0, 1920, 1080)).to(dev)
    ctx.bind_device(d.data_ptr(), 1080, 1920, 256)
    for _ in range(2):
        ctx.lsd_field(fd.LsdParams(20.0, 1))
    ctx.sync()
    del d
print("done")
