"""Experiment: the headline step in L2-sized sub-batches -- k contexts, each bound to 1/k of the 1024 frames, all on ONE
stream, FAST -> selection -> BRIEF per sub-batch, so that BRIEF's window reads find the sub-batch in the 126 MB L2."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_b200 as fd
from bench import make_frames
n = 1024
frames = make_frames(n, 0)
dev = torch.device("cuda", 0)
d = torch.from_numpy(frames).to(dev)
det = fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9)
brief = fd.BriefParams(256, 8)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
def timed(fn, reps=40):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for parts in (1, 2, 4, 8):
    ctxs = [fd.Context(0) for _ in range(parts)]
    per = n // parts
    for i, c in enumerate(ctxs):
        c.set_stream(stream.cuda_stream)
        c.bind_device(d.data_ptr() + i * per * 480 * 752, 480, 752, per)
    def step():
        for c in ctxs:
            c.detect(det, 65536)
            c.describe_selected(brief)
    print("%d sub-batch(es) of %d frames (%.0f MB), one stream: %.4f ms per 1024 frames" % (parts, per, per * 480 * 752 / 1e6, timed(step)), flush=True)
    for c in ctxs: c.close()
