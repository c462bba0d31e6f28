"""Experiment: the headline step on one context (1024 frames) vs split over two contexts / streams (512 frames each, launches interleaved)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import feature_detector_b200 as fd
from bench import make_frames
n = 1024
frames = make_frames(n, 0)
d = torch.from_numpy(frames).cuda()
det = fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9)
brief = fd.BriefParams(256, 8)
cap = 16384
def timed(fn, sync, reps=50):
    for _ in range(3): fn()
    sync(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    sync(); return (time.perf_counter() - t0) / reps * 1e3
one = fd.Context(0)
one.bind_device(d.data_ptr(), 480, 752, n)
def step1():
    one.detect(det, cap); one.describe_selected(brief)
print("one context, 1024 frames: %.4f ms" % timed(step1, one.sync))
for parts in (2, 4):
    ctxs = [fd.Context(0) for _ in range(parts)]
    per = n // parts
    for i, c in enumerate(ctxs):
        c.bind_device(d.data_ptr() + i * per * 480 * 752, 480, 752, per)
    def stepn():
        for c in ctxs: c.detect(det, cap)
        for c in ctxs: c.describe_selected(brief)
    def syncn():
        for c in ctxs: c.sync()
    print("%d contexts, %d frames each: %.4f ms" % (parts, per, timed(stepn, syncn)))
    for c in ctxs: c.close()
