"""FAST at the reference's default options (thr 0.1, d 15, N 200, kN 12) on 256 synthetic 752x480 frames: every pixel past the 10 000th is a candidate."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import feature_detector_b200 as fd
from bench import make_frames
n = 256
frames = make_frames(1024, 0)[:n]
d = torch.from_numpy(frames).cuda()
ctx = fd.Context(0)
ctx.bind_device(d.data_ptr(), 480, 752, n)
prm = fd.DetectParams(fd.FAST, 0.1, 15, 200, fast_n=12)
reps = 1 if len(sys.argv) > 1 else 5
for _ in range(2): ctx.detect(prm, 0)
ctx.sync(); t0 = time.perf_counter()
for _ in range(reps): ctx.detect(prm, 0)
ctx.sync(); dt = (time.perf_counter() - t0) / reps
print("fast default detect %.3f ms per %d frames" % (dt * 1e3, n), "kept", ctx.keypoint_counts().mean())
