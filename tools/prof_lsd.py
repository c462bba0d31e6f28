"""LSD field on 64 synthetic 1920x1080 frames: timing with CUDA events (and the launch ncu attaches to)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import feature_detector_b200 as fd
from feature_detector_b200.synth import synth
count = 64
fr = np.stack([synth(1920, 1080, i) for i in range(4)])
fr = np.concatenate([fr] * (count // 4))
d = torch.from_numpy(fr).cuda()
ctx = fd.Context(0)
ctx.bind_device(d.data_ptr(), 1080, 1920, count)
for want_sorted in (0, 1):
    prm = fd.LsdParams(20.0, want_sorted)
    for _ in range(3): ctx.lsd_field(prm)
    ctx.sync()
    reps = 1 if len(sys.argv) > 1 else 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); 
    import time; t0 = time.perf_counter()
    for _ in range(reps): ctx.lsd_field(prm)
    ctx.sync(); dt = (time.perf_counter() - t0) / reps
    px = count * 1920 * 1080
    print(f"lsd sorted={want_sorted}: {dt*1e3:.3f} ms per {count} frames, {px/dt/1e9:.1f} Gpx/s, {px*9/dt/1e9:.0f} GB/s algorithmic")
