import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import feature_detector_b200 as fd
from bench import make_frames
n=256
frames=make_frames(1024,0)[:n]
d=torch.from_numpy(frames).cuda()
ctx=fd.Context(0)
ctx.bind_device(d.data_ptr(),480,752,n)
for name,prm,cap in (("fast default (thr .1, d 15, N 200, kN 12)", fd.DetectParams(fd.FAST,0.1,15,200,fast_n=12),0),
                 ("harris default (thr .1, d 15, N 200)", fd.DetectParams(fd.HARRIS,0.1,15,200),0),
                 ("shi default (thr .1, d 15, N 200)", fd.DetectParams(fd.SHI_TOMAS,0.1,15,200),0)):
    for part in ("cand","detect"):
        fn=(lambda: ctx.compute_candidates(prm,cap)) if part=="cand" else (lambda: ctx.detect(prm,cap))
        for _ in range(2): fn()
        ctx.sync(); t0=time.perf_counter()
        for _ in range(5): fn()
        ctx.sync(); dt=(time.perf_counter()-t0)/5
        print(name,part,"%.3f ms per %d frames"%(dt*1e3,n),"%.1f Gpx/s"%(n*480*752/dt/1e9), "cands/frame", int(ctx.candidate_counts().mean()))
