"""Times the stages of the headline step (FAST candidates, selection, BRIEF) and of the other detectors separately, CUDA events on the
context's stream.  python tools/stage_times.py [frames]   (knobs come from the FD_B200_* environment variables)"""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_b200 as fd
from bench import make_frames

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
frames = make_frames(n, 0)
dev = torch.device("cuda", 0)
d = torch.from_numpy(frames).to(dev)
ctx = fd.Context(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
ctx.bind_device(d.data_ptr(), 480, 752, n)


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {}
brief = fd.BriefParams(256, 8)
for name, det, cap in (("fast9", fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9), 65536), ("fast12", fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=12), 65536),
                       ("harris", fd.DetectParams(fd.HARRIS, 30.0, 20, 200), 65536), ("shi", fd.DetectParams(fd.SHI_TOMAS, 40.0, 20, 200), 65536)):
    t_c = timed(lambda: ctx.compute_candidates(det, cap))
    t_d = timed(lambda: ctx.detect(det, cap))
    ctx.detect(det, cap)
    t_b = timed(lambda: ctx.describe_selected(brief))
    t_all = timed(lambda: (ctx.detect(det, cap), ctx.describe_selected(brief)))
    ctx.sync()
    out[name] = {"candidates_ms": round(t_c, 4), "select_ms": round(t_d - t_c, 4), "brief_ms": round(t_b, 4), "step_ms": round(t_all, 4),
                 "mean_cand": float(ctx.candidate_counts().mean()), "max_cand": int(ctx.candidate_counts().max()), "mean_kp": float(ctx.keypoint_counts().mean())}
# one frame at a time (the drop-in classes' call pattern): kernel time only
ctx.bind_device(d.data_ptr(), 480, 752, 1)
det = fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9)
t_c = timed(lambda: ctx.compute_candidates(det, 65536), 200)
t_d = timed(lambda: ctx.detect(det, 65536), 200)
t_b = timed(lambda: ctx.describe_selected(brief), 200)
out["one_frame_us"] = {"candidates": round(t_c * 1e3, 1), "select": round((t_d - t_c) * 1e3, 1), "brief": round(t_b * 1e3, 1)}
print(json.dumps(out, indent=1))
