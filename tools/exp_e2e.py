"""Experiment: raw pinned H2D rate vs the HostPipeline at several chunk sizes / context counts."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import feature_detector_b200 as fd
from feature_detector_b200.pipeline import HostPipeline
from bench import make_frames
n = 1024
host = torch.from_numpy(make_frames(n, 0)).pin_memory()
dev = torch.empty_like(host, device='cuda')
for _ in range(3): dev.copy_(host, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): dev.copy_(host, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print(f"raw pinned H2D of {host.numel()/1e6:.0f} MB: {dt*1e3:.3f} ms, {host.numel()/dt/1e9:.1f} GB/s")
det = fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9)
brief = fd.BriefParams(256, 8)
kp = np.zeros((n, 200), fd.KEYPOINT_DTYPE); cnt = np.zeros(n, np.int32); desc = np.zeros((n, 200, 32), np.uint8)
kp_t = torch.from_numpy(kp.view(np.uint8)).pin_memory(); cnt_t = torch.from_numpy(cnt).pin_memory(); desc_t = torch.from_numpy(desc).pin_memory()
kp = kp_t.numpy().view(fd.KEYPOINT_DTYPE).reshape(n, 200); cnt = cnt_t.numpy(); desc = desc_t.numpy()
for chunk, nctx in ((128, 2), (64, 2), (256, 2), (128, 3), (64, 3), (32, 4)):
    with HostPipeline(0, chunk, nctx) as pipe:
        def step(): pipe.run(host.data_ptr(), 480, 752, n, det, brief, kp, cnt, desc, 65536)
        for _ in range(3): step()
        t0 = time.perf_counter()
        for _ in range(10): step()
        dt = (time.perf_counter() - t0) / 10
        print(f"pipeline chunk {chunk} x {nctx} contexts: {dt*1e3:.3f} ms, {n*480*752/dt/1e9:.2f} Gpx/s")
