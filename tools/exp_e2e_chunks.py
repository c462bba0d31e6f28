"""e2e leg of bench.py (HostPipeline over pinned frames) for several chunk sizes and context counts."""
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_detector_b200 as fd
from feature_detector_b200.pipeline import HostPipeline
from bench import make_frames
n, H, W = 1024, 480, 752
frames = make_frames(n, 0)
host = torch.from_numpy(frames).pin_memory()
det, brief = fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9), fd.BriefParams(256, 8)
kp = torch.zeros((n, 200, 16), dtype=torch.uint8).pin_memory()
cnt = torch.zeros((n,), dtype=torch.int32).pin_memory()
desc = torch.zeros((n, 200, 32), dtype=torch.uint8).pin_memory()
kpv = kp.numpy().view(fd.KEYPOINT_DTYPE).reshape(n, 200)
d = torch.empty((n, H, W), dtype=torch.uint8, device="cuda")
def plain():
    d.copy_(host, non_blocking=True)
for _ in range(3): plain()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): plain()
torch.cuda.synchronize(); tp = (time.perf_counter() - t0) / 10
print("plain pinned H2D of the batch: %.3f ms = %.1f GB/s" % (tp * 1e3, n * H * W / tp / 1e9), flush=True)
for n_ctx in (2, 3):
    for chunk in (32, 64, 128, 256):
        pipe = HostPipeline(0, chunk_frames=chunk, n_contexts=n_ctx)
        for _ in range(3): pipe.run(host.data_ptr(), H, W, n, det, brief, kpv, cnt.numpy(), desc.numpy(), 65536)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): pipe.run(host.data_ptr(), H, W, n, det, brief, kpv, cnt.numpy(), desc.numpy(), 65536)
        torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 10
        print("contexts %d chunk %4d: %.3f ms per step = %.1f Gpx/s (%.3f of the plain copy)" % (n_ctx, chunk, t * 1e3, n * H * W / t / 1e9, tp / t), flush=True)
        pipe.close()
