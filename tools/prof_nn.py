"""NN post-processing on 1024 synthetic 752x480 heat maps (+ 256-channel descriptor volumes at 1/8 resolution): timing."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import feature_detector_b200 as fd
from feature_detector_b200.synth import synth_heatmap, synth_descriptor_volume
n, w, h = 1024, 752, 480
base = np.stack([synth_heatmap(w, h, i) for i in range(8)])
maps = torch.from_numpy(base).cuda().repeat(n // 8, 1, 1).contiguous()
vol = torch.from_numpy(np.stack([synth_descriptor_volume(256, h // 8, w // 8, i) for i in range(2)])).cuda().repeat(n // 2, 1, 1, 1).contiguous()
ctx = fd.Context(0)
prm = fd.NnParams(0.1, 3, 15, 240)
reps = 1 if len(sys.argv) > 1 else 20
for what in ("select", "select+descriptors"):
    def step():
        ctx.nn_select(maps.data_ptr(), h, w, n, prm, 65536)
        if what != "select": ctx.nn_sample_descriptors(vol.data_ptr(), 256, h // 8, w // 8)
    for _ in range(3): step()
    ctx.sync(); t0 = time.perf_counter()
    for _ in range(reps): step()
    ctx.sync(); dt = (time.perf_counter() - t0) / reps
    print(f"nn {what}: {dt*1e3:.3f} ms per {n} maps, {n*w*h/dt/1e9:.1f} Gpx/s, heat-map read {n*w*h*4/dt/1e9:.0f} GB/s; kept {ctx.keypoint_counts().mean():.1f}")
# reference point: how fast the same buffer can be read at all (torch reductions)
for name, fn in (("torch.sum", lambda: maps.sum()), ("torch.amax", lambda: maps.amax()), ("torch (maps > 0.1).sum", lambda: (maps > 0.1).sum())):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"{name}: {dt*1e3:.3f} ms, {maps.numel()*4/dt/1e9:.0f} GB/s")
