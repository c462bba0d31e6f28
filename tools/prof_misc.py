"""One launch each of the kernels that are not on the headline step, for ncu: corner (Harris, Shi-Tomasi), per-cell selection (FAST at the
reference's default threshold), LSD field + seed order, NN post-processing."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import feature_detector_b200 as fd
from bench import make_frames
from feature_detector_b200.synth import synth, synth_heatmap, synth_descriptor_volume
n = 256
d = torch.from_numpy(make_frames(1024, 0)[:n]).cuda()
ctx = fd.Context(0)
ctx.bind_device(d.data_ptr(), 480, 752, n)
for prm, cap in ((fd.DetectParams(fd.HARRIS, 30.0, 20, 200), 65536), (fd.DetectParams(fd.SHI_TOMAS, 40.0, 20, 200), 65536),
                 (fd.DetectParams(fd.FAST, 0.1, 15, 200, fast_n=12), 0)):
    for _ in range(2): ctx.detect(prm, cap)
ctx.sync()
fr = np.stack([synth(1920, 1080, i) for i in range(4)]); fr = np.concatenate([fr] * 16)
d2 = torch.from_numpy(fr).cuda()
ctx.bind_device(d2.data_ptr(), 1080, 1920, 64)
for _ in range(2): ctx.lsd_field(fd.LsdParams(20.0, 1))
ctx.sync()
maps = torch.from_numpy(np.stack([synth_heatmap(752, 480, i) for i in range(8)])).cuda().repeat(32, 1, 1).contiguous()
vol = torch.from_numpy(np.stack([synth_descriptor_volume(256, 60, 94, i) for i in range(2)])).cuda().repeat(128, 1, 1, 1).contiguous()
for _ in range(2):
    ctx.nn_select(maps.data_ptr(), 480, 752, 256, fd.NnParams(0.1, 3, 15, 240), 65536)
    ctx.nn_sample_descriptors(vol.data_ptr(), 256, 60, 94)
ctx.sync()
print("done")
