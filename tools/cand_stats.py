import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import feature_detector_b200 as fd
from bench import make_frames
n=1024
d=torch.from_numpy(make_frames(n,0)).cuda()
ctx=fd.Context(0); ctx.bind_device(d.data_ptr(),480,752,n)
ctx.detect(fd.DetectParams(fd.FAST,10.0,20,200,fast_n=9),65536); ctx.sync()
c=ctx.candidate_counts(); k=ctx.keypoint_counts()
print("cands mean %.0f median %.0f p90 %.0f p99 %.0f max %d min %d"%(c.mean(),np.median(c),np.percentile(c,90),np.percentile(c,99),c.max(),c.min()))
print("kept mean %.0f max %d"%(k.mean(),k.max()))
