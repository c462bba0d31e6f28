import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import feature_detector_b200 as fd
from feature_detector_b200.synth import synth
from oracle import bindings as ob
chk=ob.best_checker()
im=synth(1920,1080,3)
ctx=fd.Context(0)
for kind,okind,thr,d,n,fn in ((fd.HARRIS,ob.HARRIS,30.0,20,200,12),(fd.FAST,ob.FAST,10.0,20,500,9),(fd.SHI_TOMAS,ob.SHI_TOMAS,40.0,12,1000,12)):
    o=chk.detect(okind, im, thr, d, n, fast_n=fn)
    ctx.upload(im); ctx.detect(fd.DetectParams(kind,thr,d,n,fast_n=fn))
    kp,cnt=ctx.keypoints(n)
    got=np.stack([kp["x"][0,:cnt[0]],kp["y"][0,:cnt[0]]],1).astype(np.float32)
    same=np.array_equal(got,o["features"])
    ts=[]
    for i in range(60):
        ctx.sync(); t0=time.perf_counter(); ctx.detect(fd.DetectParams(kind,thr,d,n,fast_n=fn)); ctx.sync(); ts.append(time.perf_counter()-t0)
    tc=[]
    for i in range(60):
        ctx.sync(); t0=time.perf_counter(); ctx.compute_candidates(fd.DetectParams(kind,thr,d,n,fast_n=fn)); ctx.sync(); tc.append(time.perf_counter()-t0)
    print(kind, 'same', same, 'cand', o['n_cand'], 'kp', int(cnt[0]), 'detect us', round(np.median(ts)*1e6,1), 'candidates us', round(np.median(tc)*1e6,1))
