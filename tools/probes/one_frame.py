import sys; sys.path.insert(0,'.')
import numpy as np, torch
from stereo_depth_ruler_b200 import capi, synth
ctx=capi.Context(0)
l,r=synth.stereo_pair(1280,720,128,1)
L=torch.from_numpy(l[None]).cuda(); R=torch.from_numpy(r[None]).cuda(); O=torch.empty((1,720,1280),dtype=torch.int16,device="cuda")
for mode in (0,2):
    m=capi.Sgbm(ctx,capi.make_params(mode=mode)); m.set_profiling(True)
    acc={}
    for rep in range(6):
        m.compute_batch(L,R,O); ctx.synchronize()
        if rep:
            for k,v in m.stage_times(): acc[k]=acc.get(k,0)+v/5
    print("mode",mode,"one 1280x720x128 frame: total %.3f ms"%sum(acc.values()),{k:round(v,4) for k,v in acc.items()})
    m.close()
