import sys, time, json
sys.path.insert(0,'.')
import numpy as np
import renderer_b200 as rb
from oracle import pyport
path=pyport.model_path('chessboard.tri')
s=rb.Scene(path).UpdateBoundingVolumeHierarchy(path+'.bvh')
g=rb.Renderer(0); g.upload(s)
cams=rb.Orbit.cameras(range(20))
W,H=1920,1080
for flags,name in [(5,'C2 norefl'),(7,'defaults')]:
    ts=[]
    for k in range(20):
        f=rb.make_frame(9,W,H,cams[k],flags=flags)
        t=time.time(); img=g.render(f); dt=time.time()-t
        ts.append((g.last_kernel_ms()[0],dt*1e3))
    print(name,'kernel ms',[round(t[0],3) for t in ts[:8]],'e2e ms',[round(t[1],2) for t in ts[:8]])
    g.set_counters(True); f=rb.make_frame(9,W,H,cams[0],flags=flags); img=g.render(f); print(g.counters()); g.set_counters(False)
    t=time.time(); want=pyport.render(s,f); print('oracle port s',time.time()-t,'mismatch',int((img!=want).sum()))
