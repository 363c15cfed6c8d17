import sys, time, ctypes as C
sys.path.insert(0,'.'); sys.path.insert(0,'./tests')
import numpy as np, torch
import bench
from tendrils_b200 import _native as N
wl=bench.WORKLOADS['cfg3']
t,first,sp=bench.build_sim(wl,0,1,0,None)
first.spawn(t)
L=N.load(); ctx=t.particles._ctx
for k in range(40):
    t.timer.tick()
    torch.cuda.synchronize() if k>=30 else None
    a=time.perf_counter(); t.step(); b=time.perf_counter(); t.draw(); c=time.perf_counter()
    if k>=30:
        torch.cuda.synchronize(); d=time.perf_counter()
        print('step call %.0fus draw call %.0fus tail %.0fus total %.0fus'%((b-a)*1e6,(c-b)*1e6,(d-c)*1e6,(d-a)*1e6), t.particles.timing(reset=True))
