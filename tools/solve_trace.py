import sys; sys.path.insert(0,'.')
import numpy as np
from mcptam_b200 import synth, capi
prob = synth.make_ba_config("cfg2", 0)
g = capi.BaHandle(); g.load(prob)
g.compute(3)
g.reset_state()
g.solve_trace(arm=True)
d,s,r = g.lm_step(100.0)
tr = g.solve_trace()
T = 10; nt = T*(T+1)//2 + T
t = tr[:nt]
t0 = t[:,2].min()
print("tasks", nt, "span us", (t[:,4].max()-t0)/1e3, "backsolve end us", (tr[nt,0]-t0)/1e3)
for k in range(nt):
    i,j,a,b,c,cta = t[k][:6]
    if i==j or i==j+1 or i==T:
        extra = f" potrf_done {(t[k][6]-t0)/1e3:7.1f} inv_done {(t[k][7]-t0)/1e3:7.1f}" if i==j else ""
        print(f"task {k:3d} ({int(i)},{int(j)}) start {(a-t0)/1e3:7.1f} deps {(b-t0)/1e3:7.1f} end {(c-t0)/1e3:7.1f} cta {int(cta)}"+extra)
