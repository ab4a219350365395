import sys; sys.path.insert(0,'.')
import numpy as np
from mcptam_b200 import synth, capi
from oracle import oracle as ora
img = synth.make_frame(seed=1)
f = capi.FeHandle(640,480,max_corners_per_level=16384)
lv = f.make_keyframe(0,img)
sm = f.debug_scores(0,0)
xy = ora.fast10_detect(img,5); sc = ora.fast10_score(img,xy,5)
ref = np.zeros_like(sm); ref[xy[:,1],xy[:,0]] = np.minimum(sc,255)
print("score maps equal", np.array_equal(sm,ref), "nonzero", (sm>0).sum(), (ref>0).sum())
bad = np.argwhere(sm!=ref)
print("n bad", len(bad))
DX=[0,1,2,3,3,3,2,1,0,-1,-2,-3,-3,-3,-2,-1]; DY=[-3,-3,-2,-1,0,1,2,3,3,3,2,1,0,-1,-2,-3]
for y,x in bad[:6]:
    c=int(img[y,x]); d=[int(img[y+DY[i],x+DX[i]])-c for i in range(16)]
    print((x,y), "gpu", sm[y,x], "ref", ref[y,x], "x%32", x%32, "y%8", y%8, "d", d)
print("bad x%32 hist", np.bincount(bad[:,1]%32, minlength=32)); print("bad y%8 hist", np.bincount(bad[:,0]%8, minlength=8))
