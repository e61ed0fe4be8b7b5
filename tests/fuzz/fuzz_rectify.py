"""Differential fuzzing of camera/rectify.py (the host restatement of cv2.stereoRectify) against cv2 over random rigs: 4 / 5 / 8 /
12 / 14 distortion coefficients, both baseline directions, every alpha, with and without CALIB_ZERO_DISPARITY.

    python tests/fuzz/fuzz_rectify.py <seed> <iterations>
"""
import os, sys
import cv2
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from laser_3d_reconstruction_b200.camera.rectify import stereo_rectify
rng=np.random.default_rng(int(sys.argv[1])); n=int(sys.argv[2])
worst=0; bad=0
def rel(a,b): return float(np.abs(np.asarray(a)-np.asarray(b)).max()/max(np.abs(np.asarray(b)).max(),1e-300))
for it in range(n):
    W=int(rng.choice([320,640,1280,800])); H=int(rng.choice([240,480,720,600]))
    f=rng.uniform(0.6,1.4)*W
    K1=np.array([[f*rng.uniform(.97,1.03),0,W/2+rng.uniform(-20,20)],[0,f*rng.uniform(.97,1.03),H/2+rng.uniform(-20,20)],[0,0,1]])
    K2=np.array([[f*rng.uniform(.97,1.03),0,W/2+rng.uniform(-20,20)],[0,f*rng.uniform(.97,1.03),H/2+rng.uniform(-20,20)],[0,0,1]])
    nd=int(rng.choice([4,5,8,12,14]))
    def dist():
        d=np.zeros(nd); d[0]=rng.uniform(-.3,.2); d[1]=rng.uniform(-.1,.1); d[2]=rng.uniform(-2e-3,2e-3); d[3]=rng.uniform(-2e-3,2e-3)
        if nd>4: d[4]=rng.uniform(-.05,.05)
        if nd>=8: d[5:8]=rng.uniform(-.02,.02,3)
        if nd>=12: d[8:12]=rng.uniform(-1e-3,1e-3,4)
        if nd>=14: d[12:14]=rng.uniform(-1e-2,1e-2,2)
        return d
    d1,d2=dist(),dist()
    R=cv2.Rodrigues(rng.uniform(-0.05,0.05,3))[0]
    vertical=rng.random()<0.25
    T=np.array([rng.uniform(-0.005,0.005),-rng.uniform(0.03,0.2),rng.uniform(-0.005,0.005)]) if vertical else np.array([-rng.uniform(0.03,0.2),rng.uniform(-0.005,0.005),rng.uniform(-0.005,0.005)])
    if rng.random()<0.2: T=-T
    flags=int(rng.choice([cv2.CALIB_ZERO_DISPARITY,0])); alpha=float(rng.choice([0,-1,1,0.5,0.25]))
    newsize=(0,0)
    want=cv2.stereoRectify(K1,d1,K2,d2,(W,H),R,T,flags=flags,alpha=alpha)
    try: got=stereo_rectify(K1,d1,K2,d2,(W,H),R,T,flags=flags,alpha=alpha)
    except Exception as e: print("raises",e); bad+=1; continue
    e=max(rel(got[i],want[i]) for i in range(5)); worst=max(worst,e)
    roi_ok = got[5]==tuple(want[5]) and got[6]==tuple(want[6])
    if e>1e-11 or not roi_ok:
        bad+=1; print("MISMATCH it",it,"err",e,"roi",got[5],tuple(want[5]),got[6],tuple(want[6]),"nd",nd,"vert",vertical,"flags",flags,"alpha",alpha)
print("done bad=",bad,"worst rel",worst)
