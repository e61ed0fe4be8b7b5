"""Differential fuzzing of the C restatement of cv2.StereoSGBM / cv2.StereoBM (oracle/csrc/orc_sgbm.c, orc_bm.c) against the
installed cv2 binary over random parameter sets (disparity range and sign, block size, P1 / P2, preFilterCap, uniqueness,
disp12MaxDiff, speckle filter, all four modes) and random small images (textured pairs, quantised pairs with cost ties, pure
noise).  This is how the shifted stripe rows of SGBM_3WAY on images of a few rows were found.

    python tests/fuzz/fuzz_oracle.py [seed] [iterations] [sgbm|bm]
"""
import os, sys
import cv2
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cref
from laser_3d_reconstruction_b200 import synth

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
what = sys.argv[3] if len(sys.argv) > 3 else "sgbm"
rng = np.random.default_rng(seed)
bad = 0
for it in range(iters if what == "sgbm" else 0):
    D=int(rng.choice([16,32,48,64,96,128]))
    bs=int(rng.choice([1,3,5,7,9,11]))
    W=int(rng.integers(D+20, D+160)); H=int(rng.integers(8,60))
    mode=int(rng.integers(0,4))
    minD=int(rng.choice([0,0,-(D-1),3,-5,16,-D//2]))
    cap=int(rng.choice([63,63,31,15,1,40]))
    uq=int(rng.choice([0,10,5,15,40])); d12=int(rng.choice([1,0,-1,2,1000000,5])); sw=int(rng.choice([0,100,20,200])); sr=int(rng.choice([32,1,2,16]))
    P1=int(rng.choice([24*bs*bs, 8*bs*bs, 10, 0, 100])); P2=int(rng.choice([96*bs*bs, 32*bs*bs, 200, 1000])); 
    if P2<=P1: P2=P1+1
    l,r=synth.stereo_pair(W,H,max(D,16),int(rng.integers(0,1000)))
    lg,rg=cv2.cvtColor(l,cv2.COLOR_BGR2GRAY),cv2.cvtColor(r,cv2.COLOR_BGR2GRAY)
    q=int(rng.choice([0,0,8,32]))
    if q: lg=(lg//q*q).astype(np.uint8); rg=(rg//q*q).astype(np.uint8)
    if rng.random()<0.2: lg=rng.integers(0,256,lg.shape,dtype=np.uint8); rg=rng.integers(0,256,rg.shape,dtype=np.uint8)
    kw=dict(minDisparity=minD,numDisparities=D,blockSize=bs,P1=P1,P2=P2,disp12MaxDiff=d12,preFilterCap=cap,uniquenessRatio=uq,speckleWindowSize=sw,speckleRange=sr,mode=mode)
    try:
        want=cv2.StereoSGBM_create(**kw).compute(lg,rg)
    except cv2.error as e:
        print("cv2 error",kw,str(e)[:80]); continue
    try:
        got=cref.sgbm_compute(lg,rg,**kw)
    except Exception as e:
        print("oracle refuses",W,H,kw,str(e)[:100]); continue
    if not np.array_equal(got,want):
        bad+=1; print("MISMATCH",W,H,kw,int((got!=want).sum()))

for it in range(iters if what == "bm" else 0):
    D=int(rng.choice([16,32,48,64,96,128])); bs=int(rng.choice([5,7,9,11,15,21]))
    W=int(rng.integers(D+bs+8, D+200)); H=int(rng.integers(bs//2+2,70))
    minD=int(rng.choice([0,0,-8,-(D-1),-3,-D//2]))
    cap=int(rng.choice([31,63,15,1,7])); tex=int(rng.choice([10,0,50,200])); uq=int(rng.choice([15,0,5,40])); sw=int(rng.choice([0,100,30])); sr=int(rng.choice([32,0,2])); d12=int(rng.choice([-1,-1,1,0,5]))
    l,r=synth.stereo_pair(W,H,max(D,16),int(rng.integers(0,1000)))
    lg,rg=cv2.cvtColor(l,cv2.COLOR_BGR2GRAY),cv2.cvtColor(r,cv2.COLOR_BGR2GRAY)
    q=int(rng.choice([0,0,16,32]))
    if q: lg=(lg//q*q).astype(np.uint8); rg=(rg//q*q).astype(np.uint8)
    if rng.random()<0.2: lg=rng.integers(0,256,lg.shape,dtype=np.uint8); rg=rng.integers(0,256,rg.shape,dtype=np.uint8)
    m=cv2.StereoBM_create(numDisparities=D,blockSize=bs)
    m.setMinDisparity(minD); m.setPreFilterCap(cap); m.setTextureThreshold(tex); m.setUniquenessRatio(uq); m.setSpeckleWindowSize(sw); m.setSpeckleRange(sr); m.setDisp12MaxDiff(d12)
    try: want=m.compute(lg,rg)
    except cv2.error as e: print("cv2 error",str(e)[:60]); continue
    try: got=cref.bm_compute(lg,rg,D,bs,minD,cap,tex,uq,sw,sr,d12)
    except Exception as e: print("oracle refuses",(W,H,D,bs,minD,cap,tex,uq,sw,sr,d12),str(e)[:60]); continue
    if not np.array_equal(got,want): bad+=1; print("MISMATCH",(W,H,D,bs,minD,cap,tex,uq,sw,sr,d12),int((got!=want).sum()))

print("done bad =", bad)
