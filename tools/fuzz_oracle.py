import sys, numpy as np, cv2
sys.path.insert(0,'/root/repo')
from oracle import cref
from laser_3d_reconstruction_b200 import synth
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
bad=0
for it in range(int(sys.argv[2]) if len(sys.argv)>2 else 60):
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
print("done bad=",bad)
