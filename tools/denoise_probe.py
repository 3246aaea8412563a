"""Scratch: reconstruction-only timing on 4K synthetic feature buffers (configs[4])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, hijiki_b200 as hj, sys
dw,dh=3840,2160
rng=np.random.default_rng(5)
rad=np.exp(rng.standard_normal((dh,dw,4),dtype=np.float32)); rad[...,3]=1
mode=sys.argv[1] if len(sys.argv)>1 else 'random'
nrm=rng.standard_normal((dh,dw,4),dtype=np.float32); nrm[...,:3]/=np.linalg.norm(nrm[...,:3],axis=2,keepdims=True)
if mode=='flat': nrm[...,:3]=(0,0,1)
blocks=hj.ImageBlockGenerator(dw,dh,128,1).blocks()
ctx=hj.Context(0); ctx.frame_begin(dw,dh); ctx.denoise_upload(rad,nrm,blocks)
p=hj.make_params()
ctx.denoise_resident(p,3)
reps=int(sys.argv[2]) if len(sys.argv)>2 else 20
ms=ctx.denoise_resident(p,reps)/reps
print(mode,'ms/pass',ms,'GB/s',64*dw*dh/ms/1e6)
