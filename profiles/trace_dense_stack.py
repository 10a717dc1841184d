import sys, numpy as np
sys.path.insert(0,'solving-poisson-s-equation-through-dl-for-cfd-apllications_b200')
import psm_b200
rng=np.random.default_rng(0)
dims=[128,512,512,512,128]
x=rng.standard_normal((121,128)).astype(np.float32)
ks=[(rng.standard_normal((dims[i],dims[i+1]))/np.sqrt(dims[i])).astype(np.float32) for i in range(4)]
bs=[np.zeros(dims[i+1],np.float32) for i in range(4)]
psm_b200.debug_dense_stack(x,ks,bs)
