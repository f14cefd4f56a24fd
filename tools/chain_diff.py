import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from conftest import Oracle
from svdsolver_b200 import capi
from svdsolver_b200.synth import uniform_matrix
o=Oracle('oracle/libsvd_oracle.so')
for dt in (np.float32, np.float64):
  for n,b in [(64,32),(128,16),(128,32),(256,32),(512,64)]:
    a=np.stack([uniform_matrix(n,n,1000+i,0.0,5.0,dt) for i in range(3)])
    with capi.Handle(n,b,dt) as h:
        sig=h.svdvals_batched(a,b)
        s1,_=h.svdvals(a[0],b)
    band=o.brd_p1_panel(a[0],b); _,d,e=o.brd_p2(band,b)
    ref=np.linalg.svd(np.diag(d.astype(np.float64))+np.diag(e.astype(np.float64),1),compute_uv=False)
    true=np.linalg.svd(a[0].astype(np.float64),compute_uv=False)
    f=lambda x,y: np.abs(x.astype(np.float64)-y).max()/y[0]
    print(np.dtype(dt).name,n,b,"batched-oracle %.2e single-oracle %.2e batched-single %.2e oracle-true %.2e batched-true %.2e"%(f(sig[0],ref),f(s1,ref),f(sig[0],s1.astype(np.float64)),f(ref,true),f(sig[0],true)))
