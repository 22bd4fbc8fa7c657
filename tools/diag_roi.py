import numpy as np, sys
sys.path.insert(0, '.')
from oracle import orc
from soc_b200 import backend
from tests.cases import _oct, run_ps, with_roi_save
opts=dict(no_ps=1, with_roi_save=1, roi=[3, 4, 2, 4, 1, 3], roi_step=1, roi_nside=1)
cloud=_oct(6,3)()
K=12
A=[];Bv=[]
O=orc.Oracle(cloud, mirror_exact=1, **opts)
B=backend.Backend(cloud, rng_mode=backend.RNG_PACKET, **opts)
for k in range(K):
    s=0.05+0.9*(k+0.5)/K
    A.append(with_roi_save(run_ps([(1.3,1.2,4.9)],batch=96,seed=s))(O)["roi_save"].astype(np.float64))
    Bv.append(with_roi_save(run_ps([(1.3,1.2,4.9)],batch=96,seed=s))(B)["roi_save"].astype(np.float64))
A=np.array(A);Bv=np.array(Bv)
ma,mb=A.mean(0),Bv.mean(0)
print("tot",ma.sum(),mb.sum())
ne=21
print("per element (sum over pixels): oracle / gpu")
ea=ma.reshape(ne,12).sum(1); eb=mb.reshape(ne,12).sum(1)
for i in range(ne): print(i, "%.4f %.4f  ratio %.3f"%(ea[i],eb[i],eb[i]/max(ea[i],1e-30)))
print("per pixel:", (ma.reshape(ne,12).sum(0)), (mb.reshape(ne,12).sum(0)))
