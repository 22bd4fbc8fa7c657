import numpy as np, sys
sys.path.insert(0, '.')
from oracle import orc
from soc_b200 import backend
from tests.cases import _oct, _reg, run_ps, run_bg, with_roi_save
from tests.stats import chi2_per_dof
K=16
def go(name, cloud, opts, fac, **okw):
    O=orc.Oracle(cloud, mirror_exact=1, **opts, **okw)
    B=backend.Backend(cloud, rng_mode=backend.RNG_PACKET, **opts)
    A=[];Bv=[];TA=[];TB=[]
    for k in range(K):
        s=0.05+0.9*(k+0.5)/K
        o=fac(s)(O); A.append(o["roi_save"].astype(np.float64)); TA.append(o["tabs"].astype(np.float64))
        o=fac(s)(B); Bv.append(o["roi_save"].astype(np.float64)); TB.append(o["tabs"].astype(np.float64))
    A=np.array(A);Bv=np.array(Bv);TA=np.array(TA);TB=np.array(TB)
    c=chi2_per_dof(Bv,A,min_rel=1e-4); ct=chi2_per_dof(TB,TA,min_rel=1e-4)
    print(name,"roi chi2 %.3f dof %d tot %.2e sig %.2e | tabs chi2 %.3f tot %.2e sig %.2e"%(c[0],c[1],c[2],c[3],ct[0],ct[2],ct[3]))
    B.close()
r1=dict(with_roi_save=1, roi=[3, 4, 2, 4, 1, 3], roi_step=1, roi_nside=1)
go("oct PS", _oct(6,3)(), dict(no_ps=1, **r1), lambda s: with_roi_save(run_ps([(1.3,1.2,4.9)],batch=96,seed=s)))
go("oct BG", _oct(6,3)(), r1, lambda s: with_roi_save(run_bg(batch=16,seed=s)))
go("reg PS", _reg(6)(), dict(no_ps=1, **r1), lambda s: with_roi_save(run_ps([(1.3,1.2,4.9)],batch=96,seed=s)))
go("reg BG", _reg(6)(), r1, lambda s: with_roi_save(run_bg(batch=16,seed=s)))
