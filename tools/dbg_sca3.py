import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import orc
from soc_b200 import backend
from tests.cases import run_sca, _reg, _oct
from tests.stats import chi2_per_dof
cases = {
 "sca_ps_reg16": (_reg(16), dict(no_ps=2), lambda s: run_sca("ps", pspos=[(8.3, 8.3, 8.3), (4.1, 10.7, 12.2)], batch=64, glob=2048, seed=s)),
 "sca_bg_reg12": (_reg(12), {}, lambda s: run_sca("bg", batch=24, seed=s)),
 "sca_bg_oct6": (_oct(6, 3), {}, lambda s: run_sca("bg", batch=48, dirs=((45.0, 45.0),), seed=s)),
}
K = 24
for name, (make, opts, fac) in cases.items():
    cloud = make()
    def rep(X, off):
        return np.array([fac(0.03 + 0.94 * (k + off) / K)(X)["out"] for k in range(K)]).reshape(K, -1)
    O = orc.Oracle(cloud, **opts)
    a, a2 = rep(O, 0.5), rep(O, 0.21)
    print(name, "oracle-oracle", chi2_per_dof(a2, a, min_rel=0.02))
    for geo in (1, 0):
        for mode in ((backend.RNG_REFERENCE,) if geo == 1 else ()) + (backend.RNG_PACKET,):
            B = backend.Backend(cloud, rng_mode=mode, **opts)
            B.dev.set_geometry(geo)
            b = rep(B, 0.37)
            print(name, "gpu geo", geo, "rng", mode, chi2_per_dof(b, a, min_rel=0.02), "vs a2", chi2_per_dof(b, a2, min_rel=0.02)[0])
            B.close()
