import sys, numpy as np
sys.path.insert(0, '/root/repo')
from tests.cases import CASES, run_sca, _reg
from soc_b200 import backend
from oracle import orc
cloud = _reg(16)()
for name, kw in (("ps", dict(kind="ps", pspos=[(8.3, 8.3, 8.3)], batch=64, glob=4096)), ("bg", dict(kind="bg", batch=8))):
    kind = kw.pop("kind")
    run = run_sca(kind, **kw)
    O = orc.Oracle(cloud, no_ps=1)
    o = run(O)["out"].astype(np.float64)
    print(name, "oracle", o.sum(), O.counters.packets, O.counters.steps, O.counters.scatterings, O.counters.peels)
    for mode in (0, 1):
        B = backend.Backend(cloud, rng_mode=mode, no_ps=1)
        g = run(B)["out"].astype(np.float64)
        c = B.counters
        print(name, "gpu mode", mode, g.sum(), c.packets, c.steps, c.scatterings, c.peels, c.reserved[0])
        B.close()
