import numpy as np, sys
sys.path.insert(0, '.')
from oracle import orc
from soc_b200 import backend
from tests.cases import CASES
for name in ("roi_oct6_load", "roi_reg12_load"):
    make, opts, run = CASES[name]
    cloud = make()
    O = orc.Oracle(cloud, **opts); orc.set_threads(1)
    a = run(O)["tabs"].astype(np.float64)
    B = backend.Backend(cloud, rng_mode=backend.RNG_REFERENCE, **opts)
    b = run(B)["tabs"].astype(np.float64)
    co, cg = O.counters, B.counters
    rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-30)
    print(name, "packets", co.packets, cg.packets, "steps", co.steps, cg.steps, "scat", co.scatterings, cg.scatterings)
    print("  sum", a.sum(), b.sum(), "max rel", rel.max(), "n(rel>1e-4)", (rel > 1e-4).sum(), "n(rel>1e-3)", (rel > 1e-3).sum(), "of", a.size)
    B.close()
