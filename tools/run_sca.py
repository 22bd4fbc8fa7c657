#!/usr/bin/env python
"""One scattered-light background launch (plus warm-up) on the bench octree or a regular grid, for ncu.
tools/run_sca.py [octree|regular] [packets]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soc_b200 import backend, synth
from soc_b200.hostmath import observer_directions
kind = sys.argv[1] if len(sys.argv) > 1 else "octree"
packets = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0e6
n = 64 if kind == "octree" else 128
cloud = synth.octree_cloud(64, 6, refine_fraction=0.22, seed=12345) if kind == "octree" else synth.regular_cloud(n)
dsc, csc = synth.hg_tables(0.6, 2500)
m = cloud.DENS[:n ** 3]
k = 2.0 / (n * float(np.mean(np.where(m > 0, m, 1.0))))
_, od, ra, de = observer_directions([0.0, 60.0], [0.0, 30.0])
B = backend.Backend(cloud, rng_mode=backend.RNG_PACKET, ffs=1)
dev = B.dev
for b, v in ((backend.BUF_DSC, dsc), (backend.BUF_CSC, csc), (backend.BUF_ODIR, od.reshape(-1)), (backend.BUF_ORA, ra.reshape(-1)), (backend.BUF_ODE, de.reshape(-1))):
    dev.upload(b, v)
glob = 8 * cloud.AREA
batch = max(1, int(round(packets / glob)))
npix = 512
centre = np.array([0.5 * n] * 3, np.float32)
for r in range(2):
    dev.sca_zero_out(2, npix, npix)
    dev.reset_counters()
    dev.sca_pb(1, glob * batch, batch, 0.3 + 0.01 * r, k, k, 1.0, 2, npix, npix, n / float(npix), centre, glob)
    ms = dev.last_launch_ms()
c = dev.counters()
print("%s: %.2f ms  %.3e steps/s  %.3e peels/s  %.1f steps/packet" % (kind, ms, c.steps / ms * 1e3, c.peels / ms * 1e3, c.steps / c.packets))
B.close()
