#!/usr/bin/env python
"""One background launch (plus warm-up) on the bench octree, for ncu.  tools/run_octree.py [bg_batch] [geometry]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soc_b200 import backend, synth
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 50
geo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cloud = synth.octree_cloud(64, 6, refine_fraction=0.22, seed=12345)
dsc, csc = synth.hg_tables(0.6, 2500)
m = cloud.DENS[:64 ** 3]
k = 2.0 / (64 * float(np.mean(np.where(m > 0, m, 1.0))))
B = backend.Backend(cloud, rng_mode=backend.RNG_PACKET, noabsorbed=0)
B.dev.set_geometry(geo)
B.dev.upload(backend.BUF_DSC, dsc), B.dev.upload(backend.BUF_CSC, csc)
glob = 8 * cloud.AREA
for r in range(2):
    B.dev.reset_counters()
    B.dev.sim_pb(1, glob * batch, batch, 0.3 + 0.01 * r, k, k, 1.0, 1.0, glob)
    ms = B.dev.last_launch_ms()
c = B.dev.counters()
print("geometry %d: %.2f ms  %.3e steps/s  %.1f steps/packet" % (geo, ms, c.steps / ms * 1e3, c.steps / c.packets))
B.close()
