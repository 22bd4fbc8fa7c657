#!/usr/bin/env python
"""Development tool: background launch on the bench octree (root 64^3 + 5 levels) for a set of refill thresholds
(SOC_SC_BATCH / SOC_NAV_HOPS from the environment).  Prints one line per setting."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soc_b200 import backend, synth  # noqa: E402

cloud = synth.octree_cloud(64, 6, refine_fraction=0.22, seed=12345)
dsc, csc = synth.hg_tables(0.6, 2500)
root_mean = float(np.mean(np.where(cloud.DENS[:64 ** 3] > 0, cloud.DENS[:64 ** 3], 1.0)))
k = 2.0 / (64 * root_mean)
glob = 8 * cloud.AREA
B = backend.Backend(cloud, rng_mode=backend.RNG_PACKET, noabsorbed=0)
dev = B.dev
dev.upload(backend.BUF_DSC, dsc), dev.upload(backend.BUF_CSC, csc)
for refill in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8").split(",")]:
    dev.set_tuning(0, refill, 24)
    ms = []
    for r in range(3):
        dev.zero_amc(0), dev.zero_amc(1)
        if r == 1:
            dev.reset_counters()
        dev.sim_pb(1, glob * 20, 20, 0.3 + 0.01 * r, k, k, 1.0, 1.0, glob)
        ms.append(dev.last_launch_ms())
    c = dev.counters()
    t = np.mean(ms[1:])
    print("refill %2d sc_batch %s: %.2f ms, %.3e cell-steps/s, %.3e packets/s (%s)" % (
        refill, os.environ.get("SOC_SC_BATCH", "-"), t, c.steps / 2 / t * 1e3, c.packets / 2 / t * 1e3, dev.last_kernel()), flush=True)
B.close()
