#!/usr/bin/env python
"""Development tool: one background launch on an nx*ny*nz regular grid (Plummer-like density), timed with the
settings of the environment (SOC_DOMAINS, SOC_SCRAMBLE, ...).  Used to separate layout effects of the domain mode."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soc_b200 import backend, synth  # noqa: E402
from soc_b200.formats import Cloud  # noqa: E402

nx, ny, nz = [int(v) for v in sys.argv[1:4]]
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 8
c = [np.arange(n, dtype=np.float32) + 0.5 - 0.5 * n for n in (nx, ny, nz)]
z, y, x = np.meshgrid(c[2], c[1], c[0], indexing="ij")
d = (1.0 / (1.0 + (x * x + y * y + z * z) / (0.1 * max(nx, ny, nz)) ** 2)).astype(np.float32)
cloud = Cloud(nx, ny, nz, [nx * ny * nz], d.ravel())
dsc, csc = synth.hg_tables(0.6)
B = backend.Backend(cloud, rng_mode=backend.RNG_PACKET, noabsorbed=0)
dev = B.dev
dev.upload(backend.BUF_DSC, dsc), dev.upload(backend.BUF_CSC, csc)
glob = 8 * cloud.AREA
k = 5.0 / max(nx, ny, nz)
ms = []
for r in range(3):
    dev.zero_amc(0), dev.zero_amc(1)
    if r == 1:
        dev.reset_counters()
    dev.sim_pb(1, glob * batch, batch, 0.3 + 0.01 * r, k, k, 1.0, 1.0, glob)
    ms.append(dev.last_launch_ms())
cnt = dev.counters()
t = np.mean(ms[1:])
print("%dx%dx%d BG %d packets: %.2f ms, %.3e cell-steps/s, sum %.6e" % (nx, ny, nz, glob * batch, t, cnt.steps / 2 / t * 1e3,
                                                                         float(B.tabs.astype(np.float64).sum())), flush=True)
B.close()
