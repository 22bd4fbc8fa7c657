#!/usr/bin/env python
"""Profiling target: point-source launches of the bench workload at 256^3 (SOC_TWO_PASS / SOC_TILE_PASS from the environment)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from soc_b200 import backend  # noqa: E402

w = bench.make_workload(256)
B = backend.Backend(w["cloud"], rng_mode=backend.RNG_PACKET, **bench.REF_OPTS)
dev = B.dev
for b, a in ((backend.BUF_PSPOS, w["pspos"]), (backend.BUF_PS, w["ps"]), (backend.BUF_DSC, w["dsc"]), (backend.BUF_CSC, w["csc"])):
    dev.upload(b, a)
dev.set_tuning(2, 8, 24)
for r in range(2):
    dev.zero_amc(0), dev.zero_amc(1)
    dev.reset_counters()
    dev.sim_pb(0, w["ps_batch"] * w["ps_glob"], w["ps_batch"], 0.3 + 0.01 * r, w["kabs"], w["ksca"], 0.0, w["tw"], w["ps_glob"])
    ms = dev.last_launch_ms()
c = dev.counters()
print("PS launch: %.2f ms, %d packets, %d cell-steps, kernel %s" % (ms, c.packets, c.steps, dev.last_kernel()))
B.close()
