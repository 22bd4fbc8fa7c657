#!/usr/bin/env python
"""Profiling target: ONE background launch of the bench workload at 512^3 (domain-tiled propagation unless
SOC_DOMAINS=-1), preceded by one warm-up launch.  Prints the launch time and the work counters."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from soc_b200 import backend  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
w = bench.make_workload(n)
B = backend.Backend(w["cloud"], rng_mode=backend.RNG_PACKET, **bench.REF_OPTS)
dev = B.dev
dev.upload(backend.BUF_DSC, w["dsc"]), dev.upload(backend.BUF_CSC, w["csc"])
for r in range(2):
    dev.zero_amc(0), dev.zero_amc(1)
    dev.reset_counters()
    dev.sim_pb(1, w["bg_batch"] * w["bg_glob"], w["bg_batch"], 0.3 + 0.01 * r, w["kabs"], w["ksca"], w["bg"], w["tw"], w["bg_glob"])
    ms = dev.last_launch_ms()
c = dev.counters()
print("%d^3 BG launch: %.2f ms, %d packets, %d cell-steps, %.3e cell-steps/s, kernel %s" % (n, ms, c.packets, c.steps, c.steps / ms * 1e3, dev.last_kernel()))
B.close()
