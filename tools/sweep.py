#!/usr/bin/env python
"""Times the two launches of the bench step (point source, background) on the 256^3 bench grid for a set of
accumulation-engine / refill settings.  Development tool; prints one line per setting."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from soc_b200 import backend  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--deposit", default="0,1,2")
    ap.add_argument("--refill", default="8")
    ap.add_argument("--agg", default="24")
    ap.add_argument("--opts", default="noabsorbed=0")
    args = ap.parse_args()
    w = bench.make_workload(args.n)
    opts = dict(no_ps=1)
    for kv in args.opts.split(","):
        if kv:
            k, v = kv.split("=")
            opts[k] = int(v)
    B = backend.Backend(w["cloud"], rng_mode=backend.RNG_PACKET, **opts)
    dev = B.dev
    for b, a in ((backend.BUF_PSPOS, w["pspos"]), (backend.BUF_PS, w["ps"]), (backend.BUF_DSC, w["dsc"]), (backend.BUF_CSC, w["csc"])):
        dev.upload(b, a)
    if opts.get("with_abu"):          # per-cell opacities holding the same constants: same physics through the general kernel
        n = w["cloud"].CELLS
        o = np.empty((n, 2), np.float32)
        o[:, 0], o[:, 1] = w["kabs"], w["ksca"]
        dev.upload(backend.BUF_OPT, o.reshape(-1))
    for dep in [int(x) for x in args.deposit.split(",")]:
        for refill in [int(x) for x in args.refill.split(",")]:
            for agg in [int(x) for x in args.agg.split(",")]:
                dev.set_tuning(dep, refill, agg)
                res = {}
                for name, src, batch, glob, bg in (("ps", 0, w["ps_batch"], w["ps_glob"], 0.0), ("bg", 1, w["bg_batch"], w["bg_glob"], w["bg"])):
                    ms = []
                    dev.reset_counters()
                    for r in range(args.reps + 1):
                        dev.zero_amc(0), dev.zero_amc(1)
                        if r == 1:
                            dev.reset_counters()
                        dev.sim_pb(src, batch * glob, batch, 0.3 + 0.01 * r, w["kabs"], w["ksca"], bg, w["tw"], glob)
                        ms.append(dev.last_launch_ms())
                    c = dev.counters()
                    res[name] = (np.mean(ms[1:]), c.steps / args.reps, c.packets / args.reps, float(B.tabs.astype(np.float64).sum()))
                print("deposit=%d refill=%2d agg=%3d | PS %8.2f ms %.3e steps/s (%.0f steps/pkt) sum %.6e | BG %8.2f ms %.3e steps/s sum %.6e" % (
                    dep, refill, agg, res["ps"][0], res["ps"][1] / res["ps"][0] * 1e3, res["ps"][1] / res["ps"][2], res["ps"][3],
                    res["bg"][0], res["bg"][1] / res["bg"][0] * 1e3, res["bg"][3]), flush=True)
    B.close()


if __name__ == "__main__":
    main()
