#!/usr/bin/env python
"""Secondary measurements on a BASELINE.json configs[2]-like model: octree cloud (root 64^3 + 5 refinement
levels of a log-normal turbulent field, ~1e7 cells): (a) absorption run, isotropic background, stream kernel;
(b) scattered light with peel-off towards 2 observers (ASOCS kernels), point source + background.
Prints one JSON line per measurement; the CPU column is the oracle port on a sample of the work items."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soc_b200 import backend, synth  # noqa: E402
from soc_b200.hostmath import observer_directions  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", type=int, default=64)
    ap.add_argument("--levels", type=int, default=6)
    ap.add_argument("--frac", type=float, default=0.22)
    ap.add_argument("--bg-batch", type=int, default=20)
    ap.add_argument("--sca-packets", type=float, default=1.0e7)
    ap.add_argument("--npix", type=int, default=512)
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    args = ap.parse_args()
    t0 = time.time()
    cloud = synth.octree_cloud(args.root, args.levels, refine_fraction=args.frac, seed=12345)
    print("# octree: LCELLS %s  CELLS %d  (%.1f s)" % (list(cloud.LCELLS), cloud.CELLS, time.time() - t0), flush=True)
    dsc, csc = synth.hg_tables(0.6, 2500)
    root_mean = float(np.mean(np.where(cloud.DENS[:args.root ** 3] > 0, cloud.DENS[:args.root ** 3], 1.0)))
    k = 2.0 / (args.root * root_mean)                  # tau ~ 2 across the root grid at mean density
    glob = 8 * cloud.AREA

    # ---- (a) absorption run on the octree ---------------------------------------------------------------------
    B = backend.Backend(cloud, rng_mode=backend.RNG_PACKET, noabsorbed=0)
    dev = B.dev
    dev.upload(backend.BUF_DSC, dsc), dev.upload(backend.BUF_CSC, csc)
    ms = []
    for r in range(3):
        dev.zero_amc(0), dev.zero_amc(1)
        if r == 1:
            dev.reset_counters()
        dev.sim_pb(1, glob * args.bg_batch, args.bg_batch, 0.3 + 0.01 * r, k, k, 1.0, 1.0, glob)
        ms.append(dev.last_launch_ms())
    c = dev.counters()
    t = np.mean(ms[1:]) * 1e-3
    line = {"workload": "octree %d^3 + %d levels, %d cells, isotropic background, TABS+INT" % (args.root, args.levels - 1, cloud.CELLS),
            "kernel": dev.last_kernel(), "packets_per_s": c.packets / 2 / t, "cell_steps_per_s": c.steps / 2 / t,
            "steps_per_packet": c.steps / max(1, c.packets), "ms": t * 1e3, "stuck": int(c.reserved[0])}
    B.close()
    from oracle import orc
    O = orc.Oracle(cloud, noabsorbed=0)
    orc.set_chunk(4)
    g = 4096
    while True:
        O.zero(0), O.zero(1)
        s0, p0 = O.counters.steps, O.counters.packets
        t1 = time.perf_counter()
        O.sim_pb(g, 1, g * args.bg_batch, args.bg_batch, 0.3, 1.0, 1.0, abs_=k, sca=k, dsc=dsc, csc=csc)
        dt = time.perf_counter() - t1
        if dt > 0.5 * args.cpu_seconds or g >= glob:
            break
        g = int(min(glob, g * max(2.0, 0.8 * args.cpu_seconds / max(dt, 1e-3))))
    line["cpu"] = {"kind": "port", "cores": orc.threads(), "packets_per_s": (O.counters.packets - p0) / dt,
                   "cell_steps_per_s": (O.counters.steps - s0) / dt, "sample": "%d of %d work items (%.1f s)" % (g, glob, dt)}
    print(json.dumps(line), flush=True)

    # ---- (b) scattered light ---------------------------------------------------------------------------------------
    _, od, ra, de = observer_directions([0.0, 60.0], [0.0, 30.0])
    n = args.root
    centre = np.array([0.5 * n] * 3, np.float32)
    pspos = np.array([0.5 * n + 0.3] * 3, np.float32)
    map_dx = n / float(args.npix)
    B = backend.Backend(cloud, rng_mode=backend.RNG_PACKET, no_ps=1, ffs=1)
    dev = B.dev
    for b, v in ((backend.BUF_DSC, dsc), (backend.BUF_CSC, csc), (backend.BUF_PSPOS, pspos), (backend.BUF_PS, np.ones(1, np.float32)),
                 (backend.BUF_ODIR, od.reshape(-1)), (backend.BUF_ORA, ra.reshape(-1)), (backend.BUF_ODE, de.reshape(-1))):
        dev.upload(b, v)
    gl_ps = 65536
    ps_batch = max(1, int(args.sca_packets / gl_ps))
    bg_batch = max(1, int(round(args.sca_packets / glob)))
    for name in ("ps", "bg"):
        ms = []
        for r in range(3):
            dev.sca_zero_out(2, args.npix, args.npix)
            if r == 1:
                dev.reset_counters()
            if name == "ps":
                dev.sca_ps(gl_ps * ps_batch, ps_batch, 0.3 + 0.01 * r, k, k, 2, args.npix, args.npix, map_dx, centre, gl_ps)
            else:
                dev.sca_pb(1, glob * bg_batch, bg_batch, 0.3 + 0.01 * r, k, k, 1.0, 2, args.npix, args.npix, map_dx, centre, glob)
            ms.append(dev.last_launch_ms())
        c = dev.counters()
        t = np.mean(ms[1:]) * 1e-3
        line = {"workload": "scattered light (%s), same octree, 2 observers, %dx%d px" % (name, args.npix, args.npix),
                "kernel": "sca_link_kernel", "packets_per_s": c.packets / 2 / t, "cell_steps_per_s": c.steps / 2 / t,
                "peel_rays_per_s": c.peels / 2 / t, "steps_per_packet": c.steps / max(1, c.packets), "ms": t * 1e3,
                "stuck": int(c.reserved[0])}
        O = orc.Oracle(cloud, no_ps=1, ffs=1)
        g = 1024
        while True:
            s0, p0 = O.counters.steps, O.counters.packets
            t1 = time.perf_counter()
            if name == "ps":
                O.sca_ps(g, g * ps_batch, ps_batch, 0.3, 2, args.npix, args.npix, map_dx, centre, od, ra, de, abs_=k, sca=k,
                         dsc=dsc, csc=csc, pspos=pspos, ps=np.ones(1, np.float32))
            else:
                O.sca_pb(g, 1, g * bg_batch, bg_batch, 0.3, 1.0, 2, args.npix, args.npix, map_dx, centre, od, ra, de, abs_=k,
                         sca=k, dsc=dsc, csc=csc)
            dt = time.perf_counter() - t1
            full = gl_ps if name == "ps" else glob
            if dt > 0.5 * args.cpu_seconds or g >= full:
                break
            g = int(min(full, g * max(2.0, 0.8 * args.cpu_seconds / max(dt, 1e-3))))
        line["cpu"] = {"kind": "port", "cores": orc.threads(), "packets_per_s": (O.counters.packets - p0) / dt,
                       "cell_steps_per_s": (O.counters.steps - s0) / dt, "sample": "%d work items (%.1f s)" % (g, dt)}
        print(json.dumps(line), flush=True)
    B.close()


if __name__ == "__main__":
    main()
