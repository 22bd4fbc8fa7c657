#!/usr/bin/env python
"""Benchmark of the photon-packet hot path (BASELINE.json: photon packets/s and cell-steps/s at 256^3).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                          # the reference's CPU path, host cores

Workload (BASELINE.json configs[1], absorption part): 256^3 regular grid, n = 1/(1+(r/0.1N)^2), one point
source inside the cloud + isotropic background, anisotropic scattering (HG g=0.6 .dsc tables), absorbed
energy per cell and frequency (TABS and INT).  One "step" = one frequency of ASOC.py's simulation loop
(ASOC.py:1120-1497): upload of the scattering tables, ZeroAMC, SimRAM_PB for the point source, SimRAM_PB for
the background, read-back of the per-frequency absorptions.

  value  : packets/s with everything resident in HBM (CUDA events on the library's stream, max over ranks)
  e2e    : the same step through the host-facing API with host buffers -- DSC/CSC uploaded from pinned
           memory and the INT array (4*CELLS bytes) read back to the host inside the timed region
  N > 1  : strong scaling -- the packets of the step are sharded over the ranks (packet q on rank q % N, grid
           replicated); INT of frequency f is reduced to rank 0 (ncclReduce) and read back on a second stream while
           the kernels of frequency f+1 run (two INT buffers), all inside the timed region.
  roofline        : both launches of the step, `kernel` = the one with the larger share of the step
  e2e_driver      : the same configuration through the driver itself (bin/ASOC.py: files in, files out), wall clock with the
                    driver's breakdown (context, uploads, kernels, solve, maps, file writes); N = 1 only
  extra_workloads : short runs of the other BASELINE.json configurations outside the timed headline -- 512^3
                    (configs 4/5), 256^3 with per-cell opacities (config 4 physics), the ~1e7-cell octree absorption
                    run and its scattered-light launch (config 3) -- each with its roofline fraction and the
                    reference kernels' rate on the host cores (kind "reference"); and configs[4] itself: 512^3 with
                    1e9 packets per frequency (one step, ~5 s on one GPU).  N > 1: the two 512^3 lines only (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_GRID = 256
BINS = 2500
GLOBAL_PS = 32768                 # ASOC.py:86 GLOBAL_0
PSPAC_REQ = 3.0e7                 # `pspackets 3e7`
BGPAC_REQ = 3.0e7                 # `bgpackets 3e7`
SEED = 0.4
ALG_BYTES_PER_STEP = 20           # DENS gather 4 B + TABS RMW 8 B + INT RMW 8 B (NOABSORBED=0)
REF_OPTS = dict(no_ps=1, noabsorbed=0)


def make_workload(n=N_GRID, pspac=PSPAC_REQ, bgpac_req=BGPAC_REQ):
    from soc_b200 import synth, hostmath
    cloud = synth.regular_cloud(n)
    dsc, csc = synth.hg_tables(0.6, BINS)
    k = 5.0 / n                                   # ABS = SCA = 5/N per cell and unit density
    ps_batch = int(max(1, pspac / GLOBAL_PS))                 # ASOC.py:1039
    bg_batch, bgpac, _, bg_glob = hostmath.source_weights_bg(bgpac_req, cloud.AREA)
    w = dict(cloud=cloud, dsc=dsc, csc=csc, kabs=k, ksca=k, pspos=np.array([0.5 * n + 0.3] * 3, np.float32),
             ps=np.array([1.0], np.float32), ps_batch=ps_batch, ps_glob=GLOBAL_PS, bg_batch=bg_batch, bg_glob=bg_glob,
             packets=ps_batch * GLOBAL_PS + bgpac, bg=1.0, tw=1.0)
    return w


def ref_cfg(n=N_GRID):
    return dict(NX=n, NY=n, NZ=n, LEVELS=1, CELLS=n ** 3, BINS=BINS, GL=0.01, NO_PS=1, NOABSORBED=0)


def build_reference_lib():
    from oracle import build_ref
    return build_ref.build(ref_cfg())


def host_cores():
    """All host cores this process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for every rank)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(r[2 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """DRAM bytes per launch of the bench kernels from the committed `ncu --set full` captures (profiles/)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------
# the reference's CPU implementation (oracle/_ref, or the oracle port where no library exists)
def cpu_device(cloud, opts, tuned=False):
    """(X, kind): the reference kernels compiled in place (prebuilt .so on the GPU box), else the plain-C port."""
    from oracle import orc, build_ref, ref
    cfg = dict(NX=cloud.NX, NY=cloud.NY, NZ=cloud.NZ, LEVELS=cloud.LEVELS, CELLS=cloud.CELLS, BINS=BINS, GL=0.01)
    cfg.update({k.upper(): v for k, v in opts.items()})
    if tuned:
        cfg["TUNED"] = 1
    try:
        if build_ref.build(cfg) is not None:
            X = ref.Reference(cloud, **dict(opts, **({"tuned": 1} if tuned else {})))
            X.set_threads(host_cores())
            X.set_chunk(4)
            return X, "reference"
    except Exception:
        pass
    if tuned:
        return None, None
    X = orc.Oracle(cloud, **opts)
    orc.set_threads(host_cores())
    orc.set_chunk(4)
    return X, "port"


def cpu_sample(X, kind, launches, seconds_target):
    """Times a bounded sample of `launches` = [(full work items, packets per work item, call(n items))]: n work items
    of every launch, spread over the whole launch with a stride (the reference build; the port takes the first n),
    each with the job's own BATCH so that the per-work-item MWC64X seeding keeps its real share.  The sample is sized
    from a ~1 s calibration.  Returns packets/s, cell-steps/s, description."""
    from oracle import orc
    cores = X.threads() if kind == "reference" else orc.threads()

    def run(frac):
        X.zero(0), X.zero(1)
        if kind == "reference":
            X.atomic_count(reset=True)
        s0 = None if kind == "reference" else X.counters.steps
        packets, parts = 0, []
        t0 = time.perf_counter()
        for full, per_item, call in launches:
            n = int(min(full, max(cores, round(full * frac))))
            if kind == "reference":
                X.set_sampling(max(1, full // n), 0, full)
            call(n)
            packets += n * per_item
            parts.append("%d of %d work items x %d" % (n, full, per_item))
        dt = time.perf_counter() - t0
        if kind == "reference":
            X.set_sampling(1, 0, 0)
        return dt, packets, (X.atomic_count() if kind == "reference" else X.counters.steps - s0), parts

    frac = max(4.0 * cores / min(l[0] for l in launches), 2e-4)
    dt, _, _, _ = run(frac)
    frac = min(1.0, frac * max(1.0, seconds_target / max(dt, 1e-3)))
    dt, packets, adds, parts = run(frac)
    return packets / dt, adds, dt, cores, "; ".join(parts) + " (stride over the launch, %.1f s)" % dt


def cpu_leg(w, seconds_target=12.0, tuned=False, opts=REF_OPTS, opt=None, adds_per_step=2):
    """The bench step (point-source + background launch) on the host cores.  Returns a cpu_baseline dict."""
    X, kind = cpu_device(w["cloud"], opts, tuned)
    if X is None:
        return None
    common = dict(abs_=w["kabs"], sca=w["ksca"], dsc=w["dsc"], csc=w["csc"], opt=opt)
    launches = [
        (w["bg_glob"], w["bg_batch"], lambda n: X.sim_pb(n, 1, n * w["bg_batch"], w["bg_batch"], SEED, w["bg"], w["tw"], **common)),
        (w["ps_glob"], w["ps_batch"], lambda n: X.sim_pb(n, 0, n * w["ps_batch"], w["ps_batch"], SEED, 0.0, w["tw"],
                                                        pspos=w["pspos"], ps=w["ps"], **common)),
    ]
    pps, adds, dt, cores, sample = cpu_sample(X, kind, launches, seconds_target)
    steps = adds / adds_per_step if kind == "reference" else adds           # TABS + INT add per cell-step
    from oracle import build_ref
    return {"value": pps, "unit": "packets/s", "cell_steps_per_s": steps / dt, "cores": cores, "kind": kind,
            "build": " ".join(build_ref.OPT_FLAGS[bool(tuned)]) if kind == "reference" else "gcc -O2 (oracle port)", "sample": sample}


# ---------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload()
    vals, steps_rate = [], []
    info = None
    for i in range(args.warmup + args.steps):
        # each step is a bounded sample sized so that the whole run ends within a few minutes
        b = cpu_leg(w, seconds_target=max(4.0, min(12.0, 120.0 / (args.warmup + args.steps))), tuned=True) or cpu_leg(w, 6.0)
        if i >= args.warmup:
            vals.append(b["value"])
            steps_rate.append(b["cell_steps_per_s"])
        info = b
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "photon_packets_per_s", "value": v, "unit": "packets/s",
        "cell_steps_per_s": float(np.mean(steps_rate)),
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * w["packets"] / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(w),
        "cpu_baseline": dict(info, value=v),
        "e2e": {"value": v, "unit": "packets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(w):
    return {"workload": "256^3 regular grid, point source + isotropic background, HG g=0.6 dsc tables, "
                        "TABS+INT absorptions, one frequency per step",
            "grid": "%d^3" % w["cloud"].NX, "packets_per_step": int(w["packets"]),
            "ps_packets": int(w["ps_batch"] * w["ps_glob"]), "bg_packets": int(w["packets"] - w["ps_batch"] * w["ps_glob"]),
            "kabs=ksca": "5/N per cell", "l2": "DENS+TABS+INT = 192 MiB > 126 MB L2 (inputs larger than L2)"}


# ---------------------------------------------------------------------------------------------------------
def _upload_sources(dev, w, backend):
    dev.upload(backend.BUF_PSPOS, w["pspos"])
    dev.upload(backend.BUF_PS, w["ps"])
    dev.upload(backend.BUF_DSC, w["dsc"])
    dev.upload(backend.BUF_CSC, w["csc"])


def _launch_rooflines(dev, w, alg_bytes, peak, reps=2):
    """Each launch of the step alone: library CUDA events around the launch, cell-steps from the work counters."""
    out = []
    for name, src, batch, glob, bg in (("point-source launch", 0, w["ps_batch"], w["ps_glob"], 0.0),
                                       ("background launch", 1, w["bg_batch"], w["bg_glob"], w["bg"])):
        dev.zero_amc(1)
        dev.sim_pb(src, batch * glob, batch, SEED + 0.05, w["kabs"], w["ksca"], bg, w["tw"], glob)      # warm
        dev.sync()
        dev.reset_counters()
        ms = []
        for i in range(reps):
            dev.sim_pb(src, batch * glob, batch, SEED + 0.01 * i, w["kabs"], w["ksca"], bg, w["tw"], glob)
            ms.append(dev.last_launch_ms())
        c = dev.counters()
        steps, t = c.steps / reps, float(np.mean(ms))
        ach = steps * alg_bytes / (t * 1e-3) / 1e9
        out.append({"launch": name, "kernel": dev.last_kernel(), "kernel_ms": t, "cell_steps_per_launch": steps,
                    "cell_steps_per_s": steps / (t * 1e-3), "achieved": ach, "frac": ach / peak})
    return out


def extra_grid(backend, local, rank, world, n, peak, with_abu=False, steps=1, cpu_seconds=4.0, allreduce=None, pac_req=None, warm=True):
    """One step (PS + BG launch) of the bench workload on an n^3 grid, optionally with per-cell opacities.
    pac_req = (point-source, background) packets requested per step instead of the bench's 3e7 + 3e7."""
    w = make_workload(n) if pac_req is None else make_workload(n, pac_req[0], pac_req[1])
    cloud = w["cloud"]
    opts = dict(REF_OPTS, **({"with_abu": 1} if with_abu else {}))
    B = backend.Backend(cloud, ordinal=local, rng_mode=backend.RNG_PACKET, **opts)
    dev = B.dev
    dev.set_shard(rank, world)
    _upload_sources(dev, w, backend)
    opt = None
    if with_abu:          # smooth abundance gradient: kabs, ksca vary by +-30 % across the cloud
        z = np.linspace(0.7, 1.3, n, dtype=np.float32)
        opt = np.empty((n, n, n, 2), np.float32)
        opt[..., 0] = w["kabs"] * z[:, None, None]
        opt[..., 1] = w["ksca"] * z[None, :, None]
        opt = opt.reshape(-1)
        dev.upload(backend.BUF_OPT, opt)
    alg = ALG_BYTES_PER_STEP + (8 if with_abu else 0)

    def step(seed):
        dev.zero_amc(1)
        dev.sim_pb(0, w["ps_batch"] * w["ps_glob"], w["ps_batch"], seed, w["kabs"], w["ksca"], 0.0, w["tw"], w["ps_glob"])
        t = dev.last_launch_ms() if world == 1 else 0.0
        dev.sim_pb(1, w["bg_batch"] * w["bg_glob"], w["bg_batch"], seed, w["kabs"], w["ksca"], w["bg"], w["tw"], w["bg_glob"])
        t += dev.last_launch_ms() if world == 1 else 0.0
        if allreduce is not None:
            allreduce(dev, cloud.CELLS)
        return t

    if warm:
        step(0.9)
    else:                 # a long step: warm up on one short launch per source (same kernels, same buffers)
        dev.zero_amc(1)
        dev.sim_pb(0, w["ps_glob"], 1, 0.9, w["kabs"], w["ksca"], 0.0, w["tw"], w["ps_glob"])
        dev.sim_pb(1, w["bg_glob"], 1, 0.9, w["kabs"], w["ksca"], w["bg"], w["tw"], w["bg_glob"])
    dev.sync()
    dev.reset_counters()
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kms = sum(step(SEED + 0.001 * i) for i in range(steps))
    dev.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = (time.perf_counter() - t0) * 1e3
    c = dev.counters()
    counts = torch.tensor([c.packets, c.steps, wall], dtype=torch.float64, device=torch.device("cuda", local))
    if world > 1:
        tmax = counts[2:3].clone()
        dist.all_reduce(counts[:2])
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        counts[2] = tmax[0]
    packets, csteps, wall = [float(x) for x in counts.tolist()]
    ms = kms if world == 1 else wall             # 1 GPU: the library's CUDA events around the launches; N GPUs: wall clock between barriers
    line = {"workload": "%d^3 regular grid, point source + isotropic background%s, TABS+INT%s" % (
                n, ", per-cell opacities (WITH_ABU)" if with_abu else "",
                "" if pac_req is None else ", %.3g packets per frequency (BASELINE.json configs[4])" % w["packets"]),
            "packets_per_step": w["packets"],
            "kernel": dev.last_kernel(), "n_gpus": world, "packets_per_s": packets / (ms * 1e-3), "cell_steps_per_s": csteps / (ms * 1e-3),
            "ms_per_step": ms / steps, "steps": steps,
            "roofline": {"bound": "hbm", "achieved": csteps * alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": csteps * alg / (ms * 1e-3) / 1e9 / peak, "alg_bytes_per_cell_step": alg}}
    B.close()
    if rank == 0 and world == 1 and cpu_seconds > 0:
        line["cpu"] = cpu_leg(w, cpu_seconds, tuned=False, opts=opts, opt=opt)
    return line


def extra_octree(backend, local, peak, cpu_seconds=4.0):
    """BASELINE.json configs[2]: octree, root 64^3 + 5 levels (~1e7 cells): (a) absorption run, isotropic background;
    (b) scattered light with peel-off towards 2 observers of 512^2 pixels, point source (ASOCS.py kernels)."""
    from soc_b200 import synth
    from soc_b200.hostmath import observer_directions
    cloud = synth.octree_cloud(64, 6, refine_fraction=0.22, seed=12345)
    n = cloud.NX
    dsc, csc = synth.hg_tables(0.6, BINS)
    root_mean = float(np.mean(np.where(cloud.DENS[:n ** 3] > 0, cloud.DENS[:n ** 3], 1.0)))
    k = 2.0 / (n * root_mean)
    glob, batch = 8 * cloud.AREA, 20
    lines = []
    B = backend.Backend(cloud, ordinal=local, rng_mode=backend.RNG_PACKET, noabsorbed=0)
    dev = B.dev
    dev.upload(backend.BUF_DSC, dsc), dev.upload(backend.BUF_CSC, csc)
    ms = []
    for r in range(3):
        dev.zero_amc(0), dev.zero_amc(1)
        if r == 1:
            dev.reset_counters()
        dev.sim_pb(1, glob * batch, batch, 0.3 + 0.01 * r, k, k, 1.0, 1.0, glob)
        ms.append(dev.last_launch_ms())
    c = dev.counters()
    t = float(np.mean(ms[1:])) * 1e-3
    lines.append({"workload": "octree 64^3 + 5 levels (%d cells), isotropic background, TABS+INT" % cloud.CELLS,
                  "kernel": dev.last_kernel(), "packets_per_s": c.packets / 2 / t, "cell_steps_per_s": c.steps / 2 / t,
                  "ms_per_launch": t * 1e3,
                  "roofline": {"bound": "hbm", "achieved": c.steps / 2 * 20 / t / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": c.steps / 2 * 20 / t / 1e9 / peak, "alg_bytes_per_cell_step": 20}})
    B.close()
    if cpu_seconds > 0:
        X, kind = cpu_device(cloud, dict(no_ps=1, noabsorbed=0))
        pps, adds, dt, cores, sample = cpu_sample(X, kind, [(glob, batch, lambda m: X.sim_pb(m, 1, m * batch, batch, 0.3, 1.0, 1.0, abs_=k, sca=k, dsc=dsc, csc=csc))], cpu_seconds)
        lines[-1]["cpu"] = {"value": pps, "unit": "packets/s", "cell_steps_per_s": (adds / 2 if kind == "reference" else adds) / dt,
                            "cores": cores, "kind": kind, "sample": sample}
    # scattered light
    _, od, ra, de = observer_directions([0.0, 60.0], [0.0, 30.0])
    npix = 512
    centre = np.array([0.5 * n] * 3, np.float32)
    pspos = np.array([0.5 * n + 0.3] * 3, np.float32)
    map_dx = n / float(npix)
    B = backend.Backend(cloud, ordinal=local, rng_mode=backend.RNG_PACKET, no_ps=1, ffs=1)
    dev = B.dev
    for b, v in ((backend.BUF_DSC, dsc), (backend.BUF_CSC, csc), (backend.BUF_PSPOS, pspos), (backend.BUF_PS, np.ones(1, np.float32)),
                 (backend.BUF_ODIR, od.reshape(-1)), (backend.BUF_ORA, ra.reshape(-1)), (backend.BUF_ODE, de.reshape(-1))):
        dev.upload(b, v)
    gl_ps, ps_batch = 65536, 152
    ms = []
    for r in range(3):
        dev.sca_zero_out(2, npix, npix)
        if r == 1:
            dev.reset_counters()
        dev.sca_ps(gl_ps * ps_batch, ps_batch, 0.3 + 0.01 * r, k, k, 2, npix, npix, map_dx, centre, gl_ps)
        ms.append(dev.last_launch_ms())
    c = dev.counters()
    t = float(np.mean(ms[1:])) * 1e-3
    lines.append({"workload": "scattered light (ASOCS SimRAM_PS), same octree, point source, 2 observers of 512x512 px",
                  "kernel": "sca_link_kernel", "packets_per_s": c.packets / 2 / t, "cell_steps_per_s": c.steps / 2 / t,
                  "peel_rays_per_s": c.peels / 2 / t, "ms_per_launch": t * 1e3,
                  "roofline": {"bound": "hbm", "achieved": c.steps / 2 * 8 / t / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": c.steps / 2 * 8 / t / 1e9 / peak, "alg_bytes_per_cell_step": 8,
                               "note": "4 B density + 4 B neighbour-table entry per cell-step (SURVEY 8d: 4 B + 4 B per level change)"}})
    B.close()
    if cpu_seconds > 0:
        X, kind = cpu_device(cloud, dict(no_ps=1, ffs=1))
        t0 = time.perf_counter()
        m = max(64, 4 * host_cores())
        while True:
            if kind == "reference":
                X.set_sampling(max(1, gl_ps // m), 0, gl_ps)
            t0 = time.perf_counter()
            X.sca_ps(m, m * ps_batch, ps_batch, 0.3, 2, npix, npix, map_dx, centre, od, ra, de, abs_=k, sca=k, dsc=dsc, csc=csc,
                     pspos=pspos, ps=np.ones(1, np.float32))
            dt = time.perf_counter() - t0
            if dt > 0.4 * cpu_seconds or m >= gl_ps:
                break
            m = int(min(gl_ps, m * max(2.0, 0.8 * cpu_seconds / max(dt, 1e-3))))
        if kind == "reference":
            X.set_sampling(1, 0, 0)
        lines[-1]["cpu"] = {"value": m * ps_batch / dt, "unit": "packets/s", "cores": host_cores(), "kind": kind,
                            "sample": "%d of %d work items x %d (stride over the launch, %.1f s)" % (m, gl_ps, ps_batch, dt)}
    return lines


def driver_e2e(n=N_GRID, nfreq=44, workdir=None, device_factory=None):
    """BASELINE.json configs[1] through the reference-facing entry point, the ASOC driver itself (bin/ASOC.py <ini>, files
    in, files out): 256^3 cloud file, 44 frequencies, point source + background with 3e7 packets each per frequency.
    Two runs, as the reference is used (the driver, like ASOC.py, refuses to write the absorbed file and solve in one go):
      absorbed : `nosolve`, `absorbed abs.data` -- the [CELLS, NFREQ] absorptions for an external dust solver (A2E);
      maps     : `noabsorbed`, CLT / CLE -- absorptions integrated on the fly, equilibrium temperatures and emission on
                 the device, three 256^2 maps per frequency.
    Wall clock of each run in this process with the driver's own breakdown."""
    import contextlib
    import io
    import shutil
    import tempfile
    from soc_b200 import asoc, synth
    d = workdir or tempfile.mkdtemp(prefix="soc_c2_")
    runs = {}
    try:
        for name in ("absorbed", "maps"):
            t0 = time.perf_counter()
            extra = "" if name == "absorbed" else ("CLT\nCLE\nmapping %d %d 1.0\ndirections 0.0 0.0\ndirections 90.0 0.0\ndirections 60.0 30.0\n" % (n, n))
            ini, cloud = synth.write_model(d, n=n, nfreq=nfreq, bgpac=int(BGPAC_REQ), pspac=int(PSPAC_REQ), noabsorbed=(name == "maps"),
                                           absorbed=(name == "absorbed"), maps=False, extra=extra)
            if name == "maps":          # write_model(maps=False) writes `nomap`; this run has its own three directions
                txt = open(ini).read().replace("nomap\n", "")
                open(ini, "w").write(txt)
            t_model = time.perf_counter() - t0
            cwd = os.getcwd()
            os.chdir(d)
            try:
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(io.StringIO()):
                    asoc.main(["ASOC.py", "model.ini"], device_factory=device_factory)
                wall = time.perf_counter() - t0
            finally:
                os.chdir(cwd)
            t = dict(asoc.LAST_TIMINGS)
            outputs = ("abs.data",) if name == "absorbed" else ("emit.data", "model.T", "map_dir_00.bin", "map_dir_01.bin", "map_dir_02.bin")
            nbytes = int(sum(os.path.getsize(os.path.join(d, f)) for f in outputs if os.path.exists(os.path.join(d, f))))
            packets = t.pop("packets", 0)
            runs[name] = {"wall_s": round(wall, 3), "packets": packets, "packets_per_s": packets / wall, "cell_steps": t.pop("cell_steps", 0),
                          "breakdown_s": {k: round(v, 3) for k, v in t.items()}, "output_bytes": nbytes,
                          "model_files_written_in_s": round(t_model, 2)}
    finally:
        if workdir is None:
            shutil.rmtree(d, ignore_errors=True)
    runs["workload"] = ("bin/ASOC.py: %d^3 cloud, %d frequencies x (3e7 point-source + 3e7 background packets); run `absorbed` writes the "
                        "absorbed file, run `maps` solves temperatures / emission and writes 3 maps of %dx%d px per frequency" % (n, nfreq, n, n))
    return runs


# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from soc_b200 import backend

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local)
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=device)
    elif not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")

    w = make_workload()
    cloud = w["cloud"]
    B = backend.Backend(cloud, ordinal=local, rng_mode=backend.RNG_PACKET, **REF_OPTS)
    dev = B.dev
    dev.set_shard(rank, world)
    dev.set_tuning(deposit=args.deposit, refill=args.refill, aggregate_steps=args.agg_steps)
    stream = torch.cuda.ExternalStream(dev.stream(), device=device)
    side = torch.cuda.Stream(device=device)          # reduction to rank 0 and read-back of frequency f under the kernels of f+1
    n = cloud.CELLS

    # device-resident inputs
    _upload_sources(dev, w, backend)
    dev.zero_amc(0)
    dev.zero_amc(1)
    dev.sync()

    class _Raw:          # wraps a device buffer of the library as a torch tensor (no copy)
        def __init__(self, ptr, count):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    int_ptr, _ = dev.device_ptr(backend.BUF_INT)
    int_t = torch.as_tensor(_Raw(int_ptr, n), device=device)
    int_copy = [torch.empty(n, dtype=torch.float32, device=device) for _ in range(2)]      # INT of the last two frequencies
    copied = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    # pinned host buffers of the end-to-end leg
    tables = torch.empty(2 * BINS, dtype=torch.float32).pin_memory()
    tables[:BINS] = torch.from_numpy(w["dsc"])
    tables[BINS:] = torch.from_numpy(w["csc"])
    tables_np = tables.numpy()
    int_host = torch.empty(n, dtype=torch.float32).pin_memory()
    state = {"i": 0}

    def step(e2e, seed):
        i = state["i"] & 1
        state["i"] += 1
        if e2e:
            dev.upload(backend.BUF_DSC, tables_np[:BINS])
            dev.upload(backend.BUF_CSC, tables_np[BINS:])
        dev.zero_amc(1)
        dev.sim_pb(0, w["ps_batch"] * w["ps_glob"], w["ps_batch"], seed, w["kabs"], w["ksca"], 0.0, w["tw"], w["ps_glob"])
        dev.sim_pb(1, w["bg_batch"] * w["bg_glob"], w["bg_batch"], seed, w["kabs"], w["ksca"], w["bg"], w["tw"], w["bg_glob"])
        if world > 1 or e2e:
            # INT of this frequency leaves the library's buffer with one device copy (0.03 ms); the collective and the
            # read-back run on the side stream while the main stream goes on with the next frequency
            with torch.cuda.stream(stream):
                stream.wait_event(freed[i])
                int_copy[i].copy_(int_t, non_blocking=True)
                copied[i].record(stream)
            with torch.cuda.stream(side):
                side.wait_event(copied[i])
                if world > 1:
                    dist.reduce(int_copy[i], dst=0)
                if e2e and rank == 0:
                    int_host.copy_(int_copy[i], non_blocking=True)
                freed[i].record(side)

    def timed(e2e, nsteps, sample_clocks):
        sampler = None
        dev.sync()
        side.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sample_clocks and rank == 0:
            sampler = ClockSampler(local)
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
        t0 = time.perf_counter()
        for i in range(nsteps):
            step(e2e, SEED + 0.001 * i)
        with torch.cuda.stream(stream):
            stream.wait_stream(side)                 # the step is done when its absorptions have arrived
            e1.record(stream)
        dev.sync()
        side.synchronize()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if e2e:
            ms = max(ms, 1e3 * wall)          # host copies are part of the end-to-end step
        clocks = sampler.summary() if sampler else None
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), clocks

    for i in range(args.warmup):
        step(False, 0.9 - 0.001 * i)
    dev.sync()
    side.synchronize()
    dev.reset_counters()
    ms, clocks = timed(False, args.steps, True)
    c = dev.counters()
    counts = torch.tensor([c.packets, c.steps, c.scatterings, c.reserved[0]], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(counts)
    packets, csteps = counts[0].item(), counts[1].item()
    launches_per_run = int(c.launches)
    # end to end
    step(True, 0.77)
    ems, _ = timed(True, args.steps, False)
    # the two launches of the step alone
    peak, peak_src = measured_peak()
    per_launch = _launch_rooflines(dev, w, ALG_BYTES_PER_STEP, peak)
    B.close()

    # the other configurations (short runs, outside the timed headline)
    extras = []
    if not args.no_extras:
        def reduce_int(d, cells):
            p, _ = d.device_ptr(backend.BUF_INT)
            t = torch.as_tensor(_Raw(p, cells), device=device)
            with torch.cuda.stream(torch.cuda.ExternalStream(d.stream(), device=device)):
                dist.reduce(t, dst=0)
        try:
            extras.append(extra_grid(backend, local, rank, world, 512, peak, steps=1, cpu_seconds=0 if args.no_cpu else 4.0,
                                     allreduce=reduce_int if world > 1 else None))
            # configs[4]: 512^3, 1e9 packets per frequency, packet-sharded over the ranks, INT reduced to rank 0 (strong scaling)
            extras.append(extra_grid(backend, local, rank, world, 512, peak, steps=1, cpu_seconds=0, pac_req=(5.0e8, 5.0e8), warm=False,
                                     allreduce=reduce_int if world > 1 else None))
            if world == 1:
                extras.append(extra_grid(backend, local, rank, world, 256, peak, with_abu=True, steps=2, cpu_seconds=0 if args.no_cpu else 4.0))
                extras += extra_octree(backend, local, peak, cpu_seconds=0 if args.no_cpu else 4.0)
        except Exception as e:          # an extra must not take the headline down
            extras.append({"error": "%s: %s" % (type(e).__name__, e)})

    if rank == 0:
        traffic = measured_traffic() if world == 1 else None
        top = max(per_launch, key=lambda r: r["kernel_ms"])
        line = {
            "metric": "photon_packets_per_s", "value": packets / (ms * 1e-3), "unit": "packets/s",
            "cell_steps_per_s": csteps / (ms * 1e-3), "steps_per_packet": csteps / max(packets, 1.0),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(w),
            "clocks": clocks,
            "e2e": {"value": packets / (ems * 1e-3), "unit": "packets/s", "h2d_bytes_per_step": 2 * BINS * 4,
                    "d2h_bytes_per_step": 4 * n, "ms_per_step": ems / args.steps},
            "gpu_launches": launches_per_run * world,          # counted by the library: 2 packet kernels + 2 folds per step and rank
            "roofline": {"bound": "hbm", "achieved": top["achieved"], "peak": peak, "unit": "GB/s", "frac": top["frac"],
                         "traffic": ((traffic or {}).get(top["launch"]) or {}).get("dram_bytes_per_launch"),
                         "traffic_source": ((traffic or {}).get(top["launch"]) or {}).get("source"),
                         "kernel": "%s (%s, %.0f %% of the step)" % (top["kernel"], top["launch"],
                                                                    100.0 * top["kernel_ms"] / sum(r["kernel_ms"] for r in per_launch)),
                         "kernel_ms": top["kernel_ms"], "cell_steps_per_launch": top["cell_steps_per_launch"],
                         "alg_bytes_per_cell_step": ALG_BYTES_PER_STEP, "peak_source": peak_src,
                         "step_frac": csteps / args.steps * ALG_BYTES_PER_STEP / (ms / args.steps * 1e-3) / 1e9 / peak,
                         "launches": per_launch},
            "stuck_packets": counts[3].item(),
            "extra_workloads": extras,
        }
        if world == 1 and not args.no_driver:
            try:
                line["e2e_driver"] = driver_e2e()
            except Exception as e:
                line["e2e_driver"] = {"error": "%s: %s" % (type(e).__name__, e)}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_leg(w, seconds_target=10.0, tuned=False)
            tuned = cpu_leg(w, seconds_target=8.0, tuned=True)
            if tuned is not None:
                line["cpu_baseline_tuned"] = tuned
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--deposit", type=int, default=2)
    ap.add_argument("--refill", type=int, default=8)
    ap.add_argument("--agg-steps", type=int, default=24)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-driver", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
