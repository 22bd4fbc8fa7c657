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
           replicated) and INT is combined with one NCCL all-reduce per step, inside the timed region.
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


def make_workload(n=N_GRID):
    from soc_b200 import synth, hostmath
    cloud = synth.regular_cloud(n)
    dsc, csc = synth.hg_tables(0.6, BINS)
    k = 5.0 / n                                   # ABS = SCA = 5/N per cell and unit density
    ps_batch = int(max(1, PSPAC_REQ / GLOBAL_PS))             # ASOC.py:1039
    bg_batch, bgpac, _, bg_glob = hostmath.source_weights_bg(BGPAC_REQ, cloud.AREA)
    w = dict(cloud=cloud, dsc=dsc, csc=csc, kabs=k, ksca=k, pspos=np.array([0.5 * n + 0.3] * 3, np.float32),
             ps=np.array([1.0], np.float32), ps_batch=ps_batch, ps_glob=GLOBAL_PS, bg_batch=bg_batch, bg_glob=bg_glob,
             packets=ps_batch * GLOBAL_PS + bgpac, bg=1.0, tw=1.0)
    return w


def ref_cfg(n=N_GRID):
    return dict(NX=n, NY=n, NZ=n, LEVELS=1, CELLS=n ** 3, BINS=BINS, GL=0.01, NO_PS=1, NOABSORBED=0)


def build_reference_lib():
    from oracle import build_ref
    return build_ref.build(ref_cfg())


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
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def cpu_leg(w, seconds_target=12.0, use_ref=True):
    """Times the reference's CPU implementation of the same step on a bounded sample of its work items.
    Returns packets/s, cell-steps/s, cores, kind, sample description."""
    from oracle import orc, build_ref
    cloud = w["cloud"]
    kind, X = "port", None
    if use_ref:
        path = build_ref.build(ref_cfg(cloud.NX))
        if path is not None:
            from oracle import ref
            X = ref.Reference(cloud, **REF_OPTS)
            kind = "reference"
    if X is None:
        X = orc.Oracle(cloud, **REF_OPTS)
    # all host cores this process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for every rank)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    (X.set_threads if kind == "reference" else orc.set_threads)(max(1, ncpu))
    cores = X.threads() if kind == "reference" else orc.threads()
    (X.set_chunk if kind == "reference" else orc.set_chunk)(4)
    common = dict(abs_=w["kabs"], sca=w["ksca"], dsc=w["dsc"], csc=w["csc"])

    def sample(g, gp):
        """g background work items and gp point-source work items, each with the job's own BATCH (so the
        per-work-item MWC64X seeding keeps its real share).  Returns seconds, packets, cell-steps."""
        X.zero(0), X.zero(1)
        if kind == "reference":
            X.atomic_count(reset=True)
        else:
            s0 = X.counters.steps
        t0 = time.perf_counter()
        X.sim_pb(g, 1, g * w["bg_batch"], w["bg_batch"], SEED, w["bg"], w["tw"], **common)
        X.sim_pb(gp, 0, gp * w["ps_batch"], w["ps_batch"], SEED, 0.0, w["tw"], pspos=w["pspos"], ps=w["ps"], **common)
        dt = time.perf_counter() - t0
        steps = X.atomic_count() // 2 if kind == "reference" else X.counters.steps - s0   # TABS + INT add per step
        return dt, g * w["bg_batch"] + gp * w["ps_batch"], steps

    # calibrate on a ~1 s sample, then size the timed sample for `seconds_target`; equal packets from both launches
    gp = 4 * cores
    g = max(4 * cores, gp * w["ps_batch"] // w["bg_batch"])
    dt, pk, _ = sample(g, gp)
    scale = max(1.0, seconds_target / max(dt, 1e-3))
    gp = int(min(w["ps_glob"], gp * scale))
    g = int(min(w["bg_glob"], g * scale))
    dt, packets, steps = sample(g, gp)
    sample = "%d of %d background work items x %d packets + %d of %d point-source work items x %d packets (%.1f s)" % (
        g, w["bg_glob"], w["bg_batch"], gp, w["ps_glob"], w["ps_batch"], dt)
    return packets / dt, steps / dt, cores, kind, sample


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
        pps, sps, cores, kind, sample = cpu_leg(w, seconds_target=max(4.0, min(12.0, 120.0 / (args.warmup + args.steps))))
        if i >= args.warmup:
            vals.append(pps)
            steps_rate.append(sps)
        info = (cores, kind, sample)
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "photon_packets_per_s", "value": v, "unit": "packets/s",
        "cell_steps_per_s": float(np.mean(steps_rate)),
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * w["packets"] / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(w),
        "cpu_baseline": {"value": v, "unit": "packets/s", "cores": info[0], "kind": info[1], "sample": info[2]},
        "e2e": {"value": v, "unit": "packets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(w):
    return {"workload": "256^3 regular grid, point source + isotropic background, HG g=0.6 dsc tables, "
                        "TABS+INT absorptions, one frequency per step",
            "grid": "%d^3" % w["cloud"].NX, "packets_per_step": int(w["packets"]),
            "ps_packets": int(w["ps_batch"] * w["ps_glob"]), "bg_packets": int(w["packets"] - w["ps_batch"] * w["ps_glob"]),
            "kabs=ksca": "5/N per cell", "l2": "DENS+TABS+INT = 192 MiB > 126 MB L2 (inputs larger than L2)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from soc_b200 import backend

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")

    w = make_workload()
    cloud = w["cloud"]
    B = backend.Backend(cloud, ordinal=local, rng_mode=backend.RNG_PACKET, **REF_OPTS)
    dev = B.dev
    dev.set_shard(rank, world)
    dev.set_tuning(deposit=args.deposit, refill=args.refill, aggregate_steps=args.agg_steps)
    stream = torch.cuda.ExternalStream(dev.stream(), device=torch.device("cuda", local))
    n = cloud.CELLS

    # device-resident inputs
    dev.upload(backend.BUF_PSPOS, w["pspos"])
    dev.upload(backend.BUF_PS, w["ps"])
    dev.upload(backend.BUF_DSC, w["dsc"])
    dev.upload(backend.BUF_CSC, w["csc"])
    dev.zero_amc(0)
    dev.zero_amc(1)
    dev.sync()

    class _Raw:          # wraps the library's INT buffer as a torch tensor for NCCL (no copy)
        def __init__(self, ptr, count):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    int_ptr, _ = dev.device_ptr(backend.BUF_INT)
    int_t = torch.as_tensor(_Raw(int_ptr, n), device=torch.device("cuda", local)) if world > 1 else None

    # pinned host buffers of the end-to-end leg
    tables = torch.empty(2 * BINS, dtype=torch.float32).pin_memory()
    tables[:BINS] = torch.from_numpy(w["dsc"])
    tables[BINS:] = torch.from_numpy(w["csc"])
    tables_np = tables.numpy()
    int_host = torch.empty(n, dtype=torch.float32).pin_memory()
    int_host_np = int_host.numpy()

    def step(e2e, seed):
        if e2e:
            dev.upload(backend.BUF_DSC, tables_np[:BINS])
            dev.upload(backend.BUF_CSC, tables_np[BINS:])
        dev.zero_amc(1)
        dev.sim_pb(0, w["ps_batch"] * w["ps_glob"], w["ps_batch"], seed, w["kabs"], w["ksca"], 0.0, w["tw"], w["ps_glob"])
        dev.sim_pb(1, w["bg_batch"] * w["bg_glob"], w["bg_batch"], seed, w["kabs"], w["ksca"], w["bg"], w["tw"], w["bg_glob"])
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(int_t)
        if e2e and rank == 0:
            dev.download(backend.BUF_INT, n, out=int_host_np)

    def timed(e2e, nsteps, sample_clocks):
        sampler = None
        dev.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sample_clocks and rank == 0:
            sampler = ClockSampler(local)
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kernel_ms = 0.0
        with torch.cuda.stream(stream):
            e0.record(stream)
        t0 = time.perf_counter()
        for i in range(nsteps):
            step(e2e, SEED + 0.001 * i)
            if args.kernel_times:
                kernel_ms += dev.last_launch_ms()
        with torch.cuda.stream(stream):
            e1.record(stream)
        dev.sync()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if e2e:
            ms = max(ms, 1e3 * wall)          # host copies are part of the end-to-end step
        clocks = sampler.summary() if sampler else None
        t = torch.tensor([ms], dtype=torch.float64, device=torch.device("cuda", local))
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), clocks

    for i in range(args.warmup):
        step(False, 0.9 - 0.001 * i)
    dev.sync()
    dev.reset_counters()
    ms, clocks = timed(False, args.steps, True)
    c = dev.counters()
    counts = torch.tensor([c.packets, c.steps, c.scatterings, c.reserved[0]], dtype=torch.float64,
                          device=torch.device("cuda", local))
    if world > 1:
        dist.all_reduce(counts)
    packets, csteps = counts[0].item(), counts[1].item()
    # dominant kernel alone (the background launch): CUDA events of the library around that launch
    dev.zero_amc(1)
    dev.reset_counters()
    kms = []
    for i in range(max(1, min(3, args.steps))):
        dev.sim_pb(1, w["bg_batch"] * w["bg_glob"], w["bg_batch"], SEED + 0.01 * i, w["kabs"], w["ksca"], w["bg"], w["tw"], w["bg_glob"])
        kms.append(dev.last_launch_ms())
    ck = dev.counters()
    ksteps = ck.steps / len(kms)
    kavg = float(np.mean(kms))
    # end to end
    step(True, 0.77)
    ems, _ = timed(True, args.steps, False)

    if rank == 0:
        peak, peak_src = measured_peak()
        traffic = measured_traffic() if world == 1 else None
        achieved = ksteps * ALG_BYTES_PER_STEP / (kavg * 1e-3) / 1e9
        line = {
            "metric": "photon_packets_per_s", "value": packets / (ms * 1e-3), "unit": "packets/s",
            "cell_steps_per_s": csteps / (ms * 1e-3), "steps_per_packet": csteps / max(packets, 1.0),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(w),
            "clocks": clocks,
            "e2e": {"value": packets / (ems * 1e-3), "unit": "packets/s", "h2d_bytes_per_step": 2 * BINS * 4,
                    "d2h_bytes_per_step": 4 * n, "ms_per_step": ems / args.steps},
            "gpu_launches": int(c.launches) * world,          # counted by the library: 2 packet kernels + 2 folds per step and rank
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_source": (traffic or {}).get("source"),
                         "kernel": "%s<DEP_RED,%s> (background launch)" % ("sim_lean_kernel" if os.environ.get("SOC_AHEAD", "1") == "0" else "sim_ahead_kernel",
                                                                              "brick" if os.environ.get("SOC_LAYOUT", "1") != "0" else "linear"),
                         "kernel_ms": kavg, "cell_steps_per_launch": ksteps, "alg_bytes_per_cell_step": ALG_BYTES_PER_STEP,
                         "peak_source": peak_src},
            "stuck_packets": counts[3].item(),
        }
        if world == 1 and not args.no_cpu:
            pps, sps, cores, kind, sample = cpu_leg(w, seconds_target=12.0)
            line["cpu_baseline"] = {"value": pps, "unit": "packets/s", "cell_steps_per_s": sps, "cores": cores,
                                    "kind": kind, "sample": sample}
        print(json.dumps(line))
    B.close()
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
    ap.add_argument("--kernel-times", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
