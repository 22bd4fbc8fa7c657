#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/libsocref_<tag>.so.

Compiles the reference's own OpenCL kernel sources (kernel_ASOC.c,
kernel_ASOC_map.c, kernel_ASOC_sca.c and the files they include) *from where
they lie* under /root/reference into a host shared object, through the C++
compatibility header oracle/ref_shim/cl_shim.h.  The only textual change is the
OpenCL vector literal ``(uint2)(a,b)`` -> ``uint2(a,b)``, applied with sed into
a temporary directory that is deleted after the build; no reference source is
ever written into this repository.

Like the reference's run-time JIT (ASOC.py:344-396), the grid shape and all
options are compile-time macros, hence one shared object per configuration
("tag").  Outputs go to oracle/_ref/ only (git-ignored, but shipped to the GPU
box with the working tree).

Used by: tests/ (parity oracle), tests/golden/make_golden.py (fixtures) and
bench.py --impl reference / cpu_baseline (CPU baseline).  Never by soc_b200/.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.environ.get("SOC_REFERENCE_DIR", "/root/reference")
OUTDIR = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "ref_shim")

KERNEL_FILES = ["kernel_ASOC.c", "kernel_ASOC_aux.c", "kernel_ASOC_map.c", "kernel_ASOC_sca.c",
                "mwc64x_rng.cl", "skip_mwc.cl"]

PARSEC = 3.08567758e+18
FACTOR = 1.0e20


def reference_available():
    return all(os.path.exists(os.path.join(REFDIR, f)) for f in KERNEL_FILES)


def macro_flags(cfg):
    """The -D list the reference host builds at ASOC.py:344-362 / ASOCS.py:133-147,
    with the same printf formats (so LENGTH etc. carry the same 6-digit rounding)."""
    nx, ny, nz = cfg["NX"], cfg["NY"], cfg["NZ"]
    area = 2.0 * (nx * ny + ny * nz + nz * nx)
    gl = cfg.get("GL", 0.01)
    d = []
    add = d.append
    add("-DNX=%d" % nx); add("-DNY=%d" % ny); add("-DNZ=%d" % nz)
    add("-DBINS=%d" % cfg.get("BINS", 2500))
    add("-DWITH_ALI=%d" % cfg.get("WITH_ALI", 0))
    add("-DPS_METHOD=%d" % cfg.get("PS_METHOD", 0))
    add("-DFACTOR=%.4ef" % FACTOR)
    add("-DCELLS=%d" % cfg["CELLS"])
    add("-DAREA=%.0f" % area)
    add("-DNO_PS=%d" % max(1, cfg.get("NO_PS", 1)))
    add("-DWITH_ABU=%d" % cfg.get("WITH_ABU", 0))
    add("-DROI_MAP=%d" % cfg.get("ROI_MAP", 0)); add("-DMAX_SPLIT=4300"); add("-DSELEM=0")
    add("-DROI_STEP=%d" % cfg.get("ROI_STEP", 0)); add("-DROI_NSIDE=%d" % cfg.get("ROI_NSIDE", 16))
    add("-DWITH_ROI_LOAD=%d" % cfg.get("WITH_ROI_LOAD", 0)); add("-DWITH_ROI_SAVE=%d" % cfg.get("WITH_ROI_SAVE", 0))
    add("-DAXY=%.5ff" % (nx * ny / area)); add("-DAXZ=%.5ff" % (nx * nz / area)); add("-DAYZ=%.5ff" % (ny * nz / area))
    add("-DLEVELS=%d" % cfg["LEVELS"])
    add("-DLENGTH=%.5ef" % (gl * PARSEC))
    add("-DDO_SPLIT=0"); add("-DPOLSTAT=0")
    add("-DSW_A=%.3ef" % cfg.get("SW_A", 0.0)); add("-DSW_B=%.3ef" % cfg.get("SW_B", 0.0))
    add("-DSTEP_WEIGHT=%d" % cfg.get("STEP_WEIGHT", -1))
    add("-DDIR_WEIGHT=-1"); add("-DDW_A=0.000e+00f")
    add("-DLEVEL_THRESHOLD=%d" % cfg.get("LEVEL_THRESHOLD", 0))
    add("-DPOLRED=0"); add("-Dp00=0.2000f"); add("-DMINLOS=-1.000e+00f"); add("-DMAXLOS=1.000e+10f")
    add("-DFFS=%d" % cfg.get("FFS", 1)); add("-DNODIR=%d" % cfg.get("NODIR", 1))
    add("-DUSE_EMWEIGHT=%d" % cfg.get("USE_EMWEIGHT", 0))
    add("-DSAVE_INTENSITY=%d" % cfg.get("SAVE_INTENSITY", 0))
    add("-DNOABSORBED=%d" % cfg.get("NOABSORBED", 1))
    add("-DINTERPOLATE=0")
    add("-DADHOC=%.5ef" % 1.0)
    add("-DHPBG_WEIGHTED=%d" % cfg.get("HPBG_WEIGHTED", 0))
    add("-DWITH_MSF=%d" % cfg.get("WITH_MSF", 0)); add("-DNDUST=%d" % cfg.get("NDUST", 1))
    add("-DOPT_IS_HALF=0"); add("-DPOL_RHO_WEIGHT=0")
    add("-DMAP_INTERPOLATION=%d" % cfg.get("MAP_INTERPOLATION", 0))
    add("-DMIRROR=%d" % cfg.get("MIRROR", 0))
    add("-DCR_HEATING=0"); add("-DCR_HEATING_RATE=0.000e+00f")
    add("-DNVIDIA=0")
    add("-DWITH_COLDEN=0"); add("-DBG_METHOD=0")
    return d


def tag_of(cfg):
    keys = ["NX", "NY", "NZ", "LEVELS", "CELLS"]
    opt = ["NO_PS", "PS_METHOD", "WITH_ABU", "NOABSORBED", "SAVE_INTENSITY", "USE_EMWEIGHT", "WITH_ALI",
           "HPBG_WEIGHTED", "FFS", "BINS", "MAP_NSIDE", "STEP_WEIGHT", "MIRROR", "MAP_INTERPOLATION",
           "LEVEL_THRESHOLD", "GL", "WITH_MSF", "NDUST", "ROI_MAP", "ROI_STEP", "ROI_NSIDE", "WITH_ROI_LOAD", "WITH_ROI_SAVE"]
    s = "_".join(str(cfg[k]) for k in keys)
    for k in opt:
        if k in cfg:
            s += "_%s%s" % (re.sub("[^A-Z]", "", k)[:3] + k[-1], cfg[k])
    if cfg.get("TUNED"):
        s += "_tuned"
    return re.sub(r"[^A-Za-z0-9_.]", "", s)


# Compiler settings.  The default build is conservative (-O2, no contraction: the operation order of the reference
# text, used for the bit-level comparisons of the tests).  TUNED=1 is what a CPU OpenCL driver would make of the
# reference's `-cl-mad-enable` build line (ASOC.py:385): -O3, FMA contraction, AVX2-class vectors (x86-64-v3: the
# library is built in the container and runs on the GPU box's host, so no -march=native).  Never -ffast-math: the
# hierarchy links are denormal floats.
OPT_FLAGS = {False: ["-O2", "-ffp-contract=off"], True: ["-O3", "-march=x86-64-v3", "-ffp-contract=fast"]}


def lib_path(cfg):
    return os.path.join(OUTDIR, "libsocref_%s.so" % tag_of(cfg))


def build(cfg, force=False, verbose=False):
    """Build (or reuse) the reference shared object for configuration `cfg`.
    Returns the path, or None when /root/reference is not present (GPU box) and
    no prebuilt library exists."""
    out = lib_path(cfg)
    if os.path.exists(out) and not force:
        return out
    if not reference_available():
        return None
    os.makedirs(OUTDIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="socref_")
    try:
        for f in KERNEL_FILES:
            txt = open(os.path.join(REFDIR, f)).read()
            txt = re.sub(r"\(uint2\)\(", "uint2(", txt)
            open(os.path.join(tmp, f), "w").write(txt)
        flags = macro_flags(cfg)
        # -ftrivial-auto-var-init=zero: the scattered-light SimRAM_CL indexes DSC with an uninitialised `idust` when
        # WITH_MSF==0 (kernel_ASOC_sca.c:1140 vs :83 where SimRAM_HP initialises it); zero is what the single-dust run means
        common = ["g++", "-std=c++17", "-fpermissive", "-w", "-fopenmp", "-fPIC"] + OPT_FLAGS[bool(cfg.get("TUNED"))] + \
                 ["-ftrivial-auto-var-init=zero", "-I", tmp, "-I", SHIM]
        objs = []
        for tu, extra in (("ref_sim.cpp", ["-DNSIDE=128"]),
                          ("ref_map.cpp", ["-DNSIDE=%d" % cfg.get("MAP_NSIDE", cfg["NX"])]),
                          ("ref_sca.cpp", [])):
            obj = os.path.join(tmp, tu.replace(".cpp", ".o"))
            cmd = common + flags + extra + ["-c", os.path.join(SHIM, tu), "-o", obj]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
            objs.append(obj)
        subprocess.check_call(["g++", "-shared", "-fopenmp", "-o", out] + objs)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def lib_path_levels(cfg):
    return os.path.join(OUTDIR, "libsocrefH_%s_a%d_c%d.so" % (tag_of(cfg), int(cfg.get("USE_ABU", 0)), int(cfg.get("WITH_COLDEN", 0))))


def build_levels(cfg, force=False, verbose=False):
    """Per-level map kernel (kernel_ASOC_map_H.c: Mapping) as a library of its own.  The file is cut after Mapping:
    its HealpixMapping refers to OPT without having the argument under USE_ABU, the polarisation kernels use macros
    the map run of ASOC.py never defines.  USE_ABU and
    WITH_COLDEN are the file's own switches; the reference host defines neither (ASOC.py:3329-3330)."""
    out = lib_path_levels(cfg)
    if os.path.exists(out) and not force:
        return out
    src = os.path.join(REFDIR, "kernel_ASOC_map_H.c")
    if not os.path.exists(src):
        return None
    os.makedirs(OUTDIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="socrefh_")
    try:
        txt = open(src).read()
        cut = txt.index("__kernel void HealpixMapping(")   # keep Mapping only: with USE_ABU the file's HealpixMapping does not compile
        open(os.path.join(tmp, "kernel_ASOC_map_H.c"), "w").write(txt[:cut])
        cmd = ["g++", "-std=c++17", "-fpermissive", "-w", "-O2", "-fopenmp", "-fPIC", "-ffp-contract=off", "-shared",
               "-I", tmp, "-I", SHIM] + macro_flags(cfg) + ["-DNSIDE=%d" % cfg.get("MAP_NSIDE", cfg["NX"])]
        if cfg.get("USE_ABU", 0):
            cmd.append("-DUSE_ABU=1")
        cmd += ["-DWITH_COLDEN=%d" % int(cfg.get("WITH_COLDEN", 0)), os.path.join(SHIM, "ref_maph.cpp"), "-o", out]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def build_a2e(nfreq, ndust, force=False):
    """kernel_A2E_MABU_aux.c (split_absorbed) as a library of its own; NFREQ / NDUST are its compile-time macros
    (A2E_MABU.py:701)."""
    out = os.path.join(OUTDIR, "libsocrefA_%d_%d.so" % (nfreq, ndust))
    if os.path.exists(out) and not force:
        return out
    src = os.path.join(REFDIR, "kernel_A2E_MABU_aux.c")
    if not os.path.exists(src):
        return None
    os.makedirs(OUTDIR, exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-fpermissive", "-w", "-O2", "-fopenmp", "-fPIC", "-ffp-contract=off", "-shared",
                           "-I", REFDIR, "-I", SHIM, "-DNFREQ=%d" % nfreq, "-DNDUST=%d" % ndust,
                           os.path.join(SHIM, "ref_a2e.cpp"), "-o", out])
    return out


if __name__ == "__main__":
    # tiny CLI: build_ref.py NX NY NZ LEVELS CELLS [KEY=VAL ...]
    a = sys.argv[1:]
    cfg = dict(NX=int(a[0]), NY=int(a[1]), NZ=int(a[2]), LEVELS=int(a[3]), CELLS=int(a[4]))
    for kv in a[5:]:
        k, v = kv.split("=")
        cfg[k] = float(v) if "." in v else int(v)
    print(build(cfg, force=True, verbose=True))
