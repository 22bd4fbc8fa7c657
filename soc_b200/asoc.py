"""ASOC driver: dust-continuum radiative transfer (absorptions, equilibrium temperatures, emission, maps) on
the B200 library.  Command line, ini keywords, input files and output files are those of the reference
ASOC.py; the pyopencl layer (context, buffers, kernel launches -- ASOC.py:336-586, 1028-1560, 1594-2250,
2898-3180) is replaced by soc_b200.backend.Device, a ctypes binding of the C ABI in include/soc_b200.h.

    python -m soc_b200.asoc my.ini [gpu_flag]          (bin/ASOC.py is the same entry point)

Host-side arithmetic that must be restated exactly follows the reference: packet-count rounding
(ASOC.py:234-251, 1036-1066), source weights WPS/WBG (:1041, 1057-1063), trapezoid weights (:1218-1223),
seeds (:1247, 1807), the E->T table (:643-689), absorbed-file scaling (:2793-2809), map scaling (:2997).

Multi-GPU: launched under torchrun (WORLD_SIZE > 1) every rank simulates packets q with q % world == rank on
its own replica of the grid; INT is all-reduced per frequency and TABS per source pass (NCCL), rank 0 writes
the files.  With WORLD_SIZE unset nothing but ctypes and numpy is imported.
"""
import os
import sys
import time

import numpy as np

from . import backend as bk
from .constants import FACTOR, PLANCK, PARSEC, ADHOC, SEED0, SEED1, GLOBAL_0, HPBG_NPIX, f2um
from .fits import write_fits
from .formats import read_cloud, read_otfile, write_cloud, cut_levels
from .hostmath import fix, observer_directions_rad, energy_temperature_table, trapezoid_weights
from .ini import User


# ---------------------------------------------------------------------------------------------------------------
# input readers that need the User object (ASOC_aux.py:557-646, 1092-1123, 853-862)
# ---------------------------------------------------------------------------------------------------------------
def read_dusts(USER):
    afg, afabs, afsca, ffreq = [], [], [], None
    for filename in USER.file_optical:
        lines = open(filename).readlines()
        gd, gs = float(lines[1].split()[0]), float(lines[2].split()[0])
        coeff = gd * np.pi * gs ** 2.0 * USER.GL * PARSEC
        d = np.loadtxt(filename, skiprows=4, ndmin=2)
        f = np.asarray(d[:, 0], np.float32)
        if ffreq is not None and len(ffreq) != len(f):
            print("*** Error in optical parameters: dusts must have same frequency grid")
            sys.exit()
        ffreq = f
        afg.append(np.asarray(d[:, 1], np.float32))
        afabs.append(np.asarray(d[:, 2] * coeff, np.float32))
        afsca.append(np.asarray(d[:, 3] * coeff, np.float32))
    if ffreq is None:
        print("*** No dust defined: keyword optical")
        sys.exit()
    USER.NFREQ = len(ffreq)
    return ffreq, afg, afabs, afsca


def read_scattering_functions(USER):
    """FDSC, FCSC [ndust, NFREQ, BINS]; ndust is 1 or the number of dusts (=> WITH_MSF) -- ASOC_aux.py:619-650."""
    ndust = len(USER.file_scafunc)
    if ndust < 1:
        print("*** No scattering function defined: keyword dsc")
        sys.exit()
    if ndust != len(USER.file_optical) and ndust != 1:
        print("Must have either a single scattering function (DSC file) or one for each dust!")
        sys.exit()
    dsc = np.zeros((ndust, USER.NFREQ, USER.DSC_BINS), np.float32)
    csc = np.zeros((ndust, USER.NFREQ, USER.DSC_BINS), np.float32)
    for i in range(ndust):
        with open(USER.file_scafunc[i], 'rb') as fp:
            dsc[i] = np.fromfile(fp, np.float32, USER.NFREQ * USER.DSC_BINS).reshape(USER.NFREQ, USER.DSC_BINS)
            csc[i] = np.fromfile(fp, np.float32, USER.NFREQ * USER.DSC_BINS).reshape(USER.NFREQ, USER.DSC_BINS)
    return dsc, csc


def mirror_mask(USER):
    """MIRROR bit mask of ASOC.py:319-321: lower / upper border of x, y, z = x X y Y z Z."""
    m = USER.MIRROR
    return 1 * ('x' in m) + 2 * ('X' in m) + 4 * ('y' in m) + 8 * ('Y' in m) + 16 * ('z' in m) + 32 * ('Z' in m)


def upload_scattering(dev, FDSC, FCSC, ifreq, AFABS=None, AFSCA=None):
    """DSC / CSC of one frequency: [BINS], or [NDUST*BINS] plus the ABS / SCA vectors with several scattering
    functions (ASOC.py:1162-1166, 1234-1243)."""
    if FDSC.shape[0] == 1:
        dev.upload(bk.BUF_DSC, FDSC[0, ifreq])
        dev.upload(bk.BUF_CSC, FCSC[0, ifreq])
    else:
        dev.upload(bk.BUF_DSC, np.ascontiguousarray(FDSC[:, ifreq, :]).reshape(-1))
        dev.upload(bk.BUF_CSC, np.ascontiguousarray(FCSC[:, ifreq, :]).reshape(-1))
        dev.upload(bk.BUF_ABSV, np.asarray([a[ifreq] for a in AFABS], np.float32))
        dev.upload(bk.BUF_SCAV, np.asarray([a[ifreq] for a in AFSCA], np.float32))


def read_background(USER):
    ibg = np.zeros(0, np.float32)
    if USER.BGPAC > 0:
        try:
            ibg = np.fromfile(USER.file_background, np.float32, USER.NFREQ)
        except Exception:
            ibg = np.zeros(USER.NFREQ, np.float32)
            if USER.file_hpbg == '':
                print("Error: BGPAC>0 =>  must have either isotropic or healpix background defined")
                sys.exit()
        if len(ibg) != USER.NFREQ:
            print('Optical data for %d, background intensity for %d frequencies ??' % (USER.NFREQ, len(ibg)))
            sys.exit()
        ibg = ibg * np.float32(USER.scale_background)
    return ibg


def read_sources(USER):
    if USER.NO_PS < 1:
        return np.zeros((0, USER.NFREQ), np.float32)
    lps = np.zeros((USER.NO_PS, USER.NFREQ), np.float32)
    for i in range(USER.NO_PS):
        tmp = np.fromfile(USER.file_pointsource[i], np.float32, USER.NFREQ)
        if len(tmp) != USER.NFREQ:
            print('Source %d: optical data for %d, intensity for %d frequencies ??' % (i, USER.NFREQ, len(tmp)))
            sys.exit()
        lps[i] = tmp * USER.PS_SCALING[i]
    return lps


def read_cloud_cut(USER, comm=None):
    """Cloud file, cut to the `levels` of the ini file if it has more (ASOC_aux.py:749-762: a new file
    <cloud>.MAX<levels> is written and used)."""
    with open(USER.file_cloud, "rb") as fp:
        levels = int(np.fromfile(fp, np.int32, 5)[3])
    name = USER.file_cloud
    if levels > USER.LEVELS:
        name = '%s.MAX%d' % (USER.file_cloud, USER.LEVELS)
        if comm is None or comm.rank == 0:
            print("CUT LEVELS: %s -> %s" % (USER.file_cloud, name))
            cut_levels(USER.file_cloud, name, USER.LEVELS - 1)
        if comm is not None:
            comm.barrier()
    return read_cloud(name, USER.KDENSITY)


def read_abundances(cells, ndust, USER):
    if not any(a[0] != '#' for a in USER.file_abundance):
        return np.zeros((0, 0), np.float32)
    abu = np.ones((cells, ndust), np.float32)
    for i, a in enumerate(USER.file_abundance):
        if a[0] != '#':
            abu[:, i] = np.fromfile(a, np.float32)
    return abu


def external_point_sources(nx, ny, nz, pspos, no_ps, ps_method):
    """Visible cloud sides of sources outside the model (ASOC_aux.py:1538-1632)."""
    nside = np.zeros(max(1, no_ps), np.int32)
    side = np.zeros(3 * max(1, no_ps), np.int32)
    area = np.zeros(3 * max(1, no_ps), np.float32)
    axis = np.zeros(3, np.float32)
    for i in range(no_ps):
        p = pspos[i]
        if 0.0 <= p[0] <= nx and 0.0 <= p[1] <= ny and 0.0 <= p[2] <= nz:
            continue
        no = 0
        for test, sd, ax in ((p[0] > nx, 0, (-1, 0, 0)), (p[0] < 0.0, 1, (1, 0, 0)), (p[1] > ny, 2, (0, -1, 0)),
                             (p[1] < 0.0, 3, (0, 1, 0)), (p[2] > nz, 4, (0, 0, -1)), (p[2] < 0.0, 5, (0, 0, 1))):
            if test:
                side[3 * i + no], area[3 * i + no] = sd, 1.0
                axis = np.asarray(ax, np.float32)
                no += 1
        nside[i] = no
        area[3 * i:3 * i + 3] /= no
    if ps_method == 5:
        for i in range(no_ps):
            cos_theta = 0.5 * np.pi
            for ii in range(8):
                vec = np.array([nx * (ii % 2 == 0) - pspos[i][0], ny * ((ii / 2) % 2 == 0) - pspos[i][1],
                                nz * ((ii / 4) % 2 == 0) - pspos[i][2]], np.float32)
                cos_theta = min(cos_theta, abs(float(np.dot(axis, vec))) / float(np.linalg.norm(vec)))
            area[3 * i] = cos_theta
    return nside, side, area


def open_emitted(USER, cells, nfreq):
    """EMITTED[cells, nfreq]: re-used when a file of the right shape exists, else created (ASOC_aux.py:884-940)."""
    ok = True
    try:
        oc, of = np.fromfile(USER.file_emitted, np.int32, 2)
        ok = (oc == cells) and (of == nfreq)
    except Exception:
        ok = False
    if not ok:
        print("Emitted file TRUNCATED:  %s !!" % USER.file_emitted)
        with open(USER.file_emitted, "wb") as fp:
            np.asarray([cells, nfreq], np.int32).tofile(fp)
            np.zeros(cells * nfreq, np.float32).tofile(fp)
    if USER.MMAP_EMITTED:
        return np.memmap(USER.file_emitted, dtype='float32', mode='r+', shape=(cells, nfreq), offset=8)
    return np.fromfile(USER.file_emitted, np.float32, offset=8).reshape(cells, nfreq)


# ---------------------------------------------------------------------------------------------------------------
class Comm:
    """Rank bookkeeping + all-reduce of device buffers.  world == 1: no torch import, everything is a no-op."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = None
        # gloo is for the CPU tests of the rank logic (host-array devices); the product path is NCCL
        self.backend = os.environ.get("SOC_DIST_BACKEND", "nccl")
        self.own_group = False
        if self.world > 1:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            if not dist.is_initialized():
                if self.backend == "nccl":
                    torch.cuda.set_device(self.local)
                    dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
                else:
                    dist.init_process_group(self.backend)
                self.own_group = True

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def broadcast_host(self, array):
        """In-place broadcast of a host float32 array from rank 0."""
        if self.world == 1:
            return
        t = self.torch.from_numpy(array)
        if self.backend == "nccl":
            g = t.cuda()
            self.dist.broadcast(g, 0)
            array[:] = g.cpu().numpy()
        else:
            self.dist.broadcast(t, 0)

    def allreduce(self, dev, buf, count):
        if self.world == 1:
            return
        torch = self.torch
        if self.backend != "nccl":
            self.dist.all_reduce(torch.from_numpy(dev.host_view(buf, count)))
            return
        ptr, nbytes = dev.device_ptr(buf)

        class _Raw:
            pass
        r = _Raw()
        r.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        t = torch.as_tensor(r, device=torch.device("cuda", self.local))
        stream = torch.cuda.ExternalStream(dev.stream(), device=torch.device("cuda", self.local))
        with torch.cuda.stream(stream):
            self.dist.all_reduce(t)

    def close(self):
        if self.world > 1 and self.own_group:
            self.dist.destroy_process_group()


def _opt_array(USER, ABU, AFABS, AFSCA, ifreq, first=0):
    """OPT[cells,2] = per-cell (KABS, KSCA) for variable abundances (ASOC.py:1146-1161)."""
    ndust = len(AFABS)
    opt = np.zeros((ABU.shape[0], 2), np.float32)
    if USER.SINGLE_ABU:
        a = ABU[:, 0]
        opt[:, 0] = a * AFABS[0][ifreq] + (1.0 - a) * AFABS[1][ifreq]
        opt[:, 1] = a * AFSCA[0][ifreq] + (1.0 - a) * AFSCA[1][ifreq]
    else:
        for idust in range(first, ndust):
            opt[:, 0] += ABU[:, idust] * AFABS[idust][ifreq]
            opt[:, 1] += ABU[:, idust] * AFSCA[idust][ifreq]
    return opt


LAST_TIMINGS = {}          # wall-clock breakdown of the most recent main() in this process (seconds)


def main(argv=None, device_factory=None):
    """`device_factory(ordinal)` returns the object that owns the device; default = the CUDA library."""
    argv = sys.argv if argv is None else argv
    t_start = time.time()
    if len(argv) < 2:
        print("\nUsage:\n    ASOC.py  ini_file   [ gpu_flag ]\n")
        sys.exit()
    USER = User(argv[1])
    if not USER.Validate():
        print("Check the inifile... exiting!")
        sys.exit()
    bad = USER.unsupported()
    if bad:
        for b in bad:
            print("*** soc_b200: " + b)
        sys.exit()
    VERBOSE = USER.VERBOSE
    comm = Comm()
    root = comm.rank == 0
    if not root:
        VERBOSE = 0

    FFREQ, AFG, AFABS, AFSCA = read_dusts(USER)
    NFREQ, NDUST = USER.NFREQ, len(AFABS)
    FDSC, FCSC = read_scattering_functions(USER)
    IBG = read_background(USER)
    LPS = read_sources(USER)
    cloud = read_cloud_cut(USER, comm)
    NX, NY, NZ, LEVELS, CELLS, LCELLS, OFF, DENS = cloud.NX, cloud.NY, cloud.NZ, cloud.LEVELS, cloud.CELLS, \
        cloud.LCELLS, cloud.OFF, cloud.DENS
    USER.AREA = cloud.AREA
    ABU = read_abundances(CELLS, NDUST, USER)
    WITH_ABU = ABU.shape[0] > 0
    WITH_MSF = FDSC.shape[0] > 1
    if WITH_ABU and USER.SINGLE_ABU and NDUST != 2:
        print("Option USER.SINGLE_ABU assumes exactly two dust components !!")
        sys.exit(0)
    if WITH_MSF and (USER.SINGLE_ABU or not WITH_ABU):
        print("Cannot have multiple scattering functions without multiple dusts with variable abundances")
        sys.exit()
    DIFFUSERAD = []
    if len(USER.file_diffuse) > 0:
        dims = np.fromfile(USER.file_diffuse, np.int32, 2)
        if dims[0] != CELLS:
            print("DIFFUSERAD has %d cells but the cloud has %d cells ??" % (dims[0], CELLS))
            sys.exit()
        DIFFUSERAD = np.memmap(USER.file_diffuse, dtype='float32', mode='r', shape=(CELLS, int(dims[1])), offset=8)

    NODIR, ODIR, RA, DE = observer_directions_rad(USER.OBS_THETA, USER.OBS_PHI) if len(USER.OBS_THETA) else \
        (0, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    m = np.nonzero((FFREQ >= USER.REMIT_F[0]) & (FFREQ <= USER.REMIT_F[1]))
    REMIT_I1, REMIT_I2 = int(m[0][0]), int(m[0][-1])
    REMIT_NFREQ = len(m[0])

    LOCAL = 32
    if 'local' in USER.KEYS:
        LOCAL = int(USER.KEYS['local'][0])
    PSPAC = fix(USER.PSPAC, LOCAL)
    BGPAC = fix(fix(USER.BGPAC, USER.AREA), LOCAL)
    DFPAC = 0
    if USER.CLPAC < 1:
        USER.USE_EMWEIGHT = 0
    if USER.USE_EMWEIGHT > 0:
        CLPAC = fix(USER.CLPAC, LOCAL)
        if USER.DFPAC > 0:
            DFPAC = fix(USER.DFPAC, LOCAL)
    else:
        CLPAC = fix(fix(USER.CLPAC, CELLS), LOCAL)
        if USER.DFPAC > 0:
            DFPAC = fix(fix(USER.DFPAC, CELLS), LOCAL)
    if VERBOSE:
        print('PACKETS: PSPAC %d   BGPAC %d  CLPAC %d  DFPAC %d' % (PSPAC, BGPAC, CLPAC, DFPAC))
    if root:
        np.asarray([BGPAC, PSPAC, DFPAC, CLPAC], np.int32).tofile('packet.info')
    if USER.ITERATIONS < 1:
        USER.NOABSORBED = True
    if NDUST > 1 and USER.ITERATIONS > 0 and USER.NOABSORBED and not USER.NOSOLVE:
        print("One cannot have NDUST>1, ITERATIONS>0, NOABSORBED, and NOSOLVE=False!")
        sys.exit()
    if not USER.NOABSORBED and not USER.NOSOLVE:
        print("*** Error: noabsorbed is not set and nosolve is not set ??")
        sys.exit()

    XPS_NSIDE, XPS_SIDE, XPS_AREA = external_point_sources(NX, NY, NZ, USER.PSPOS, USER.NO_PS, USER.PS_METHOD)
    HPBG = []
    if len(USER.file_hpbg) > 2:
        HPBG = np.fromfile(USER.file_hpbg, np.float32).reshape(NFREQ, HPBG_NPIX) * np.float32(USER.scale_background)

    # ---- device: context, parameter block (the former -D macros, ASOC.py:344-362), grid --------------------
    Tread = time.time() - t_start
    t_ctx = time.time()
    ordinal = comm.local if comm.world > 1 else int(os.environ.get("SOC_DEVICE", "0"))
    dev = (device_factory or bk.Device)(ordinal)
    length = float("%.5e" % (USER.GL * PARSEC))
    dev.set_params(bins=USER.DSC_BINS, no_ps=max(1, USER.NO_PS), ps_method=USER.PS_METHOD, with_abu=int(WITH_ABU),
                   with_ali=int(USER.WITH_ALI > 0), noabsorbed=int(bool(USER.NOABSORBED)),
                   save_intensity=USER.SAVE_INTENSITY if USER.SAVE_INTENSITY in (1, 2) else 0,
                   use_emweight=USER.USE_EMWEIGHT, hpbg_weighted=int(USER.HPBG_WEIGHTED), step_weight=USER.STEP_WEIGHT[0],
                   sw_a=float("%.3e" % USER.STEP_WEIGHT[1]), sw_b=float("%.3e" % USER.STEP_WEIGHT[2]),
                   level_threshold=USER.LEVEL_THRESHOLD, length=length, factor=FACTOR, adhoc=ADHOC,
                   with_msf=int(WITH_MSF), ndust=NDUST, mirror=mirror_mask(USER),
                   map_interpolation=USER.MAP_INTERPOLATION, opt_is_half=int(bool(USER.OPT_IS_HALF)),
                   with_roi_load=int(USER.WITH_ROI_LOAD), with_roi_save=int(USER.WITH_ROI_SAVE), roi_map=int(USER.ROI_MAP),
                   ref_quirks=3 if 'REFQUIRKS' in USER.KEYS else 0)
    dev.set_grid(cloud)
    dev.sync()
    Tctx = time.time() - t_ctx          # CUDA context, library load, grid upload, parent / neighbour / brick tables
    # region of interest (ASOC.py:906-945): the external field to load, the file of photons entering ROI
    ROI_LOAD = ROI_SAVE = None
    ROI_LOAD_NELEM = ROI_SAVE_NPIX = 0
    roi_dim = (1, 1, 1)
    if USER.WITH_ROI_LOAD:
        hdr = np.fromfile(USER.FILE_ROI_LOAD, np.int32, 5)                   # (nx, ny, nz, nside, nfreq)
        if hdr[3] != USER.ROI_NSIDE:
            print("ROI file %s has nside %d, ini-file has %d" % (USER.FILE_ROI_LOAD, hdr[3], USER.ROI_NSIDE))
            sys.exit()
        if hdr[4] != NFREQ:
            print("ROI file %s has %d, current run %d frequencies" % (USER.FILE_ROI_LOAD, hdr[4], NFREQ))
            sys.exit()
        roi_dim = tuple(int(v) for v in hdr[:3])
        ROI_LOAD_NELEM = roi_dim[0] * roi_dim[1] + roi_dim[1] * roi_dim[2] + roi_dim[2] * roi_dim[0]
        ROI_LOAD = np.memmap(USER.FILE_ROI_LOAD, dtype='float32', mode='r', offset=20,
                             shape=(NFREQ, ROI_LOAD_NELEM * 12 * USER.ROI_NSIDE * USER.ROI_NSIDE))
    if USER.WITH_ROI_SAVE:
        rn = [(int(USER.ROI[2 * k + 1]) - int(USER.ROI[2 * k]) + 1) * USER.ROI_STEP for k in range(3)]
        ROI_SAVE_NPIX = (rn[0] * rn[1] + rn[1] * rn[2] + rn[2] * rn[0]) * 12 * USER.ROI_NSIDE * USER.ROI_NSIDE
        if root:
            np.asarray(rn + [USER.ROI_NSIDE, NFREQ], np.int32).tofile(USER.FILE_ROI_SAVE)
            ROI_SAVE = np.memmap(USER.FILE_ROI_SAVE, dtype='float32', mode='r+', offset=20, shape=(NFREQ, ROI_SAVE_NPIX))
            ROI_SAVE[:, :] = 0.0
    if USER.WITH_ROI_LOAD or USER.WITH_ROI_SAVE or USER.ROI_MAP:
        dev.set_roi(USER.ROI, USER.ROI_STEP, USER.ROI_NSIDE, roi_dim)
    if WITH_ABU:          # once: OPT of every frequency is built from it on the device (soc_build_opt)
        dev.upload(bk.BUF_ABU, np.ascontiguousarray(ABU, np.float32).reshape(-1))
    dev.set_rng_mode(bk.RNG_REFERENCE if 'REFSTREAMS' in USER.KEYS else bk.RNG_PACKET)
    dev.set_geometry(1 if 'REFGEOMETRY' in USER.KEYS else 0)
    dev.set_shard(comm.rank, comm.world)
    if USER.NO_PS > 0:
        dev.upload(bk.BUF_PSPOS, np.ascontiguousarray(USER.PSPOS[:USER.NO_PS].reshape(-1)))
        dev.upload(bk.BUF_XPS_NSIDE, XPS_NSIDE, np.int32)
        dev.upload(bk.BUF_XPS_SIDE, XPS_SIDE, np.int32)
        dev.upload(bk.BUF_XPS_AREA, XPS_AREA)
    Tkernel = Tpush = Tpull = Tsolve = Tmap = 0.0
    use_int = (not USER.NOABSORBED) or USER.SAVE_INTENSITY in (1, 2)

    # the [CELLS, NFREQ] emission array is needed by the cell-emission iterations, the temperature solution and the maps; a
    # pure absorption run (nosolve, nomap, no cell packets) neither reads nor writes it (the reference creates the file anyway)
    need_emitted = not (USER.NOSOLVE and USER.NOMAP and CLPAC < 1 and not USER.LOAD_TEMPERATURE)
    EMITTED = open_emitted(USER, CELLS, REMIT_NFREQ) if (root and need_emitted) else None
    comm.barrier()
    if not root and need_emitted:
        EMITTED = np.fromfile(USER.file_emitted, np.float32, offset=8).reshape(CELLS, REMIT_NFREQ)
    emission_on_device = False          # BUF_TNEW holds the temperatures EMITTED was computed from (single dust)
    FABSORBED = None
    # [CELLS, NFREQ] absorptions: kept on the device and scaled / transposed there when it fits
    # (soc_absorbed_*), else accumulated on the host per frequency like the reference (ASOC.py:1482-1497)
    fabs_on_device = False
    if not USER.NOABSORBED and USER.MMAP_ABSORBED == 0 and 'HOSTABSORBED' not in USER.KEYS:
        try:
            dev.absorbed_begin(NFREQ)
            fabs_on_device = True
        except bk.SocError as e:
            if VERBOSE:
                print("absorptions stay on the host: %s" % e)
    if not USER.NOABSORBED and root and not fabs_on_device:
        if USER.MMAP_ABSORBED > 0:
            with open(USER.file_absorbed, "wb") as fp:
                np.asarray([CELLS, NFREQ], np.int32).tofile(fp)
            FABSORBED = np.memmap(USER.file_absorbed, dtype='float32', mode='r+', offset=8, shape=(CELLS, NFREQ))
            FABSORBED[:, :] = 0.0
        else:
            FABSORBED = np.zeros((CELLS, NFREQ), np.float32)

    Emin = kE = TTT = None
    NE = 30000
    if not USER.NOSOLVE:
        if NDUST > 1:
            print("*** Error: If emission is solved inside SOC, there must be only a single dust population!")
            sys.exit()
        Emin, kE, TTT = energy_temperature_table(FFREQ, AFABS[0], USER.GL, NE)

    TNEW = None
    if USER.LOAD_TEMPERATURE:
        try:
            TNEW = read_otfile(USER.file_temperature)
            if len(TNEW) != CELLS:
                raise ValueError
        except Exception:
            print("*** Failed to read old temperatures !!")
            TNEW = 15.0 * np.ones(CELLS, np.float32)
            TNEW[np.nonzero(DENS < 1e-7)] = 0.0
    EMIT = np.zeros(CELLS, np.float32)
    TMP = np.zeros(CELLS, np.float32)
    FF_ALL = trapezoid_weights(FFREQ)

    def emission_from_temperature(T):
        """EMITTED[cells, freq] from temperatures.  Default: kernel Emission2 in batches of cells, all frequencies
        per launch and one contiguous copy per batch (ASOC.py:2157-2180, the reference's EBATCH path); with the
        key EMISSION1 one launch of kernel Emission and one strided host copy per frequency (ASOC.py:2185-2197)."""
        nonlocal emission_on_device
        dev.upload(bk.BUF_TNEW, np.ascontiguousarray(T, np.float32))
        emission_on_device = True
        if 'EMISSION1' in USER.KEYS:
            for ifreq in range(REMIT_I1, REMIT_I2 + 1):
                dev.emission(float(FFREQ[ifreq]), float(AFABS[0][ifreq]))
                EMITTED[:, ifreq - REMIT_I1] = dev.download(bk.BUF_EMIT, CELLS, out=TMP)
            return
        batch = max(1, min(CELLS, (1 << 28) // max(1, REMIT_NFREQ)))        # <= 1 GiB of floats per launch
        for a in range(0, CELLS, batch):
            b = min(a + batch, CELLS)
            EMITTED[a:b, :] = dev.emission2(a, b, FFREQ[REMIT_I1:REMIT_I2 + 1], AFABS[0][REMIT_I1:REMIT_I2 + 1])

    if USER.LOAD_TEMPERATURE and USER.ITERATIONS < 1:
        emission_from_temperature(TNEW)

    OEMITTED = OTABS = XEM = OXEM = OXAB = None
    if USER.WITH_REFERENCE > 0:
        OEMITTED = np.zeros((CELLS, REMIT_NFREQ), np.float32)
        OTABS = np.zeros(CELLS, np.float32)
    if USER.WITH_ALI:
        XEM = np.zeros(CELLS, np.float32)
        if USER.WITH_REFERENCE:
            OXEM, OXAB = np.zeros(CELLS, np.float32), np.zeros(CELLS, np.float32)
    EMWEI = np.zeros(CELLS, np.float32) if USER.USE_EMWEIGHT > 0 else None
    INTENSITY = None
    if USER.SAVE_INTENSITY == 1 and root:
        INTENSITY = np.memmap(USER.SAVE_INTENSITY_FILE, dtype='float32', mode="w+", shape=(CELLS, NFREQ), offset=8)
        INTENSITY[:, :] = 0.0
    if USER.SAVE_INTENSITY == 2 and root:
        INTENSITY = np.memmap(USER.SAVE_INTENSITY_FILE, dtype='float32', mode="w+", shape=(CELLS, NFREQ, 4), offset=12)
        INTENSITY[:, :, :] = 0.0
    # host random numbers (seeds when `seed` <= 0, emission-weight roulette): the same stream on every rank
    entropy = np.array([np.random.default_rng().random() if USER.SEED <= 0 else USER.SEED], np.float32)
    comm.broadcast_host(entropy)
    host_rng = np.random.default_rng(int(float(entropy[0]) * 2 ** 31))

    def set_opacity(ifreq, first=0):
        """Scalar ABS/SCA, or OPT upload for variable abundances.  Returns (abs, sca)."""
        if WITH_ABU:
            if 'HOSTOPT' in USER.KEYS:          # the reference's way: numpy loop over the species + 8*CELLS bytes over PCIe
                o = _opt_array(USER, ABU, AFABS, AFSCA, ifreq, first).reshape(-1)
                dev.upload(bk.BUF_OPT, o, np.float16 if USER.OPT_IS_HALF else np.float32)      # ASOC.py:1155-1158
            else:
                dev.build_opt([a[ifreq] for a in AFABS], [s[ifreq] for s in AFSCA], first, bool(USER.SINGLE_ABU))
            return 0.0, 0.0
        return float(sum(a[ifreq] for a in AFABS)), float(sum(s[ifreq] for s in AFSCA))

    def harvest(ifreq, kabs):
        """Per-frequency absorptions / intensities after a launch (ASOC.py:1476-1519, 1879-1905)."""
        nonlocal Tpull
        t0 = time.time()
        if fabs_on_device:
            dev.absorbed_add(ifreq)                 # this rank's share; ranks are combined once at the end
        host_int = USER.SAVE_INTENSITY in (1, 2) or (not USER.NOABSORBED and not fabs_on_device)
        if use_int and host_int:
            comm.allreduce(dev, bk.BUF_INT, CELLS)
        if (USER.SAVE_INTENSITY == 1 or (not USER.NOABSORBED and not fabs_on_device)) and root:
            dev.download(bk.BUF_INT, CELLS, out=TMP)
            if not USER.NOABSORBED and not fabs_on_device:
                FABSORBED[:, ifreq] += TMP
            if USER.SAVE_INTENSITY == 1:
                for level in range(LEVELS):
                    a, b = OFF[level], OFF[level] + LCELLS[level]
                    INTENSITY[a:b, ifreq] += (PLANCK * FFREQ[ifreq] / kabs) * (8.0 ** level) * TMP[a:b] / DENS[a:b]
        if USER.SAVE_INTENSITY == 2:
            for icomp, b_ in enumerate((bk.BUF_INT, bk.BUF_INTX, bk.BUF_INTY, bk.BUF_INTZ)):
                if icomp > 0:
                    comm.allreduce(dev, b_, CELLS)
                if root:
                    dev.download(b_, CELLS, out=TMP)
                    for level in range(LEVELS):
                        a, b = OFF[level], OFF[level] + LCELLS[level]
                        INTENSITY[a:b, ifreq, icomp] += (PLANCK * FFREQ[ifreq] / kabs) * (8.0 ** level) * TMP[a:b] / DENS[a:b]
        Tpull += time.time() - t0

    def harvest_roi(ifreq, scale):
        """Photons that entered ROI during the last launch -> the roisave file (this rank's share reduced to rank 0)."""
        if not USER.WITH_ROI_SAVE:
            return
        comm.allreduce(dev, bk.BUF_ROI_SAVE, ROI_SAVE_NPIX)
        if root:
            ROI_SAVE[ifreq, :] += dev.download(bk.BUF_ROI_SAVE, ROI_SAVE_NPIX) * np.float32(scale)

    # =============================================================================================================
    # constant sources: point sources, background, diffuse emission (ASOC.py:1004-1549)
    # =============================================================================================================
    # Several ranks: the packets of every launch are dealt out over the ranks (packet q on rank q % world), or -- for the
    # constant sources, whose frequencies are independent -- rank r takes frequencies r, r + world, ... whole:
    # no per-frequency collective at all, one all-reduce of TABS per source and of the [CELLS, NFREQ] absorptions at the end.
    # Frequency sharding needs every per-frequency product to stay on the device of the rank that made it.
    n_sim = int(np.sum((FFREQ >= USER.SIM_F[0]) & (FFREQ <= USER.SIM_F[1])))
    freq_shard = (comm.world > 1 and n_sim >= comm.world and USER.SAVE_INTENSITY == 0 and not USER.WITH_ROI_SAVE
                  and (USER.NOABSORBED or fabs_on_device) and USER.USE_EMWEIGHT == 0 and 'PACKETSHARD' not in USER.KEYS)
    if 'FREQSHARD' in USER.KEYS and comm.world > 1 and not freq_shard:
        print("*** FREQSHARD: not possible with these options (saveint, roisave, emweight, HOSTABSORBED or fewer frequencies than ranks)")
    if VERBOSE and comm.world > 1:
        print("%d ranks: %s" % (comm.world, "constant sources sharded by frequency" if freq_shard else "packets of every launch sharded"))
    CTABS = np.zeros(CELLS, np.float32)
    if len(USER.file_constant_load) > 0:
        CTABS = np.fromfile(USER.file_constant_load, np.float32, CELLS)
    else:
        skip = USER.EMWEIGHT_SKIP - 1
        for II in range(4):                 # PSPAC, BGPAC, DFPAC, ROI background (ASOC.py:1028)
            if USER.ITERATIONS < 1:
                continue
            WPS = WBG = 0.0
            if II == 0:
                GLOBAL = GLOBAL_0
                if PSPAC < 1 or USER.NO_PS < 1:
                    continue
                BATCH = int(max([1, PSPAC / GLOBAL]))
                pspac = GLOBAL * BATCH
                WPS = 1.0 / (PLANCK * pspac * ((USER.GL * PARSEC) ** 2.0))
                BATCH *= USER.NO_PS
                PACKETS = pspac * USER.NO_PS
                if VERBOSE:
                    print("=== PS  GLOBAL %d x BATCH %d = %d" % (GLOBAL, BATCH, PACKETS))
            elif II == 1:
                if BGPAC < 1:
                    continue
                if len(HPBG) > 0:
                    BATCH = 100
                    GLOBAL = fix(BGPAC / BATCH, 64)
                    bgpac = GLOBAL * BATCH
                    WBG = np.pi / PLANCK
                    WBG /= (GLOBAL * BATCH) / (2 * (NX * NY + NX * NZ + NY * NZ))
                else:
                    BATCH = max([1, int(round(BGPAC / (8 * USER.AREA)))])
                    bgpac = int(8 * USER.AREA * BATCH)
                    WBG = np.pi / (PLANCK * 8 * BATCH)
                    GLOBAL = fix(int(8 * USER.AREA), 64)
                PACKETS = bgpac
                if VERBOSE:
                    print("=== BG: BGPAC %d, BATCH %d, GLOBAL %d" % (bgpac, BATCH, GLOBAL))
            elif II == 2:
                GLOBAL = GLOBAL_0
                if len(DIFFUSERAD) < 1 or DFPAC < 1:
                    continue
                BATCH = int(DFPAC / CELLS)
                PACKETS = DFPAC
                if VERBOSE:
                    print("=== DFPAC %d, GLOBAL %d, BATCH %d" % (DFPAC, GLOBAL, BATCH))
            else:                               # ROI background: the stored field re-emitted from the surface (ASOC.py:1093-1110)
                if USER.ROIPAC < 1 or not USER.WITH_ROI_LOAD:
                    continue
                npix_roi = 12 * USER.ROI_NSIDE * USER.ROI_NSIDE
                GLOBAL = fix(100 * ROI_LOAD_NELEM, LOCAL)
                BATCH = max([1, int(USER.ROIPAC / (100.0 * npix_roi * ROI_LOAD_NELEM))]) * npix_roi
                PACKETS = ROI_LOAD_NELEM
                if VERBOSE:
                    print("=== ROI: GLOBAL %d, BATCH %d, ROI_LOAD_NELEM %d" % (GLOBAL, BATCH, ROI_LOAD_NELEM))
            dev.zero_amc(0)
            if freq_shard:
                dev.set_shard(0, 1)                 # this rank runs all packets of its frequencies
            i_sim = -1
            for IFREQ in range(NFREQ):
                T000 = time.time()
                FREQ = FFREQ[IFREQ]
                if FREQ < USER.SIM_F[0] or FREQ > USER.SIM_F[1]:
                    continue
                i_sim += 1
                if freq_shard and i_sim % comm.world != comm.rank:
                    if USER.SEED <= 0:
                        host_rng.random()           # keep the seed sequence of the frequencies independent of the number of ranks
                    continue
                t0 = time.time()
                kabs, ksca = set_opacity(IFREQ)
                dev.zero_amc(1)
                BG = 0.0
                if II == 0:
                    dev.upload(bk.BUF_PS, np.asarray(LPS[:, IFREQ] * WPS / FREQ, np.float32))
                if len(IBG) == NFREQ:
                    BG = float(IBG[IFREQ] * WBG / FREQ)
                if II == 1 and len(HPBG) > 0:
                    if USER.HPBG_WEIGHTED:
                        tmp = np.asarray(HPBG[IFREQ, :], np.float64)
                        if max(tmp) < 1.0e-40:
                            continue
                        tmp /= np.mean(tmp)
                        tmp = np.clip(tmp, 1.0e-3, 1.0e4)
                        tmp /= np.sum(tmp)
                        HPBGW = (1.0 / 49152.0) / tmp
                        HPBGP = np.cumsum(tmp)
                        HPBGP[-1] = 1.00001
                        dev.upload(bk.BUF_HPBG, np.asarray((WBG / FREQ) * HPBG[IFREQ, :] * HPBGW, np.float32))
                        dev.upload(bk.BUF_HPBGP, np.asarray(HPBGP, np.float32))
                    else:
                        dev.upload(bk.BUF_HPBG, np.asarray((WBG / FREQ) * HPBG[IFREQ, :], np.float32))
                FF = float(FF_ALL[IFREQ])
                upload_scattering(dev, FDSC, FCSC, IFREQ, AFABS, AFSCA)
                if USER.SEED > 0:
                    seed = float(np.fmod(USER.SEED + SEED0 + IFREQ * SEED1, 1.0))
                else:
                    seed = float(host_rng.random())
                if II == 2:
                    dr_ind = IFREQ + (DIFFUSERAD.shape[1] - NFREQ)
                    if dr_ind >= DIFFUSERAD.shape[1] or dr_ind < 0:
                        continue
                    for level in range(LEVELS):
                        coeff = USER.GL * PARSEC / (8.0 ** level) * USER.K_DIFFUSE
                        a, b = OFF[level], OFF[level] + LCELLS[level]
                        EMIT[a:b] = DIFFUSERAD[a:b, dr_ind] * coeff
                    dev.upload(bk.BUF_EMIT, EMIT)
                    if USER.USE_EMWEIGHT > 0:
                        skip += 1
                        if skip % USER.EMWEIGHT_SKIP == 0:
                            tmp = np.asarray(EMIT, np.float64)
                            tmp[~np.isfinite(tmp)] = 0.0
                            tmp[:] = DFPAC * tmp / (np.sum(tmp) + 1.0e-32)
                            EMWEI[:] = np.clip(tmp, USER.EMWEIGHT_LIM[0], USER.EMWEIGHT_LIM[1])
                            EMWEI[np.nonzero(host_rng.random(CELLS) > EMWEI)] = 0.0
                            comm.broadcast_host(EMWEI)
                            dev.upload(bk.BUF_EMWEI, EMWEI)
                if USER.WITH_ROI_SAVE:
                    dev.clear(bk.BUF_ROI_SAVE, 4 * ROI_SAVE_NPIX)               # per frequency (ASOC.py:1301-1302)
                if II == 3:
                    dev.upload(bk.BUF_ROI_LOAD, np.asarray(ROI_LOAD[IFREQ, :] * USER.ROI_LOAD_SCALE / (USER.GL * USER.GL), np.float32))
                Tpush += time.time() - t0
                t0 = time.time()
                if II == 3:
                    dev.sim_pb(3, PACKETS, BATCH, seed, kabs, ksca, BG, FF, GLOBAL)
                elif II == 2:
                    dev.sim_cl(II, PACKETS, BATCH, seed, kabs, ksca, FF, GLOBAL)
                elif II == 1 and len(HPBG) > 0:
                    dev.sim_hp(PACKETS, BATCH, seed, kabs, ksca, FF, GLOBAL)
                else:
                    dev.sim_pb(II, PACKETS, BATCH, seed, kabs, ksca, BG, FF, GLOBAL)
                dev.sync()
                Tkernel += time.time() - t0
                harvest(IFREQ, kabs)
                harvest_roi(IFREQ, USER.GL * USER.GL)                            # ASOC.py:1468-1476
                if VERBOSE:
                    sys.stdout.write("  FREQ %3d/%3d  %10.3e   BG %12.4e   TW %10.3e   %7.2f\n" % (IFREQ + 1, NFREQ, FREQ, BG, FF, time.time() - T000))
            if freq_shard:
                dev.set_shard(comm.rank, comm.world)
            comm.allreduce(dev, bk.BUF_TABS, CELLS)
            CTABS += dev.download(bk.BUF_TABS, CELLS, out=TMP)
            if VERBOSE and CELLS < 1e8:
                print("******  CONSTANT   %10s   CTABS -> %12.4e" % (['PS', 'BG', 'DE'][II], np.mean(CTABS)))
        if len(USER.file_constant_save) > 0 and root:
            CTABS.tofile(USER.file_constant_save)

    # =============================================================================================================
    # emission from the dust itself, iterated with the temperature solution (ASOC.py:1594-2250)
    # =============================================================================================================
    scale = (6.62607e-27 * FACTOR) / (USER.GL * PARSEC)
    for iteration in range(USER.ITERATIONS):
        if VERBOSE:
            print("ITERATION %d/%d" % (iteration + 1, USER.ITERATIONS))
        dev.zero_amc(0)
        beta = None
        if USER.WITH_ALI:
            XEM[:] = 1.0e-32
        if USER.WITH_REFERENCE:
            if USER.WITH_REFERENCE == 1:
                k = iteration / float(USER.ITERATIONS)
            else:
                k = (iteration + int(USER.WITH_REFERENCE % 100)) / float(int(np.floor(0.01 * USER.WITH_REFERENCE)))
            OEMITTED[:, :] *= k
            OTABS[:] *= k
        GLOBAL, BATCH = GLOBAL_0, max([1, int(CLPAC / CELLS)])
        skip = USER.EMWEIGHT_SKIP - 1
        if CLPAC > 0:
            for IFREQ in range(NFREQ):
                FREQ = FFREQ[IFREQ]
                dev.zero_amc(1)
                if FREQ < USER.SIM_F[0] or FREQ > USER.SIM_F[1]:
                    continue
                t0 = time.time()
                # sic: with variable abundances the reference sums the opacities from dust 1 on here (ASOC.py:1673)
                kabs, ksca = set_opacity(IFREQ, first=0 if USER.SINGLE_ABU else (1 if WITH_ABU else 0))
                FF = float(FF_ALL[IFREQ])
                upload_scattering(dev, FDSC, FCSC, IFREQ, AFABS, AFSCA)
                if IFREQ < REMIT_I1 or IFREQ > REMIT_I2:
                    continue
                if USER.WITH_REFERENCE:
                    EMIT[:] = EMITTED[:, IFREQ - REMIT_I1] - OEMITTED[:, IFREQ - REMIT_I1]
                    OEMITTED[:, IFREQ - REMIT_I1] = 1.0 * EMITTED[:, IFREQ - REMIT_I1]
                else:
                    EMIT[:] = EMITTED[:, IFREQ - REMIT_I1]
                for level in range(LEVELS):
                    coeff = USER.GL * PARSEC / (8.0 ** level) / FACTOR
                    a, b = OFF[level], OFF[level] + LCELLS[level]
                    EMIT[a:b] *= coeff * DENS[a:b]
                EMIT[np.nonzero(DENS < 1.0e-10)] = 0.0
                if USER.WITH_ALI:
                    XEM += EMIT * FF
                if USER.USE_EMWEIGHT > 0:
                    skip += 1
                    if skip % USER.EMWEIGHT_SKIP == 0:
                        tmp = np.asarray(EMITTED[:, IFREQ - REMIT_I1].copy(), np.float64)
                        tmp[~np.isfinite(tmp)] = 0.0
                        tmp[:] = CLPAC * tmp / (np.sum(tmp) + 1.0e-65)
                        EMWEI[:] = np.clip(tmp, USER.EMWEIGHT_LIM[0], USER.EMWEIGHT_LIM[1])
                        EMWEI[np.nonzero(host_rng.random(CELLS) > EMWEI)] = 0.0
                        if USER.EMWEIGHT_LIM[2] > 0.0:
                            EMWEI[np.nonzero(EMWEI < USER.EMWEIGHT_LIM[2])] = 0.0
                        comm.broadcast_host(EMWEI)       # every rank must use the same roulette outcome
                        dev.upload(bk.BUF_EMWEI, EMWEI)
                dev.upload(bk.BUF_EMIT, EMIT)
                if USER.SEED > 0:
                    seed = float(np.fmod(USER.SEED + IFREQ * SEED1, 1.0))
                else:
                    seed = float(host_rng.random())
                if USER.WITH_ROI_SAVE:
                    dev.clear(bk.BUF_ROI_SAVE, 4 * ROI_SAVE_NPIX)
                Tpush += time.time() - t0
                t0 = time.time()
                dev.sim_cl(2, CLPAC, BATCH, seed, kabs, ksca, FF, GLOBAL)
                dev.sync()
                Tkernel += time.time() - t0
                if iteration == USER.ITERATIONS - 1:
                    harvest(IFREQ, kabs)
                    harvest_roi(IFREQ, 1.0)                                      # sic: no GL^2 here (ASOC.py:1908-1914)
                if VERBOSE:
                    print("  FREQ %3d/%3d  %10.3e" % (IFREQ + 1, NFREQ, FREQ))
            if USER.WITH_ALI:
                comm.allreduce(dev, bk.BUF_XAB, CELLS)
                xab = dev.download(bk.BUF_XAB, CELLS)
                if USER.WITH_REFERENCE:
                    OXAB += xab
                    OXEM += XEM
                    beta = (OXEM - OXAB) / OXEM
                else:
                    beta = (XEM - xab) / XEM
            if USER.NOABSORBED:
                comm.allreduce(dev, bk.BUF_TABS, CELLS)
                dev.download(bk.BUF_TABS, CELLS, out=EMIT)
                if USER.WITH_REFERENCE:
                    EMIT[:] += OTABS
                    OTABS[:] = 1.0 * EMIT
                EMIT[:] += CTABS
        elif USER.NOABSORBED:
            EMIT[:] = 1.0 * CTABS

        if not USER.NOSOLVE:
            t0 = time.time()
            if VERBOSE:
                print('Calculate temperatures')
            TNEW = np.zeros(CELLS, np.float32)
            if beta is None:
                # EqTemperature kernel on every level (the reference does this with the raw key CLT, ASOC.py:2027-2040)
                dev.upload(bk.BUF_TTT, TTT)
                dev.upload(bk.BUF_EMIT, EMIT)
                for l in range(LEVELS):
                    dev.eq_temperature(l, ADHOC, float(kE), float(Emin), NE)
                dev.download(bk.BUF_TNEW, CELLS, out=TNEW)
                if 'CLT' not in USER.KEYS:
                    TNEW[np.nonzero(DENS < 1.0e-10)] = 0.0        # host solver of the reference: links hold 0
            else:
                # with ALI the escape probability enters the energy balance: host solve (ASOC.py:2044-2061)
                oplgkE = 1.0 / np.log10(kE)
                for level in range(LEVELS):
                    a, b = OFF[level], OFF[level] + LCELLS[level]
                    ok = DENS[a:b] >= 1.0e-10
                    Ein = (scale / ADHOC) * EMIT[a:b] * (8.0 ** level) / np.where(ok, DENS[a:b], 1.0) / np.where(ok, beta[a:b], 1.0)
                    iE = np.clip(np.floor(oplgkE * np.log10(np.maximum(Ein, 1e-300) / Emin)), 0, NE - 2).astype(np.int64)
                    wi = (Emin * kE ** (iE + 1) - Ein) / (Emin * kE ** (iE + 1) - kE ** iE)
                    TNEW[a:b] = np.where(ok, wi * TTT[iE] + (1.0 - wi) * TTT[iE + 1], 0.0)
            mok = np.nonzero(DENS > 1.0e-8)
            TNEW[~np.isfinite(TNEW)] = 10.0
            TNEW[mok] = np.clip(TNEW[mok], 3.0, 1600.0)
            if len(USER.file_temperature) > 0 and root:
                write_cloud(USER.file_temperature, cloud, TNEW)
            if VERBOSE:
                print("Calculate emission")
            tt = TNEW
            if 'CLE' not in USER.KEYS and 'CLT' not in USER.KEYS:
                tt = np.where(TNEW < 3.0, 10.0, TNEW).astype(np.float32)     # avoid 1/0 for parent cells
            emission_from_temperature(tt)
            Tsolve += time.time() - t0
        if VERBOSE:
            print("--- End of iteration ---")

    # =============================================================================================================
    # absorbed file (ASOC.py:2782-2878), emitted file (:3971-3975)
    # =============================================================================================================
    t_files = time.time()
    if fabs_on_device:
        comm.allreduce(dev, bk.BUF_FABS, CELLS * NFREQ)
        if root:
            out = np.empty((CELLS, NFREQ), np.float32)
            dev.absorbed_finish(float(FACTOR / (USER.GL * PARSEC)), float(USER.NNNLIMIT), True, out)
            with open(USER.file_absorbed, 'wb') as fpa:
                np.asarray([CELLS, NFREQ], np.int32).tofile(fpa)
                out.tofile(fpa)
            del out
    elif not USER.NOABSORBED and root:
        for level in range(LEVELS):
            a, b = OFF[level], OFF[level] + LCELLS[level]
            coeff = (8.0 ** level) * (FACTOR / (USER.GL * PARSEC))
            with np.errstate(all='ignore'):      # links have DENS <= 0; those rows are overwritten below
                FABSORBED[a:b, :] *= (coeff / DENS[a:b].reshape(b - a, 1)).astype(np.float32)
            mm = np.nonzero(DENS[a:b] <= USER.NNNLIMIT)
            FABSORBED[a + mm[0], :] = -1.0e20
        if USER.MMAP_ABSORBED == 0:
            with open(USER.file_absorbed, 'wb') as fpa:
                np.asarray([CELLS, NFREQ], np.int32).tofile(fpa)
                FABSORBED.tofile(fpa)
        del FABSORBED
    Tfiles = time.time() - t_files
    if ROI_SAVE is not None:
        ROI_SAVE.flush()
        del ROI_SAVE
    if INTENSITY is not None:
        hdr = [CELLS, NFREQ] if USER.SAVE_INTENSITY == 1 else [CELLS, NFREQ, 4]
        if USER.SAVE_INTENSITY == 2:
            for icomp in (1, 2, 3):
                INTENSITY[:, :, icomp] /= np.where(INTENSITY[:, :, 0] != 0.0, INTENSITY[:, :, 0], 1.0)
        INTENSITY.flush()
        del INTENSITY
        with open(USER.SAVE_INTENSITY_FILE, "r+b") as fp:
            np.asarray(hdr, np.int32).tofile(fp)

    # =============================================================================================================
    # maps (ASOC.py:2898-3177; Healpix :3185-3318)
    # =============================================================================================================
    t0 = time.time()
    KK = (1.0e23 / FACTOR) * PLANCK / (4.0 * np.pi) * USER.GL * PARSEC
    centre = USER.MAPCENTRE
    if centre[0] < -1e7:
        centre = np.array([0.5 * NX, 0.5 * NY, 0.5 * NZ], np.float32)             # ASOC_aux.py:791-793
    if not USER.NOMAP and root and USER.NPIX['y'] <= 0:
        # Healpix map seen by an internal observer (ASOC.py:3185-3318): map_dir_00_H.bin
        print('Write maps')
        nside = USER.NPIX['x']
        freqs = [i for i in range(REMIT_I1, REMIT_I2 + 1) if USER.MAP_FREQ[0] <= FFREQ[i] <= USER.MAP_FREQ[1]]
        with open("map_dir_00_H.bin", "wb") as fp:
            np.asarray([nside, USER.NPIX['y']], np.int32).tofile(fp)
            np.asarray([len(freqs), LEVELS], np.int32).tofile(fp)
            for n_done, IFREQ in enumerate(freqs):
                FREQ = FFREQ[IFREQ]
                kabs, ksca = set_opacity(IFREQ)
                EMIT[:] = EMITTED[:, IFREQ - REMIT_I1] * KK * FREQ
                dev.upload(bk.BUF_EMIT, EMIT)
                dev.healpix_mapping(nside, kabs, ksca, USER.INTOBS, 1)
                dev.download(bk.BUF_MAP, 12 * nside * nside).tofile(fp)
                if n_done == 0 and len(USER.file_savetau) > 0:
                    with open("%s.%d" % (USER.file_savetau, 0), "wb") as fq:
                        np.asarray([nside, USER.NPIX['y']], np.int32).tofile(fq)
                        dev.download(bk.BUF_SAVETAU, 12 * nside * nside).tofile(fq)
    elif not USER.NOMAP and root and len(USER.OBS_THETA) > 0 and USER.FAST_MAP >= 999:
        # one image per hierarchy level (ASOC.py:3320-3440, kernel_ASOC_map_H.c): map_dir_%02d_H.bin holds
        # [npx, npy], [nfreq, LEVELS], then per frequency LEVELS images of npy x npx pixels
        print('Write maps')
        NDIR = len(USER.OBS_THETA)
        npx, npy = USER.NPIX['x'], USER.NPIX['y']
        freqs = [i for i in range(REMIT_I1, REMIT_I2 + 1) if USER.MAP_FREQ[0] <= FFREQ[i] <= USER.MAP_FREQ[1]]
        fpmap = []
        for idir in range(NDIR):
            fpmap.append(open("map_dir_%02d_H.bin" % idir, "wb"))
            np.asarray([npx, npy], np.int32).tofile(fpmap[idir])
            np.asarray([len(freqs), LEVELS], np.int32).tofile(fpmap[idir])
        for n_done, IFREQ in enumerate(freqs):
            kabs, ksca = set_opacity(IFREQ)
            EMIT[:] = EMITTED[:, IFREQ - REMIT_I1] * KK * FFREQ[IFREQ]
            dev.upload(bk.BUF_EMIT, EMIT)
            want_colden = 1 if (n_done == 0 and len(USER.file_savetau) > 0) else 0
            for idir in range(NDIR):
                dev.mapping_levels(USER.MAP_DX, npx, npy, ODIR[idir], RA[idir], DE[idir], kabs, ksca, centre, USER.INTOBS, want_colden)
                dev.download(bk.BUF_MAP, LEVELS * npx * npy).tofile(fpmap[idir])
                if want_colden:                                             # column density of the first frequency (ASOC.py:3419-3425)
                    with open("%s.%d" % (USER.file_savetau, idir), "wb") as fq:
                        np.asarray([npx, npy], np.int32).tofile(fq)
                        dev.download(bk.BUF_SAVETAU, npx * npy).tofile(fq)
        for fp in fpmap:
            fp.close()
    elif not USER.NOMAP and root and len(USER.OBS_THETA) > 0:
        print('Write maps')
        NDIR = len(USER.OBS_THETA)
        npx, npy = USER.NPIX['x'], USER.NPIX['y']
        npix = npx * npy
        fpmap = []
        # FITS files per direction and wavelength when `fits` and `singlemap`-style frequencies are given
        # (ASOC.py:2984-2996), else one binary file per direction
        using_fits = USER.FITS > 0 and len(USER.SINGLE_MAP_FREQ) > 0
        fits_pix = USER.GL * USER.MAP_DX / (USER.DISTANCE if USER.DISTANCE > 0.0 else 1000.0)
        if not using_fits:
            for idir in range(NDIR):
                fpmap.append(open("map_dir_%02d.bin" % idir, "wb"))
                np.asarray([npx, npy], np.int32).tofile(fpmap[idir])
        first_freq = True
        savetau_freq = np.asarray(USER.savetau_freq, np.float64)
        for IFREQ in range(NFREQ):
            save_spe = REMIT_I1 <= IFREQ <= REMIT_I2
            FREQ = FFREQ[IFREQ]
            um = f2um(FREQ)
            ums = ('%.0f' % um) if um > 20.0 else (('%.1f' % um) if um > 2.0 else ('%.2f' % um))
            if FREQ < USER.MAP_FREQ[0] or FREQ > USER.MAP_FREQ[1]:
                continue
            save_tau, save_colden = 0, 0
            if len(savetau_freq) > 0:
                if np.min(np.abs((savetau_freq - FREQ) / FREQ)) < 0.001:
                    save_tau = 1
                if np.min(np.abs((-savetau_freq - FREQ) / FREQ)) < 0.001:
                    save_tau, save_colden = 0, 1
                if save_tau == 0 and first_freq and np.min(savetau_freq) <= 0.0:
                    save_colden = 1
            first_freq = False
            if len(USER.SINGLE_MAP_FREQ) > 0:
                ii = np.argmin(np.abs(FREQ - USER.SINGLE_MAP_FREQ))
                if abs(FREQ - USER.SINGLE_MAP_FREQ[ii]) / FREQ > 0.005:
                    save_spe = False
            if not save_spe and save_tau == 0 and save_colden == 0:
                continue
            kabs, ksca = set_opacity(IFREQ)
            if save_spe and emission_on_device and 'HOSTEMIT' not in USER.KEYS:
                # the emission of this frequency straight from the temperatures on the device (kernel Emission is linear in
                # its FABS argument) instead of a strided host gather of EMITTED[:, f] and a 4*CELLS-byte upload per frequency
                dev.emission(float(FREQ), float(np.float32(AFABS[0][IFREQ]) * np.float32(KK * FREQ)))
            elif save_spe:
                EMIT[:] = KK * FREQ * EMITTED[:, IFREQ - REMIT_I1]
                dev.upload(bk.BUF_EMIT, EMIT)
            elif dev.device_ptr(bk.BUF_EMIT)[1] != 4 * CELLS:
                dev.upload(bk.BUF_EMIT, EMIT)
            for idir in range(NDIR):
                suffix = '_dir%d' % idir if NDIR > 1 else ''
                dev.mapping(USER.MAP_DX, npx, npy, ODIR[idir], RA[idir], DE[idir], kabs, ksca, centre, USER.INTOBS, save_colden)
                if save_spe:
                    if using_fits:
                        name = "%s_%s.fits" % (USER.FITS_PREFIX, ums) if NDIR == 1 else "%s_%s_%03d.fits" % (USER.FITS_PREFIX, ums, idir)
                        write_fits(name, dev.download(bk.BUF_MAP, npix).reshape(npy, npx), USER.FITS_RA, USER.FITS_DE, fits_pix)
                    else:
                        dev.download(bk.BUF_MAP, npix).tofile(fpmap[idir])
                if save_colden > 0:                                         # always FITS (ASOC.py:3152-3159)
                    name = '%s_colden%s.fits' % (USER.file_savetau, suffix) if NDIR == 1 else \
                        '%s_colden%s_%03d.fits' % (USER.file_savetau, suffix, idir)
                    write_fits(name, dev.download(bk.BUF_SAVETAU, npix).reshape(npy, npx), USER.FITS_RA, USER.FITS_DE, fits_pix)
                if save_tau > 0:
                    ext = 'fits' if using_fits else 'bin'
                    name = '%s_tau_%s%s.%s' % (USER.file_savetau, ums, suffix, ext) if NDIR == 1 else \
                        '%s_tau_%s%s_%03d.%s' % (USER.file_savetau, ums, suffix, idir, ext)
                    if using_fits:
                        write_fits(name, dev.download(bk.BUF_SAVETAU, npix).reshape(npy, npx), USER.FITS_RA, USER.FITS_DE, fits_pix)
                    else:
                        dev.download(bk.BUF_SAVETAU, npix).tofile(name)
            if VERBOSE:
                print("IFREQ=%3d/%3d  %9.2f um -- save_spe %d, save_tau %d, save_colden %d" % (IFREQ, NFREQ, um, save_spe, save_tau, save_colden))
        for fp in fpmap:
            fp.close()
    if USER.NO_PS > 0 and USER.pssavetau_freq > 0.0 and USER.NPIX['y'] > 0 and root and len(USER.OBS_THETA) > 0:
        # column density and optical depth from each point source towards the observers (ASOC.py:3576-3644)
        IFREQ = int(np.argmin(np.abs(FFREQ - USER.pssavetau_freq)))
        if abs(FFREQ[IFREQ] - USER.pssavetau_freq) > 0.001 * FFREQ[IFREQ]:
            print("*** Requested frequency for PSSAVETAU is not in the frequency grid")
        kabs, ksca = set_opacity(IFREQ)
        dev.upload(bk.BUF_PSPOS, np.ascontiguousarray(USER.PSPOS[:USER.NO_PS].reshape(-1)))
        for idir in range(len(USER.OBS_THETA)):
            pscolden, pstau = dev.ps_tau(USER.NO_PS, ODIR[idir], kabs, ksca)
            with open("%s_%d.dat" % (USER.file_pssavetau, idir), "w") as fp:
                for i in range(USER.NO_PS):
                    fp.write('%6d  %12.4e  %12.4e\n' % (i, pscolden[i], pstau[i]))
    Tmap = time.time() - t0

    t_files = time.time()
    if root and EMITTED is not None and not USER.MMAP_EMITTED and (not USER.NOSOLVE or USER.LOAD_TEMPERATURE):
        with open(USER.file_emitted, "wb") as fp:
            np.asarray([CELLS, REMIT_NFREQ], np.int32).tofile(fp)
            np.asarray(EMITTED, np.float32).tofile(fp)
    Tfiles += time.time() - t_files
    c = dev.counters()
    if VERBOSE:
        print("        PUSH     %9.4f seconds" % Tpush)
        print("        KERNEL   %9.4f seconds   (%d packets, %d cell-steps on this rank)" % (Tkernel, c.packets, c.steps))
        print("        PULL     %9.4f seconds" % Tpull)
        print("        SOLVE    %9.4f seconds" % Tsolve)
        print("        MAPS     %9.4f seconds" % Tmap)
    dev.close()
    comm.close()
    LAST_TIMINGS.clear()
    LAST_TIMINGS.update(total=time.time() - t_start, read_inputs=Tread, context_and_grid=Tctx, push=Tpush, kernel=Tkernel, pull=Tpull,
                        solve=Tsolve, maps=Tmap, write_files=Tfiles, packets=int(c.packets), cell_steps=int(c.steps))
    if root:
        print("@@ ASOC.py %.2f seconds WC" % (time.time() - t_start))
    return 0


if __name__ == "__main__":
    main()
