"""ASOCS driver: images of scattered light (peel-off method) on the B200 library.  Command line, ini keywords,
inputs and the `outcoming.socs` output follow the reference ASOCS.py; its pyopencl layer (ASOCS.py:129-381,
416-720) is replaced by the C ABI of include/soc_b200.h.

    python -m soc_b200.asocs my.ini            (bin/ASOCS.py is the same entry point)

Sources: point sources (kernel SimRAM_PS) and the isotropic background (SimRAM_PB) for orthographic observers.
The Healpix background, diffuse/cell emission as sources and the Healpix observer (`perspective`) of the
reference's scattered-light kernels are not implemented; an ini file that asks for them is rejected.
Weights follow ASOCS.py:437-457 (WPS, WBG), the final scaling ASOCS.py:874-877.
"""
import sys
import time

import numpy as np

from . import backend as bk
from .asoc import read_dusts, read_scattering_functions, read_background, read_sources, read_abundances, Comm, _opt_array
from .constants import PLANCK, PARSEC, FACTOR, ADHOC, SEED0, SEED1, GLOBAL_0_SCA
from .formats import read_cloud
from .hostmath import fix, observer_directions_rad
from .ini import User


def main(argv=None, device_factory=None):
    argv = sys.argv if argv is None else argv
    t_start = time.time()
    if len(argv) < 2:
        print(" ASOCS.py input_file")
        sys.exit()
    USER = User(argv[1])
    if not USER.Validate():
        print("Check the inifile... exiting!")
        sys.exit()
    bad = USER.unsupported()
    if USER.INTOBS[0] > -10000.0:
        bad.append("perspective: Healpix images of scattered light (NDIR<0) are not implemented")
    if len(USER.file_hpbg) > 2:
        bad.append("hpbg: the Healpix background is not implemented for scattered light")
    if len(USER.file_diffuse) > 0 or USER.CLPAC > 0:
        bad.append("diffuse / cellpackets: emission from the medium is not implemented for scattered light")
    if USER.PS_METHOD not in (0, 1):
        bad.append("psmethod %d: the reference's scattered-light kernels read XPS_* through mistyped pointers" % USER.PS_METHOD)
    if bad:
        for b in bad:
            print("*** soc_b200: " + b)
        sys.exit()
    comm = Comm()
    root = comm.rank == 0
    VERBOSE = USER.VERBOSE if root else 0

    FFREQ, AFG, AFABS, AFSCA = read_dusts(USER)
    NFREQ, NDUST = USER.NFREQ, len(AFABS)
    FDSC, FCSC = read_scattering_functions(USER)
    IBG = read_background(USER)
    LPS = read_sources(USER)
    cloud = read_cloud(USER.file_cloud, USER.KDENSITY)
    NX, NY, NZ, CELLS = cloud.NX, cloud.NY, cloud.NZ, cloud.CELLS
    USER.AREA = cloud.AREA
    ABU = read_abundances(CELLS, NDUST, USER)
    WITH_ABU = ABU.shape[0] > 0
    if len(USER.OBS_THETA) < 1:
        print("*** No observer directions: keyword directions")
        sys.exit()
    NDIR, ODIR, RA, DE = observer_directions_rad(USER.OBS_THETA, USER.OBS_PHI)
    npx, npy = USER.NPIX['x'], USER.NPIX['y']
    LOCAL = 32
    if 'local' in USER.KEYS:
        LOCAL = int(USER.KEYS['local'][0])
    GLOBAL_0 = GLOBAL_0_SCA
    if 'global' in USER.KEYS:
        GLOBAL_0 = int(USER.KEYS['global'][0])
    GLOBAL_0 = fix(GLOBAL_0, 32 * LOCAL)
    PSPAC = fix(USER.PSPAC, LOCAL)
    BGPAC = fix(fix(USER.BGPAC, USER.AREA), LOCAL)
    if root:
        np.asarray([BGPAC, PSPAC, 0, 0], np.int32).tofile('packet.info')
    centre = USER.MAPCENTRE
    if centre[0] < -1e7:
        centre = np.array([0.5 * NX, 0.5 * NY, 0.5 * NZ], np.float32)

    ordinal = comm.local if comm.world > 1 else 0
    dev = (device_factory or bk.Device)(ordinal)
    dev.set_params(bins=USER.DSC_BINS, no_ps=max(1, USER.NO_PS), ps_method=USER.PS_METHOD, with_abu=int(WITH_ABU),
                   ffs=USER.FFS, length=float("%.5e" % (USER.GL * PARSEC)), factor=FACTOR, adhoc=ADHOC)
    dev.set_grid(cloud)
    dev.set_rng_mode(bk.RNG_REFERENCE if 'REFSTREAMS' in USER.KEYS else bk.RNG_PACKET)
    dev.set_shard(comm.rank, comm.world)
    if USER.NO_PS > 0:
        dev.upload(bk.BUF_PSPOS, np.ascontiguousarray(USER.PSPOS[:USER.NO_PS].reshape(-1)))
    for b, v in ((bk.BUF_ODIR, ODIR), (bk.BUF_ORA, RA), (bk.BUF_ODE, DE)):
        dev.upload(b, np.ascontiguousarray(np.asarray(v, np.float32)[:, :3].reshape(-1)))

    OUTCOMING = np.zeros((NFREQ, NDIR, npy, npx), np.float32)
    Tkernel = 0.0
    host_rng = np.random.default_rng()
    for II in range(2):
        if II == 0:
            GLOBAL = GLOBAL_0
            if PSPAC < 1 or USER.NO_PS < 1:
                continue
            BATCH = int(max([1, PSPAC / GLOBAL]))
            PACKETS = GLOBAL * BATCH
            WPS = 1.0 / (PLANCK * PACKETS * ((USER.GL * PARSEC) ** 2.0))
            BATCH *= USER.NO_PS
            PACKETS = GLOBAL * BATCH
        else:
            if BGPAC < 1:
                continue
            BATCH = max([1, int(round(BGPAC / (8 * USER.AREA)))])
            PACKETS = int(8 * USER.AREA * BATCH)
            WBG = np.pi / (PLANCK * 8 * BATCH)
            GLOBAL = fix(int(8 * USER.AREA), 64)
        for IFREQ in range(NFREQ):
            FREQ = FFREQ[IFREQ]
            if FREQ < USER.SIM_F[0] or FREQ > USER.SIM_F[1]:
                continue
            dev.sca_zero_out(NDIR, npx, npy)
            if WITH_ABU:
                dev.upload(bk.BUF_OPT, _opt_array(USER, ABU, AFABS, AFSCA, IFREQ).reshape(-1))
                kabs = ksca = 0.0
            else:
                kabs, ksca = float(sum(a[IFREQ] for a in AFABS)), float(sum(s[IFREQ] for s in AFSCA))
            BG = float(IBG[IFREQ] * WBG / FREQ) if (II == 1 and len(IBG) == NFREQ) else 0.0
            if II == 0:
                dev.upload(bk.BUF_PS, np.asarray(LPS[:, IFREQ] * WPS / FREQ, np.float32))
            dev.upload(bk.BUF_DSC, FDSC[IFREQ])
            dev.upload(bk.BUF_CSC, FCSC[IFREQ])
            seed = float(np.fmod(USER.SEED + SEED0 + IFREQ * SEED1, 1.0)) if USER.SEED > 0 else float(host_rng.random())
            t0 = time.time()
            if II == 0:
                dev.sca_ps(PACKETS, BATCH, seed, kabs, ksca, NDIR, npx, npy, USER.MAP_DX, centre, GLOBAL)
            else:
                dev.sca_pb(1, PACKETS, BATCH, seed, kabs, ksca, BG, NDIR, npx, npy, USER.MAP_DX, centre, GLOBAL)
            dev.sync()
            Tkernel += time.time() - t0
            comm.allreduce(dev, bk.BUF_OUT, NDIR * npx * npy)
            if root:
                OUTCOMING[IFREQ] += dev.download(bk.BUF_OUT, NDIR * npx * npy).reshape(NDIR, npy, npx)
            if VERBOSE:
                print("  %s FREQ %3d/%3d  %10.3e  ABS %.3e  SCA %.3e" % (["PS", "BG"][II], IFREQ + 1, NFREQ, FREQ, kabs, ksca))
    if root:
        for IFREQ in range(NFREQ):
            OUTCOMING[IFREQ] *= np.float32(FFREQ[IFREQ] * 1.0e23 * PLANCK / (USER.MAP_DX * USER.MAP_DX))
        with open('outcoming.socs', 'wb') as fp:
            np.asarray([npy, npx, NFREQ], np.int32).tofile(fp)
            np.asarray(FFREQ, np.float32).tofile(fp)
            OUTCOMING.tofile(fp)
    c = dev.counters()
    if VERBOSE:
        print("        KERNEL   %9.4f seconds   (%d packets, %d cell-steps, %d peel-off rays on this rank)" % (Tkernel, c.packets, c.steps, c.peels))
    dev.close()
    comm.close()
    if root:
        print("@@ ASOCS.py %.2f seconds WC" % (time.time() - t_start))
    return 0


if __name__ == "__main__":
    main()
