"""ASOCS driver: images of scattered light (peel-off method) on the B200 library.  Command line, ini keywords,
inputs and the `outcoming.socs` / `<scattering>.fits` outputs follow the reference ASOCS.py; its pyopencl layer
(ASOCS.py:129-381, 416-720) is replaced by the C ABI of include/soc_b200.h.

    python -m soc_b200.asocs my.ini            (bin/ASOCS.py is the same entry point)

Sources, as in ASOCS.py:416-870: point sources (kernel SimRAM_PS), the isotropic or Healpix background (SimRAM_PB /
SimRAM_HP), a diffuse field from a file and the emission of the dust itself read from the `emitted` file
(SimRAM_CL).  Observers: orthographic maps (`directions`) or one Healpix image seen from inside/outside the model
(`perspective x y z` + `outnside`, ASOCS.py:44-47).  `roiload` + `roipac` add the stored external field (II == 3).
Weights follow ASOCS.py:437-475 (WPS, WBG), the final scaling ASOCS.py:874-884.
"""
import os
import sys
import time

import numpy as np

from . import backend as bk
from .asoc import (read_dusts, read_scattering_functions, read_background, read_sources, read_abundances, Comm,
                   _opt_array, mirror_mask, upload_scattering, read_cloud_cut)
from .constants import PLANCK, PARSEC, FACTOR, ADHOC, SEED0, SEED1, GLOBAL_0_SCA, HPBG_NPIX
from .fits import write_fits
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
    cloud = read_cloud_cut(USER, comm)
    NX, NY, NZ, CELLS = cloud.NX, cloud.NY, cloud.NZ, cloud.CELLS
    USER.AREA = cloud.AREA
    ABU = read_abundances(CELLS, NDUST, USER)
    WITH_ABU = ABU.shape[0] > 0
    WITH_MSF = FDSC.shape[0] > 1
    if WITH_MSF and not WITH_ABU:
        print("Multiple scattering functions given but abundances do not vary =>")
        print("Calculate a single scattering function outside SOC!")
        sys.exit()
    LEVELS, LCELLS, OFF, DENS = cloud.LEVELS, cloud.LCELLS, cloud.OFF, cloud.DENS
    HEALPIX_OUT = USER.INTOBS[0] > -10000.0                    # ASOCS.py:44-50
    if HEALPIX_OUT:
        NDIR = -int(USER.OUT_NSIDE)
        ODIR = np.asarray(USER.INTOBS, np.float32).reshape(1, 3)
        RA = DE = np.zeros((1, 3), np.float32)
        NOUT = 12 * NDIR * NDIR
    else:
        if len(USER.OBS_THETA) < 1:
            print("*** No observer directions: keyword directions")
            sys.exit()
        NDIR, ODIR, RA, DE = observer_directions_rad(USER.OBS_THETA, USER.OBS_PHI)
    npx, npy = USER.NPIX['x'], USER.NPIX['y']
    if not HEALPIX_OUT:
        NOUT = NDIR * npx * npy
    LOCAL = 32
    if 'local' in USER.KEYS:
        LOCAL = int(USER.KEYS['local'][0])
    GLOBAL_0 = GLOBAL_0_SCA
    if 'global' in USER.KEYS:
        GLOBAL_0 = int(USER.KEYS['global'][0])
    GLOBAL_0 = fix(GLOBAL_0, 32 * LOCAL)
    PSPAC = fix(USER.PSPAC, LOCAL)
    BGPAC = fix(fix(USER.BGPAC, USER.AREA), LOCAL)
    DFPAC = CLPAC = 0
    if USER.USE_EMWEIGHT > 0:                                    # ASOCS.py:82-95
        CLPAC = fix(USER.CLPAC, LOCAL)
        if USER.DFPAC > 0:
            DFPAC = fix(USER.DFPAC, LOCAL)
    else:
        CLPAC = fix(fix(USER.CLPAC, CELLS), LOCAL)
        if USER.DFPAC > 0:
            DFPAC = fix(fix(USER.DFPAC, CELLS), LOCAL)
    if root:
        np.asarray([BGPAC, PSPAC, DFPAC, CLPAC], np.int32).tofile('packet.info')
    centre = USER.MAPCENTRE
    if centre[0] < -1e7:
        centre = np.array([0.5 * NX, 0.5 * NY, 0.5 * NZ], np.float32)
    HPBG = []
    if len(USER.file_hpbg) > 2:
        HPBG = np.fromfile(USER.file_hpbg, np.float32).reshape(NFREQ, HPBG_NPIX) * np.float32(USER.scale_background)
    DIFFUSERAD = []
    if len(USER.file_diffuse) > 0:
        dims = np.fromfile(USER.file_diffuse, np.int32, 2)
        if dims[0] != CELLS:
            print("DIFFUSERAD has %d cells but the cloud has %d cells ??" % (dims[0], CELLS))
            sys.exit()
        DIFFUSERAD = np.memmap(USER.file_diffuse, dtype='float32', mode='r', shape=(CELLS, int(dims[1])), offset=8)
    m = np.nonzero((FFREQ >= USER.REMIT_F[0]) & (FFREQ <= USER.REMIT_F[1]))
    REMIT_I1, REMIT_I2 = int(m[0][0]), int(m[0][-1])
    EMITTED = None
    if CLPAC > 0:                                                # emission of the dust, solved earlier by ASOC.py
        hdr = np.fromfile(USER.file_emitted, np.int32, 2)
        if hdr[0] != CELLS or hdr[1] != REMIT_I2 - REMIT_I1 + 1:
            print("*** emitted file %s has %d cells x %d frequencies ??" % (USER.file_emitted, hdr[0], hdr[1]))
            sys.exit()
        EMITTED = np.memmap(USER.file_emitted, dtype='float32', mode='r', shape=(CELLS, int(hdr[1])), offset=8)

    ordinal = comm.local if comm.world > 1 else int(os.environ.get("SOC_DEVICE", "0"))
    dev = (device_factory or bk.Device)(ordinal)
    dev.set_params(bins=USER.DSC_BINS, no_ps=max(1, USER.NO_PS), ps_method=USER.PS_METHOD, with_abu=int(WITH_ABU),
                   ffs=USER.FFS, hpbg_weighted=int(USER.HPBG_WEIGHTED), use_emweight=USER.USE_EMWEIGHT,
                   with_msf=int(WITH_MSF), ndust=NDUST, mirror=mirror_mask(USER), opt_is_half=int(bool(USER.OPT_IS_HALF)),
                   with_roi_load=int(USER.WITH_ROI_LOAD), ref_quirks=3 if 'REFQUIRKS' in USER.KEYS else 0,
                   length=float("%.5e" % (USER.GL * PARSEC)), factor=FACTOR, adhoc=ADHOC)
    dev.set_grid(cloud)
    ROI_LOAD, ROI_LOAD_NELEM = None, 0
    if USER.WITH_ROI_LOAD:                                       # ASOCS.py:176-197
        hdr = np.fromfile(USER.FILE_ROI_LOAD, np.int32, 5)       # (nx, ny, nz, nside, nfreq)
        if hdr[3] != USER.ROI_NSIDE or hdr[4] != NFREQ:
            print("ROI file %s: nside %d, %d frequencies; the run has nside %d, %d frequencies" %
                  (USER.FILE_ROI_LOAD, hdr[3], hdr[4], USER.ROI_NSIDE, NFREQ))
            sys.exit()
        roi_dim = tuple(int(v) for v in hdr[:3])
        ROI_LOAD_NELEM = roi_dim[0] * roi_dim[1] + roi_dim[1] * roi_dim[2] + roi_dim[2] * roi_dim[0]
        ROI_LOAD = np.memmap(USER.FILE_ROI_LOAD, dtype='float32', mode='r', offset=20,
                             shape=(NFREQ, ROI_LOAD_NELEM * 12 * USER.ROI_NSIDE * USER.ROI_NSIDE))
        dev.set_roi(USER.ROI, USER.ROI_STEP, USER.ROI_NSIDE, roi_dim)
    dev.set_rng_mode(bk.RNG_REFERENCE if 'REFSTREAMS' in USER.KEYS else bk.RNG_PACKET)
    dev.set_shard(comm.rank, comm.world)
    if USER.NO_PS > 0:
        dev.upload(bk.BUF_PSPOS, np.ascontiguousarray(USER.PSPOS[:USER.NO_PS].reshape(-1)))
    if WITH_ABU:          # once: OPT of every frequency is built from it on the device (soc_build_opt)
        dev.upload(bk.BUF_ABU, np.ascontiguousarray(ABU, np.float32).reshape(-1))
    for b, v in ((bk.BUF_ODIR, ODIR), (bk.BUF_ORA, RA), (bk.BUF_ODE, DE)):
        dev.upload(b, np.ascontiguousarray(np.asarray(v, np.float32)[:, :3].reshape(-1)))

    OUTCOMING = np.zeros((NFREQ, NOUT), np.float32)
    EMIT = np.zeros(CELLS, np.float32)
    EMWEI = np.zeros(CELLS, np.float32) if USER.USE_EMWEIGHT > 0 else None
    Tkernel = 0.0
    entropy = np.array([np.random.default_rng().random() if USER.SEED <= 0 else USER.SEED], np.float32)
    comm.broadcast_host(entropy)
    host_rng = np.random.default_rng(int(float(entropy[0]) * 2 ** 31))

    def opacities(IFREQ):
        if WITH_ABU:
            if 'HOSTOPT' in USER.KEYS:
                o = _opt_array(USER, ABU, AFABS, AFSCA, IFREQ).reshape(-1)
                dev.upload(bk.BUF_OPT, o, np.float16 if USER.OPT_IS_HALF else np.float32)
            else:
                dev.build_opt([a[IFREQ] for a in AFABS], [s_[IFREQ] for s_ in AFSCA], 0, bool(USER.SINGLE_ABU))
            return 0.0, 0.0
        return float(sum(a[IFREQ] for a in AFABS)), float(sum(s_[IFREQ] for s_ in AFSCA))

    def emission_weights(IFREQ, npac):
        """Packets per cell from the emission of the current frequency, Russian roulette below one (ASOCS.py:554-575)."""
        tmp = np.asarray(EMITTED[:, IFREQ - REMIT_I1], np.float64).copy() if EMITTED is not None else np.asarray(EMIT, np.float64)
        tmp[~np.isfinite(tmp)] = 0.0
        tmp[:] = npac * tmp / (np.sum(tmp) + 1.0e-65)
        EMWEI[:] = np.clip(tmp, USER.EMWEIGHT_LIM[0], USER.EMWEIGHT_LIM[1])
        EMWEI[np.nonzero(host_rng.random(CELLS) > EMWEI)] = 0.0
        if USER.EMWEIGHT_LIM[2] > 0.0:
            EMWEI[np.nonzero(EMWEI < USER.EMWEIGHT_LIM[2])] = 0.0
        comm.broadcast_host(EMWEI)
        dev.upload(bk.BUF_EMWEI, EMWEI)

    def harvest(IFREQ):
        comm.allreduce(dev, bk.BUF_OUT, NOUT)
        if root:
            OUTCOMING[IFREQ] += dev.download(bk.BUF_OUT, NOUT)

    # ---- constant sources: point sources, background, diffuse field (ASOCS.py:416-720) ----------------------
    for II in range(4):
        WPS = WBG = 0.0
        if II == 0:
            GLOBAL = GLOBAL_0
            if PSPAC < 1 or USER.NO_PS < 1:
                continue
            BATCH = int(max([1, PSPAC / GLOBAL]))
            PACKETS = GLOBAL * BATCH
            WPS = 1.0 / (PLANCK * PACKETS * ((USER.GL * PARSEC) ** 2.0))
            BATCH *= USER.NO_PS
            PACKETS = GLOBAL * BATCH
        elif II == 1:
            if BGPAC < 1:
                continue
            if len(HPBG) < 1:
                BATCH = max([1, int(round(BGPAC / (8 * USER.AREA)))])
                PACKETS = int(8 * USER.AREA * BATCH)
                WBG = np.pi / (PLANCK * 8 * BATCH)
                GLOBAL = fix(int(8 * USER.AREA), 64)
            else:                                                # Healpix sky: packets aimed at a sphere of radius Rout
                BATCH = 1
                GLOBAL = fix(int(BGPAC / BATCH), 64)
                PACKETS = GLOBAL * BATCH
                Rout = 0.5 * np.sqrt(NX * NX + NY * NY + NZ * NZ)
                WBG = np.pi * 4.0 * np.pi * Rout ** 2.0 / (PLANCK * PACKETS)
        elif II == 2:
            GLOBAL = GLOBAL_0
            if len(DIFFUSERAD) < 1 or DFPAC < 1:
                continue
            BATCH = int(DFPAC / CELLS)
            PACKETS = DFPAC
        else:                                                    # the stored external field (ASOCS.py:483-499)
            if USER.ROIPAC < 1 or not USER.WITH_ROI_LOAD:
                continue
            npix_roi = 12 * USER.ROI_NSIDE * USER.ROI_NSIDE
            GLOBAL = fix(100 * ROI_LOAD_NELEM, LOCAL)
            BATCH = max([1, int(USER.ROIPAC / (100.0 * npix_roi * ROI_LOAD_NELEM))]) * npix_roi
            PACKETS = ROI_LOAD_NELEM
        skip = 2
        for IFREQ in range(NFREQ):
            FREQ = FFREQ[IFREQ]
            if FREQ < USER.SIM_F[0] or FREQ > USER.SIM_F[1]:
                continue
            dev.sca_zero_out(NDIR, npx, npy)
            kabs, ksca = opacities(IFREQ)
            BG = float(IBG[IFREQ] * WBG / FREQ) if (II == 1 and len(IBG) == NFREQ) else 0.0
            if II == 0:
                dev.upload(bk.BUF_PS, np.asarray(LPS[:, IFREQ] * WPS / FREQ, np.float32))
            if II == 1 and len(HPBG) > 0:
                if USER.HPBG_WEIGHTED:
                    tmp = np.asarray(HPBG[IFREQ, :], np.float64)
                    tmp /= np.mean(tmp)
                    tmp = np.clip(tmp, 1.0e-2, 1.0e4)
                    tmp /= np.sum(tmp)
                    HPBGW = (1.0 / 49152.0) / tmp
                    HPBGP = np.cumsum(tmp)
                    HPBGP[-1] = 1.00001
                    dev.upload(bk.BUF_HPBG, np.asarray((WBG / FREQ) * HPBG[IFREQ, :] * HPBGW, np.float32))
                    dev.upload(bk.BUF_HPBGP, np.asarray(HPBGP, np.float32))
                else:
                    dev.upload(bk.BUF_HPBG, np.asarray((WBG / FREQ) * HPBG[IFREQ, :], np.float32))
            upload_scattering(dev, FDSC, FCSC, IFREQ, AFABS, AFSCA)
            seed = float(np.fmod(USER.SEED + SEED0 + IFREQ * SEED1, 1.0)) if USER.SEED > 0 else float(host_rng.random())
            if II == 2:
                if IFREQ >= DIFFUSERAD.shape[1]:
                    continue
                for level in range(LEVELS):
                    coeff = USER.GL * PARSEC / (8.0 ** level) * USER.K_DIFFUSE
                    a_, b_ = OFF[level], OFF[level] + LCELLS[level]
                    EMIT[a_:b_] = DIFFUSERAD[a_:b_, IFREQ] * coeff
                EMIT[np.nonzero(DENS < 1.0e-10)] = 0.0
                dev.upload(bk.BUF_EMIT, EMIT)
                if USER.USE_EMWEIGHT > 0:
                    skip += 1
                    if skip == 3:
                        skip = 0
                        emission_weights(IFREQ, CLPAC)         # sic: the reference scales with CLPAC here too (ASOCS.py:566)
            if II == 3:
                dev.upload(bk.BUF_ROI_LOAD, np.asarray(ROI_LOAD[IFREQ, :] * USER.ROI_LOAD_SCALE / (USER.GL * USER.GL), np.float32))
            t0 = time.time()
            if II == 3:
                dev.sca_pb(3, PACKETS, BATCH, seed, kabs, ksca, 0.0, NDIR, npx, npy, USER.MAP_DX, centre, GLOBAL)
            elif II == 0:
                dev.sca_ps(PACKETS, BATCH, seed, kabs, ksca, NDIR, npx, npy, USER.MAP_DX, centre, GLOBAL)
            elif II == 1 and len(HPBG) > 0:
                dev.sca_hp(PACKETS, BATCH, seed, kabs, ksca, NDIR, npx, npy, USER.MAP_DX, centre, GLOBAL)
            elif II == 1:
                dev.sca_pb(1, PACKETS, BATCH, seed, kabs, ksca, BG, NDIR, npx, npy, USER.MAP_DX, centre, GLOBAL)
            else:
                dev.sca_cl(II, PACKETS, BATCH, seed, kabs, ksca, NDIR, npx, npy, USER.MAP_DX, centre, GLOBAL)
            dev.sync()
            Tkernel += time.time() - t0
            harvest(IFREQ)
            if VERBOSE:
                print("  %s FREQ %3d/%3d  %10.3e  ABS %.3e  SCA %.3e" % (["PS", "BG", "DF", "ROI"][II], IFREQ + 1, NFREQ, FREQ, kabs, ksca))

    # ---- emission of the dust itself, read from the emitted file (ASOCS.py:723-868) -------------------------
    if CLPAC > 0:
        GLOBAL, BATCH = GLOBAL_0, max([1, int(CLPAC / CELLS)])
        skip = 2
        for IFREQ in range(NFREQ):
            FREQ = FFREQ[IFREQ]
            if FREQ < USER.SIM_F[0] or FREQ > USER.SIM_F[1]:
                continue
            if IFREQ < REMIT_I1 or IFREQ > REMIT_I2:
                continue
            dev.sca_zero_out(NDIR, npx, npy)
            kabs, ksca = opacities(IFREQ)
            upload_scattering(dev, FDSC, FCSC, IFREQ, AFABS, AFSCA)
            EMIT[:] = EMITTED[:, IFREQ - REMIT_I1]
            for level in range(LEVELS):
                coeff = 1.0e-20 * USER.GL * PARSEC / (8.0 ** level)
                a_, b_ = OFF[level], OFF[level] + LCELLS[level]
                EMIT[a_:b_] *= coeff * DENS[a_:b_]
            EMIT[np.nonzero(DENS < 1.0e-10)] = 0.0
            dev.upload(bk.BUF_EMIT, EMIT)
            if USER.USE_EMWEIGHT > 0:
                skip += 1
                if skip == 3:
                    skip = 0
                    emission_weights(IFREQ, CLPAC)
            seed = float(np.fmod(USER.SEED + IFREQ * SEED1, 1.0)) if USER.SEED > 0 else float(host_rng.random())
            t0 = time.time()
            dev.sca_cl(2, CLPAC, BATCH, seed, kabs, ksca, NDIR, npx, npy, USER.MAP_DX, centre, GLOBAL)
            dev.sync()
            Tkernel += time.time() - t0
            harvest(IFREQ)
            if VERBOSE:
                print("  CL FREQ %3d/%3d  %10.3e  ABS %.3e  SCA %.3e" % (IFREQ + 1, NFREQ, FREQ, kabs, ksca))

    # ---- surface brightness and files (ASOCS.py:874-898) -------------------------------------------------------
    if root:
        for IFREQ in range(NFREQ):
            if HEALPIX_OUT:
                OUTCOMING[IFREQ] *= np.float32(FFREQ[IFREQ] * 1.0e23 * PLANCK / (4.0 * np.pi / (12.0 * NDIR * NDIR)))
            else:
                OUTCOMING[IFREQ] *= np.float32(FFREQ[IFREQ] * 1.0e23 * PLANCK / (USER.MAP_DX * USER.MAP_DX))
        if HEALPIX_OUT:
            with open('outcoming.socs', 'wb') as fp:
                np.asarray([-NDIR, NFREQ], np.int32).tofile(fp)
                np.asarray(FFREQ, np.float32).tofile(fp)
                OUTCOMING.tofile(fp)
        elif NDIR == 1 and USER.FITS > 0:
            pix = USER.GL * USER.MAP_DX / (USER.DISTANCE if USER.DISTANCE > 0.0 else 1000.0)
            write_fits('%s.fits' % USER.file_scattering, OUTCOMING.reshape(NFREQ, npy, npx) if NFREQ > 1 else
                       OUTCOMING.reshape(npy, npx), USER.FITS_RA, USER.FITS_DE, pix, freq=FFREQ if NFREQ > 1 else ())
        else:
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
