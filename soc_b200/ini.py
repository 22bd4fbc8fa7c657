"""The ini-file parser of the SOC drivers: same keywords, same prefix matching and defaults as the reference
`User` class (ASOC_aux.py:79-551), without its pyopencl vector types.

Keywords are matched on the first token of a line with `key.find(prefix) == 0`; `#` starts a comment; every
raw key is also kept in `KEYS` because the drivers probe a few of them ad hoc (CLT, CLE, local, ...).
Options of the reference that this implementation does not provide are parsed (so that ini files written for
the reference load) and rejected later with a clear message by `unsupported()`.
"""
import sys

import numpy as np

from .constants import MAXPS, D2R, um2f


class User:
    def __init__(self, filename):
        # input files
        self.file_cloud = ''
        self.file_diffuse = ''
        self.file_background = ''
        self.file_constant_load = ''
        self.file_optical = []
        self.file_scafunc = []
        self.file_abundance = []
        self.file_hpbg = ''
        self.HPBG_WEIGHTED = False
        # output files
        self.file_absorbed = 'default.absorbed'
        self.file_emitted = 'soc.emitted'
        self.file_temperature = ''
        self.file_savetau = ''
        self.file_pssavetau = ''
        self.file_scattering = 'scattering'
        self.file_constant_save = ''
        self.kernel_defs = ''
        # run parameters
        self.GL = 0.0
        self.MAP_DX = 1.0
        self.KDENSITY = 1.0
        self.DISTANCE = 0.0
        self.ITERATIONS = 1
        self.STEP_WEIGHT = [-1, 0, 0]
        self.DIR_WEIGHT = [-1, 0, 0]
        self.NPIX = {'x': 10, 'y': 10}
        self.FAST_MAP = -1
        self.REMIT_F = [0.0, 1e30]
        self.SIM_F = [1.0e8, 1.0e17]
        self.LEVEL_THRESHOLD = 0
        self.INTOBS = np.array([-1e12, 0.0, 0.0], np.float32)
        self.MAPCENTRE = np.array([-1e12, 0.0, 0.0], np.float32)
        self.DEVICES = 'g'
        self.DSC_BINS = 0
        self.LOCAL = -1
        self.GLOBAL = -1
        self.BATCH = 30
        self.OBS_THETA = []
        self.OBS_PHI = []
        self.PSPAC = 0
        self.PS_METHOD = 0
        self.BGPAC = 0
        self.CLPAC = 0
        self.DFPAC = 0
        self.NO_PS = 0
        self.file_pointsource = []
        self.PS_SCALING = np.ones(MAXPS, np.float32)
        self.PSPOS = np.zeros((MAXPS, 3), np.float32)
        self.PSPOS[:, 0] = -1e10
        self.DO_SPLIT = 0
        self.POLMAP = 0
        self.POLSTAT = 0
        self.NOSOLVE = 0
        self.LOAD_TEMPERATURE = 0
        self.NOMAP = 0
        self.NOABSORBED = 0
        self.SAVE_INTENSITY = 0
        self.SAVE_INTENSITY_FILE = 'ISRF.DAT'
        self.USE_EMWEIGHT = 0
        self.EMWEIGHT_SKIP = 3
        self.EMWEIGHT_LIM = [0.0, 1e10, 0.0]
        self.SEED = np.pi / 4.0
        self.MAP_FREQ = [1.0e6, 1e18]
        self.SINGLE_MAP_FREQ = np.asarray([], np.float32)
        self.FFS = 1
        self.BG_METHOD = 0
        self.WITH_ALI = 0
        self.WITH_REFERENCE = 0
        self.scale_background = 1.0
        self.LEVELS = 999
        self.KEYS = {}
        self.K_DIFFUSE = 1.0
        self.SINGLE_ABU = 0
        self.OPT_IS_HALF = 0
        self.savetau_freq = []
        self.pssavetau_freq = -1.0
        self.ROI = np.zeros(6, np.int32)
        self.ROI_STEP = 0
        self.ROI_NSIDE = 16
        self.ROI_LOAD_SCALE = 1.0
        self.FILE_ROI_SAVE = ''
        self.FILE_ROI_LOAD = ''
        self.WITH_ROI_SAVE = 0
        self.WITH_ROI_LOAD = 0
        self.ROI_MAP = 0
        self.ROIPAC = 0
        self.OUT_NSIDE = 128
        self.FSELECT = []
        self.LIB_ABS = False
        self.LIB_MAPS = False
        self.MAP_INTERPOLATION = 0
        self.FITS = 0
        self.FITS_PREFIX = 'map'
        self.FITS_RA = 0.0
        self.FITS_DE = 0.0
        self.MIRROR = ''
        self.VERBOSE = 1
        self.MMAP_ABSORBED = 0
        self.MMAP_EMITTED = 0
        self.POLSIM = 0
        self.CR_HEATING = 0.0
        self.ABSTHIN = -1
        self.NNNLIMIT = 0.0
        self.AREA = 0

        for line in open(filename).readlines():
            s = line.split('#')[0].split()
            if len(s) < 1:
                continue
            if s[0] == 'DEFS':
                self.kernel_defs = line[4:].split('#')[0]
            if s[0].find('mapum') == 0:
                for ss in s[1:]:
                    self.SINGLE_MAP_FREQ = np.concatenate((self.SINGLE_MAP_FREQ, np.asarray([um2f(float(ss))], np.float32)))
                if len(self.SINGLE_MAP_FREQ) > 1:
                    self.SINGLE_MAP_FREQ = np.sort(self.SINGLE_MAP_FREQ)
            if s[0] == 'singleabu':
                self.SINGLE_ABU = 1
            if s[0] == 'optishalf':
                self.OPT_IS_HALF = 1
            self.KEYS.update({s[0]: s[1:]})
            # keywords without arguments
            key = s[0].lower()
            if key.find('nosolve') == 0:
                self.NOSOLVE = 1
            if key.find('loadtemp') == 0:
                self.LOAD_TEMPERATURE = 1
            if key.find('nomap') == 0:
                self.NOMAP = 1
            if key.find('noabs') == 0:
                self.NOABSORBED = 1
            if key.find('dustem') == 0:
                self.NOABSORBED = 1
                self.SAVE_INTENSITY = 1
            if key.find('roimap') == 0:
                self.ROI_MAP = 1
            if key.find('savetau') == 0 and len(s) > 2:
                self.file_savetau = s[1]
                for x in s[2:]:
                    self.savetau_freq.append(0.0 if float(x) < 0.0 else um2f(float(x)))
            if key.find('pssavetau') == 0:
                self.file_pssavetau = s[1]
                self.pssavetau_freq = um2f(float(s[2]))
            if key.find('fits') == 0:
                self.FITS = 1
                if len(s) >= 3:
                    self.FITS_RA, self.FITS_DE = float(s[1]), float(s[2])
                    if len(s) >= 4:
                        self.FITS_PREFIX = s[3]
            if key.find('mirror') == 0 and len(s) > 1:
                self.MIRROR = s[1]
            if len(s) < 2:
                continue
            # keywords with a single argument
            key, a = s[0], s[1]
            if key.find('device') == 0:
                self.DEVICES = a.lower()
            if key.find('verbose') == 0:
                self.VERBOSE = int(a)
            if key.find('mmapabs') == 0:
                self.MMAP_ABSORBED = int(a)
            if key.find('mmapemit') == 0:
                self.MMAP_EMITTED = int(a)
            if key.find('tempera') == 0:
                self.file_temperature = a
            if key.find('cloud') == 0:
                self.file_cloud = a
            if key.find('absorb') == 0:
                self.file_absorbed = a
            if key.find('scatter') == 0:
                self.file_scattering = a
            if key.find('emit') == 0:
                self.file_emitted = a
            if key.find('split') == 0:
                self.DO_SPLIT = int(a)
            if key.find('mapint') == 0:
                self.MAP_INTERPOLATION = int(a)
            if key.find('polstat') == 0:
                self.POLSTAT = int(a)
            if key.find('absthin') == 0:
                self.ABSTHIN = int(a)
            if key.find('nnnlimit') == 0:
                self.NNNLIMIT = float(a)
            if key.find('libabs') == 0:
                self.FSELECT = np.atleast_1d(np.asarray(np.loadtxt(a), np.float32))
                self.LIB_ABS = True
            if key.find('libmap') == 0:
                self.FSELECT = np.atleast_1d(np.asarray(np.loadtxt(a), np.float32))
                self.LIB_MAPS = True
            if key.find('diffus') == 0:
                self.file_diffuse = a
                if len(s) > 2:
                    self.K_DIFFUSE = float(s[2])
            if key.find('optic') == 0:
                self.file_optical.append(a)
                self.file_abundance.append(s[2] if (len(s) > 2 and s[2][0:1] != '#') else '#')
            if key.find('backg') == 0:
                self.file_background = a
                if len(s) > 2:
                    self.scale_background = float(s[2])
            if key.find('hpbg') == 0:
                self.file_hpbg = a
                if len(s) > 2:
                    self.scale_background = float(s[2])
                if len(s) > 3:
                    self.HPBG_WEIGHTED = int(s[3])
            if key.find('cload') == 0:
                self.file_constant_load = a
            if key.find('csave') == 0:
                self.file_constant_save = a
            if key.find('iterations') == 0:
                self.ITERATIONS = int(a)
            if key.find('threshold') == 0:
                self.LEVEL_THRESHOLD = int(a)
            if key.find('gridlen') == 0:
                self.GL = float(a)
            if key.find('distance') == 0:
                self.DISTANCE = float(a)
            if key.find('bgpac') == 0:
                self.BGPAC = int(float(a))
            if key.find('pspac') == 0:
                self.PSPAC = int(float(a))
            if key.find('psmetho') == 0:
                self.PS_METHOD = int(a)
            if key.find('cellpac') == 0:
                self.CLPAC = int(round(float(a)))
            if key.find('roipac') == 0:
                self.ROIPAC = int(round(float(a)))
            if key.find('roinside') == 0:
                self.ROI_NSIDE = int(round(float(a)))
            if key.find('diffpac') == 0:
                self.DFPAC = int(a)
            if key.find('seed') == 0:
                self.SEED = float(np.clip(float(a), -1.0, 1.0))
            if key.find('dens') == 0:
                self.KDENSITY = float(a)
            if key.find('CR_HEATING') == 0:
                self.CR_HEATING = float(a)
            if key.find('batch') == 0:
                self.BATCH = int(a)
            if key.find('local') == 0:
                self.LOCAL = int(a)
            if key.find('global') == 0:
                self.GLOBAL = int(a)
            if key.find('forcedfirst') == 0:
                self.FFS = int(a)
            if key.find('ffs') == 0:
                self.FFS = int(a)
            if key.find('bgmethod') == 0:
                self.BG_METHOD = int(a)
            if key.find('ali') == 0:
                self.WITH_ALI = int(a)
            if key.find('reference') == 0:
                self.WITH_REFERENCE = int(a)
            if key.find('saveint') == 0:
                self.SAVE_INTENSITY = int(a)
                if len(s) > 2:
                    self.SAVE_INTENSITY_FILE = s[2]
            if key.find('levels') == 0:
                self.LEVELS = int(a)
            if key.find('outnside') == 0:
                self.OUT_NSIDE = int(a)
            if key.find('emwei') == 0:
                self.USE_EMWEIGHT = int(a)
                if len(s) > 3:
                    self.EMWEIGHT_LIM = [float(s[2]), float(s[3]), 0.0]
                    if len(s) > 4:
                        self.EMWEIGHT_LIM[2] = float(s[4])
                        if len(s) > 5:
                            self.EMWEIGHT_SKIP = int(s[5])
            if len(s) < 3:
                continue
            key, a, b = s[0], s[1], s[2]
            # keywords with two arguments
            if key.find('remit') == 0:
                self.REMIT_F = [um2f(float(b)), um2f(float(a))]
            if key.find('simum') == 0:
                self.SIM_F = [um2f(float(b)), um2f(float(a))]
            if key.find('dsc') == 0:
                self.file_scafunc.append(s[1])
                if len(self.file_scafunc) == 1:
                    self.DSC_BINS = int(s[2])
                elif self.DSC_BINS != int(s[2]):
                    print("*** Error in scattering functions: number of bins must be the same for all dusts")
                    sys.exit()
            if key.find('direwei') == 0:
                self.DIR_WEIGHT = [int(a), float(b)]
            if key.find('direct') == 0:
                if len(self.OBS_THETA) >= 10:
                    print("** ERROR - cannot have more than 10 directions -- ABORT !!")
                    sys.exit()
                self.OBS_THETA.append(float(a) * D2R)
                self.OBS_PHI.append(float(b) * D2R)
            if key.find('wavelen') == 0:
                self.MAP_FREQ = [um2f(float(b)), um2f(float(a))]
            if key.find('roisave') == 0:                   # roisave <file> <step>      (ASOC_aux.py:448-451)
                self.WITH_ROI_SAVE = 1
                self.FILE_ROI_SAVE = a
                self.ROI_STEP = int(b)
            if key.find('roiload') == 0:                   # roiload <file> <scale>     (ASOC_aux.py:452-455)
                self.WITH_ROI_LOAD = 1
                self.FILE_ROI_LOAD = a
                self.ROI_LOAD_SCALE = float(b)
            if len(s) < 4:
                continue
            # keywords with three arguments
            key, a, b, c = s[0], s[1], s[2], s[3]
            if key.find('polsim') == 0:
                self.POLSIM = 1
            if key.find('polmap') == 0:
                self.POLMAP = 1
            if key.find('perspec') == 0:
                self.INTOBS = np.array([float(a), float(b), float(c)], np.float32)
            if key.find('stepwei') == 0:
                self.STEP_WEIGHT = [int(a), float(b), float(c)]
            if key.find('mapping') == 0:
                self.NPIX = {'x': int(a), 'y': int(b)}
                self.MAP_DX = float(c)
                if len(s) > 4:
                    try:
                        self.FAST_MAP = int(s[4])
                    except ValueError:
                        pass
            if key == 'roi' and len(s) >= 7:                 # roi x0 x1 y0 y1 z0 z1 (inclusive root cells, ASOC_aux.py:527)
                self.ROI = np.asarray([int(v) for v in s[1:7]], np.int32)
            if key.find('mapcent') == 0:
                self.MAPCENTRE = np.array([float(a), float(b), float(c)], np.float32)
            if key.find('mapview') == 0:
                self.OBS_THETA = [float(s[1]) * np.pi / 180.0]
                self.OBS_PHI = [float(s[2]) * np.pi / 180.0]
                if len(s) >= 5:
                    self.NPIX = {'x': int(s[3]), 'y': int(s[4])}
                    if len(s) >= 6:
                        self.MAP_DX = float(s[5])
                        if len(s) >= 9:
                            self.MAPCENTRE = np.array([float(s[6]), float(s[7]), float(s[8])], np.float32)
            if len(s) < 5:
                continue
            if key.find('pointsou') == 0:
                if self.NO_PS < MAXPS:
                    self.PSPOS[self.NO_PS] = [float(s[1]), float(s[2]), float(s[3])]
                    self.file_pointsource.append(s[4])
                    if len(s) > 5 and s[5] != '#':
                        self.PS_SCALING[self.NO_PS] = float(s[5])
                    self.NO_PS += 1
                else:
                    print("Reached maximum number of point sources = %d" % MAXPS)
                    sys.exit()
        if self.CLPAC > 0:
            self.DFPAC = self.CLPAC

    def Validate(self):
        ok = True
        if len(self.file_cloud) < 1:
            print("*** Cloud model not definied: keyword cloud")
            ok = False
        if self.CLPAC < 1 and self.WITH_ALI > 0:
            print("*** WARNING:  CLPAC=0 and WITH_ALI=%d" % self.WITH_ALI)
            print("***           Cannot use ALI, we set WITH_ALI=0")
            self.WITH_ALI = 0
        if self.PSPAC < 1:
            self.NO_PS = 0
        return ok

    def unsupported(self):
        """Options of the reference that this implementation does not provide (SURVEY.md appendix B /
        DESIGN.md 'out of scope').  Returns a list of messages; empty = fine."""
        bad = []
        if self.DO_SPLIT:
            bad.append("split: packet splitting (SimBgSplit/SimHpSplit) is not implemented")
        if self.POLMAP or self.POLSIM or self.POLSTAT:
            bad.append("polmap/polsim/polstat: polarisation maps are not implemented")
        if self.DIR_WEIGHT[0] > 0:
            bad.append("dirweight: does not compile in the reference either (undeclared pweight)")
        if self.PS_METHOD == 3:
            bad.append("psmethod 3: not implemented in the reference either")
        if 1 < self.FAST_MAP < 999:
            bad.append("mapping ... fast: kernel_ASOC_map_X.c does not exist in the reference either (ASOC.py:3442)")
        if self.FAST_MAP >= 999 and self.NPIX['y'] <= 0:
            bad.append("mapping ... 999 with a Healpix map: per-level Healpix maps are not implemented")
        if self.USE_EMWEIGHT > 1:
            bad.append("emweight 2 is not implemented")
        if self.LIB_ABS or self.LIB_MAPS:
            bad.append("libabs/libmaps: the library method is not implemented")
        if self.ABSTHIN > 1:
            bad.append("absthin is not implemented")
        if self.CR_HEATING > 0.0:
            bad.append("CR_HEATING is not implemented")
        return bad
