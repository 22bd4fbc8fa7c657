"""The ini-file reader of the SOC drivers.  The keyword set, the prefix matching and the defaults are the drop-in
contract with the reference (`User`, ASOC_aux.py:79-551: an ini file written for ASOC.py / ASOCS.py must mean the same
here), so they are kept; the implementation is a keyword table, and the pyopencl vector types are gone.

A line is `keyword arguments... # comment`.  A keyword matches a table entry when it *starts with* the entry's prefix
(`mapping`, `mappin`... all select `mapping`); entries marked `fold` compare in lower case, `exact` entries need the
whole word.  Every entry whose prefix matches and whose minimum number of arguments is present is applied, in table
order.  All raw keywords are also kept in `KEYS` because the drivers probe a few of them ad hoc (CLT, CLE, REFSTREAMS).
Options of the reference that this implementation does not provide are read (so that such ini files load) and
rejected later with a message by `unsupported()`.
"""
import sys

import numpy as np

from .constants import MAXPS, D2R, um2f


def _defaults():
    """Attribute -> default value (fresh objects on every call)."""
    pspos = np.zeros((MAXPS, 3), np.float32)
    pspos[:, 0] = -1e10
    return dict(
        # input files
        file_cloud='', file_diffuse='', file_background='', file_constant_load='', file_optical=[], file_scafunc=[],
        file_abundance=[], file_hpbg='', HPBG_WEIGHTED=False, file_pointsource=[],
        # output files
        file_absorbed='default.absorbed', file_emitted='soc.emitted', file_temperature='', file_savetau='',
        file_pssavetau='', file_scattering='scattering', file_constant_save='', kernel_defs='',
        SAVE_INTENSITY_FILE='ISRF.DAT', FITS_PREFIX='map', FILE_ROI_SAVE='', FILE_ROI_LOAD='',
        # model and run parameters
        GL=0.0, MAP_DX=1.0, KDENSITY=1.0, DISTANCE=0.0, ITERATIONS=1, STEP_WEIGHT=[-1, 0, 0], DIR_WEIGHT=[-1, 0, 0],
        NPIX={'x': 10, 'y': 10}, FAST_MAP=-1, REMIT_F=[0.0, 1e30], SIM_F=[1.0e8, 1.0e17], LEVEL_THRESHOLD=0,
        INTOBS=np.array([-1e12, 0.0, 0.0], np.float32), MAPCENTRE=np.array([-1e12, 0.0, 0.0], np.float32),
        DEVICES='g', DSC_BINS=0, LOCAL=-1, GLOBAL=-1, BATCH=30, OBS_THETA=[], OBS_PHI=[],
        PSPAC=0, PS_METHOD=0, BGPAC=0, CLPAC=0, DFPAC=0, ROIPAC=0, NO_PS=0,
        PS_SCALING=np.ones(MAXPS, np.float32), PSPOS=pspos,
        DO_SPLIT=0, POLMAP=0, POLSTAT=0, POLSIM=0, NOSOLVE=0, LOAD_TEMPERATURE=0, NOMAP=0, NOABSORBED=0,
        SAVE_INTENSITY=0, USE_EMWEIGHT=0, EMWEIGHT_SKIP=3, EMWEIGHT_LIM=[0.0, 1e10, 0.0], SEED=np.pi / 4.0,
        MAP_FREQ=[1.0e6, 1e18], SINGLE_MAP_FREQ=np.asarray([], np.float32), FFS=1, BG_METHOD=0, WITH_ALI=0,
        WITH_REFERENCE=0, scale_background=1.0, LEVELS=999, K_DIFFUSE=1.0, SINGLE_ABU=0, OPT_IS_HALF=0,
        savetau_freq=[], pssavetau_freq=-1.0, ROI=np.zeros(6, np.int32), ROI_STEP=0, ROI_NSIDE=16, ROI_LOAD_SCALE=1.0,
        WITH_ROI_SAVE=0, WITH_ROI_LOAD=0, ROI_MAP=0, OUT_NSIDE=128, FSELECT=[], LIB_ABS=False, LIB_MAPS=False,
        MAP_INTERPOLATION=0, FITS=0, FITS_RA=0.0, FITS_DE=0.0, MIRROR='', VERBOSE=1, MMAP_ABSORBED=0, MMAP_EMITTED=0,
        CR_HEATING=0.0, ABSTHIN=-1, NNNLIMIT=0.0, AREA=0, KEYS={})


def _int_of_float(a):
    return int(float(a))


def _rounded(a):
    return int(round(float(a)))


def _seed(a):
    return float(np.clip(float(a), -1.0, 1.0))


def _um_range(a):
    """`lambda_min lambda_max` [um] -> [f_min, f_max] [Hz]."""
    return [um2f(float(a[1])), um2f(float(a[0]))]


# ---- keywords whose arguments need more than a conversion ------------------------------------------------------------
def _k_defs(u, a, line):
    u.kernel_defs = line[4:].split('#')[0]


def _k_mapum(u, a, line):
    add = np.asarray([um2f(float(x)) for x in a], np.float32)
    u.SINGLE_MAP_FREQ = np.sort(np.concatenate((u.SINGLE_MAP_FREQ, add))) if len(u.SINGLE_MAP_FREQ) + len(add) > 1 \
        else np.concatenate((u.SINGLE_MAP_FREQ, add))


def _k_savetau(u, a, line):
    u.file_savetau = a[0]
    u.savetau_freq += [0.0 if float(x) < 0.0 else um2f(float(x)) for x in a[1:]]


def _k_pssavetau(u, a, line):
    u.file_pssavetau, u.pssavetau_freq = a[0], um2f(float(a[1]))


def _k_fits(u, a, line):
    u.FITS = 1
    if len(a) >= 2:
        u.FITS_RA, u.FITS_DE = float(a[0]), float(a[1])
    if len(a) >= 3:
        u.FITS_PREFIX = a[2]


def _k_library(attr):
    def f(u, a, line):
        u.FSELECT = np.atleast_1d(np.asarray(np.loadtxt(a[0]), np.float32))
        setattr(u, attr, True)
    return f


def _k_diffuse(u, a, line):
    u.file_diffuse = a[0]
    if len(a) > 1:
        u.K_DIFFUSE = float(a[1])


def _k_optical(u, a, line):
    u.file_optical.append(a[0])
    u.file_abundance.append(a[1] if (len(a) > 1 and a[1][0:1] != '#') else '#')


def _k_background(u, a, line):
    u.file_background = a[0]
    if len(a) > 1:
        u.scale_background = float(a[1])


def _k_hpbg(u, a, line):
    u.file_hpbg = a[0]
    if len(a) > 1:
        u.scale_background = float(a[1])
    if len(a) > 2:
        u.HPBG_WEIGHTED = int(a[2])


def _k_saveint(u, a, line):
    u.SAVE_INTENSITY = int(a[0])
    if len(a) > 1:
        u.SAVE_INTENSITY_FILE = a[1]


def _k_emweight(u, a, line):
    u.USE_EMWEIGHT = int(a[0])
    if len(a) > 2:
        u.EMWEIGHT_LIM = [float(a[1]), float(a[2]), float(a[3]) if len(a) > 3 else 0.0]
        if len(a) > 4:
            u.EMWEIGHT_SKIP = int(a[4])


def _k_dsc(u, a, line):
    u.file_scafunc.append(a[0])
    if len(u.file_scafunc) == 1:
        u.DSC_BINS = int(a[1])
    elif u.DSC_BINS != int(a[1]):
        print("*** Error in scattering functions: number of bins must be the same for all dusts")
        sys.exit()


def _k_direction(u, a, line):
    if len(u.OBS_THETA) >= 10:
        print("** ERROR - cannot have more than 10 directions -- ABORT !!")
        sys.exit()
    u.OBS_THETA.append(float(a[0]) * D2R)
    u.OBS_PHI.append(float(a[1]) * D2R)


def _k_roisave(u, a, line):                 # roisave <file> <step>      (ASOC_aux.py:448-451)
    u.WITH_ROI_SAVE, u.FILE_ROI_SAVE, u.ROI_STEP = 1, a[0], int(a[1])


def _k_roiload(u, a, line):                 # roiload <file> <scale>     (ASOC_aux.py:452-455)
    u.WITH_ROI_LOAD, u.FILE_ROI_LOAD, u.ROI_LOAD_SCALE = 1, a[0], float(a[1])


def _k_mapping(u, a, line):
    u.NPIX = {'x': int(a[0]), 'y': int(a[1])}
    u.MAP_DX = float(a[2])
    if len(a) > 3:
        try:
            u.FAST_MAP = int(a[3])
        except ValueError:
            pass


def _k_roi(u, a, line):                     # roi x0 x1 y0 y1 z0 z1 (inclusive root cells, ASOC_aux.py:527)
    if len(a) >= 6:
        u.ROI = np.asarray([int(v) for v in a[:6]], np.int32)


def _k_mapview(u, a, line):
    u.OBS_THETA, u.OBS_PHI = [float(a[0]) * np.pi / 180.0], [float(a[1]) * np.pi / 180.0]
    if len(a) >= 4:
        u.NPIX = {'x': int(a[2]), 'y': int(a[3])}
    if len(a) >= 5:
        u.MAP_DX = float(a[4])
    if len(a) >= 8:
        u.MAPCENTRE = np.array([float(a[5]), float(a[6]), float(a[7])], np.float32)


def _k_pointsource(u, a, line):
    if u.NO_PS >= MAXPS:
        print("Reached maximum number of point sources = %d" % MAXPS)
        sys.exit()
    u.PSPOS[u.NO_PS] = [float(a[0]), float(a[1]), float(a[2])]
    u.file_pointsource.append(a[3])
    if len(a) > 4 and a[4] != '#':
        u.PS_SCALING[u.NO_PS] = float(a[4])
    u.NO_PS += 1


def _set(**values):
    def f(u, a, line):
        for k, v in values.items():
            setattr(u, k, v)
    return f


def _arg(attr, conv=str):
    def f(u, a, line):
        setattr(u, attr, conv(a[0]))
    return f


def _args(attr, conv):
    def f(u, a, line):
        setattr(u, attr, conv(a))
    return f


def _vec3(attr):
    return _args(attr, lambda a: np.array([float(a[0]), float(a[1]), float(a[2])], np.float32))


# (prefix, minimum number of arguments, action, how the keyword is compared)
KEYWORDS = [
    ('DEFS', 0, _k_defs, 'exact'), ('mapum', 0, _k_mapum, ''), ('singleabu', 0, _set(SINGLE_ABU=1), 'exact'),
    ('optishalf', 0, _set(OPT_IS_HALF=1), 'exact'),
    # flags and keywords that are recognised in any case
    ('nosolve', 0, _set(NOSOLVE=1), 'fold'), ('loadtemp', 0, _set(LOAD_TEMPERATURE=1), 'fold'),
    ('nomap', 0, _set(NOMAP=1), 'fold'), ('noabs', 0, _set(NOABSORBED=1), 'fold'),
    ('dustem', 0, _set(NOABSORBED=1, SAVE_INTENSITY=1), 'fold'), ('roimap', 0, _set(ROI_MAP=1), 'fold'),
    ('savetau', 2, _k_savetau, 'fold'), ('pssavetau', 2, _k_pssavetau, 'fold'), ('fits', 0, _k_fits, 'fold'),
    ('mirror', 1, _arg('MIRROR'), 'fold'),
    # one argument
    ('device', 1, _arg('DEVICES', str.lower), ''), ('verbose', 1, _arg('VERBOSE', int), ''),
    ('mmapabs', 1, _arg('MMAP_ABSORBED', int), ''), ('mmapemit', 1, _arg('MMAP_EMITTED', int), ''),
    ('tempera', 1, _arg('file_temperature'), ''), ('cloud', 1, _arg('file_cloud'), ''),
    ('absorb', 1, _arg('file_absorbed'), ''), ('scatter', 1, _arg('file_scattering'), ''),
    ('emit', 1, _arg('file_emitted'), ''), ('split', 1, _arg('DO_SPLIT', int), ''),
    ('mapint', 1, _arg('MAP_INTERPOLATION', int), ''), ('polstat', 1, _arg('POLSTAT', int), ''),
    ('absthin', 1, _arg('ABSTHIN', int), ''), ('nnnlimit', 1, _arg('NNNLIMIT', float), ''),
    ('libabs', 1, _k_library('LIB_ABS'), ''), ('libmap', 1, _k_library('LIB_MAPS'), ''),
    ('diffus', 1, _k_diffuse, ''), ('optic', 1, _k_optical, ''), ('backg', 1, _k_background, ''),
    ('hpbg', 1, _k_hpbg, ''), ('cload', 1, _arg('file_constant_load'), ''), ('csave', 1, _arg('file_constant_save'), ''),
    ('iterations', 1, _arg('ITERATIONS', int), ''), ('threshold', 1, _arg('LEVEL_THRESHOLD', int), ''),
    ('gridlen', 1, _arg('GL', float), ''), ('distance', 1, _arg('DISTANCE', float), ''),
    ('bgpac', 1, _arg('BGPAC', _int_of_float), ''), ('pspac', 1, _arg('PSPAC', _int_of_float), ''),
    ('psmetho', 1, _arg('PS_METHOD', int), ''), ('cellpac', 1, _arg('CLPAC', _rounded), ''),
    ('roipac', 1, _arg('ROIPAC', _rounded), ''), ('roinside', 1, _arg('ROI_NSIDE', _rounded), ''),
    ('diffpac', 1, _arg('DFPAC', int), ''), ('seed', 1, _arg('SEED', _seed), ''), ('dens', 1, _arg('KDENSITY', float), ''),
    ('CR_HEATING', 1, _arg('CR_HEATING', float), ''), ('batch', 1, _arg('BATCH', int), ''),
    ('local', 1, _arg('LOCAL', int), ''), ('global', 1, _arg('GLOBAL', int), ''),
    ('forcedfirst', 1, _arg('FFS', int), ''), ('ffs', 1, _arg('FFS', int), ''), ('bgmethod', 1, _arg('BG_METHOD', int), ''),
    ('ali', 1, _arg('WITH_ALI', int), ''), ('reference', 1, _arg('WITH_REFERENCE', int), ''),
    ('saveint', 1, _k_saveint, ''), ('levels', 1, _arg('LEVELS', int), ''), ('outnside', 1, _arg('OUT_NSIDE', int), ''),
    ('emwei', 1, _k_emweight, ''),
    # two arguments
    ('remit', 2, _args('REMIT_F', _um_range), ''), ('simum', 2, _args('SIM_F', _um_range), ''), ('dsc', 2, _k_dsc, ''),
    ('direwei', 2, _args('DIR_WEIGHT', lambda a: [int(a[0]), float(a[1])]), ''), ('direct', 2, _k_direction, ''),
    ('wavelen', 2, _args('MAP_FREQ', _um_range), ''), ('roisave', 2, _k_roisave, ''), ('roiload', 2, _k_roiload, ''),
    # three and more
    ('polsim', 3, _set(POLSIM=1), ''), ('polmap', 3, _set(POLMAP=1), ''), ('perspec', 3, _vec3('INTOBS'), ''),
    ('stepwei', 3, _args('STEP_WEIGHT', lambda a: [int(a[0]), float(a[1]), float(a[2])]), ''),
    ('mapping', 3, _k_mapping, ''), ('roi', 3, _k_roi, 'exact'), ('mapcent', 3, _vec3('MAPCENTRE'), ''),
    ('mapview', 3, _k_mapview, ''), ('pointsou', 4, _k_pointsource, ''),
]


class User:
    def __init__(self, filename):
        for k, v in _defaults().items():
            setattr(self, k, v)
        for line in open(filename).readlines():
            s = line.split('#')[0].split()
            if len(s) < 1:
                continue
            key, args = s[0], s[1:]
            self.KEYS[key] = args
            for prefix, nargs, action, how in KEYWORDS:
                if len(args) < nargs:
                    continue
                if how == 'exact':
                    hit = key == prefix
                else:
                    hit = (key.lower() if how == 'fold' else key).startswith(prefix)
                if hit:
                    action(self, args, line)
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
        checks = [
            (self.DO_SPLIT, "split: packet splitting (SimBgSplit/SimHpSplit) is not implemented"),
            (self.POLMAP or self.POLSIM or self.POLSTAT, "polmap/polsim/polstat: polarisation maps are not implemented"),
            (self.DIR_WEIGHT[0] > 0, "dirweight: does not compile in the reference either (undeclared pweight)"),
            (self.PS_METHOD == 3, "psmethod 3: not implemented in the reference either"),
            (1 < self.FAST_MAP < 999, "mapping ... fast: kernel_ASOC_map_X.c does not exist in the reference either (ASOC.py:3442)"),
            (self.FAST_MAP >= 999 and self.NPIX['y'] <= 0, "mapping ... 999 with a Healpix map: per-level Healpix maps are not implemented"),
            (self.USE_EMWEIGHT > 1, "emweight 2 is not implemented"),
            (self.LIB_ABS or self.LIB_MAPS, "libabs/libmaps: the library method is not implemented"),
            (self.ABSTHIN > 1, "absthin is not implemented"),
            (self.CR_HEATING > 0.0, "CR_HEATING is not implemented"),
            (self.SAVE_INTENSITY == 3, "saveint 3 (intensity file derived from the absorbed file, ASOC.py:2839-2862) is not implemented"),
            (self.SAVE_INTENSITY in (1, 2) and any(a != '#' for a in self.file_abundance),
             "saveint 1/2 with abundance files: the reference divides by a stale scalar ABS here (ASOC.py:1503); not supported"),
        ]
        return [msg for cond, msg in checks if cond]
