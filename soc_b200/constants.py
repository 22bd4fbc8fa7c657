"""Physical constants and fixed numbers of the SOC host code (values as in ASOC_aux.py:28-54)."""
FACTOR = 1.0e20            # scaling of absorbed/emitted photon numbers in the on-disk files
C_LIGHT = 2.99792458e10
PLANCK = 6.62606957e-27
H_K = 4.79924335e-11
D2R = 0.0174532925
PARSEC = 3.08567758e+18
H_CC20 = 7.372496678e-28
SEED0 = 0.8150982470475214
SEED1 = 0.1393378751427912
MAXPS = 4000
ADHOC = 1.0                # ASOC.py:81
GLOBAL_0 = 32768           # ASOC.py:86 (work items of the PS / cell-emission launches)
GLOBAL_0_SCA = 65536       # ASOCS.py:79
HPBG_NPIX = 49152          # Healpix background is fixed to NSIDE=64 (ASOC.py:297)


def um2f(um):
    return C_LIGHT / (1.0e-4 * um)


def f2um(f):
    return 1.0e4 * C_LIGHT / f
