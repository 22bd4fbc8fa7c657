"""Host-side arithmetic of the SOC drivers that has to be restated exactly: observer geometry, packet
count rounding, source weights, the trapezoid frequency weights, the E->T table."""
import numpy as np

from .constants import PLANCK, PARSEC, FACTOR, H_K, H_CC20


def fix(n, l):
    """Smallest integer >= n divisible by l (ASOC_aux.py:1504-1506)."""
    return int(int(np.floor((n + l - 1) / l)) * l)


def observer_directions(theta_deg, phi_deg):
    """Observer geometry from ini-file angles in degrees (theta from +Z, phi from +X)."""
    theta = np.atleast_1d(np.asarray(theta_deg, np.float64)) * 0.0174532925       # D2R of ASOC_aux.py:37
    phi = np.atleast_1d(np.asarray(phi_deg, np.float64)) * 0.0174532925
    return observer_directions_rad(theta, phi)


def observer_directions_rad(theta, phi):
    """Unit vectors (towards observer, map right, map up) per direction, ASOC_aux.py:1129-1183:
    (ODIR, RA, DE) = R (x, y, z) with b = latitude = pi/2 - theta, a = phi; ODIR components with
    |x| < 1e-5 are set to 1e-5.  Returns NDIR and three [NDIR,3] float32 arrays."""
    theta = np.atleast_1d(np.asarray(theta, np.float64))
    phi = np.atleast_1d(np.asarray(phi, np.float64))
    n = len(theta)
    od = np.zeros((n, 3), np.float32)
    ra = np.zeros((n, 3), np.float32)
    de = np.zeros((n, 3), np.float32)
    R = np.zeros((3, 3), np.float32)
    for i in range(n):
        b, a = 0.5 * np.pi - theta[i], phi[i]
        R[0, :] = [np.cos(a) * np.cos(b), -np.sin(a), -np.cos(a) * np.sin(b)]
        R[1, :] = [np.sin(a) * np.cos(b), np.cos(a), -np.sin(a) * np.sin(b)]
        R[2, :] = [np.sin(b), 0.0, np.cos(b)]
        od[i] = np.matmul(R, [1, 0, 0])
        ra[i] = np.matmul(R, [0, 1, 0])
        de[i] = np.matmul(R, [0, 0, 1])
        for j in range(3):
            if abs(od[i, j]) < 1.0e-5:
                od[i, j] = 1.0e-5
    return n, od, ra, de


def trapezoid_weights(ffreq):
    """FF[i] = f_i * 0.5*(f_{i+1}-f_{i-1}) with one-sided ends (ASOC.py:1218-1223)."""
    f = np.asarray(ffreq, np.float32)
    n = len(f)
    w = np.empty(n, np.float64)
    if n == 1:
        return np.asarray(f, np.float64)
    for i in range(n):
        ff = f[i]
        if i == 0:
            ff = ff * (0.5 * (f[1] - f[0]))
        elif i == n - 1:
            ff = ff * (0.5 * (f[n - 1] - f[n - 2]))
        else:
            ff = ff * (0.5 * (f[i + 1] - f[i - 1]))
        w[i] = ff
    return w


def planck_safe(f, t):
    return 2.0e-20 * ((H_CC20 * f) * f) * f / (np.exp(np.clip(H_K * f / t, -100, +100)) - 1.0)


def energy_temperature_table(ffreq, fabs, gl, ne=30000):
    """E -> T lookup for the equilibrium-temperature solve (ASOC.py:643-689).
    Returns Emin, kE, TTT[ne] with TTT[i] = T(E = Emin*kE^i)."""
    ffreq = np.asarray(ffreq, np.float32)
    f64 = np.asarray(ffreq, np.float64)
    fabs = np.asarray(fabs, np.float32)
    tstep = 1600.0 / ne
    tt = 1.0 + tstep * np.arange(ne, dtype=np.float64)
    df = (ffreq[2:] - ffreq[:-2])
    # integral of FABS*B(T) over frequency for every T at once (trapezoid rule of the reference)
    x = np.clip(H_K * f64[None, :] / tt[:, None], -100, 100)
    tmp = fabs[None, :] * (2.0e-20 * ((H_CC20 * f64) * f64) * f64)[None, :] / (np.exp(x) - 1.0)
    res = tmp[:, 0] * (ffreq[1] - ffreq[0]) + tmp[:, -1] * (ffreq[-1] - ffreq[-2])
    res = res + np.sum(tmp[:, 1:-1] * df[None, :], axis=1)
    eout = (4.0 * np.pi * FACTOR / (gl * PARSEC)) * 0.5 * res
    emin, emax = eout[0], eout[ne - 1] * 0.9999
    ke = (emax / emin) ** (1.0 / (ne - 1.0))
    ttt = np.interp(emin * ke ** np.arange(ne), eout, tt).astype(np.float32)
    return emin, ke, ttt


def source_weights_ps(pspac_requested, no_ps, gl, global_0):
    """PS launch decomposition and weight (ASOC.py:1036-1045): returns BATCH (per work item, over all
    sources), total packets, WPS."""
    batch = int(max(1, pspac_requested / global_0))
    pspac = global_0 * batch
    wps = 1.0 / (PLANCK * pspac * ((gl * PARSEC) ** 2.0))
    return batch * no_ps, pspac * no_ps, wps


def source_weights_bg(bgpac_requested, area):
    """Isotropic background decomposition (ASOC.py:1061-1064): BATCH, BGPAC, WBG, GLOBAL."""
    batch = max(1, int(round(bgpac_requested / (8 * area))))
    bgpac = int(8 * area * batch)
    wbg = np.pi / (PLANCK * 8 * batch)
    glob = fix(int(8 * area), 64)
    return batch, bgpac, wbg, glob
