"""Synthetic model generators used by tests and bench.py: regular-grid and
octree clouds, grey-ish dust with Henyey-Greenstein scattering tables, source spectra.

The .dsc recipe follows DustLib.write_eqdust_dsc (DustLib.py:2532-2563): DSC = HG phase function per
steradian on cos(theta)=linspace(-1,1,BINS); CSC = cos(theta) at cumulative probability
linspace(0,1,BINS), starting from +1 (forward).
"""
import os

import numpy as np

from .constants import C_LIGHT, H_K, H_CC20
from .formats import Cloud, links_to_float, write_cloud, write_dust, write_dsc


def plummer_density(n, r0_frac=0.1, n0=1.0):
    """n(r) = n0 / (1 + (r/r0)^2) on an n^3 grid, r0 = r0_frac*n cells (SURVEY.md section 6 probe cloud)."""
    c = (np.arange(n, dtype=np.float32) + 0.5 - 0.5 * n)
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    r2 = x * x + y * y + z * z
    return (n0 / (1.0 + r2 / (r0_frac * n) ** 2)).astype(np.float32)


def regular_cloud(n, r0_frac=0.1, n0=1.0):
    d = plummer_density(n, r0_frac, n0)
    return Cloud(n, n, n, [n ** 3], d.ravel())


def lognormal_field(n, sigma=1.5, slope=-3.7, seed=12345):
    """Turbulent-looking log-normal density on an n^3 grid (power-law spectrum k^slope, sigma of ln n)."""
    rng = np.random.default_rng(seed)
    k = np.fft.fftfreq(n) * n
    kz, ky, kx = np.meshgrid(k, k, k, indexing="ij")
    kk = np.sqrt(kx * kx + ky * ky + kz * kz)
    kk[0, 0, 0] = 1.0
    amp = kk ** (0.5 * slope)
    amp[0, 0, 0] = 0.0
    f = np.fft.ifftn(amp * np.exp(2j * np.pi * rng.random((n, n, n)))).real
    f = (f - f.mean()) / f.std()
    return np.exp(sigma * f - 0.5 * sigma * sigma).astype(np.float32)


def box_cloud(nx, ny, nz, levels=1, refine_fraction=0.2, seed=7):
    """Non-cubic test cloud: smooth random density on an nx*ny*nz root grid (x fastest), optionally refined like
    octree_cloud()."""
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    d0 = (0.4 + np.exp(-((x - 0.4 * nx) ** 2 / (0.3 * nx) ** 2 + (y - 0.6 * ny) ** 2 / (0.3 * ny) ** 2
                          + (z - 0.5 * nz) ** 2 / (0.3 * nz) ** 2)) + 0.2 * rng.random((nz, ny, nx))).astype(np.float32)
    if levels == 1:
        return Cloud(nx, ny, nz, [nx * ny * nz], d0.ravel())
    c = octree_cloud(1, levels, refine_fraction=refine_fraction, seed=seed, base=d0.ravel())
    return Cloud(nx, ny, nz, c.LCELLS, c.DENS)


def octree_cloud(nroot, levels, refine_fraction=0.15, sigma=1.5, seed=12345, base=None):
    """Hierarchical cloud in SOC's format (ASOC_aux.py:734-744): level 0 is an x-fastest nroot^3 grid,
    deeper levels are runs of 8 children ordered sid = 4z+2y+x; a refined cell stores the link
    -bitcast(first child index within the next level).  On every level the densest `refine_fraction`
    of the leaves is refined; children scatter log-normally around the parent value."""
    rng = np.random.default_rng(seed)
    d0 = (lognormal_field(nroot, sigma=sigma, seed=seed) if base is None else np.asarray(base, np.float32)).ravel()
    vals = [d0.copy()]
    for l in range(levels - 1):
        cur = vals[l]
        leaves = np.nonzero(cur > 0.0)[0]
        if l > 0:
            pass
        nref = int(refine_fraction * len(leaves))
        if nref < 1:
            break
        order = leaves[np.argsort(cur[leaves])[::-1][:nref]]
        order.sort()
        child = np.empty(8 * nref, np.float32)
        pert = np.exp(0.35 * rng.standard_normal((nref, 8))).astype(np.float32)
        pert /= pert.mean(axis=1, keepdims=True)
        child[:] = (cur[order][:, None] * pert).ravel()
        cur[order] = links_to_float(8 * np.arange(nref, dtype=np.int32))
        vals.append(child)
    lcells = [len(v) for v in vals]
    return Cloud(nroot, nroot, nroot, lcells, np.concatenate(vals))


def hg_phase(cos_theta, g):
    """Henyey-Greenstein probability per solid angle (DustLib.py:123-127)."""
    return (1.0 / (4.0 * np.pi)) * (1.0 - g * g) / (1.0 + g * g - 2.0 * g * cos_theta) ** 1.5


def hg_tables(g, bins=2500):
    """DSC[bins], CSC[bins] for asymmetry parameter g (recipe of DustLib.py:2546-2563)."""
    g = float(g)
    if abs(g) < 1e-4:
        g = 1e-4
    cos_theta = np.linspace(-1.0, 1.0, bins)
    x = hg_phase(cos_theta, g)
    dsc = np.clip(x, 1e-5 * x.max(), 1e20).astype(np.float32)
    theta = np.linspace(0.0, np.pi, 5 * bins)
    y = 2.0 * np.pi * np.sin(theta) * hg_phase(np.cos(theta), g)
    p = np.cumsum(y) + 1e-7 * np.cumsum(np.ones(len(y)))
    p -= p[0]
    p /= p[-1]
    p[0] = -1.0e-7
    p[-1] = 1.0 + 1.0e-7
    csc = np.interp(np.linspace(0.0, 1.0, bins), p, np.cos(theta)).astype(np.float32)
    return dsc, csc


def synthetic_dust(nfreq=44, fmin=1.5e11, fmax=3.0e15, gmax=0.6):
    """Frequency grid and Q factors of a toy silicate/carbon-like grain: Qabs ~ nu^1.7 saturating at 1,
    Qsca ~ nu^4 saturating at 1.2, g ramping 0 -> gmax (cf. tmp.dust of soc_example.zip)."""
    freq = np.logspace(np.log10(fmin), np.log10(fmax), nfreq)
    x = freq / 3.0e14
    qabs = 1.0 * x ** 1.7 / (1.0 + x ** 1.7)
    qsca = 1.2 * x ** 4 / (1.0 + x ** 4)
    g = gmax * x ** 1.2 / (1.0 + x ** 1.2)
    return freq, g, qabs, qsca


def planck(f, t):
    return 2.0e-20 * ((H_CC20 * f) * f) * f / (np.exp(np.clip(H_K * f / t, -80, 80)) - 1.0)


def isrf_like_background(freq, scale=1.0):
    """Diluted black bodies, a crude interstellar radiation field [erg/s/cm2/sr/Hz]."""
    return (scale * (1e-14 * planck(freq, 7500.0) + 1e-13 * planck(freq, 4000.0)
                     + 4e-13 * planck(freq, 3000.0) + 1.0e-5 * planck(freq, 250.0)
                     + planck(freq, 2.73))).astype(np.float32)


def blackbody_source(freq, t=10000.0, lum_lsun=1.0):
    """Luminosity per Hz [erg/s/Hz] of a black body of temperature t normalised to lum_lsun."""
    b = planck(freq, t)
    lum = 3.846e33 * lum_lsun * b / np.trapezoid(b, freq)
    return lum.astype(np.float32)


def write_model(path, n=12, octree=False, nfreq=8, bgpac=40000, pspac=0, cellpac=0, iterations=1, extra="",
                noabsorbed=True, absorbed=False, seed=0.4321, maps=True, abundance=False, two_dusts=False, hpbg=False):
    """Returns the ini file name.  Frequencies 3e11..3e15 Hz, silicate-like toy dust, HG scattering tables."""
    os.makedirs(path, exist_ok=True)
    cloud = octree_cloud(n, 3, refine_fraction=0.2, seed=5) if octree else regular_cloud(n, 0.25)
    write_cloud(os.path.join(path, "model.cloud"), cloud)
    freq, g, qabs, qsca = synthetic_dust(nfreq, 3.0e11, 3.0e15)
    write_dust(os.path.join(path, "toy.dust"), freq, g, qabs, qsca, grain_density=1.0e-7, grain_size=1.0e-5)
    freq32 = np.loadtxt(os.path.join(path, "toy.dust"), skiprows=4)[:, 0]
    bins = 500
    dsc = np.zeros((nfreq, bins), np.float32)
    csc = np.zeros((nfreq, bins), np.float32)
    for i in range(nfreq):
        dsc[i], csc[i] = hg_tables(g[i], bins)
    write_dsc(os.path.join(path, "toy.dsc"), dsc, csc)
    isrf_like_background(freq32, 1.0).tofile(os.path.join(path, "bg.bin"))
    blackbody_source(freq32, 6000.0, 0.05).tofile(os.path.join(path, "ps.bin"))
    if abundance:
        rng = np.random.default_rng(3)
        (0.5 + rng.random(cloud.CELLS)).astype(np.float32).tofile(os.path.join(path, "abu.bin"))
    if two_dusts:
        # second species: same frequency grid, different albedo and asymmetry, own abundance file and dsc file
        write_dust(os.path.join(path, "toy2.dust"), freq, 0.3 * g, 1.4 * qabs, 0.5 * qsca, grain_density=1.0e-7, grain_size=1.0e-5)
        dsc2 = np.zeros((nfreq, bins), np.float32)
        csc2 = np.zeros((nfreq, bins), np.float32)
        for i in range(nfreq):
            dsc2[i], csc2[i] = hg_tables(0.3 * g[i], bins)
        write_dsc(os.path.join(path, "toy2.dsc"), dsc2, csc2)
        rng = np.random.default_rng(4)
        (0.5 + rng.random(cloud.CELLS)).astype(np.float32).tofile(os.path.join(path, "abu1.bin"))
        (0.2 + rng.random(cloud.CELLS)).astype(np.float32).tofile(os.path.join(path, "abu2.bin"))
    if hpbg:
        rng = np.random.default_rng(6)
        sky = (0.5 + rng.random(49152)).astype(np.float32)
        sky[10000:12000] *= 5.0
        np.outer(isrf_like_background(freq32, 1.0), sky / sky.mean()).astype(np.float32).tofile(os.path.join(path, "hpbg.bin"))
    ini = os.path.join(path, "model.ini")
    with open(ini, "w") as fp:
        fp.write("cloud        model.cloud\n")
        if two_dusts:
            fp.write("optical      toy.dust  abu1.bin\noptical      toy2.dust  abu2.bin\n")
            fp.write("dsc          toy.dsc %d\ndsc          toy2.dsc %d\n" % (bins, bins))
        else:
            fp.write("optical      toy.dust%s\n" % ("  abu.bin" if abundance else ""))
            fp.write("dsc          toy.dsc %d\n" % bins)
        if hpbg:
            fp.write("hpbg         hpbg.bin 1.0 %d\n" % (1 if hpbg == 2 else 0))
        fp.write("gridlength   0.02\ndensity      %.3e   # scaling of densities\n" % (6.0 / n))   # tau_V across the model ~ 10
        fp.write("background   bg.bin  1.0\n")
        fp.write("bgpackets    %d\n" % bgpac)
        if pspac > 0:
            c = 0.5 * n + 0.3
            fp.write("pointsource  %.2f %.2f %.2f  ps.bin  1.0\npspackets    %d\n" % (c, c, c, pspac))
        if cellpac > 0:
            fp.write("cellpackets  %d\n" % cellpac)
        fp.write("iterations   %d\nseed         %.4f\n" % (iterations, seed))
        if noabsorbed:
            fp.write("noabsorbed\n")
        else:
            fp.write("nosolve\n")
        if absorbed:
            fp.write("absorbed     abs.data\n")
        fp.write("emitted      emit.data\ntemperature  model.T\n")
        if maps:
            fp.write("mapping      %d %d 1.0\ndirections   0.0 0.0\ndirections   70.0 30.0\n" % (n, n))
        else:
            fp.write("nomap\n")
        fp.write("verbose 0\n")
        fp.write(extra)
    return ini, cloud
