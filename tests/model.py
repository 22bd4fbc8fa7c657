"""Builds small complete SOC models on disk (cloud, dust, dsc, sources, ini) for end-to-end driver tests."""
import os

import numpy as np

from soc_b200 import synth
from soc_b200.formats import write_cloud, write_dust, write_dsc


def write_model(path, n=12, octree=False, nfreq=8, bgpac=40000, pspac=0, cellpac=0, iterations=1, extra="",
                noabsorbed=True, absorbed=False, seed=0.4321, maps=True, abundance=False, two_dusts=False, hpbg=False):
    """Returns the ini file name.  Frequencies 3e11..3e15 Hz, silicate-like toy dust, HG scattering tables."""
    os.makedirs(path, exist_ok=True)
    cloud = synth.octree_cloud(n, 3, refine_fraction=0.2, seed=5) if octree else synth.regular_cloud(n, 0.25)
    write_cloud(os.path.join(path, "model.cloud"), cloud)
    freq, g, qabs, qsca = synth.synthetic_dust(nfreq, 3.0e11, 3.0e15)
    write_dust(os.path.join(path, "toy.dust"), freq, g, qabs, qsca, grain_density=1.0e-7, grain_size=1.0e-5)
    freq32 = np.loadtxt(os.path.join(path, "toy.dust"), skiprows=4)[:, 0]
    bins = 500
    dsc = np.zeros((nfreq, bins), np.float32)
    csc = np.zeros((nfreq, bins), np.float32)
    for i in range(nfreq):
        dsc[i], csc[i] = synth.hg_tables(g[i], bins)
    write_dsc(os.path.join(path, "toy.dsc"), dsc, csc)
    synth.isrf_like_background(freq32, 1.0).tofile(os.path.join(path, "bg.bin"))
    synth.blackbody_source(freq32, 6000.0, 0.05).tofile(os.path.join(path, "ps.bin"))
    if abundance:
        rng = np.random.default_rng(3)
        (0.5 + rng.random(cloud.CELLS)).astype(np.float32).tofile(os.path.join(path, "abu.bin"))
    if two_dusts:
        # second species: same frequency grid, different albedo and asymmetry, own abundance file and dsc file
        write_dust(os.path.join(path, "toy2.dust"), freq, 0.3 * g, 1.4 * qabs, 0.5 * qsca, grain_density=1.0e-7, grain_size=1.0e-5)
        dsc2 = np.zeros((nfreq, bins), np.float32)
        csc2 = np.zeros((nfreq, bins), np.float32)
        for i in range(nfreq):
            dsc2[i], csc2[i] = synth.hg_tables(0.3 * g[i], bins)
        write_dsc(os.path.join(path, "toy2.dsc"), dsc2, csc2)
        rng = np.random.default_rng(4)
        (0.5 + rng.random(cloud.CELLS)).astype(np.float32).tofile(os.path.join(path, "abu1.bin"))
        (0.2 + rng.random(cloud.CELLS)).astype(np.float32).tofile(os.path.join(path, "abu2.bin"))
    if hpbg:
        rng = np.random.default_rng(6)
        sky = (0.5 + rng.random(49152)).astype(np.float32)
        sky[10000:12000] *= 5.0
        np.outer(synth.isrf_like_background(freq32, 1.0), sky / sky.mean()).astype(np.float32).tofile(os.path.join(path, "hpbg.bin"))
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
