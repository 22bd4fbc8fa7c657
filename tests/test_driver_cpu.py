"""End-to-end tests of the ASOC driver's host logic on the CPU: the driver runs unchanged, its device is the
oracle-backed stand-in of tests/oracle_device.py.  Checks the process contract (files, formats, headers), the
unit chain (sane temperatures), and the rank logic with two gloo ranks."""
import os
import subprocess
import sys

import numpy as np
import pytest

from soc_b200 import asoc
from soc_b200.formats import read_otfile, read_map_file, read_cells_freq_file
from soc_b200.ini import User
from tests.model import write_model
from tests.oracle_device import OracleDevice

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tmp, **kw):
    ini, cloud = write_model(str(tmp), **kw)
    cwd = os.getcwd()
    os.chdir(str(tmp))
    try:
        asoc.main(["ASOC.py", "model.ini"], device_factory=OracleDevice)
    finally:
        os.chdir(cwd)
    return cloud


def test_ini_keywords_prefix_matching(tmp_path):
    ini, _ = write_model(str(tmp_path), pspac=1000, extra="DEFS -D X=1 # c\nmapum 100 250\nCLT\nsavetau tau -1 250\nstepweight 2 0.5 0.3\n")
    U = User(ini)
    assert U.file_cloud == "model.cloud" and U.DSC_BINS == 500 and U.BGPAC == 40000 and U.PSPAC == 1000
    assert U.NO_PS == 1 and abs(U.PSPOS[0][0] - 6.3) < 1e-6 and U.NOABSORBED == 1 and "CLT" in U.KEYS
    assert len(U.SINGLE_MAP_FREQ) == 2 and U.SINGLE_MAP_FREQ[0] < U.SINGLE_MAP_FREQ[1]
    assert U.savetau_freq[0] == 0.0 and U.STEP_WEIGHT == [2, 0.5, 0.3]
    assert len(U.OBS_THETA) == 2 and U.unsupported() == []
    open(ini, "a").write("split 1\npolmap a b c\n")
    assert len(User(ini).unsupported()) == 2


def test_background_run_writes_the_reference_files(tmp_path):
    cloud = _run(tmp_path, n=10, bgpac=30000)
    n = cloud.CELLS
    assert (np.fromfile(tmp_path / "packet.info", np.int32) == [30016, 0, 0, 0]).all()
    T = read_otfile(str(tmp_path / "model.T"))
    assert T.shape == (n,) and (T >= 3.0).all() and (T < 60.0).all() and T.std() > 0.01
    hdr = np.fromfile(tmp_path / "model.T", np.int32, 6)
    assert list(hdr[:5]) == [10, 10, 10, 1, n] and hdr[5] == n
    em = read_cells_freq_file(str(tmp_path / "emit.data"))
    assert em.shape == (n, 8) and (em >= 0).all() and em.max() > 0
    for idir in range(2):
        m = read_map_file(str(tmp_path / ("map_dir_%02d.bin" % idir)))
        assert m.shape == (8, 10, 10) and np.isfinite(m).all() and m[:4].min() > 0.0


def test_pssavetau_writes_source_optical_depths(tmp_path):
    _run(tmp_path, n=8, bgpac=10000, pspac=33000, extra="pssavetau pstau 100.0\n")
    rows = np.loadtxt(str(tmp_path / "pstau_0.dat"), ndmin=2)
    assert rows.shape == (1, 3) and rows[0, 1] > 0 and rows[0, 2] > 0
    assert os.path.exists(str(tmp_path / "pstau_1.dat"))


def test_per_level_maps(tmp_path):
    """`mapping nx ny dx 999` (ASOC.py:3320-3440): map_dir_%02d_H.bin = [npx, npy], [nfreq, LEVELS], then LEVELS images per
    frequency; their sum over the levels is the ordinary map of the same run."""
    cloud = _run(tmp_path / "lev", n=6, octree=True, bgpac=20000, extra="mapping 6 6 1.0 999\nsavetau colden -1\n")
    _run(tmp_path / "all", n=6, octree=True, bgpac=20000, extra="mapping 6 6 1.0\n")
    raw = np.fromfile(str(tmp_path / "lev" / "map_dir_01_H.bin"), np.int32, 4)
    assert list(raw[:2]) == [6, 6] and raw[3] == cloud.LEVELS
    nfreq = int(raw[2])
    lev = np.fromfile(str(tmp_path / "lev" / "map_dir_01_H.bin"), np.float32, offset=16).reshape(nfreq, cloud.LEVELS, 6, 6)
    full = read_map_file(str(tmp_path / "all" / "map_dir_01.bin"))
    assert full.shape == (nfreq, 6, 6)
    assert (lev[:, 1:] > 0).any() and np.allclose(lev.sum(axis=1), full, rtol=2e-5, atol=0)
    col = np.fromfile(str(tmp_path / "lev" / "colden.1"), np.float32, offset=8)
    assert col.shape == (36,) and (col > 0).all()


def test_absorbed_file_and_octree(tmp_path):
    cloud = _run(tmp_path, n=6, octree=True, bgpac=20000, pspac=33000, noabsorbed=False, absorbed=True, maps=False)
    a = read_cells_freq_file(str(tmp_path / "abs.data"))
    assert a.shape == (cloud.CELLS, 8)
    parents = cloud.DENS <= 0.0
    assert (a[parents] == np.float32(-1.0e20)).all()
    assert (a[~parents] >= 0.0).all() and (a[~parents].sum(axis=0) > 0).all()


def test_cell_emission_iterations_with_reference_field(tmp_path):
    cloud = _run(tmp_path, n=8, bgpac=20000, cellpac=8 ** 3 * 4, iterations=2, maps=False, extra="reference 1\n")
    T = read_otfile(str(tmp_path / "model.T"))
    assert (T >= 3.0).all() and (T < 80.0).all()


def test_two_gloo_ranks_agree_with_one(tmp_path):
    """world_size 2 over gloo: both ranks run, rank 0 writes; temperatures agree with the 1-rank run within noise."""
    one = tmp_path / "one"
    two = tmp_path / "two"
    _run(one, n=8, bgpac=60000, pspac=33000)
    write_model(str(two), n=8, bgpac=60000, pspac=33000)
    (two / "run.py").write_text(
        "import sys\nsys.path.insert(0, %r)\nfrom soc_b200 import asoc\nfrom tests.oracle_device import OracleDevice\n"
        "asoc.main(['ASOC.py', 'model.ini'], device_factory=OracleDevice)\n" % ROOT)
    env = dict(os.environ, SOC_DIST_BACKEND="gloo", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", "run.py"], cwd=str(two), env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    T1, T2 = read_otfile(str(one / "model.T")), read_otfile(str(two / "model.T"))
    assert np.abs(T2 / T1 - 1.0).max() < 0.05
    m1, m2 = read_map_file(str(one / "map_dir_00.bin")), read_map_file(str(two / "map_dir_00.bin"))
    assert np.abs(m2[:2] / m1[:2] - 1.0).max() < 0.1


@pytest.mark.parametrize("mode", ["frequencies", "packets"])
def test_two_gloo_ranks_absorbed_file(tmp_path, mode):
    """world_size 2 over gloo, absorbed file on: by default the constant sources are sharded by frequency (rank r runs
    frequencies r, r+2, ... whole; no per-frequency collective), with PACKETSHARD by packet index.  Both must give the
    1-rank absorbed file within Monte Carlo noise -- frequency sharding runs exactly the 1-rank launches, so with a fixed
    seed it reproduces that file to rounding."""
    from soc_b200.formats import read_cells_freq_file
    kw = dict(n=8, bgpac=60000, pspac=33000, noabsorbed=False, absorbed=True, maps=False)
    one, two = tmp_path / "one", tmp_path / "two"
    _run(one, **kw)
    write_model(str(two), extra="verbose 1\n" + ("PACKETSHARD\n" if mode == "packets" else ""), **kw)
    (two / "run.py").write_text(
        "import sys\nsys.path.insert(0, %r)\nfrom soc_b200 import asoc\nfrom tests.oracle_device import OracleDevice\n"
        "asoc.main(['ASOC.py', 'model.ini'], device_factory=OracleDevice)\n" % ROOT)
    env = dict(os.environ, SOC_DIST_BACKEND="gloo", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29537" if mode == "packets" else "29539", "run.py"],
                       cwd=str(two), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert ("sharded by frequency" in r.stdout) == (mode == "frequencies"), r.stdout[-1500:]
    a1 = read_cells_freq_file(str(one / "abs.data")).astype(np.float64)
    a2 = read_cells_freq_file(str(two / "abs.data")).astype(np.float64)
    assert a1.shape == a2.shape
    t1, t2 = a1.sum(axis=0), a2.sum(axis=0)
    ok = t1 > 0
    if mode == "frequencies":
        assert np.abs(t2[ok] / t1[ok] - 1.0).max() < 1e-5
        assert np.abs(a2 - a1).max() <= 1e-5 * a1.max()
    else:
        assert np.abs(t2[ok] / t1[ok] - 1.0).max() < 0.03


def test_absorbed_file_hand_off_per_species(tmp_path):
    """soc_b200.a2e_handoff: the absorbed file of a two-species run split into the files the dust solver of each species
    reads (A2E_MABU.py:700-705); abundance-weighted shares add up to the mixture's absorptions."""
    from soc_b200.a2e_handoff import split_absorbed_file
    from soc_b200.formats import read_cells_freq_file
    cloud = _run(tmp_path, n=8, bgpac=40000, two_dusts=True, noabsorbed=False, absorbed=True, maps=False)
    a = read_cells_freq_file(str(tmp_path / "abs.data")).astype(np.float64)
    rng = np.random.default_rng(2)
    rabs = 1e-21 * (0.5 + rng.random((a.shape[1], 2)))
    abus = [str(tmp_path / "abu1.bin"), str(tmp_path / "abu2.bin")]
    parts = []
    for d in range(2):
        out = str(tmp_path / ("abs_%d.data" % d))
        split_absorbed_file(str(tmp_path / "abs.data"), rabs, d, out, abus, batch=200, device_factory=OracleDevice)
        parts.append(read_cells_freq_file(out).astype(np.float64))
        assert parts[-1].shape == a.shape
    abu = np.stack([np.fromfile(f, np.float32, cloud.CELLS) for f in abus], axis=1).astype(np.float64)
    ok = a > 0
    tot = parts[0] * abu[:, :1] + parts[1] * abu[:, 1:]
    assert np.abs(tot[ok] / a[ok] - 1.0).max() < 1e-5


def test_scattered_light_driver_writes_outcoming(tmp_path):
    from soc_b200 import asocs
    from soc_b200.formats import read_outcoming
    ini, cloud = write_model(str(tmp_path), n=8, bgpac=20000, pspac=66000)
    cwd = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        asocs.main(["ASOCS.py", "model.ini"], device_factory=OracleDevice)
    finally:
        os.chdir(cwd)
    freq, out = read_outcoming(str(tmp_path / "outcoming.socs"))
    assert out.shape == (8, 2, 8, 8) and len(freq) == 8
    assert np.isfinite(out).all() and (out >= 0).all() and out[4:].max() > 0


def _run_asocs(tmp_path):
    from soc_b200 import asocs
    cwd = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        asocs.main(["ASOCS.py", "model.ini"], device_factory=OracleDevice)
    finally:
        os.chdir(cwd)


def test_two_scattering_functions_mirror_and_fits(tmp_path):
    """Two dust species with their own abundance and dsc files (=> WITH_MSF), reflecting borders, FITS products."""
    from soc_b200.fits import read_fits
    kw = dict(n=8, bgpac=20000, pspac=33000, two_dusts=True, noabsorbed=False, absorbed=True)
    write_model(str(tmp_path), **kw)
    um = 2.99792458e14 / np.loadtxt(str(tmp_path / "toy.dust"), skiprows=4)[:, 0]       # wavelengths of the grid
    cloud = _run(tmp_path, extra="mirror xY\nfits 12.5 -3.0 img\nmapum %.4f %.4f\nsavetau tau -1 %.4f\n" % (um[1], um[2], um[2]), **kw)
    a = read_cells_freq_file(str(tmp_path / "abs.data"))          # several dusts: the solve is external (A2E)
    assert a.shape == (cloud.CELLS, 8) and (a >= 0).all() and (a.sum(axis=0) > 0).all()
    names = sorted(f for f in os.listdir(str(tmp_path)) if f.endswith(".fits"))
    assert any(f.startswith("img_") for f in names) and any("colden" in f for f in names) and any("_tau_" in f for f in names)
    hdr, img = read_fits(str(tmp_path / [f for f in names if f.startswith("img_") and f.endswith("_000.fits")][0]))
    assert img.shape == (8, 8) and hdr["CRVAL1"] == 12.5 and hdr["CTYPE1"] == "RA---TAN" and np.isfinite(img).all()
    hdr, col = read_fits(str(tmp_path / [f for f in names if "colden" in f][0]))
    assert col.min() > 0


def test_scattered_light_healpix_sky_cell_emission_and_healpix_observer(tmp_path):
    """ASOCS with every source of ASOCS.py:416-870 -- point source, Healpix sky, emission of the dust read from
    the emitted file -- and a Healpix image seen by an internal observer."""
    _run(tmp_path, n=8, bgpac=20000, pspac=33000, hpbg=2)                                # ASOC first: writes emit.data
    ini, cloud = write_model(str(tmp_path), n=8, bgpac=20000, pspac=66000, cellpac=8 ** 3 * 2, hpbg=2,
                             extra="perspective 3.3 4.1 2.7\noutnside 4\n")
    _run_asocs(tmp_path)
    raw = np.fromfile(str(tmp_path / "outcoming.socs"), np.int32, 2)
    assert list(raw) == [4, 8]
    data = np.fromfile(str(tmp_path / "outcoming.socs"), np.float32, offset=4 * (2 + 8)).reshape(8, 12 * 16)
    assert np.isfinite(data).all() and (data >= 0).all() and (data[4:] > 0).mean() > 0.9


def test_scattered_light_fits_cube(tmp_path):
    from soc_b200.fits import read_fits
    ini, cloud = write_model(str(tmp_path), n=8, bgpac=20000, pspac=66000)
    txt = open(ini).read().replace("directions   70.0 30.0\n", "") + "fits\nscattering scat\n"
    open(ini, "w").write(txt)
    _run_asocs(tmp_path)
    hdr, cube = read_fits(str(tmp_path / "scat.fits"))
    assert cube.shape == (8, 8, 8) and hdr["NAXIS3"] == 8 and np.isfinite(cube).all() and cube[4:].max() > 0


def test_levels_keyword_cuts_the_hierarchy(tmp_path):
    """`levels 2` on a 3-level cloud: <cloud>.MAX2 is written with the removed octets averaged into their parents
    (ASOC_aux.py:749-762, kernel_OT_tools.c:5-22) and the run uses it."""
    from soc_b200.formats import read_cloud
    cloud = _run(tmp_path, n=6, octree=True, bgpac=20000, maps=False, extra="levels 2\n")
    cut = read_cloud(str(tmp_path / "model.cloud.MAX2"))
    assert cloud.LEVELS == 3 and cut.LEVELS == 2 and cut.CELLS == cloud.LCELLS[0] + cloud.LCELLS[1]
    assert (cut.DENS[cut.OFF[1]:] > 0).all()                       # former parents of level 1 are leaves now
    lv1 = cloud.DENS[cloud.OFF[1]:cloud.OFF[2]]
    par = np.nonzero(lv1 <= 0)[0]
    first = (-lv1[par]).view(np.int32)
    lv2 = cloud.DENS[cloud.OFF[2]:]
    want = np.array([lv2[f:f + 8].astype(np.float32).sum(dtype=np.float32) / np.float32(8) for f in first])
    got = cut.DENS[cut.OFF[1]:][par]
    assert np.allclose(got, want, rtol=1e-6)
    T = read_otfile(str(tmp_path / "model.T"))
    assert T.shape == (cut.CELLS,)


def test_region_of_interest_save_then_load(tmp_path):
    """roisave: the photons entering ROI are written per frequency, surface element and direction; a second run loads
    that file as an external field (roiload + roipac) on a model of the ROI's size; roimap restricts the maps."""
    big = tmp_path / "big"
    cloud = _run(big, n=8, bgpac=40000, pspac=0, maps=True,
                 extra="roi 2 5 2 5 2 5\nroisave roi.save 1\nroinside 1\nroimap\n")
    hdr = np.fromfile(str(big / "roi.save"), np.int32, 5)
    assert list(hdr) == [4, 4, 4, 1, 8]
    data = np.fromfile(str(big / "roi.save"), np.float32, offset=20).reshape(8, 48 * 12)
    assert np.isfinite(data).all() and (data >= 0).all() and (data[2:].sum(axis=1) > 0).all()
    m = read_map_file(str(big / "map_dir_00.bin"))
    assert m[0].max() > 0 and (m[0][0, :] == 0).all() and (m[0][:, 0] == 0).all()      # lines of sight that miss ROI are empty
    small = tmp_path / "small"
    write_model(str(small), n=4, bgpac=0, pspac=0, maps=False)
    import shutil
    shutil.copy(str(big / "roi.save"), str(small / "roi.load"))
    ini = str(small / "model.ini")
    txt = open(ini).read().replace("bgpackets    0\n", "bgpackets    0\nroiload roi.load 1.0\nroinside 1\nroipac 200000\n")
    open(ini, "w").write(txt)
    cwd = os.getcwd()
    os.chdir(str(small))
    try:
        asoc.main(["ASOC.py", "model.ini"], device_factory=OracleDevice)
    finally:
        os.chdir(cwd)
    T = read_otfile(str(small / "model.T"))
    assert T.shape == (64,) and (T > 3.0).all() and (T < 80.0).all() and T.std() > 0.0
