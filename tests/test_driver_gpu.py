"""End-to-end: the ASOC driver on the CUDA library against the same driver on the oracle-backed device.
Same ini, same input files; output files compared (temperatures, emitted, absorbed, maps)."""
import os

import numpy as np
import pytest

from soc_b200 import asoc
from soc_b200.formats import read_otfile, read_map_file, read_cells_freq_file
from tests.model import write_model
from tests.oracle_device import OracleDevice

pytestmark = pytest.mark.gpu


def _run(path, factory, **kw):
    ini, cloud = write_model(str(path), **kw)
    cwd = os.getcwd()
    os.chdir(str(path))
    try:
        asoc.main(["ASOC.py", "model.ini"], device_factory=factory)
    finally:
        os.chdir(cwd)
    return cloud


@pytest.mark.parametrize("octree", [False, True])
def test_reference_streams_reproduce_the_oracle_run(tmp_path, octree):
    """REFSTREAMS: same MWC64X streams as the reference => temperatures and maps equal to rounding."""
    kw = dict(n=8, octree=octree, bgpac=30000, pspac=33000, extra="REFSTREAMS\nCLT\n")
    _run(tmp_path / "gpu", None, **kw)
    _run(tmp_path / "cpu", OracleDevice, **kw)
    Tg, Tc = read_otfile(str(tmp_path / "gpu" / "model.T")), read_otfile(str(tmp_path / "cpu" / "model.T"))
    # identical streams: nearly every cell agrees to rounding; a few small cells see a packet whose path flipped
    d = np.abs(Tg - Tc) / Tc.max()
    assert np.median(d) < 1e-5 and (d > 1e-3).mean() < 0.02 and d.max() < 0.03
    for idir in range(2):
        mg = read_map_file(str(tmp_path / "gpu" / ("map_dir_%02d.bin" % idir)))
        mc = read_map_file(str(tmp_path / "cpu" / ("map_dir_%02d.bin" % idir)))
        ok = mc[:3] > 0
        assert np.abs(mg[:3][ok] / mc[:3][ok] - 1.0).max() < 5e-3


def test_production_streams_agree_within_noise(tmp_path):
    kw = dict(n=12, bgpac=400000, pspac=330000, noabsorbed=False, absorbed=True, maps=False)
    cloud = _run(tmp_path / "gpu", None, **kw)
    _run(tmp_path / "cpu", OracleDevice, **kw)
    ag = read_cells_freq_file(str(tmp_path / "gpu" / "abs.data")).astype(np.float64)
    ac = read_cells_freq_file(str(tmp_path / "cpu" / "abs.data")).astype(np.float64)
    assert ag.shape == ac.shape == (cloud.CELLS, 8)
    # total absorbed photons per frequency: ~1e6 packets => well below 1 %
    tg, tc = ag.sum(axis=0), ac.sum(axis=0)
    ok = tc > 0
    assert np.abs(tg[ok] / tc[ok] - 1.0).max() < 0.01
    # per cell: relative noise of a cell ~ 1/sqrt(hits) ~ 2-3 %
    f = np.argmax(tc)
    rel = np.abs(ag[:, f] / ac[:, f] - 1.0)
    assert np.median(rel) < 0.05 and rel.max() < 0.5


def test_temperatures_with_cell_emission(tmp_path):
    kw = dict(n=10, bgpac=200000, cellpac=10 ** 3 * 20, iterations=2, maps=True)
    _run(tmp_path / "gpu", None, **kw)
    _run(tmp_path / "cpu", OracleDevice, **kw)
    Tg, Tc = read_otfile(str(tmp_path / "gpu" / "model.T")), read_otfile(str(tmp_path / "cpu" / "model.T"))
    assert np.abs(Tg / Tc - 1.0).mean() < 0.01 and np.abs(Tg / Tc - 1.0).max() < 0.06
    mg = read_map_file(str(tmp_path / "gpu" / "map_dir_00.bin"))
    mc = read_map_file(str(tmp_path / "cpu" / "map_dir_00.bin"))
    assert np.abs(mg[:2] / mc[:2] - 1.0).max() < 0.1


def test_scattered_light_driver(tmp_path):
    from soc_b200 import asocs
    from soc_b200.formats import read_outcoming
    res = []
    for name, fac in (("gpu", None), ("cpu", OracleDevice)):
        d = tmp_path / name
        write_model(str(d), n=10, bgpac=200000, pspac=655360)
        cwd = os.getcwd()
        os.chdir(str(d))
        try:
            asocs.main(["ASOCS.py", "model.ini"], device_factory=fac)
        finally:
            os.chdir(cwd)
        res.append(read_outcoming(str(d / "outcoming.socs"))[1].astype(np.float64))
    g, c = res
    for f in range(8):
        if c[f].sum() > 0:
            assert abs(g[f].sum() / c[f].sum() - 1.0) < 0.02, f
    f = int(np.argmax(c.sum(axis=(1, 2, 3))))
    ok = c[f] > 0.2 * c[f].max()
    assert np.median(np.abs(g[f][ok] / c[f][ok] - 1.0)) < 0.1


def test_device_resident_absorbed_array_matches_host_accumulation(tmp_path):
    """soc_absorbed_* (FABS kept, scaled and marked on the device) == the reference's host-side loop."""
    kw = dict(n=6, octree=True, bgpac=40000, pspac=33000, noabsorbed=False, absorbed=True, maps=False)
    cloud = _run(tmp_path / "dev", None, **kw)
    _run(tmp_path / "host", None, extra="HOSTABSORBED\n", **kw)
    a = read_cells_freq_file(str(tmp_path / "dev" / "abs.data")).astype(np.float64)
    b = read_cells_freq_file(str(tmp_path / "host" / "abs.data")).astype(np.float64)
    parents = cloud.DENS <= 0.0
    assert (a[parents] == float(np.float32(-1.0e20))).all() and (b[parents] == float(np.float32(-1.0e20))).all()
    leaves = ~parents
    assert np.abs(a[leaves] - b[leaves]).max() <= 2e-5 * b[leaves].max()
    assert np.abs(a[leaves].sum() / b[leaves].sum() - 1.0) < 1e-5


def test_two_dusts_msf_mirror_driver(tmp_path):
    """Two dust species with their own scattering functions (WITH_MSF) and reflecting borders through the driver:
    the absorbed file of the CUDA run agrees with the oracle-backed run within Monte Carlo noise."""
    kw = dict(n=10, bgpac=300000, pspac=330000, two_dusts=True, noabsorbed=False, absorbed=True, maps=False, extra="mirror xY\n")
    cloud = _run(tmp_path / "gpu", None, **kw)
    _run(tmp_path / "cpu", OracleDevice, **kw)
    ag = read_cells_freq_file(str(tmp_path / "gpu" / "abs.data")).astype(np.float64)
    ac = read_cells_freq_file(str(tmp_path / "cpu" / "abs.data")).astype(np.float64)
    tg, tc = ag.sum(axis=0), ac.sum(axis=0)
    ok = tc > 0
    assert np.abs(tg[ok] / tc[ok] - 1.0).max() < 0.015
    f = np.argmax(tc)
    rel = np.abs(ag[:, f] / ac[:, f] - 1.0)
    assert np.median(rel) < 0.06


def test_scattered_light_all_sources_healpix_observer(tmp_path):
    """ASOCS with the Healpix sky, the emission of the dust (emitted file) and a Healpix image for an outside
    observer: CUDA run vs oracle-backed run."""
    from soc_b200 import asocs
    res = []
    for name, fac in (("gpu", None), ("cpu", OracleDevice)):
        d = tmp_path / name
        _run(d, fac, n=8, bgpac=60000, pspac=33000, hpbg=1, maps=False)          # ASOC first: writes emit.data
        write_model(str(d), n=8, bgpac=200000, pspac=330000, cellpac=8 ** 3 * 40, hpbg=1,
                    extra="perspective 12.3 4.1 2.7\noutnside 2\n")
        cwd = os.getcwd()
        os.chdir(str(d))
        try:
            asocs.main(["ASOCS.py", "model.ini"], device_factory=fac)
        finally:
            os.chdir(cwd)
        res.append(np.fromfile(str(d / "outcoming.socs"), np.float32, offset=4 * (2 + 8)).reshape(8, 48).astype(np.float64))
    g, c = res
    assert np.isfinite(g).all()
    for f in range(8):
        if c[f].sum() > 0:
            assert abs(g[f].sum() / c[f].sum() - 1.0) < 0.05, (f, g[f].sum(), c[f].sum())


def test_per_level_maps_sum_to_the_ordinary_map(tmp_path):
    """`mapping nx ny dx 999`: the per-level images of the CUDA library add up to its ordinary map (same emission:
    REFSTREAMS makes the two runs identical up to the map step)."""
    kw = dict(n=6, octree=True, bgpac=20000)
    cloud = _run(tmp_path / "lev", None, extra="REFSTREAMS\nmapping 6 6 1.0 999\n", **kw)
    _run(tmp_path / "all", None, extra="REFSTREAMS\nmapping 6 6 1.0\n", **kw)
    hdr = np.fromfile(str(tmp_path / "lev" / "map_dir_01_H.bin"), np.int32, 4)
    assert list(hdr[:2]) == [6, 6] and hdr[3] == cloud.LEVELS
    lev = np.fromfile(str(tmp_path / "lev" / "map_dir_01_H.bin"), np.float32, offset=16).reshape(int(hdr[2]), cloud.LEVELS, 6, 6)
    full = read_map_file(str(tmp_path / "all" / "map_dir_01.bin"))
    assert (lev[:, 1:] > 0).any() and np.allclose(lev.sum(axis=1), full, rtol=3e-5, atol=0)


@pytest.mark.parametrize("extra", ["reference 1\n", "ali 1\n", "reference 1\nali 1\nemweight 1\n"])
def test_cell_emission_iterations_reference_field_and_ali(tmp_path, extra):
    """Iterated dust re-emission with the reference-field bookkeeping (`reference 1`: only the change of the emission is
    simulated), accelerated lambda iterations (`ali 1`: absorptions in the emitting cell go to XAB) and emission weighting,
    through the driver: CUDA library vs oracle device (ASOC.py:1594-1990)."""
    kw = dict(n=10, bgpac=200000, cellpac=10 ** 3 * 20, iterations=3, maps=False, extra=extra)
    _run(tmp_path / "gpu", None, **kw)
    _run(tmp_path / "cpu", OracleDevice, **kw)
    Tg, Tc = read_otfile(str(tmp_path / "gpu" / "model.T")), read_otfile(str(tmp_path / "cpu" / "model.T"))
    assert np.isfinite(Tg).all() and (Tg > 2.0).all()
    assert np.abs(Tg / Tc - 1.0).mean() < 0.01 and np.abs(Tg / Tc - 1.0).max() < 0.08
