"""Parity at the sizes of BASELINE.json's configurations (`-m gpu`): the CUDA library through its C ABI against the
REFERENCE's own kernels (oracle/_ref: kernel_ASOC.c / kernel_ASOC_map.c / kernel_ASOC_sca.c compiled in place, shipped
to the GPU box as prebuilt .so files by __graft_entry__.build()) and against the oracle-backed driver.

  C2  256^3, point source + isotropic background (the bench workload):
        * production streams vs the reference kernels on all host cores, >= 3e7 packets per source: chi^2/dof <= 1.1
          on 8^3-cell blocks and on radial shells, total absorbed energy within max(4 sigma, 1e-4);
        * reference streams (MWC64X per work item) on the GPU vs the reference kernels: same numbers to rounding;
        * production vs reference streams on the GPU alone at >= 1e9 packets per source (the CPU could not get the
          noise that low): total absorbed energy within 1e-4;
        * three 256^2 maps against the reference `Mapping` (the NX >= 200 ray set-up branch): 1e-5 relative.
  C1  the reference's runnable example (soc_example.zip, 64^3, 44 frequencies) through bin/ASOC.py vs the same driver
      on the oracle device: dust temperatures and the map.
  C3  octree, root 64^3 + 5 levels (~1e7 cells): absorptions and one scattered-light image vs the reference kernels.
"""
import os

import numpy as np
import pytest

from soc_b200 import synth
from soc_b200.hostmath import observer_directions
from tests.stats import chi2_per_dof

pytestmark = pytest.mark.gpu


def _ncpu():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _reference(cloud, **opts):
    from oracle import ref, build_ref
    cfg = dict(NX=cloud.NX, NY=cloud.NY, NZ=cloud.NZ, LEVELS=cloud.LEVELS, CELLS=cloud.CELLS, BINS=2500, GL=0.01)
    cfg.update({k.upper(): v for k, v in opts.items()})
    if build_ref.build(cfg) is None:
        pytest.skip("reference library %s neither prebuilt nor buildable here" % build_ref.tag_of(cfg))
    R = ref.Reference(cloud, **opts)
    R.set_threads(_ncpu())
    R.set_chunk(4)
    return R


def _backend(cloud, rng_mode, **opts):
    from soc_b200 import backend
    return backend.Backend(cloud, rng_mode=rng_mode, **opts)


# ---- C2: 256^3 ---------------------------------------------------------------------------------------------------------
N2 = 256


def _c2():
    import bench
    return bench.make_workload(N2)


def _blocks(a, n=N2, b=8):
    """Sums over b^3-cell blocks and over radial shells of one cell width (float64)."""
    a = a.astype(np.float64).reshape(n // b, b, n // b, b, n // b, b)
    return a.sum(axis=(1, 3, 5)).ravel()


_SHELL = {}


def _shells(a, n=N2):
    if n not in _SHELL:
        c = np.arange(n, dtype=np.float32) + 0.5 - 0.5 * n
        z, y, x = np.meshgrid(c, c, c, indexing="ij")
        _SHELL[n] = np.minimum(np.sqrt(x * x + y * y + z * z).astype(np.int32), n // 2).ravel()
    return np.bincount(_SHELL[n], weights=a.astype(np.float64), minlength=n // 2 + 1)


def _launch(X, w, source, items, batch, seed):
    X.zero(0)
    X.zero(1)
    common = dict(abs_=w["kabs"], sca=w["ksca"], dsc=w["dsc"], csc=w["csc"])
    if source == 1:
        X.sim_pb(items, 1, items * batch, batch, seed, w["bg"], w["tw"], **common)
    else:
        X.sim_pb(items, 0, items * batch, batch, seed, 0.0, w["tw"], pspos=w["pspos"], ps=w["ps"], **common)


@pytest.mark.parametrize("source", [1, 0], ids=["background", "pointsource"])
def test_c2_absorptions_match_the_reference_kernels(source):
    """>= 3e7 packets per source on both sides, K = 6 independent repetitions each (the repetitions give the variances)."""
    from soc_b200 import backend
    w = _c2()
    cloud = w["cloud"]
    R = _reference(cloud, no_ps=1, noabsorbed=0)
    B = _backend(cloud, backend.RNG_PACKET, no_ps=1, noabsorbed=0)
    K = 6
    if source == 1:
        items, batch = cloud.AREA, 13              # every surface element once per work-item round (id % AREA): 5.1e6 packets
    else:
        items, batch = w["ps_glob"], 153           # 5.0e6 packets
    gpu_mult = 4                                   # the GPU side is free: four times the packets per repetition
    assert K * items * batch >= 3.0e7
    ra, rb, sa, sb = [], [], [], []
    # The reference adds every absorption to its float32 cell with an atomic: the cell of the point source takes one add
    # per packet, and a float32 sum of 5e6 small terms is off by ~1e-4 -- more than the Monte Carlo noise of that cell
    # (0.25 % at 3e7 adds).  Its repetition is therefore run as `sub` launches of 1/sub of the packets,
    # summed here in float64, so that the comparison sees the kernels and not the accumulator.
    sub = 1 if source == 1 else 9
    assert batch % sub == 0
    for k in range(K):
        seed = 0.05 + 0.9 * (k + 0.5) / K
        acc = np.zeros(cloud.CELLS, np.float64)
        for j in range(sub):
            _launch(R, w, source, items, batch // sub, seed + 0.0007 * j)
            acc += R.int_
            if k == 0 and j == 0:                  # one frequency, TW = 1, ADHOC = 1: TABS == INT
                assert np.allclose(R.tabs, R.int_, rtol=1e-6, atol=0)
        ra.append(_blocks(acc)), sa.append(_shells(acc))
        _launch(B, w, source, items, batch * gpu_mult, seed)
        g = B.int_
        if k == 0:
            assert np.allclose(B.tabs, g, rtol=2e-6, atol=0)
        rb.append(_blocks(g) / gpu_mult), sb.append(_shells(g) / gpu_mult)
    assert B.counters.reserved[0] == 0
    B.close()
    chi2, dof, tot, tot_sigma = chi2_per_dof(rb, ra, min_rel=1e-6, var_ratio=1.0 / gpu_mult)
    assert dof > (20000 if source == 1 else 5000)       # the point source sits in the dense core: most of its energy stays there
    assert chi2 <= 1.1, "blocks: chi2/dof = %.3f over %d blocks" % (chi2, dof)
    chi2s, dofs, _, _ = chi2_per_dof(sb, sa, var_ratio=1.0 / gpu_mult)
    assert dofs > 100
    # Few shells (129): the mean of t^2 itself scatters by sqrt(2/dof) ~ 0.12
    assert chi2s <= 1.1 + 3.0 * np.sqrt(2.0 / dofs), "shells: chi2/dof = %.3f over %d shells" % (chi2s, dofs)
    assert tot <= max(4.0 * tot_sigma, 1e-4), "total energy differs by %.2e (sigma %.2e)" % (tot, tot_sigma)
    print("C2 %s: chi2/dof blocks %.3f (%d) shells %.3f (%d), total %.2e (sigma %.2e)" % (
        "BG" if source else "PS", chi2, dof, chi2s, dofs, tot, tot_sigma))


def test_c2_reference_streams_reproduce_the_reference_kernels():
    """REFSTREAMS layout at 256^3: thread = reference work item with its MWC64X stream => the same packets as the
    reference kernels; the absorbed energy agrees to rounding except for the few paths that flip at a cell face
    because expf / sincosf / acosf differ in the last bit between CUDA and glibc."""
    from soc_b200 import backend
    w = _c2()
    cloud = w["cloud"]
    R = _reference(cloud, no_ps=1, noabsorbed=0)
    B = _backend(cloud, backend.RNG_REFERENCE, no_ps=1, noabsorbed=0)
    for source, items, batch in ((1, 120000, 2), (0, 4096, 40)):
        _launch(R, w, source, items, batch, 0.4)
        _launch(B, w, source, items, batch, 0.4)
        a, b = B.int_.astype(np.float64), R.int_.astype(np.float64)
        assert abs(a.sum() - b.sum()) <= 1e-4 * b.sum(), (source, a.sum(), b.sum())
        ab, bb = _blocks(a), _blocks(b)
        bad = np.abs(ab - bb) > 1e-5 * bb.max() + 1e-3 * bb
        assert bad.mean() < 0.02, "source %d: %d of %d blocks differ" % (source, bad.sum(), bad.size)
        cells = np.abs(a - b) > 1e-5 * b.max() + 1e-4 * b
        assert cells.mean() < 0.01
    B.close()


@pytest.mark.parametrize("source", [1, 0], ids=["background", "pointsource"])
def test_c2_total_energy_production_vs_reference_streams(source):
    """The 1e-4 total-energy gate of BASELINE.json needs ~1e9 packets per arm, out of reach of the host cores.  The GPU
    runs both the production kernels (Philox per packet, DDA, series/ex2 absorption) and the reference layout (MWC64X,
    the reference's own GetStep arithmetic and draw order -- pinned to the reference kernels by the test above): K launches
    each, sigma from the scatter of the launch totals."""
    from soc_b200 import backend
    w = _c2()
    cloud = w["cloud"]
    K = 8
    if source == 1:
        items, batch = 8 * cloud.AREA, 48          # 1.5e8 packets per launch, 1.2e9 in all
    else:
        items, batch = w["ps_glob"], 4600
    tot = {}
    for mode in (backend.RNG_PACKET, backend.RNG_REFERENCE):
        B = _backend(cloud, mode, no_ps=1)
        t = []
        for k in range(K):
            _launch(B, w, source, items, batch, 0.03 + 0.9 * (k + 0.5) / K)
            t.append(B.tabs.astype(np.float64).sum())
        assert B.counters.reserved[0] == 0
        B.close()
        tot[mode] = np.array(t)
    a, b = tot[backend.RNG_PACKET], tot[backend.RNG_REFERENCE]
    sigma = np.sqrt(a.var(ddof=1) / K + b.var(ddof=1) / K) / b.mean()
    diff = abs(a.mean() - b.mean()) / b.mean()
    print("C2 total energy %s: production %.8e reference streams %.8e diff %.2e sigma %.2e (%.2e packets per arm)" % (
        "BG" if source else "PS", a.mean(), b.mean(), diff, sigma, K * items * batch))
    assert sigma < 5e-5
    assert diff <= 1e-4


def test_c2_maps_match_the_reference_kernel():
    """Three 256^2 maps of the 256^3 model (directions 0 0 / 90 0 / 60 30) against the reference's Mapping kernel:
    1e-5 relative per pixel.  NX >= 200 takes the second ray set-up branch (kernel_ASOC_map.c:571-626) and the
    double-precision Index (NX > 100)."""
    from soc_b200 import backend
    w = _c2()
    cloud = w["cloud"]
    R = _reference(cloud, no_ps=1, noabsorbed=0)
    B = _backend(cloud, backend.RNG_PACKET, no_ps=1, noabsorbed=0)
    rng = np.random.default_rng(9)
    emit = (cloud.DENS * (0.5 + rng.random(cloud.CELLS))).astype(np.float32)
    _, od, ra, de = observer_directions([0.0, 90.0, 60.0], [0.0, 0.0, 30.0])
    centre = np.array([0.5 * N2] * 3, np.float32)
    k = 3.0 / N2
    for i in range(3):
        mr, tr = R.mapping(1.0, N2, N2, emit, od[i], ra[i], de[i], 0.6 * k, 0.4 * k, centre)
        mg, tg = B.mapping(1.0, N2, N2, emit, od[i], ra[i], de[i], 0.6 * k, 0.4 * k, centre)
        for name, a, b in (("map", mg, mr), ("tau", tg, tr)):
            a, b = a.astype(np.float64), b.astype(np.float64)
            nz = b != 0.0
            assert nz.mean() > 0.5
            assert (a[~nz] == 0.0).all()
            rel = np.abs(a[nz] - b[nz]) / np.abs(b[nz])
            assert rel.max() <= 1e-5, "direction %d %s: max rel %.3e" % (i, name, rel.max())
    B.close()


# ---- C1: the reference's own example -----------------------------------------------------------------------------------
def _run_example(path, factory, extra):
    from soc_b200 import asoc
    from tests.soc_example import write_example
    write_example(str(path), edits={"seed": "seed 0.4", "device": None, "verbose": "verbose 0", **extra})
    cwd = os.getcwd()
    os.chdir(str(path))
    try:
        asoc.main(["ASOC.py", "my.ini"], device_factory=factory)
    finally:
        os.chdir(cwd)


def test_c1_example_through_the_driver(tmp_path):
    """soc_example.zip's my.ini (64^3, 44 frequencies, 1e6 background packets each, CLT / CLE, one 64^2 map) through the
    ASOC driver: CUDA library vs oracle device.  With the reference streams the two runs see the same packets; the
    production streams agree within the Monte Carlo noise of the run."""
    from soc_b200.formats import read_otfile, read_map_file
    from tests.oracle_device import OracleDevice
    _run_example(tmp_path / "cpu", OracleDevice, {})
    _run_example(tmp_path / "ref", None, {"refstreams": "REFSTREAMS"})
    _run_example(tmp_path / "gpu", None, {})
    Tc = read_otfile(str(tmp_path / "cpu" / "tmp.T")).astype(np.float64)
    Tr = read_otfile(str(tmp_path / "ref" / "tmp.T")).astype(np.float64)
    Tg = read_otfile(str(tmp_path / "gpu" / "tmp.T")).astype(np.float64)
    assert Tc.shape == (64 ** 3,) and 5.0 < Tc.min() and Tc.max() < 40.0
    d = np.abs(Tr - Tc) / Tc
    assert np.median(d) < 2e-6 and (d > 1e-3).mean() < 0.01 and d.max() < 0.02
    # production streams: a cell sees ~ 44 x 1e6 x 64 / 64^3 ~ 1e4 packets => ~1 % in the absorbed energy, T ~ E^(1/5.5)
    d = Tg / Tc - 1.0
    assert abs(d.mean()) < 2e-4 and d.std() < 8e-3 and np.abs(d).max() < 0.04
    mc = read_map_file(str(tmp_path / "cpu" / "map_dir_00.bin")).astype(np.float64)
    mr = read_map_file(str(tmp_path / "ref" / "map_dir_00.bin")).astype(np.float64)
    mg = read_map_file(str(tmp_path / "gpu" / "map_dir_00.bin")).astype(np.float64)
    assert mc.shape == (44, 64, 64)
    ok = mc > 1e-6 * mc.max(axis=(1, 2), keepdims=True)
    assert np.abs(mr[ok] / mc[ok] - 1.0).max() < 5e-3
    far = slice(0, 11)          # lambda >= 250 um: towards the Wien side (h nu / kT ~ 10 at 100 um) 0.5 % of temperature noise is 5 % of emission
    r = np.abs(mg[far][ok[far]] / mc[far][ok[far]] - 1.0)
    assert np.median(r) < 3e-3 and r.max() < 0.03, (np.median(r), r.max())


# ---- C3: ~1e7-cell octree -----------------------------------------------------------------------------------------------
def c3_cloud():
    return synth.octree_cloud(64, 6, refine_fraction=0.22, seed=12345)


def _root_of_cells(cloud):
    """Root cell above every cell of the hierarchy."""
    from soc_b200.formats import float_to_links
    n0 = cloud.NX * cloud.NY * cloud.NZ
    root = [np.arange(n0, dtype=np.int64)]
    for l in range(1, cloud.LEVELS):
        lo, hi = cloud.OFF[l - 1], cloud.OFF[l - 1] + cloud.LCELLS[l - 1]
        up = cloud.DENS[lo:hi]
        parents = np.nonzero(up <= 0.0)[0]
        first = float_to_links(up[parents])
        r = np.empty(cloud.LCELLS[l], np.int64)
        for s in range(8):
            r[first + s] = root[l - 1][parents]
        root.append(r)
    return np.concatenate(root)


def test_c3_octree_absorptions_and_image_match_the_reference_kernels():
    from soc_b200 import backend
    cloud = c3_cloud()
    assert cloud.LEVELS == 6 and 9.0e6 < cloud.CELLS < 1.1e7
    R = _reference(cloud, no_ps=1, noabsorbed=0)
    B = _backend(cloud, backend.RNG_PACKET, no_ps=1, noabsorbed=0)
    dsc, csc = synth.hg_tables(0.6, 2500)
    n = cloud.NX
    root_mean = float(np.mean(np.where(cloud.DENS[:n ** 3] > 0, cloud.DENS[:n ** 3], 1.0)))
    k = 2.0 / (n * root_mean)
    root = _root_of_cells(cloud)
    rz, ry, rx = root // (n * n), (root // n) % n, root % n
    block = ((rz // 4) * 16 + ry // 4) * 16 + rx // 4             # 4^3 root cells and everything below them
    level = np.repeat(np.arange(cloud.LEVELS), cloud.LCELLS)
    K, items, batch, mult = 6, cloud.AREA, 48, 4
    a, b, la, lb = [], [], [], []
    for i in range(K):
        seed = 0.07 + 0.9 * (i + 0.5) / K
        for X, m, out, lout in ((R, 1, a, la), (B, mult, b, lb)):
            X.zero(0), X.zero(1)
            X.sim_pb(items, 1, items * batch * m, batch * m, seed, 1.0, 1.0, abs_=k, sca=k, dsc=dsc, csc=csc)
            v = X.int_.astype(np.float64) / m
            out.append(np.bincount(block, weights=v, minlength=4096))
            lout.append(np.bincount(level, weights=v, minlength=cloud.LEVELS))
    assert B.counters.reserved[0] == 0
    chi2, dof, tot, tot_sigma = chi2_per_dof(b, a, min_rel=1e-4, var_ratio=1.0 / mult)
    assert dof > 3000
    assert chi2 <= 1.1, "octree blocks: chi2/dof = %.3f over %d" % (chi2, dof)
    assert tot <= max(4.0 * tot_sigma, 1e-4), "octree total energy differs by %.2e (sigma %.2e)" % (tot, tot_sigma)
    la, lb = np.array(la), np.array(lb)
    sig = np.sqrt(la.var(0, ddof=1) / K + lb.var(0, ddof=1) / K)
    assert (np.abs(la.mean(0) - lb.mean(0)) <= 4.5 * sig + 1e-4 * la.mean(0)).all(), (la.mean(0), lb.mean(0), sig)
    print("C3 absorptions: chi2/dof %.3f (%d), total %.2e (sigma %.2e)" % (chi2, dof, tot, tot_sigma))
    B.close()
    # Scattered light: point source, one observer, 32^2 pixels of two root cells.  The reference converts the scattering
    # position with the level of the *next* cell (kernel_ASOC_sca.c:958); on this 6-level cloud that moves the total image
    # flux by ~0.7 % (oracle, literal vs exact variant).  So: (i) the library's reference-geometry kernels (soc_set_geometry 1,
    # the quirk included) against the reference kernels, (ii) the production kernel (exact position) against the oracle's
    # exact-level variant, which differs from the pinned literal one in that single line.
    from oracle import orc
    _, od, ra, de = observer_directions([60.0], [30.0])
    centre = np.array([0.5 * n] * 3, np.float32)
    pspos = np.array([0.5 * n + 0.3] * 3, np.float32)
    ps = np.ones(1, np.float32)
    K, items, batch, mult = 6, 8192, 36, 8
    args = (1, 32, 32, 2.0, centre, od, ra, de)
    kw = dict(abs_=k, sca=k, dsc=dsc, csc=csc, pspos=pspos, ps=ps)
    orc.set_threads(_ncpu())
    for name, make_cpu, literal in (("reference kernels vs reference-geometry kernel", lambda: _reference(cloud, no_ps=1, ffs=1), True),
                                    ("oracle (exact level) vs production kernel", lambda: orc.Oracle(cloud, no_ps=1, ffs=1, sca_exact_level=1), False)):
        X = make_cpu()
        B = _backend(cloud, backend.RNG_PACKET, no_ps=1, ffs=1)
        if literal:
            B.dev.set_geometry(1)
        a, b = [], []
        for i in range(K):
            seed = 0.11 + 0.9 * (i + 0.5) / K
            a.append(X.sca_ps(items, items * batch, batch, seed, *args, **kw).astype(np.float64).ravel())
            b.append(B.sca_ps(items, items * batch * mult, batch * mult, seed, *args, **kw).astype(np.float64).ravel() / mult)
        assert B.counters.reserved[0] == 0
        B.close()
        chi2, dof, tot, tot_sigma = chi2_per_dof(b, a, min_rel=0.01, var_ratio=1.0 / mult)
        print("C3 scattered light, %s: chi2/dof %.3f (%d), total %.2e (sigma %.2e)" % (name, chi2, dof, tot, tot_sigma))
        assert dof > 200
        assert chi2 <= 1.1 + 3.0 * np.sqrt(2.0 / dof), "%s: chi2/dof = %.3f over %d pixels" % (name, chi2, dof)
        assert tot <= max(4.0 * tot_sigma, 1e-4), "%s: total flux differs by %.2e (sigma %.2e)" % (name, tot, tot_sigma)
