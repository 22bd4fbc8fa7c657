"""GPU parity tests proper: the CUDA library (through its C ABI) against the oracle and the golden vectors.

  * reference stream layout (MWC64X, thread = reference work item): same random numbers and the same order
    of operations as the oracle => the Monte Carlo outputs must agree to float rounding, except for the few
    packets whose path flips at a cell face because expf/sincosf/acosf differ in the last bit between CUDA
    and glibc;
  * maps: 1e-5 relative per pixel (the contract of BASELINE.json);
  * production layout (Philox per packet): statistical parity, chi^2/dof <= 1.1 per cell and total energy
    within Monte Carlo noise.
"""
import os

import numpy as np
import pytest

from tests.cases import CASES, run_bg, run_ps, run_abu, run_hp, run_cl, run_sca, run_bg_msf, run_roi_load, with_roi_save, _reg, _oct
from tests.stats import chi2_per_dof
from soc_b200 import synth

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MC_CASES = sorted(k for k in CASES if not k.startswith(("map_", "hpmap_")))
MAP_CASES = sorted(k for k in CASES if k.startswith(("map_", "hpmap_")))
# The shipped scattered-light SimRAM_HP / SimRAM_CL run their `#ifdef HG_TEST` branch (analytic g=0.65 phase function,
# 1-exp(-tau)); the library implements the intended #else branch.  The oracle restates both (flag hg_test, the shipped
# one pinned against the reference kernels); the goldens of these cases hold the shipped variant.
SHIPPED_HG_TEST = tuple(k for k in CASES if k.startswith(("sca_hp", "sca_cl")))


def _backend(cloud, rng_mode, **opts):
    from soc_b200 import backend
    return backend.Backend(cloud, rng_mode=rng_mode, **opts)


def _compare_mc(name, key, a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    assert a.shape == b.shape
    assert np.isfinite(a).all(), "%s/%s: non-finite values" % (name, key)
    scale = np.abs(b).max()
    if scale == 0.0:
        assert np.abs(a).max() == 0.0
        return
    # cells touched by a packet whose path flipped differ visibly; everything else agrees to rounding
    bad = np.abs(a - b) > 1e-5 * scale + 1e-4 * np.abs(b)
    assert bad.mean() < 0.02, "%s/%s: %d of %d cells differ (max %.3e of %.3e)" % (
        name, key, bad.sum(), bad.size, np.abs(a - b).max(), scale)
    assert abs(a.sum() - b.sum()) <= 2e-4 * abs(b.sum()), "%s/%s: totals %.8e vs %.8e" % (name, key, a.sum(), b.sum())


@pytest.mark.parametrize("name", MC_CASES)
def test_reference_streams_match_oracle(name):
    from oracle import orc
    from soc_b200 import backend
    make, opts, run = CASES[name]
    cloud = make()
    O = orc.Oracle(cloud, **opts)
    out_o = run(O)
    B = _backend(cloud, backend.RNG_REFERENCE, **opts)
    out_g = run(B)
    for key in out_o:
        _compare_mc(name, key, out_g[key], out_o[key])
    co, cg = O.counters, B.counters
    assert cg.packets == co.packets
    assert abs(int(cg.steps) - int(co.steps)) <= 2e-3 * co.steps
    assert cg.reserved[0] == 0, "packets killed by the step guard"
    B.close()


@pytest.mark.parametrize("name", MC_CASES)
def test_reference_streams_match_golden(name):
    """Same comparison against the vectors produced by the reference's own kernels (tests/golden)."""
    from soc_b200 import backend
    make, opts, run = CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    # the goldens of the scattered-light HP / CL cases hold the reference's HG_TEST branch: ref_quirks bit 0 reproduces it
    # (the intended branch is checked against the oracle in test_reference_streams_match_oracle)
    B = _backend(make(), backend.RNG_REFERENCE, ref_quirks=1 if name in SHIPPED_HG_TEST else 0, **opts)
    out = run(B)
    for key in gold.files:
        _compare_mc(name, key, out[key], gold[key])
    B.close()


@pytest.mark.parametrize("name", MAP_CASES)
def test_maps_match_oracle_and_golden(name):
    from oracle import orc
    from soc_b200 import backend
    make, opts, run = CASES[name]
    cloud = make()
    out_o = run(orc.Oracle(cloud, **opts))
    B = _backend(cloud, backend.RNG_REFERENCE, **opts)
    out_g = run(B)
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    refs = [(out_o, out_g), ({k: gold[k] for k in gold.files}, out_g)]
    if name.startswith("map_lev") and cloud.LEVELS > 1:
        # the goldens hold kernel_ASOC_map_H.c as shipped, whose Index() loses rays that climb into root-grid leaves;
        # the library steps like kernel_ASOC_map.c unless ref_quirks bit 1 asks for the file's own Index()
        Q = _backend(cloud, backend.RNG_REFERENCE, ref_quirks=2, **opts)
        refs[1] = (refs[1][0], run(Q))
        Q.close()
    for ref, got in refs:
        for key in ref:
            a, b = got[key].astype(np.float64), ref[key].astype(np.float64)
            nz = b != 0.0
            assert (a[~nz] == 0.0).all(), key
            rel = np.abs(a[nz] - b[nz]) / np.abs(b[nz])
            # tolerance of the contract: 1e-5 relative per pixel
            assert rel.max() <= 1e-5, "%s/%s: max rel %.3e at %d" % (name, key, rel.max(), rel.argmax())
    B.close()


def test_half_precision_opacities():
    """OPT_IS_HALF: the uploaded half values are widened on the device; the run equals a float run on the same
    half-rounded opacities bit for bit (same streams, same arithmetic)."""
    from soc_b200 import backend
    cloud = synth.regular_cloud(12)
    rng = np.random.default_rng(5)
    k = 0.1
    opt = np.empty((cloud.CELLS, 2), np.float32)
    opt[:, 0] = k * (1.0 + rng.random(cloud.CELLS))
    opt[:, 1] = k * (1.5 + 2 * rng.random(cloud.CELLS))
    dsc, csc = synth.hg_tables(0.6)
    glob = 8 * cloud.AREA
    res = []
    for half in (1, 0):
        B = _backend(cloud, backend.RNG_PACKET, with_abu=1, opt_is_half=half)
        o16 = opt.reshape(-1).astype(np.float16)
        B.dev.upload(backend.BUF_OPT, o16 if half else o16.astype(np.float32), np.float16 if half else np.float32)
        B.zero(0)
        B.sim_pb(glob, 1, glob * 2, 2, 0.3, 1.0, 1.0, dsc=dsc, csc=csc)
        res.append(B.tabs.astype(np.float64))
        B.close()
    assert res[0].sum() > 0 and np.abs(res[0] - res[1]).max() <= 1e-5 * res[1].max()


def test_emission2_matches_oracle():
    """Emission2 (cells x frequencies in one launch) against the oracle restatement (pinned to the reference kernel)."""
    from oracle import orc
    from soc_b200 import backend
    cloud = synth.octree_cloud(6, 3, refine_fraction=0.2, seed=3)
    rng = np.random.default_rng(2)
    t = (5.0 + 40.0 * rng.random(cloud.CELLS)).astype(np.float32)
    freq = np.logspace(11.5, 14.5, 9).astype(np.float32)
    fabs_ = (1e-6 * (freq / 1e12) ** 1.8).astype(np.float32)
    a = orc.Oracle(cloud).emission2(17, cloud.CELLS - 5, freq, fabs_, t)
    B = _backend(cloud, backend.RNG_PACKET)
    b = B.emission2(17, cloud.CELLS - 5, freq, fabs_, t)
    B.close()
    assert a.shape == b.shape == (cloud.CELLS - 22, 9)
    ok = a > 0
    assert (b[~ok] == 0).all() and np.abs(b[ok] / a[ok] - 1.0).max() < 2e-5       # expf differs in the last bits


@pytest.mark.parametrize("mode", ["sum", "from_second", "single_abu", "half"])
def test_opt_built_on_the_device_equals_the_host_loop(mode):
    """soc_build_opt against the host loop of ASOC.py:1146-1161 (soc_b200.asoc._opt_array): same float32 operations in
    the same order => bit-identical; OPT_IS_HALF rounds to half precision like numpy's astype(float16)."""
    from types import SimpleNamespace
    from soc_b200 import backend
    from soc_b200.asoc import _opt_array
    cloud = synth.octree_cloud(6, 3, refine_fraction=0.2, seed=3)
    rng = np.random.default_rng(8)
    ndust = 2 if mode == "single_abu" else 3
    abu = (0.05 + rng.random((cloud.CELLS, ndust))).astype(np.float32)
    afabs = [(1e-3 * (1 + d) * (0.5 + rng.random(4))).astype(np.float32) for d in range(ndust)]
    afsca = [(2e-3 * (1 + d) * (0.5 + rng.random(4))).astype(np.float32) for d in range(ndust)]
    first = 1 if mode == "from_second" else 0
    user = SimpleNamespace(SINGLE_ABU=1 if mode == "single_abu" else 0)
    B = _backend(cloud, backend.RNG_PACKET, with_abu=1, opt_is_half=1 if mode == "half" else 0)
    B.dev.upload(backend.BUF_ABU, abu.reshape(-1))
    for ifreq in range(4):
        want = _opt_array(user, abu, afabs, afsca, ifreq, first)
        if mode == "half":
            want = want.astype(np.float16).astype(np.float32)
        B.dev.build_opt([a[ifreq] for a in afabs], [s[ifreq] for s in afsca], first, mode == "single_abu")
        got = B.dev.download(backend.BUF_OPT, 2 * cloud.CELLS).reshape(-1, 2)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (mode, ifreq, np.abs(got - want).max())
    B.close()


def test_split_absorbed_equals_the_oracle():
    """soc_split_absorbed (hand-off of the absorptions to the dust solver of one species, kernel_A2E_MABU_aux.c:3-24)
    against the oracle restatement, which is pinned bit for bit to the reference kernel: identical bits."""
    from oracle import orc
    from soc_b200 import backend
    cloud = synth.octree_cloud(6, 3, refine_fraction=0.2, seed=3)
    rng = np.random.default_rng(1)
    nfreq, ndust = 9, 3
    abu = (0.1 + rng.random((cloud.CELLS, ndust))).astype(np.float32)
    rabs = 1e-21 * (0.5 + rng.random((nfreq, ndust)))
    a = (1e-3 * rng.random((cloud.CELLS, nfreq))).astype(np.float32)
    a[cloud.DENS <= 0] = -1e20
    B = _backend(cloud, backend.RNG_PACKET)
    B.dev.upload(backend.BUF_FABS, a.reshape(-1))
    B.dev.upload(backend.BUF_ABU, abu.reshape(-1))
    for idust in range(ndust):
        got = B.dev.split_absorbed(idust, rabs, cloud.CELLS)
        want = orc.split_absorbed(idust, rabs, abu, a)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (idust, np.abs(got - want).max())
    B.close()


def _repeat(X, runner_factory, K, key="tabs"):
    """K repetitions with different seeds; `key` may be a tuple of output names (returns a dict of arrays then)."""
    keys = key if isinstance(key, tuple) else (key,)
    outs = {k: [] for k in keys}
    for k in range(K):
        run = runner_factory(0.05 + 0.9 * (k + 0.5) / K)
        res = run(X)
        for kk in keys:
            outs[kk].append(res[kk])
    outs = {k: np.array(v) for k, v in outs.items()}
    return outs if isinstance(key, tuple) else outs[key]


def _abu_opt(cells):
    rng = np.random.default_rng(5)
    k = 1.0 / 12
    opt = np.empty((cells, 2), np.float32)
    opt[:, 0] = 2.0 * k * (1.0 + rng.random(cells))
    opt[:, 1] = 2.0 * k * (0.5 + rng.random(cells))
    return opt.reshape(-1)


STAT_CASES = {
    # name: (cloud, options, runner factory(seed), output key)
    "bg_reg16": (_reg(16), {}, lambda s: run_bg(batch=4, seed=s), "tabs"),
    "bg_box_20_12_8": (lambda: synth.box_cloud(20, 12, 8), {}, lambda s: run_bg(batch=6, seed=s), "tabs"),
    "bg_box_oct_10_6_4": (lambda: synth.box_cloud(10, 6, 4, levels=4, refine_fraction=0.25), {}, lambda s: run_bg(batch=24, seed=s), "tabs"),
    "ps_box_oct_10_6_4": (lambda: synth.box_cloud(10, 6, 4, levels=4, refine_fraction=0.25), dict(no_ps=1),
                          lambda s: run_ps([(4.3, 3.2, 1.7)], batch=48, seed=s), "tabs"),
    "bg_reg16_refgeo": (_reg(16), {}, lambda s: run_bg(batch=4, seed=s), "tabs"),
    "bg_reg16_iso": (_reg(16), {}, lambda s: run_bg(batch=4, seed=s, g=False, tau_s=6.0), "tabs"),
    "bg_oct8_3": (_oct(8, 3), {}, lambda s: run_bg(batch=8, seed=s), "tabs"),
    "bg_oct6_4_int": (_oct(6, 4, 0.25, 8), dict(noabsorbed=0), lambda s: run_bg(batch=8, seed=s), "int"),
    "ps_reg16": (_reg(16), dict(no_ps=2), lambda s: run_ps([(8.3, 8.3, 8.3), (3.7, 11.2, 5.1)], batch=24, seed=s), "tabs"),
    "ps_reg12_ext1": (_reg(12), dict(no_ps=1, ps_method=1), lambda s: run_ps([(-9.0, 5.0, 7.0)], batch=48, seed=s), "tabs"),
    "bg_reg12_int": (_reg(12), dict(noabsorbed=0), lambda s: run_bg(batch=8, seed=s), "int"),
    "bg_reg12_int2": (_reg(12), dict(save_intensity=2), lambda s: run_bg(batch=8, seed=s), "tabs"),
    "bg_reg12_abu": (_reg(12), dict(with_abu=1), lambda s: run_abu(batch=8, seed=s), "tabs"),
    # per-cell opacities on the look-ahead kernel: odd dimensions (x-fastest layout), point source with the tile, per-frequency array
    "bg_box_9_7_5_abu": (lambda: synth.box_cloud(9, 7, 5), dict(with_abu=1), lambda s: run_abu(batch=24, seed=s), "tabs"),
    "ps_reg12_abu": (_reg(12), dict(with_abu=1, no_ps=1), lambda s: run_ps([(6.3, 6.2, 5.9)], batch=48, seed=s, opt=_abu_opt(12 ** 3)), "tabs"),
    "bg_reg12_abu_int": (_reg(12), dict(with_abu=1, noabsorbed=0), lambda s: run_abu(batch=8, seed=s), "int"),
    "hp_reg12": (_reg(12), {}, lambda s: run_hp(False, batch=24, seed=s), "tabs"),
    "hp_reg12_w": (_reg(12), dict(hpbg_weighted=1), lambda s: run_hp(True, batch=24, seed=s), "tabs"),
    "cl_reg10": (_reg(10), {}, lambda s: run_cl(False, batch=6, seed=s), "tabs"),
    "cl_oct6_ew_ali": (_oct(6, 3), dict(use_emweight=1, with_ali=1), lambda s: run_cl(True, batch=2, seed=s), ("tabs", "xab")),
    # ALI with per-frequency absorptions: XAB takes the emitting cell's share of TABS, INT takes everything (kernel_ASOC.c:1486-1499)
    "cl_reg16_ali_int": (_reg(16), dict(with_ali=1, noabsorbed=0), lambda s: run_cl(False, batch=4, seed=s), ("tabs", "xab", "int")),
    "cl_oct6_ali_int": (_oct(6, 3), dict(with_ali=1, noabsorbed=0), lambda s: run_cl(False, batch=3, seed=s), ("tabs", "xab", "int")),
    # optically thick (tau ~ 20 per cell: packets die in the surface layers, many scatterings per cell) and very thin
    # (tau ~ 1e-4 per cell: the series branch of the absorbed fraction) media
    "bg_reg16_thick": (_reg(16), dict(noabsorbed=0), lambda s: run_bg(batch=16, seed=s, tau_a=80.0, tau_s=240.0), ("tabs", "int")),
    "ps_reg16_thick": (_reg(16), dict(no_ps=1), lambda s: run_ps([(8.3, 8.2, 7.9)], batch=48, seed=s, tau_a=2.0, tau_s=8.0), "tabs"),     # tau ~ 8 per cell at the source: few cells see the packets
    "bg_oct6_thick": (_oct(6, 3), {}, lambda s: run_bg(batch=32, seed=s, tau_a=40.0, tau_s=80.0), "tabs"),
    "bg_reg16_thin": (_reg(16), dict(noabsorbed=0), lambda s: run_bg(batch=4, seed=s, tau_a=1.0e-3, tau_s=1.0e-3), ("tabs", "int")),
    "ps_reg16_thin": (_reg(16), dict(no_ps=1), lambda s: run_ps([(8.3, 8.2, 7.9)], batch=24, seed=s, tau_a=2.0e-3, tau_s=1.0e-3), "tabs"),
    # the cone of PS_METHOD 4 (kernel_ASOC.c:377-397); positions in double precision (NX > 100 with LEVELS >= 3)
    "ps_reg12_ext4": (_reg(12), dict(no_ps=1, ps_method=4), lambda s: run_ps([(6.0, 6.0, 30.0)], batch=48, seed=s), "tabs"),
    "bg_oct101_dbl": (lambda: synth.box_cloud(101, 6, 6, levels=3, refine_fraction=0.15), dict(noabsorbed=0), lambda s: run_bg(batch=4, seed=s), "int"),
    # several scattering functions (WITH_MSF), reflecting borders (MIRROR; the expectation is the oracle's
    # mirror_exact variant: the production kernels reflect only the border that was crossed)
    "bg_reg12_msf": (_reg(12), dict(with_abu=1, with_msf=1, ndust=2), lambda s: run_bg_msf(batch=8, seed=s), "tabs"),
    "bg_oct6_msf": (_oct(6, 3), dict(with_abu=1, with_msf=1, ndust=2), lambda s: run_bg_msf(batch=16, seed=s), "tabs"),
    "bg_box_mirror": (lambda: synth.box_cloud(20, 12, 8), dict(mirror=1 + 8 + 32), lambda s: run_bg(batch=6, seed=s), "tabs"),
    "bg_reg12_mirror_abu": (_reg(12), dict(with_abu=1, mirror=2 + 4), lambda s: run_abu(batch=8, seed=s), "tabs"),
    "bg_oct6_mirror": (_oct(6, 3), dict(mirror=32 + 1), lambda s: run_bg(batch=16, seed=s), "tabs"),
    # region of interest: photons recorded on entering ROI (per surface element and direction), the stored field as a source
    "bg_reg12_roisave": (_reg(12), dict(with_roi_save=1, roi=[3, 8, 2, 7, 4, 9], roi_step=1, roi_nside=1),
                         lambda s: with_roi_save(run_bg(batch=8, seed=s)), "roi_save"),
    "ps_oct6_roisave": (_oct(6, 3), dict(no_ps=1, with_roi_save=1, roi=[3, 4, 2, 4, 1, 3], roi_step=1, roi_nside=1),
                        lambda s: with_roi_save(run_ps([(1.3, 1.2, 4.9)], batch=96, seed=s)), "roi_save"),
    "bg_reg12_roisave_tabs": (_reg(12), dict(with_roi_save=1, roi=[3, 8, 2, 7, 4, 9], roi_step=2, roi_nside=2),
                              lambda s: with_roi_save(run_bg(batch=8, seed=s)), "tabs"),
    "roi_reg12_load": (_reg(12), dict(with_roi_load=1, roi_dim=[4, 4, 4], roi_nside=2), lambda s: run_roi_load([4, 4, 4], 2, rounds=1, seed=s), "tabs"),
    "roi_oct6_load": (_oct(6, 3), dict(with_roi_load=1, roi_dim=[3, 3, 3], roi_nside=2), lambda s: run_roi_load([3, 3, 3], 2, rounds=2, seed=s), "tabs"),
}


@pytest.mark.parametrize("name", sorted(STAT_CASES))
def test_packet_streams_statistical_parity(name):
    """Production layout (Philox stream per packet; DDA stepping on regular grids) vs the oracle: per-cell
    chi^2/dof <= 1.1 and total absorbed energy within Monte Carlo noise (and 1e-4 where the noise allows)."""
    from oracle import orc
    from soc_b200 import backend
    make, opts, fac, key = STAT_CASES[name]
    cloud = make()
    K = 16
    keys = key if isinstance(key, tuple) else (key,)
    O = orc.Oracle(cloud, mirror_exact=1, **opts)
    a = _repeat(O, fac, K, keys)
    B = _backend(cloud, backend.RNG_PACKET, **opts)
    if name.endswith("_refgeo"):
        B.dev.set_geometry(1)
    b = _repeat(B, fac, K, keys)
    for kk in keys:          # every accumulator the options enable
        chi2, dof, tot, tot_sigma = chi2_per_dof(b[kk], a[kk], min_rel=1e-4)
        assert dof > (50 if name == "ps_reg16_thick" else (100 if kk == "roi_save" else 300)), "%s/%s: %d cells" % (name, kk, dof)
        assert chi2 <= 1.1, "%s/%s: chi2/dof = %.3f over %d cells" % (name, kk, chi2, dof)
        assert tot <= max(4.0 * tot_sigma, 1e-4), "%s/%s: total energy differs by %.2e (sigma %.2e)" % (name, kk, tot, tot_sigma)
    assert B.counters.reserved[0] == 0, "%s: packets killed by the step guard" % name
    B.close()


SCA_STAT_CASES = {
    "sca_ps_reg16": (_reg(16), dict(no_ps=2), lambda s: run_sca("ps", pspos=[(8.3, 8.3, 8.3), (4.1, 10.7, 12.2)], batch=64, glob=2048, seed=s)),
    "sca_ps_oct8_noffs": (_oct(8, 3), dict(no_ps=1, ffs=0), lambda s: run_sca("ps", pspos=[(4.3, 4.2, 3.9)], batch=96, glob=2048, seed=s)),
    "sca_bg_reg12": (_reg(12), {}, lambda s: run_sca("bg", batch=24, seed=s)),
    "sca_bg_oct6": (_oct(6, 3), {}, lambda s: run_sca("bg", batch=48, dirs=((45.0, 45.0),), seed=s)),
    "sca_hp_reg12": (_reg(12), {}, lambda s: run_sca("hp", batch=64, glob=1024, seed=s)),
    "sca_hp_oct6_w": (_oct(6, 3), dict(hpbg_weighted=1, ffs=0), lambda s: run_sca("hp", batch=128, glob=1024, dirs=((120.0, 200.0),), seed=s)),
    "sca_cl_reg10": (_reg(10), {}, lambda s: run_sca("cl", batch=24, glob=256, seed=s)),
    "sca_cl_oct6_ew": (_oct(6, 3), dict(use_emweight=1), lambda s: run_sca("cl", batch=1, glob=256, emweight=True, dirs=((35.0, 110.0),), seed=s)),
    "sca_ps_reg12_hpobs": (_reg(12), dict(no_ps=1), lambda s: run_sca("ps", pspos=[(6.3, 6.2, 5.9)], batch=64, glob=2048, hp_observer=(14.5, 7.31, 6.63), seed=s)),     # observer outside: 1/d^2 stays bounded
    "sca_bg_oct6_hpobs": (_oct(6, 3), {}, lambda s: run_sca("bg", batch=48, hp_observer=(7.5, 2.03, 3.47), nside=8, seed=s)),
    "sca_bg_reg12_msf": (_reg(12), dict(with_abu=1, with_msf=1, ndust=2), lambda s: run_sca("bg", batch=24, msf=True, seed=s)),
    "sca_bg_reg12_mirror": (_reg(12), dict(mirror=1 + 8), lambda s: run_sca("bg", batch=24, seed=s)),
    "sca_roi_reg12_load": (_reg(12), dict(with_roi_load=1, roi_dim=[4, 4, 4], roi_nside=2), lambda s: run_sca("roi", batch=2, seed=s)),
    # ref_quirks bit 0: the production kernel with the peel-off weight of the shipped SimRAM_HP / SimRAM_CL (HG_TEST branch)
    "sca_hp_reg12_quirk": (_reg(12), dict(ref_quirks=1), lambda s: run_sca("hp", batch=64, glob=1024, seed=s)),
    "sca_cl_oct6_quirk": (_oct(6, 3), dict(ref_quirks=1), lambda s: run_sca("cl", batch=2, glob=256, dirs=((35.0, 110.0),), seed=s)),
}


@pytest.mark.parametrize("name", sorted(SCA_STAT_CASES))
def test_scattered_light_statistical_parity(name):
    """Production scattered-light kernel (Philox per packet, incremental walker) vs the oracle: per-pixel
    chi^2/dof <= 1.1 on the well-sampled pixels and total image flux within noise."""
    from oracle import orc
    from soc_b200 import backend
    make, opts, fac = SCA_STAT_CASES[name]
    cloud = make()
    K = 24
    # On octrees the reference displaces the scattering point with the level of the *next* cell
    # (kernel_ASOC_sca.c:958); the production kernel is geometrically exact, so the expectation is the oracle's
    # exact-level variant (the reference-faithful variant is what the REFSTREAMS/REFGEOMETRY kernels are tested on).
    a = _repeat(orc.Oracle(cloud, sca_exact_level=1, mirror_exact=1, hg_test=opts.get("ref_quirks", 0) & 1, **opts), fac, K, "out").reshape(K, -1)
    B = _backend(cloud, backend.RNG_PACKET, **opts)
    b = _repeat(B, fac, K, "out").reshape(K, -1)
    assert np.isfinite(b).all()
    # Healpix observer: a peel-off direction with an exactly zero component (scattering point level with the observer
    # to the last float bit, ~1 in 1e7 scatterings) makes the reference's GetStep divide by zero and the pixel inf
    # (kernel_ASOC_aux.c:289-296); the oracle restates that, the library clamps the component.  Drop such runs.
    good = np.isfinite(a).all(1)
    assert good.sum() >= K - 6
    a, b = a[good], b[good]
    chi2, dof, tot, tot_sigma = chi2_per_dof(b, a, min_rel=0.02)
    assert dof > (12 if name.endswith("hpobs") else 60)       # a Healpix observer sees the cloud in few bright pixels
    assert chi2 <= 1.1, "%s: chi2/dof = %.3f over %d pixels" % (name, chi2, dof)
    assert tot <= max(4.0 * tot_sigma, 1e-4), "%s: total flux differs by %.2e (sigma %.2e)" % (name, tot, tot_sigma)
    assert B.counters.reserved[0] == 0
    B.close()


@pytest.mark.parametrize("deposit", [1, 2])
def test_accumulation_engines_agree(deposit):
    """Warp-aggregated and shared-memory-tile accumulation give the same sums as plain per-lane RED."""
    from soc_b200 import backend
    make, opts, _ = CASES["ps_reg16_in"]
    cloud = make()
    run = run_ps([(8.3, 8.3, 8.3)], batch=16, glob=4096)
    res = []
    for mode in (0, deposit):
        B = _backend(cloud, backend.RNG_PACKET, no_ps=1)
        B.dev.set_tuning(deposit=mode, refill=8, aggregate_steps=24)
        res.append(run(B)["tabs"].astype(np.float64))
        B.close()
    scale = res[0].max()
    # identical packets; only the order / grouping of the float32 additions differs (the source cell takes one
    # addition per packet, so its float32 sum moves in the 5th digit)
    assert np.abs(res[0] - res[1]).max() <= 1e-4 * scale
    assert abs(res[0].sum() - res[1].sum()) <= 3e-5 * res[0].sum()


@pytest.mark.parametrize("kind", ["bg", "ps"])
def test_sharded_ranks_sum_to_single_rank(kind):
    """Packet q runs on rank q % world with the same Philox stream: the sum over ranks equals the 1-rank
    result up to the order of float additions.  ps: the two-pass point-source launch (tile pass + queue-fed pass)."""
    from soc_b200 import backend
    make, opts, _ = CASES["bg_reg16"]
    cloud = make()
    run = run_bg(batch=3, seed=0.37) if kind == "bg" else run_ps([(7.3, 8.2, 6.7)], batch=20, glob=4096)
    B = _backend(cloud, backend.RNG_PACKET, **(dict(no_ps=1) if kind == "ps" else {}))
    one = run(B)["tabs"].astype(np.float64)
    steps_one = B.counters.steps
    tot = np.zeros_like(one)
    steps = 0
    for r in range(3):
        B.dev.reset_counters()
        B.dev.set_shard(r, 3)
        tot += run(B)["tabs"]
        steps += B.counters.steps
    assert steps == steps_one
    assert np.abs(tot - one).max() <= 2e-5 * one.max()
    B.close()


@pytest.mark.parametrize("kind", ["bg", "ps"])
def test_brick_layout_equals_reference_order(kind):
    """The production kernel on 2x2x2 bricks (DENS and the scratch accumulator permuted on the device) follows the
    same Philox streams through the same cells as on the x-fastest order: TABS / INT agree up to the order of the
    float additions, the work counters agree exactly.  Non-cubic grid so that a wrong stride shows."""
    from soc_b200 import backend
    from soc_b200.formats import Cloud
    from soc_b200 import synth
    nx, ny, nz = 20, 12, 16
    d = synth.plummer_density(24)[2:2 + nz, 6:6 + ny, 2:2 + nx]
    cloud = Cloud(nx, ny, nz, [nx * ny * nz], np.ascontiguousarray(d, np.float32).ravel())
    run = run_bg(batch=3, seed=0.37) if kind == "bg" else run_ps([(9.3, 5.2, 7.7)], batch=50, glob=4096)
    opts = dict(noabsorbed=0) if kind == "bg" else dict(no_ps=1, noabsorbed=0)
    res, steps = [], []
    for layout in (0, 1):
        B = _backend(cloud, backend.RNG_PACKET, **opts)
        B.dev.set_layout(layout)
        out = run(B)
        res.append((out["tabs"].astype(np.float64), out["int"].astype(np.float64)))
        steps.append((B.counters.packets, B.counters.steps, B.counters.scatterings))
        B.close()
    assert steps[0] == steps[1] and steps[0][1] > 0
    for a, b in zip(res[0], res[1]):
        assert np.abs(a - b).max() <= 1e-4 * a.max()
        assert abs(a.sum() - b.sum()) <= 3e-5 * a.sum()


@pytest.mark.parametrize("kind", ["abs", "sca"])
def test_neighbour_table_walker_equals_climb_walker(kind, monkeypatch):
    """Octree production kernels: stepping through the per-cell neighbour table (linkwalk.cuh) follows the same Philox
    streams through the same cells as the climb / cross / descend walker (walk.cuh); only the rounding of the face
    distances differs, so a few packets may part ways at a cell corner."""
    from soc_b200 import backend
    cloud = synth.box_cloud(10, 6, 4, levels=4, refine_fraction=0.25)
    res, cnt = [], []
    for nbr in ("1", "0"):
        monkeypatch.setenv("SOC_NBR", nbr)
        if kind == "abs":
            B = _backend(cloud, backend.RNG_PACKET, noabsorbed=0)
            out = run_bg(batch=12, seed=0.37)(B)["int"]
        else:
            B = _backend(cloud, backend.RNG_PACKET, no_ps=1)
            out = run_sca("ps", pspos=[(4.3, 3.2, 1.7)], batch=24, glob=1024, seed=0.37)(B)["out"]
        res.append(out.astype(np.float64))
        c = B.counters
        cnt.append((c.packets, c.steps, c.scatterings))
        assert c.reserved[0] == 0
        B.close()
    assert cnt[0][0] == cnt[1][0] and abs(int(cnt[0][1]) - int(cnt[1][1])) <= 2e-3 * cnt[1][1]
    a, b = res
    bad = np.abs(a - b) > 1e-5 * np.abs(b).max() + 1e-3 * np.abs(b)
    assert bad.mean() < 0.05, "%d of %d differ" % (bad.sum(), bad.size)
    assert abs(a.sum() - b.sum()) <= 5e-4 * b.sum()


@pytest.mark.parametrize("schedule", ["default", "boxes"])
@pytest.mark.parametrize("kind", ["bg", "ps", "bg_abu", "ps_corner"])
def test_domain_tiled_propagation_equals_whole_grid(kind, schedule, monkeypatch):
    """soc_set_domains: the grid cut into boxes, packets parked with their complete stepping state when they cross an
    interior face.  Same Philox streams, bit-identical paths: the counters agree exactly, TABS / INT up to the order of the
    float additions.  Non-cubic grid and boxes so that a wrong bound or stride shows.
    schedule "default": with this few packets the emission pass is followed by the whole-grid clean-up pass at once;
    "boxes": no clean-up pass (every packet is finished by the box kernels, parked packets staged per warp) and chunks of
    8192 packets, the last 3000 of a chunk carried over into the next one."""
    from soc_b200 import backend
    from soc_b200.formats import Cloud
    if schedule == "boxes":
        monkeypatch.setenv("SOC_DOMAIN_CLEANUP", "0")
        monkeypatch.setenv("SOC_DOMAIN_CHUNK", "8192")
        monkeypatch.setenv("SOC_DOMAIN_CARRY", "3000")
    nx, ny, nz = 24, 16, 12                      # 3 x 2 x 2 boxes of 8 x 8 x 6 cells (boxes have one size with even edges)
    d = synth.plummer_density(28)[6:6 + nz, 6:6 + ny, 2:2 + nx]
    cloud = Cloud(nx, ny, nz, [nx * ny * nz], np.ascontiguousarray(d, np.float32).ravel())
    opts = dict(noabsorbed=0)
    if kind == "bg":
        run = run_bg(batch=6, seed=0.37, tau_s=6.0)
    elif kind == "bg_abu":
        run, opts = run_abu(batch=6, seed=0.41), dict(noabsorbed=0, with_abu=1)
    elif kind == "ps":
        run, opts = run_ps([(9.3, 5.2, 7.7)], batch=50, glob=4096, tau_s=6.0), dict(no_ps=1, noabsorbed=0)
    else:       # the shared-memory tile around a source that sits on the corner of eight boxes
        run, opts = run_ps([(8.01, 7.99, 6.02)], batch=50, glob=4096, tau_s=6.0), dict(no_ps=1, noabsorbed=0)
    res, cnt = [], []
    for edge in (-1, 8):
        B = _backend(cloud, backend.RNG_PACKET, **opts)
        B.dev.set_domains(edge)
        out = run(B)
        res.append((out["tabs"].astype(np.float64), out["int"].astype(np.float64)))
        c = B.counters
        cnt.append((c.packets, c.steps, c.scatterings, c.reserved[0]))
        if edge > 0 and schedule == "boxes":
            assert "domains" in B.dev.last_kernel(), B.dev.last_kernel()
        B.close()
    if schedule == "boxes":
        assert cnt[0] == cnt[1] and cnt[0][1] > 0 and cnt[0][3] == 0, cnt
    else:       # the clean-up pass divides where the lean kernels use rcp.approx at a scattering: a path may flip at a cell face
        assert cnt[0][0] == cnt[1][0] and abs(cnt[0][1] - cnt[1][1]) <= 1e-5 * cnt[0][1] and abs(cnt[0][2] - cnt[1][2]) <= 1e-5 * cnt[0][2] \
            and cnt[0][3] == 0 and cnt[1][3] == 0, cnt
    for a, b in zip(res[0], res[1]):
        assert np.abs(a - b).max() <= 1e-4 * a.max()
        assert abs(a.sum() - b.sum()) <= 3e-5 * a.sum()


@pytest.mark.parametrize("abu", [False, True])
@pytest.mark.parametrize("pos", [(9.3, 7.2, 8.7), (1.4, 14.6, 16.5)])
def test_two_pass_point_source_launch_equals_one_pass(pos, abu, monkeypatch):
    """SOC_TWO_PASS=1: the steps of every point-source packet inside the shared-memory tile in a pass of their own (the lean
    kernel with the tile as its box), the packets parked at the border of the tile with their complete stepping state, the rest
    through the plain-add look-ahead kernel from the queue.  Same Philox streams: same packets, the steps / scatterings agree up
    to paths that flip at a cell face by rounding (lean vs look-ahead kernel), TABS and INT up to that and the order of the
    float additions.  Second position: the tile is clamped to the border of the grid; chunks of 65536 packets.
    abu: per-cell opacities (the (kabs*n, ksca*n) array in both passes)."""
    from soc_b200 import backend
    from soc_b200.formats import Cloud
    nx, ny, nz = 20, 16, 18
    d = synth.plummer_density(24)[2:2 + nz, 4:4 + ny, 2:2 + nx]
    cloud = Cloud(nx, ny, nz, [nx * ny * nz], np.ascontiguousarray(d, np.float32).ravel())
    extra, opts = {}, dict(no_ps=1, noabsorbed=0)
    if abu:
        rng = np.random.default_rng(5)
        k = 6.0 / nx
        opt = np.empty((cloud.CELLS, 2), np.float32)
        opt[:, 0] = k * (0.3 + rng.random(cloud.CELLS))
        opt[:, 1] = k * (0.8 + rng.random(cloud.CELLS))
        extra, opts = dict(opt=opt.reshape(-1)), dict(no_ps=1, noabsorbed=0, with_abu=1)
    run = run_ps([pos], batch=50, glob=4096, tau_s=6.0, **extra)
    monkeypatch.setenv("SOC_DOMAIN_CHUNK", "65536")
    res, cnt = [], []
    for two in ("0", "1"):
        monkeypatch.setenv("SOC_TWO_PASS", two)
        B = _backend(cloud, backend.RNG_PACKET, **opts)
        out = run(B)
        res.append((out["tabs"].astype(np.float64), out["int"].astype(np.float64)))
        c = B.counters
        cnt.append((c.packets, c.steps, c.scatterings, c.reserved[0]))
        assert ("sim_tile_pass_kernel" in B.dev.last_kernel()) == (two == "1"), B.dev.last_kernel()
        B.close()
    assert cnt[0][0] == cnt[1][0] and abs(cnt[0][1] - cnt[1][1]) <= 2e-5 * cnt[0][1] and abs(cnt[0][2] - cnt[1][2]) <= 2e-5 * cnt[0][2] \
        and cnt[0][3] == 0 and cnt[1][3] == 0, cnt
    for a, b in zip(res[0], res[1]):
        assert (np.abs(a - b) > 1e-4 * a.max()).mean() < 1e-3
        assert abs(a.sum() - b.sum()) <= 3e-5 * a.sum()


@pytest.mark.parametrize("kind", ["bg", "ps"])
def test_domain_tiled_propagation_at_512(kind):
    """512^3 (BASELINE.json configs 4/5): the automatic domain mode (2 x 2 x 2 boxes of 256^3) against the whole-grid
    look-ahead kernel (plain adds for the background launch, the shared-memory-tile variant for the point source, which sits
    on the corner of the eight boxes): same packets, same paths."""
    from soc_b200 import backend
    n = 512
    cloud = synth.regular_cloud(n)
    dsc, csc = synth.hg_tables(0.6)
    glob = 8 * cloud.AREA
    kabs, ksca = 3.0 / n, 5.0 / n
    res = []
    for edge in (-1, 0):
        B = _backend(cloud, backend.RNG_PACKET, **(dict(no_ps=1) if kind == "ps" else {}))
        B.dev.set_domains(edge)
        B.zero(0)
        if kind == "ps":
            B.sim_pb(32768, 0, 32768 * 100, 100, 0.7, 0.0, 1.0, abs_=kabs, sca=ksca, dsc=dsc, csc=csc,
                     pspos=np.array([0.5 * n + 0.3] * 3, np.float32), ps=np.ones(1, np.float32))
            assert ("domains" in B.dev.last_kernel()) == (edge == 0) and (edge == 0 or "sim_ahead_kernel<DEP_TILE" in B.dev.last_kernel())
        else:
            B.sim_pb(glob, 1, glob, 1, 0.7, 1.0, 1.0, abs_=kabs, sca=ksca, dsc=dsc, csc=csc)
        c = B.counters
        res.append((B.tabs.astype(np.float64), c.packets, c.steps, c.scatterings, c.reserved[0]))
        B.close()
    a, b = res
    # the last parked packets are finished by the general kernel (true division instead of rcp.approx at a scattering):
    # a path may flip at a cell face by rounding
    assert a[1] == b[1] and abs(a[2] - b[2]) <= 1e-6 * a[2] and abs(a[3] - b[3]) <= 1e-5 * a[3] and a[4] == 0 and b[4] == 0, (a[1:], b[1:])
    # (the cell of the point source takes 3e6 adds: float32 sums in a different grouping)
    assert abs(a[0].sum() - b[0].sum()) <= (2e-6 if kind == "bg" else 3e-5) * a[0].sum()
    assert (np.abs(a[0] - b[0]) > 1e-4 * a[0].max()).mean() < 1e-5


def test_invariants_at_full_size():
    """Size-independent properties at the 256^3 bench size (the oracle would need minutes here):
    no absorption opacity => TABS == 0; the absorbed energy is bounded by the injected energy and grows with
    the opacity; doubling the packet weight doubles every cell exactly (same Philox streams => same paths,
    scaling by 2 is exact in floating point); nothing is killed by the step guard."""
    from soc_b200 import backend, synth
    n = 256
    cloud = synth.regular_cloud(n)
    dsc, csc = synth.hg_tables(0.6)
    B = _backend(cloud, backend.RNG_PACKET)
    glob = 8 * cloud.AREA
    B.zero(0)
    B.sim_pb(glob, 1, glob, 1, 0.3, 1.0, 1.0, abs_=0.0, sca=8.0 / n, dsc=dsc, csc=csc)
    assert float(np.abs(B.tabs).max()) == 0.0
    totals = []
    for kabs in (0.5 / n, 4.0 / n, 64.0 / n):
        B.zero(0)
        B.sim_pb(glob, 1, glob, 1, 0.3, 1.0, 1.0, abs_=kabs, sca=0.0, dsc=dsc, csc=csc)
        totals.append(float(B.tabs.astype(np.float64).sum()))
    assert 0.0 < totals[0] < totals[1] < totals[2] <= glob * (1 + 1e-5)
    B.zero(0)
    B.sim_pb(glob, 1, glob, 1, 0.7, 1.0, 1.0, abs_=3.0 / n, sca=5.0 / n, dsc=dsc, csc=csc)
    one = B.tabs
    B.zero(0)
    B.sim_pb(glob, 1, glob, 1, 0.7, 2.0, 1.0, abs_=3.0 / n, sca=5.0 / n, dsc=dsc, csc=csc)
    two = B.tabs
    assert (one > 0).mean() > 0.99                        # every cell is reached
    # float atomics commute only up to rounding: compare to a few ulp, not bit for bit
    assert np.abs(two - 2.0 * one).max() <= 4e-6 * one.max()
    assert B.counters.reserved[0] == 0
    B.close()


def test_kernel_variants_agree_at_full_size(monkeypatch):
    """256^3, one background launch (3.1e6 packets) through the three regular-grid kernels: lean (SOC_AHEAD=0), look-ahead
    (default) and look-ahead on per-cell opacities that hold the same constants.  Same Philox streams => same paths: the
    results differ only by float rounding (stepped-back scattering positions, kabs*n formed in a different order)."""
    from soc_b200 import backend, synth
    n = 256
    cloud = synth.regular_cloud(n)
    dsc, csc = synth.hg_tables(0.6)
    glob = 8 * cloud.AREA
    kabs, ksca = 3.0 / n, 5.0 / n

    def run(with_abu, **env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        B = _backend(cloud, backend.RNG_PACKET, **(dict(with_abu=1) if with_abu else {}))
        opt = None
        if with_abu:
            opt = np.empty((cloud.CELLS, 2), np.float32)
            opt[:, 0], opt[:, 1] = kabs, ksca
            opt = opt.reshape(-1)
        B.zero(0)
        B.sim_pb(glob, 1, glob, 1, 0.7, 1.0, 1.0, abs_=kabs, sca=ksca, dsc=dsc, csc=csc, opt=opt)
        out, steps = B.tabs.astype(np.float64), B.counters.steps
        B.close()
        for k in env:
            monkeypatch.delenv(k)
        return out, steps

    lean, s0 = run(False, SOC_AHEAD="0")
    ahead, s1 = run(False)
    kappa, s2 = run(True)
    scale = lean.max()
    assert abs(s1 - s0) <= 1e-6 * s0 and abs(s2 - s0) <= 1e-6 * s0          # a path flips at a cell face only by rounding
    for other in (ahead, kappa):
        assert abs(other.sum() - lean.sum()) <= 2e-6 * lean.sum()
        assert (np.abs(other - lean) > 1e-4 * scale).mean() < 1e-4


def test_c_abi_error_paths():
    from soc_b200 import backend
    make, opts, _ = CASES["bg_reg16"]
    dev = backend.Device(0)
    with pytest.raises(backend.SocError):
        dev.sim_pb(1, 10, 1, 0.5, 0.1, 0.1, 1.0, 1.0, 64)           # no grid / params yet
    with pytest.raises(backend.SocError):
        dev.set_params(length=1.0, with_msf=1)                       # unsupported option is rejected loudly
    dev.set_params(length=3.0e16)
    dev.set_grid(make())
    with pytest.raises(backend.SocError):
        dev.sim_pb(1, 10, 1, 0.5, 0.1, 0.1, 1.0, 1.0, 64)           # CSC missing
    with pytest.raises(backend.SocError):
        dev.download(backend.BUF_TABS, 10 ** 9)                      # larger than the buffer
    dev.close()
