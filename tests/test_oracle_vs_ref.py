"""Pins the plain-C oracle (oracle/soc_oracle.c) against the reference's own kernels compiled in place
from /root/reference (oracle/_ref via oracle/build_ref.py).  Both run single-threaded with identical
MWC64X streams, so the Monte Carlo outputs must agree to float rounding, not just statistically.

Skipped where neither /root/reference nor a prebuilt oracle/_ref library exists (e.g. the GPU box);
tests/test_oracle_golden.py covers those machines with committed vectors."""
import os

import numpy as np
import pytest

from oracle import build_ref, orc, ref
from tests.cases import CASES, MAP_NSIDE

pytestmark = pytest.mark.skipif(not build_ref.reference_available(), reason="/root/reference not present")


def _run(name):
    make, opts, run = CASES[name]
    cloud = make()
    O = orc.Oracle(cloud, hg_test=1, maph_literal=1, **opts)      # the scattered-light HP / CL kernels exactly as shipped (soc_oracle.h)
    R = ref.Reference(cloud, map_nside=MAP_NSIDE.get(name), **opts)
    orc.set_threads(1)          # one thread: work items run in id order, float sums in the same order
    R.set_threads(1)
    return run(O), run(R), O, R


@pytest.mark.parametrize("name", sorted(CASES))
def test_case_matches_reference(name):
    o, r, O, R = _run(name)
    for key in r:
        a, b = o[key].astype(np.float64), r[key].astype(np.float64)
        assert a.shape == b.shape
        assert np.isfinite(a).all() and np.isfinite(b).all(), key
        scale = np.abs(b).max() + 1e-300
        if scale < 1e-290:
            assert np.abs(a).max() < 1e-290
            continue
        # identical streams + identical arithmetic: at most a handful of cells may differ by rounding
        bad = np.abs(a - b) > 2e-6 * scale
        assert bad.mean() < 2e-3, "%s/%s: %d of %d differ (max %.3e of %.3e)" % (
            name, key, bad.sum(), bad.size, np.abs(a - b).max(), scale)
        assert abs(a.sum() - b.sum()) <= 1e-5 * abs(b.sum()) + 1e-300


def test_rng_known_answers():
    """MWC64X stream seeding: oracle == reference for several ids / seeds (mwc64x_rng.cl, skip_mwc.cl)."""
    from tests.cases import CASES
    make, opts, _ = CASES["bg_reg16"]
    R = ref.Reference(make(), **opts)
    for seed in (0.4, 0.123, 0.99999):
        for id_ in (0, 1, 2, 12345, 3145727, 2 ** 25 - 1):
            s1, o1 = orc.rng_stream(seed, id_, 8)
            s2, o2 = ref.rng_stream(R.L, seed, id_, 2 ** 26, 8)
            assert (s1 == s2).all() and (o1 == o2).all()


def test_step_counter_matches_atomic_count():
    """The oracle's cell-step counter equals the number of float atomic adds the reference performs."""
    make, opts, run = CASES["bg_reg16"]
    cloud = make()
    O, R = orc.Oracle(cloud, **opts), ref.Reference(cloud, **opts)
    R.atomic_count(reset=True)          # the shared object (and its counter) may have been used before
    run(O), run(R)
    assert O.counters.steps == R.atomic_count()


def test_emission2_matches_reference():
    """Emission2 (kernel_ASOC_aux.c:862): oracle == reference kernel."""
    make, opts, _ = CASES["bg_oct8_3"]
    cloud = make()
    O, R = orc.Oracle(cloud, **opts), ref.Reference(cloud, **opts)
    rng = np.random.default_rng(2)
    t = (5.0 + 40.0 * rng.random(cloud.CELLS)).astype(np.float32)
    freq = np.logspace(11.5, 14.5, 9).astype(np.float32)
    fabs_ = (1e-6 * (freq / 1e12) ** 1.8).astype(np.float32)
    a, b = O.emission2(30, cloud.CELLS - 9, freq, fabs_, t), R.emission2(30, cloud.CELLS - 9, freq, fabs_, t)
    assert a.shape == b.shape and np.array_equal(a, b)


def test_split_absorbed_matches_reference():
    """split_absorbed (kernel_A2E_MABU_aux.c:3-24, the hand-off of the absorbed file to the dust solver of one species):
    oracle == reference kernel, bit for bit, parent markers (-1e20) included."""
    rng = np.random.default_rng(1)
    cells, nfreq, ndust = 5000, 7, 3
    abu = (0.1 + rng.random((cells, ndust))).astype(np.float32)
    rabs = 1e-21 * (0.5 + rng.random((nfreq, ndust)))
    a = (1e-3 * rng.random((cells, nfreq))).astype(np.float32)
    a[::17] = -1e20
    for idust in range(ndust):
        o, r = orc.split_absorbed(idust, rabs, abu, a), ref.split_absorbed(idust, rabs, abu, a)
        assert r is not None and np.array_equal(o.view(np.uint32), r.view(np.uint32))
    # a single species with abundance 1 is handed its absorptions unchanged; the shares of all species add up to the input
    one = orc.split_absorbed(0, rabs[:, :1], np.ones((cells, 1), np.float32), a)
    assert np.allclose(one, a, rtol=1e-6, atol=0)          # `den` is the float-rounded cross section
    tot = sum(orc.split_absorbed(d, rabs, abu, a).astype(np.float64) * abu[:, d:d + 1] for d in range(ndust))
    ok = a > 0
    assert np.abs(tot[ok] / a[ok] - 1.0).max() < 1e-6
