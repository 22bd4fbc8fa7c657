"""Oracle (plain-C restatement) against the committed golden vectors that were produced by the
reference's own kernels (tests/golden/make_golden.py).  Runs anywhere, no /root/reference needed."""
import glob
import os

import numpy as np
import pytest

from oracle import orc
from tests.cases import CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_every_case_has_a_fixture():
    have = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz"))}
    assert set(CASES) <= have, sorted(set(CASES) - have)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    make, opts, run = CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    O = orc.Oracle(make(), hg_test=1, maph_literal=1, **opts)      # scattered-light HP / CL kernels as shipped (see soc_oracle.h)
    orc.set_threads(1)
    out = run(O)
    for key in gold.files:
        a, b = out[key].astype(np.float64), gold[key].astype(np.float64)
        assert a.shape == b.shape
        scale = np.abs(b).max()
        if scale == 0.0:
            assert np.abs(a).max() == 0.0
            continue
        bad = np.abs(a - b) > 2e-6 * scale
        assert bad.mean() < 2e-3, "%s/%s: %d of %d differ" % (name, key, bad.sum(), bad.size)
        assert abs(a.sum() - b.sum()) <= 1e-5 * abs(b.sum())
