#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REFERENCE's own kernels (compiled in place from
/root/reference through oracle/build_ref.py) on the seeded cases of tests/cases.py, single-threaded.
The vectors let machines without /root/reference (the GPU box) check the oracle against the reference.

    python tests/golden/make_golden.py          # rewrites every fixture
    python tests/golden/make_golden.py NAME...  # only the named cases
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref                      # noqa: E402
from tests.cases import CASES, MAP_NSIDE    # noqa: E402


def main():
    for name in (sys.argv[1:] or sorted(CASES)):
        make, opts, run = CASES[name]
        cloud = make()
        R = ref.Reference(cloud, map_nside=MAP_NSIDE.get(name), **opts)
        R.set_threads(1)
        out = run(R)
        # float16-free, but keep files small: store float32 exactly
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("%-20s %s" % (name, {k: float(np.sum(v, dtype=np.float64)) for k, v in out.items()}))


if __name__ == "__main__":
    main()
