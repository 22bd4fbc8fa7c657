"""Hand-off of ASOC's absorbed file to a dust solver that treats the species one by one (A2E_MABU.py:660-760 of the
reference): the absorptions of the mixture are split by the species' share of the absorption cross section,
OUT[cell, f] = ABSORBED[cell, f] * RABS[f, idust] / sum_d ABU[cell, d] * RABS[f, d]  (kernel_A2E_MABU_aux.c:3-24),
on the device, in batches of cells like the reference (the absorbed file may be larger than the device).

    python -m soc_b200.a2e_handoff <absorbed file> <rabs.txt: NFREQ rows x NDUST columns> <idust> <out file> [abundance file or # per species]
"""
import sys

import numpy as np

from . import backend as bk
from .formats import Cloud


def split_absorbed_file(absorbed_file, rabs, idust, out_file, abundance_files=None, batch=1 << 22, device_factory=None, ordinal=0):
    """Writes the absorptions of species `idust` in the absorbed-file format (int32 CELLS, NFREQ; float32 [CELLS, NFREQ])."""
    rabs = np.ascontiguousarray(rabs, np.float64)
    nfreq, ndust = rabs.shape
    cells, nf = [int(v) for v in np.fromfile(absorbed_file, np.int32, 2)]
    if nf != nfreq:
        raise ValueError("absorbed file has %d frequencies, RABS %d" % (nf, nfreq))
    src = np.memmap(absorbed_file, dtype=np.float32, mode="r", offset=8, shape=(cells, nfreq))
    abu = np.ones((cells, ndust), np.float32)                      # species without a file: abundance 1 (A2E_MABU.py:672)
    for d, name in enumerate(abundance_files or []):
        if name and name != "#":
            abu[:, d] = np.fromfile(name, np.float32, cells)
    with open(out_file, "wb") as fp:
        np.asarray([cells, nfreq], np.int32).tofile(fp)
        dev = None
        for a in range(0, cells, batch):
            b = min(cells, a + batch)
            if dev is None or b - a != n_dev:
                if dev is not None:
                    dev.close()
                n_dev = b - a
                dev = (device_factory or bk.Device)(ordinal)
                dev.set_params(length=1.0)
                dev.set_grid(Cloud(n_dev, 1, 1, [n_dev], np.ones(n_dev, np.float32)))      # a flat run of cells: only CELLS matters
            dev.upload(bk.BUF_FABS, np.ascontiguousarray(src[a:b]).reshape(-1))
            dev.upload(bk.BUF_ABU, np.ascontiguousarray(abu[a:b]).reshape(-1))
            dev.split_absorbed(idust, rabs, b - a).tofile(fp)
        if dev is not None:
            dev.close()


def main(argv=None):
    argv = sys.argv if argv is None else argv
    if len(argv) < 5:
        print(__doc__)
        return 1
    rabs = np.atleast_2d(np.loadtxt(argv[2]))
    split_absorbed_file(argv[1], rabs, int(argv[3]), argv[4], argv[5:] or None)
    return 0


if __name__ == "__main__":
    sys.exit(main())
