#!/usr/bin/env python
"""Generates tests/golden/soc_example/: the input files of the reference's own runnable example
(/root/reference/soc_example.zip: my.ini, freq.dat, tmp.dust, bg_intensity.bin, tmp.dsc) as small fixtures, so that
machines without /root/reference (the GPU box) can run BASELINE.json configs[0] through the drivers.

tmp.dsc (880 kB of smooth float32 tables) is stored as the xz-compressed byte planes of the differences of the raw
32-bit patterns along each table row -- bit exact, ~90 kB.  tests/soc_example.py undoes it.  The cloud is not part of
the archive: its make_cloud.py writes a 64^3 cube of unit density, which tests/soc_example.py restates.

    python tests/golden/make_soc_example.py
"""
import lzma
import os
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "soc_example")
ZIP = os.path.join(os.environ.get("SOC_REFERENCE_DIR", "/root/reference"), "soc_example.zip")


def main():
    os.makedirs(OUT, exist_ok=True)
    z = zipfile.ZipFile(ZIP)
    for name in ("my.ini", "freq.dat", "tmp.dust", "bg_intensity.bin"):
        open(os.path.join(OUT, name), "wb").write(z.read(name))
    raw = np.frombuffer(z.read("tmp.dsc"), np.int32)
    nfreq = len(np.loadtxt(os.path.join(OUT, "freq.dat")))
    bins = raw.size // (2 * nfreq)
    assert raw.size == 2 * nfreq * bins
    d = np.diff(raw.reshape(2 * nfreq, bins), axis=1, prepend=np.int32(0)).astype(np.int32)      # wraps like the decoder's cumsum
    planes = np.ascontiguousarray(d.view(np.uint8).reshape(-1, 4).T)
    blob = lzma.compress(planes.tobytes(), preset=9)
    open(os.path.join(OUT, "tmp.dsc.delta.xz"), "wb").write(blob)
    np.array([nfreq, bins], np.int32).tofile(os.path.join(OUT, "tmp.dsc.shape"))
    print("tmp.dsc: %d -> %d bytes (%d frequencies x %d bins)" % (raw.nbytes, len(blob), nfreq, bins))


if __name__ == "__main__":
    main()
